"""Seeded random geometries through the whole chain (sample rate, audio rate, filter length, receiver count, modes,
bandwidths, offsets, ragged call sizes): n_out per call bit-exact, baseband and audio within the parity gate.  The
fixed geometry list of test_gpu_parity.py covers the named configs; this one looks for the cases nobody listed."""
import numpy as np
import pytest
import torch

from oracle import receiver_oracle as rxo
from oracle import sig_proc_oracle as odsp
from tests.util import assert_parity

pytestmark = pytest.mark.gpu

RATES = [0.25, 1.0, 1.024, 1.536, 2.0, 2.048, 2.56, 3.0, 4.0, 5.0, 6.0, 8.0, 10.0]       # Tables.py:44-45
FS_OUTS = [12, 24, 48, 96]
NFILTS = [51, 200, 301, 777, 1001, 1501]
MODES = ['AM', 'NFM', 'USB', 'LSB', 'CW', 'IQ', 'AM-Synch']
AFS = [0, 0.5, 2, 3, 5, 10]


@pytest.mark.parametrize("seed", range(14))
def test_random_geometry(seed):
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.receiver import receiver_offsets
    rng = np.random.default_rng(1000 + seed)
    fs = float(rng.choice(RATES))
    fso = int(rng.choice(FS_OUTS))
    if fso * 1e3 > fs * 1e6 / 2:
        fso = 12
    nfilt = int(rng.choice(NFILTS))
    n_rx = int(rng.integers(1, 7))
    modes = [str(rng.choice(MODES)) for _ in range(n_rx)]
    afs = [float(rng.choice(AFS)) for _ in range(n_rx)]
    fcs = [7000.0 + float(rng.uniform(-0.35, 0.35)) * fs * 1e3 for _ in range(n_rx)]
    P = RUN_TIME_PARAMS(['-fs', str(fs), '-fsout', str(fso), '-fc'] + ['%.3f' % f for f in fcs] + ['-mode'] + modes +
                        ['-foffset', '%.3f' % (0.05 * fs * 1e3), '-nfilt', str(nfilt), '-af_bw'] + [str(a) for a in afs])
    Po = rxo.make_P(P.SRATE, list(P.FC), modes, fs_out=fso * 1e3, foffset=0.05 * fs * 1e6,
                    af_bw=[a * 1e3 for a in afs], nfilt=nfilt)            # the oracle receivers get the product's offsets
    assert (P.UP, P.DOWN, P.IN_CHUNK_SIZE, P.FS_OUT) == (Po.UP, Po.DOWN, Po.IN_CHUNK_SIZE, Po.FS_OUT)
    offs = receiver_offsets(P)
    C = P.IN_CHUNK_SIZE
    n_chunks = int(rng.integers(2, 5))
    tail = int(rng.integers(0, C))                                   # ragged final call
    n = n_chunks * C + tail
    t = np.arange(n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.01
    for k, f in enumerate(offs):
        x = x + 0.05 * (1 + 0.4 * np.sin(2 * np.pi * (400.0 + 90 * k) * t / P.SRATE)) * np.exp(2j * np.pi * (f + 300.0) * t / P.SRATE)
    x = x.astype(np.complex64)
    cuts = [0]
    while cuts[-1] < n_chunks * C:                                   # calls of 1..2 whole chunks, then the ragged tail
        cuts.append(min(n_chunks * C, cuts[-1] + int(rng.integers(1, 3)) * C))
    if tail:
        cuts.append(n)
    bank = ReceiverBank(P, offs, max_in=2 * C)
    orx = [odsp.Receiver(Po, offs[r], r, str(r), fast=True) for r in range(n_rx)]
    xd = torch.from_numpy(x).cuda()
    for a, b in zip(cuts[:-1], cuts[1:]):
        am, iq, _ = bank.process(xd[a:b], want_dc=False)
        exp = odsp.n_out_total(b, P.UP, P.DOWN) - odsp.n_out_total(a, P.UP, P.DOWN)
        assert bank.n_out == exp                                     # decimation indexing: bit-exact
        for r in range(n_rx):
            refs = [np.asarray(orx[r].demod_data(x[s:min(s + C, b)])) for s in range(a, b, C)]
            ref = np.concatenate(refs)
            assert len(ref) == exp
            if exp:
                tag = "seed %d fs %.3f->%d nfilt %d rx%d %s [%d,%d)" % (seed, fs, fso, nfilt, r, modes[r], a, b)
                assert_parity(am[r].cpu().numpy(), ref, tag)
