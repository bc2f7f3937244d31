"""Tensor-core K1 (pysdr_b200/csrc/k1_mma.cu): the mix + polyphase decimation of reference receiver.py:235,822,866 as a
split-TF32 GEMM on tcgen05 (samples through TMA into tensor memory), against the oracle and against the FP32
tap-stationary kernel on the same inputs.  Mode 2 forces the tensor-core interior on every call that is long enough for one
tile; mode 0 pins the FP32 kernel."""
import numpy as np
import pytest
import torch

from oracle import receiver_oracle as rxo
from tests.util import assert_parity, make_both

pytestmark = pytest.mark.gpu


def _sig(n, P, offs, seed, amp=0.05):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = ((rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.01).astype(np.complex128)
    for k, f in enumerate(offs):
        x += amp * (1 + 0.4 * np.sin(2 * np.pi * (400.0 + 90 * k) * t / P.SRATE)) * np.exp(2j * np.pi * (f + 300.0) * t / P.SRATE)
    return x.astype(np.complex64)


def _bank(P, max_in, mode):
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.receiver import receiver_offsets
    b = ReceiverBank(P, receiver_offsets(P), max_in=max_in)
    b.set_k1_mma(mode)
    return b


@pytest.mark.parametrize("n_rx,shift", [(4, 0), (4, 1), (1, 0), (3, 1), (6, 0)])
def test_mma_k1_matches_oracle_and_fp32_kernel(n_rx, shift):
    """cfg2 geometry (8 MS/s -> 48 kHz, 3/500, FILT_LEN 1001), 6 chunks in one call: the interior super-periods run on the
    tensor cores, the first and last ones on the tap-stationary kernel.  shift = 1: the capture starts at an odd sample of its
    allocation, so the rows' 16-byte alignment falls the other way (the kernel then reads every row one sample early)."""
    fcs = [1000, 1300, 870, 1210, 940, 1100][:n_rx]
    P, Po = make_both(8, fcs, ['IQ'] * n_rx)
    from pysdr_b200.receiver import receiver_offsets
    C = P.IN_CHUNK_SIZE
    k = 6
    x = _sig(k * C + 1, P, receiver_offsets(P), 11 + n_rx)[shift:shift + k * C]
    xd = torch.from_numpy(_sig(k * C + 1, P, receiver_offsets(P), 11 + n_rx)).cuda()[shift:shift + k * C]
    b2 = _bank(P, k * C, 2)
    assert b2.k1_mma_available, "tensor-core K1 plan not built for the cfg2 geometry"
    _, iq2, _ = b2.process(xd)
    assert b2.k1_last == 2, "tensor-core K1 did not take the call (k1_last=%d)" % b2.k1_last
    got2 = [a.cpu().numpy().copy() for a in iq2]
    b0 = _bank(P, k * C, 0)
    _, iq0, _ = b0.process(xd)
    assert b0.k1_last == 1
    got0 = [a.cpu().numpy().copy() for a in iq0]
    rxo.create_receivers(Po)
    for r in range(n_rx):
        ref = np.concatenate([Po.rx[r].dec.resamp(x[c * C:(c + 1) * C], Po.rx[r].lo) for c in range(k)])
        assert_parity(got2[r], ref, "mma K1 vs oracle rx%d" % r)
        assert_parity(got2[r], got0[r], "mma K1 vs fp32 K1 rx%d" % r, rel_tol=2e-5, snr_min=90)


def test_mma_k1_streaming_calls_and_full_chain():
    """Two consecutive 5-chunk calls in mode 2 (the second call's head tiles read the carried raw history) through the whole
    AM/NFM/USB/CW chain equal the oracle's chunk-at-a-time audio."""
    P, Po = make_both(8, [-500, 700, 1400, 3100], ['AM', 'NFM', 'USB', 'CW'], af_bw_khz=[5, 10, 2, 0.5])
    from pysdr_b200.receiver import receiver_offsets
    C = P.IN_CHUNK_SIZE
    k = 5
    x = _sig(2 * k * C, P, receiver_offsets(P), 3)
    xd = torch.from_numpy(x).cuda()
    b = _bank(P, k * C, 2)
    parts = [[] for _ in range(4)]
    for call in range(2):
        am, _, _ = b.process(xd[call * k * C:(call + 1) * k * C])
        assert b.k1_last == 2
        for r in range(4):
            parts[r].append(am[r].cpu().numpy().copy())
    rxo.create_receivers(Po)
    for r in range(4):
        ref = np.concatenate([Po.rx[r].demod_data(x[c * C:(c + 1) * C]) for c in range(2 * k)])
        assert_parity(np.concatenate(parts[r]), ref, "mma chain rx%d" % r)


def test_mma_k1_impulse_indexing():
    """A unit impulse at a known sample: the tensor-core path must place every tap of the polyphase response at the same
    output index as the FP32 kernel (values agree to the split-TF32 rounding, 1e-6 of the tap)."""
    P, _ = make_both(8, [1000], ['IQ'])
    C = P.IN_CHUNK_SIZE
    n = 4 * C
    for pos in (70001, 3 * C - 7, 2 * C + 250):
        x = torch.zeros(n, dtype=torch.complex64, device="cuda")
        x[pos] = 1.0 + 0.5j
        outs = []
        for mode in (2, 0):
            b = _bank(P, n, mode)
            _, iq, _ = b.process(x)
            outs.append(iq[0].cpu().numpy().copy())
        nz2, nz0 = np.nonzero(outs[0])[0], np.nonzero(outs[1])[0]
        assert nz0.size > 0 and nz2.min() == nz0.min() and nz2.max() == nz0.max(), (pos, nz2.min(), nz0.min(), nz2.max(), nz0.max())
        assert np.max(np.abs(outs[0] - outs[1])) <= 2e-6 * np.max(np.abs(outs[1]))
        assert np.nonzero(outs[1])[0].size == np.nonzero(outs[0])[0].size


def test_mma_k1_many_tiles_per_cta_is_stable():
    """1200 chunks in one call = 3226 tiles of 127 rows on 148 persistent CTAs (22 tiles each): the barrier rings wrap hundreds
    of times.  An earlier version aliased a parity wait (a converter group saw only every other phase of a shared barrier) and
    went wrong about once in five runs at this size; repeat the call and compare with the FP32 kernel every time."""
    P, _ = make_both(8, [1000, 1300, 870, 1210], ['IQ'] * 4)
    from pysdr_b200.receiver import receiver_offsets
    C = P.IN_CHUNK_SIZE
    n = 1200 * C
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda", dtype=torch.float32) * 0.05)
    b0 = _bank(P, n, 0)
    _, iq0, _ = b0.process(x)
    ref = [a.clone() for a in iq0]
    del b0
    b2 = _bank(P, n, 1)                                   # default mode: a call this long takes the tensor-core kernel by itself
    for rep in range(6):
        b2.reset()
        _, iq2, _ = b2.process(x)
        assert b2.k1_last == 2
        for r in range(4):
            e = (iq2[r] - ref[r]).abs().max().item() / ref[r].abs().max().item()
            assert e < 5e-6, (rep, r, e)


@pytest.mark.parametrize("nfilt", [501, 751, 301, 1001, 1401])
def test_mma_k1_other_filter_lengths(nfilt):
    """The piece planner on other tap counts at 8 MS/s -> 48 kHz (3/500): FILT_LEN 501 and 301 give three windows that all lie
    inside the row (no continuation piece), 751 and 1001 split the last window, 1401 (467 taps per phase) makes two windows
    cross — five pieces, more than the kernel's four column groups: the bank must fall back to the FP32 kernel by itself."""
    P, Po = make_both(8, [1000, 1300, 870, 1210], ['IQ'] * 4, nfilt=nfilt)
    from pysdr_b200.receiver import receiver_offsets
    C = P.IN_CHUNK_SIZE
    k = 5
    x = _sig(k * C, P, receiver_offsets(P), nfilt)
    xd = torch.from_numpy(x).cuda()
    b2 = _bank(P, k * C, 2)
    _, iq2, _ = b2.process(xd)
    expect = 1 if nfilt == 1401 else 2
    assert b2.k1_last == expect and b2.k1_mma_available == (expect == 2), (nfilt, b2.k1_last, b2.k1_mma_available)
    rxo.create_receivers(Po)
    for r in range(4):
        ref = np.concatenate([Po.rx[r].dec.resamp(x[c * C:(c + 1) * C], Po.rx[r].lo) for c in range(k)])
        assert_parity(iq2[r].cpu().numpy(), ref, "nfilt %d rx%d" % (nfilt, r))
