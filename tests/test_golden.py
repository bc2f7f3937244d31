"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle on LCG inputs).
CPU: the oracle still reproduces them.  GPU: the CUDA path matches them without importing the oracle."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, run_case, run_psd
from tests.util import assert_parity, golden_input, lcg_iq

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["cfg1_usb", "cfg5rate_cw_lsb"])
def test_oracle_reproduces_golden(name):
    ref = np.load(os.path.join(G, name + ".npz"))
    got = run_case(CASES[name])
    for k in ref.files:
        if k.startswith("nout"):
            np.testing.assert_array_equal(got[k], ref[k])
        else:
            assert_parity(got[k], ref[k], name + ":" + k, rel_tol=1e-6, snr_min=110)


def test_oracle_reproduces_golden_psd():
    ref = np.load(os.path.join(G, "psd_af_panel.npz"))
    got = run_psd()
    assert_parity(got["psd_lin"], ref["psd_lin"], "psd_lin", rel_tol=1e-6, snr_min=110)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name):
    import torch
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.receiver import receiver_offsets
    c = CASES[name]
    ref = np.load(os.path.join(G, name + ".npz"))
    argv = ['-fc'] + [str(f / 1e3) for f in c['fcs']] + ['-mode'] + c['modes'] + ['-foffset', '100', '-af_bw'] + \
           [str(a / 1e3) for a in c['af']] + ['-fs', str(c['srate'] / 1e6)]
    P = RUN_TIME_PARAMS(argv)
    offs = receiver_offsets(P)
    C = P.IN_CHUNK_SIZE
    x = torch.from_numpy(golden_input(c['chunks'] * C, P.SRATE, offs, c['seed'])).cuda()
    bank = ReceiverBank(P, offs, max_in=C)
    am = [[] for _ in offs]; iq = [[] for _ in offs]; dc = [[] for _ in offs]
    for k in range(c['chunks']):
        a, q, d = bank.process(x[k * C:(k + 1) * C])
        for r in range(len(offs)):
            am[r].append(a[r].cpu().numpy().copy()); iq[r].append(q[r].cpu().numpy().copy()); dc[r].append(d[r].cpu().numpy().copy())
    for r in range(len(offs)):
        np.testing.assert_array_equal([len(a) for a in am[r]], ref["nout%d" % r])       # indexing: bit-exact
        assert_parity(np.concatenate(iq[r]), ref["iq%d" % r], "%s iq%d" % (name, r))
        assert_parity(np.concatenate(am[r]), ref["am%d" % r], "%s am%d" % (name, r))
        assert_parity(np.concatenate(dc[r]), ref["dc%d" % r], "%s dc%d" % (name, r))


@pytest.mark.gpu
def test_cuda_psd_matches_golden():
    import pysdr_b200.sig_proc as dsp
    ref = np.load(os.path.join(G, "psd_af_panel.npz"))
    x = lcg_iq(4096 * 5, 404, scale=0.02).astype(np.complex128)
    x = (x + 0.3 * np.exp(2j * np.pi * 0.0737 * np.arange(len(x)))).astype(np.complex64)
    sp = dsp.spectrum(48., 4096, 8192, 0.5)
    assert_parity(sp.psd_est(x, False), ref["psd_lin"], "psd_lin")
    top = ref["psd_db"] > ref["psd_db"].max() - 60
    assert np.max(np.abs(sp.psd_est(x, True) - ref["psd_db"])[top]) < 2e-3
    assert_parity(sp.waterfall(x, 2, False), ref["wf"], "waterfall")
