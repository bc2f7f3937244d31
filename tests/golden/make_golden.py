"""Generates the committed golden vectors from the oracle (run from the repo root:  python tests/golden/make_golden.py).

The reference ships no golden vectors for this path and its arithmetic module is absent, so
these fixtures pin OUR oracle's outputs on deterministic (LCG) inputs: the CPU suite checks the oracle still
reproduces them, the GPU suite checks the CUDA path against them without importing the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import receiver_oracle as rxo          # noqa: E402
from oracle import sig_proc_oracle as odsp         # noqa: E402
from tests.util import golden_input, lcg_iq        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "cfg1_usb": dict(srate=2.048e6, fcs=[1000e3], modes=['USB'], af=[2e3], chunks=3, seed=101),
    "cfg2_4rx": dict(srate=8e6, fcs=[-500e3, 700e3, 1400e3, 3100e3], modes=['AM', 'NFM', 'USB', 'CW'],
                     af=[5e3, 10e3, 2e3, 500.], chunks=2, seed=202),
    "cfg5rate_cw_lsb": dict(srate=10e6, fcs=[7000e3, 7030e3], modes=['CW', 'LSB'], af=[500., 3e3], chunks=2, seed=303),
}


def run_case(c):
    mode = c['modes'] if len(c['modes']) > 1 else c['modes'][0]
    af = c['af'] if len(c['af']) > 1 else c['af'][0]
    P = rxo.make_P(c['srate'], c['fcs'], mode, foffset=100e3, af_bw=af)
    offs = [P.FOFFSET + f - P.FC[0] for f in P.FC]
    n = c['chunks'] * P.IN_CHUNK_SIZE
    x = golden_input(n, P.SRATE, offs, c['seed'])
    rxo.create_receivers(P)
    out = {}
    for irx in range(P.NUM_RX):
        am, iq, dc = [], [], []
        for k in range(c['chunks']):
            d = rxo.demodulate_data(P, x[k * P.IN_CHUNK_SIZE:(k + 1) * P.IN_CHUNK_SIZE], irx)
            am.append(P.rx[irx].am.copy()); iq.append(P.rx[irx].iq.copy()); dc.append(np.asarray(d, np.float32))
        out["am%d" % irx] = np.concatenate(am).astype(np.float32)
        out["iq%d" % irx] = np.concatenate(iq).astype(np.complex64)
        out["dc%d" % irx] = np.concatenate(dc).astype(np.float32)
        out["nout%d" % irx] = np.array([len(a) for a in am], np.int64)
    return out


def run_psd():
    x = lcg_iq(4096 * 5, 404, scale=0.02).astype(np.complex128)
    t = np.arange(len(x))
    x = (x + 0.3 * np.exp(2j * np.pi * 0.0737 * t)).astype(np.complex64)
    sp = odsp.spectrum(48., 4096, 8192, 0.5)
    return dict(psd_lin=sp.psd_est(x, False).astype(np.float32), psd_db=sp.psd_est(x, True).astype(np.float32),
                wf=sp.waterfall(x, 2, False).astype(np.float32))


if __name__ == "__main__":
    for name, c in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(c))
        print("wrote", name)
    np.savez_compressed(os.path.join(HERE, "psd_af_panel.npz"), **run_psd())
    print("wrote psd_af_panel")
