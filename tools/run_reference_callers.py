"""Drives the B200 path through the REFERENCE's own, unmodified callers — the binding INTEGRATION.md describes, executed:

    python tools/run_reference_callers.py [scenario]        (needs a CUDA device AND the reference tree)

/root/reference's receiver.py (SDR_EXECUTIVE.__init__/Run/read_chunk/mode_freq_change, demodulate_data, audio_out) and
params.py (RUN_TIME_PARAMS) are loaded through tests/golden/ref_harness.py with  sys.modules['sig_proc'] =
pysdr_b200.sig_proc , i.e. `dsp.Receiver`, `dsp.signal_generator`, `dsp.up_dn` ... resolve to the CUDA implementation.
The scenario definitions and the wiring are the ones the golden generator uses with the numpy oracle behind the seam, so
the two runs differ ONLY in what stands behind `import sig_proc`; the result is compared with the committed fixture.
The reference tree is read, never written or copied.  On the driver's GPU box there is no reference tree; there the same
fixtures are compared with pysdr_b200.receiver.SDR_EXECUTIVE (tests/test_ref_callers.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(name='am2'):
    from pysdr_b200 import sig_proc as dsp
    from tests import ref_scenarios as rs
    from tests.golden import make_golden_ref_callers as mk, ref_harness as rh
    mods = rh.load(dsp)
    out = mk.run_scenario(mods, name, rs.SCENARIOS[name])
    return dict(iters=int(out['iters']), am=out['am'], players=[out['player%d' % i] for i in range(int(out['n_players']))],
                rb_af=out['rb_af'], baseband_io=out['baseband_io'])


if __name__ == "__main__":
    from tests.util import err_metrics
    name = sys.argv[1] if len(sys.argv) > 1 else 'am2'
    res = run(name)
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_callers.npz"))
    ref = G['%s/am' % name]
    worst = (0.0, np.inf)
    for c in range(res['iters']):
        for i in range(ref.shape[1]):
            rel, snr = err_metrics(res['am'][c][i], ref[c, i])
            worst = (max(worst[0], rel), min(worst[1], snr))
    print("reference receiver.py drove pysdr_b200.sig_proc through scenario %r: %d iterations, worst max-abs rel err %.2e, "
          "worst difference SNR %.1f dB vs the oracle-backed fixture" % (name, res['iters'], worst[0], worst[1]))
