#!/usr/bin/env python
"""Secondary measurement: BASELINE.json configs[0] — the reference's own CPU-runnable case: 10 s of 2.048 MS/s complex64
IQ through ONE USB receiver to 48 kHz audio (3/128, FILT_LEN 1001), resident on the device.  (CPU figures come from
bench.py's cpu_baseline / --impl reference legs, the only places that may time the oracle.)"""
import json
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as ge
    ge.build()
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.receiver import receiver_offsets
    from pysdr_b200.synth import synth_iq
    P = RUN_TIME_PARAMS(['-fs', '2.048', '-fc', '1000', '-mode', 'USB', '-af_bw', '2', '-foffset', '100'])
    C = int(P.IN_CHUNK_SIZE)
    n_chunks = int(10 * P.SRATE) // C
    n = n_chunks * C
    offs = receiver_offsets(P)
    x = synth_iq(n, P.SRATE, offs, ['USB'], seed=11, device="cuda")
    bank = ReceiverBank(P, offs, max_in=n)
    for _ in range(3):
        bank.seek(0); bank.process(x, want_dc=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = 20
    for _ in range(k):
        bank.seek(0); am, _, _ = bank.process(x, want_dc=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    lp = (int(P.FILT_LEN) + int(P.UP) - 1) // int(P.UP)
    print(json.dumps({"workload": "cfg1: 1 RX USB, 2.048 MS/s -> 48 kHz (3/128), %d samples (%.2f s)" % (n, n / P.SRATE),
                      "ms_per_pass": ms, "Msamples_per_s": n / ms / 1e3, "realtime_factor": (n / P.SRATE) / (ms / 1e3),
                      "k1_variant": bank.k1_variant, "k1_TFLOP_per_s(8 flop/tap)": 8.0 * lp * bank.n_out / ms / 1e9,
                      "hbm_algorithmic_GBps": 8.094 * n / ms / 1e6}))


if __name__ == "__main__":
    main()
