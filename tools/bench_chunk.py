#!/usr/bin/env python
"""Secondary measurement: the reference-facing per-chunk calls with HOST numpy buffers, as an unmodified pySDR loop
would make them (reference receiver.py:724-725): dsp.Receiver.demod_data per receiver, and the bank executive that
serves all receivers of a chunk from one upload (ReceiverBank.process_host = one C call: pysdr_bank_process_host).
Reports ms per chunk against the real-time budget, for a pageable numpy chunk (what the reference allocates) and for a
chunk that lives in page-locked memory."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as ge
    ge.build()
    import torch
    import pysdr_b200.sig_proc as dsp
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.receiver import receiver_offsets
    P = RUN_TIME_PARAMS(['-fs', '8', '-fc', '7000', '6500', '7900', '10100', '-mode', 'AM', 'NFM', 'USB', 'CW',
                         '-af_bw', '5', '10', '2', '.5'])
    C = int(P.IN_CHUNK_SIZE)
    rng = np.random.default_rng(0)
    x = ((rng.normal(size=C) + 1j * rng.normal(size=C)) * 0.05).astype(np.complex64)
    xp_t = torch.empty(C, dtype=torch.complex64, pin_memory=True)
    xp = xp_t.numpy()
    xp[:] = x
    offs = receiver_offsets(P)
    rxs = [dsp.Receiver(P, offs[r], r, str(r)) for r in range(4)]
    bank = ReceiverBank(P, offs, max_in=C)
    out = {"workload": "cfg2 geometry, one IN_CHUNK_SIZE chunk (%d samples = %.2f ms of signal) per call, host numpy in/out" % (C, 1e3 * C / P.SRATE)}
    cases = (("4 x Receiver.demod_data (pageable chunk)", lambda: [rx.demod_data(x) for rx in rxs]),
             ("ReceiverBank.process_host (4 RX, one upload, pageable chunk)", lambda: bank.process_host(x)),
             ("ReceiverBank.process_host (4 RX, page-locked chunk)", lambda: bank.process_host(xp)),
             ("ReceiverBank.process_host (4 RX, page-locked chunk, audio only)", lambda: bank.process_host(xp, want_dc=False, want_iq=False)))
    for name, fn in cases:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 300
        for _ in range(k):
            fn()
        dt = (time.perf_counter() - t0) / k
        out[name] = {"ms_per_chunk": dt * 1e3, "Msamples_per_s": C / dt / 1e6, "realtime_factor": (C / P.SRATE) / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
