#!/usr/bin/env python
"""Secondary measurement (not the headline bench line): BASELINE.json configs[3] — wideband stereo FM at 2.4 MS/s with
pilot recovery and 75 us de-emphasis (WFM2 chain: video FIR + discriminator at the RF rate, 3-row resampler bank on the
multiplex, L/R matrix, common AGC, de-emphasis), 10 s capture resident on the device."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mode", default="WFM2", choices=["WFM", "WFM2"])
    ap.add_argument("--direct-video", action="store_true", help="r01 route: direct-form video FIR through K1")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    import pysdr_b200.sig_proc as dsp
    from pysdr_b200.params import RUN_TIME_PARAMS
    P = RUN_TIME_PARAMS(['-fs', '2.4', '-fc', '100000', '-mode', args.mode, '-af_bw', '15', '-vid_bw', '300'], srate_hz=2.4e6)
    C = int(P.IN_CHUNK_SIZE)
    n_chunks = int(args.seconds * P.SRATE) // C
    P.WFM_MAX_CHUNKS = n_chunks
    P.DEEMPH_US = 75
    P.WFM_DIRECT_VIDEO = args.direct_video
    n = n_chunks * C
    t = torch.arange(n, device="cuda", dtype=torch.float64) / P.SRATE
    L, R = 0.4 * torch.sin(2 * np.pi * 1e3 * t), 0.4 * torch.sin(2 * np.pi * 3e3 * t + 0.5)
    mpx = 0.45 * (L + R) + 0.45 * (L - R) * torch.cos(2 * np.pi * 38e3 * t + 0.6) + 0.1 * torch.cos(2 * np.pi * 19e3 * t + 0.3)
    ph = 2 * np.pi * P.FOFFSET * t + 2 * np.pi * 75e3 * torch.cumsum(mpx, 0) / P.SRATE
    x = (0.3 * torch.exp(1j * ph)).to(torch.complex64)
    del t, L, R, mpx, ph
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1')
    chain = rx._wfm_chain()
    def restart():
        chain.rbank.seek(0)
        if chain.vbank is not None:
            chain.vbank.seek(0); chain.prev2.zero_()
        else:
            chain.vhist.zero_(); chain.vprev2.zero_(); chain.acc = 0

    for _ in range(2):
        restart()
        out, iq = chain.demod_dev(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        restart()
        out, iq = chain.demod_dev(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    lp = int(P.FILT_LEN)
    flops = 8.0 * lp * n + 8.0 * lp * (3 if args.mode == "WFM2" else 1) * out[0].numel()
    print(json.dumps({"workload": "cfg4: %s at 2.4 MS/s -> 48 kHz (1/50), %.1f s capture (%d samples), VIDEO_BW 300 kHz, AF_BW 15 kHz, de-emphasis 75 us"
                                  % (args.mode, n / P.SRATE, n), "ms_per_pass": ms, "Msamples_per_s": n / ms / 1e3,
                      "realtime_factor": (n / P.SRATE) / (ms / 1e3), "audio_samples_per_channel": int(out[0].numel()),
                      "video_stage": "direct-form FIR through K1" if chain.vbank is not None else "overlap-save FFT convolution fused with the discriminator",
                      "direct_form_equivalent_TFLOP_per_s(8 flop/tap)": flops / ms / 1e9, "channels_out": len(out),
                      "hbm_algorithmic_GBps(8B/sample in)": 8.0 * n / ms / 1e6}))


if __name__ == "__main__":
    main()
