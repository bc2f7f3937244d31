#!/usr/bin/env python
"""Secondary measurement (not the headline bench line): BASELINE.json configs[2] — panadapter PSD/waterfall,
8192-point Hann-windowed FFT, 50 % overlap, K frames averaged per waterfall line, on the 8 MS/s capture."""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--nfft", type=int, default=8192)
    ap.add_argument("--chunk", type=int, default=8192)
    ap.add_argument("--navg", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    import pysdr_b200.sig_proc as dsp
    from pysdr_b200.synth import synth_iq
    n = int(args.seconds * 8e6)
    x = synth_iq(n, 8e6, [1e5, -1.4e6], ['AM', 'CW'], seed=3, device="cuda")
    sp = dsp.spectrum(8000., args.chunk, args.nfft, 0.5)
    for _ in range(3):
        out = sp._lines(x, args.navg, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = sp._lines(x, args.navg, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    frames = 1 + (n - args.chunk) // (args.chunk // 2)
    # yardstick (north_star: "cuFFT used only as a cross-check"): the same frames through cuFFT's batched C2C alone — no window,
    # no |X|^2, no averaging, no dB — on a strided view of the capture (overlapping frames, no copy), in batches that fit memory
    hop = args.chunk // 2
    nb = 4096
    cufft_ms = None
    if args.nfft == args.chunk:
        views = [x.as_strided((min(nb, frames - f0), args.nfft), (hop, 1), f0 * hop) for f0 in range(0, frames, nb)]
        for v in views[:2]:
            torch.fft.fft(v, dim=1)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for v in views:
            torch.fft.fft(v, dim=1)
        c1.record()
        torch.cuda.synchronize()
        cufft_ms = c0.elapsed_time(c1)
    import math
    flops = frames * 5.0 * args.nfft * math.log2(args.nfft)
    peak = 6542.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"workload": "cfg3 PSD: %d-pt Hann FFT (window %d), 50%% overlap, %d frames/line, %.0f s of 8 MS/s IQ" %
                                  (args.nfft, args.chunk, args.navg, args.seconds),
                      "ms_per_pass": ms, "Msamples_per_s": n / ms / 1e3, "lines": int(out.shape[0]),
                      "fft_TFLOP_per_s(5NlogN)": flops / ms / 1e9,
                      "hbm_algorithmic_GBps(8B/sample)": 8.0 * n / ms / 1e6,
                      "frac_of_measured_hbm_peak": 8.0 * n / ms / 1e6 / peak,
                      "cufft_yardstick": None if cufft_ms is None else {
                          "ms_per_pass": cufft_ms, "Msamples_per_s": n / cufft_ms / 1e3,
                          "what": "torch.fft.fft (cuFFT batched C2C, 4096 frames per call) over the same overlapping frames: transforms only, "
                                  "spectra written to HBM and discarded; our kernel also windows, squares, averages and maps to dB"}}))


if __name__ == "__main__":
    main()
