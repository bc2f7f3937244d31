#!/usr/bin/env python
"""Secondary measurement (not the headline bench line): BASELINE.json configs[4] geometry — N channel receivers on a
9.6 kHz raster on one 10 MS/s stream (3/625 -> 48 kHz), processed in streaming blocks that are resident only while
they are being worked on.  Reports input Msamples/s through all channels, the real-time factor and the FP32 rate of the
K1 contraction (8 flop per complex tap MAC)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_sharded(args):
    """BASELINE configs[4] as a multi-GPU run (torchrun): the capture is split in time over the ranks; every rank serves all
    channels on ITS shard (raster channelizer + audio-rate stages), reading the shard from pinned HOST memory inside the
    timed region, and the AGC carry of all channels crosses the ranks in ONE all-gather of 19 doubles per channel."""
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from pysdr_b200.channelizer import ChannelBank, ShardedChannelBank, raster_offsets
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.synth import synth_iq
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P = RUN_TIME_PARAMS(['-fs', '10', '-mode', 'USB', '-fc', '7000', '-af_bw', '2'])
    C = int(P.IN_CHUNK_SIZE)
    cpr = args.block_chunks
    offs = raster_offsets(args.channels, 9600.0, 0.0)
    modes = [['AM', 'NFM', 'USB', 'CW'][k % 4] for k in range(args.channels)]
    afs = [[5e3, 10e3, 2e3, 500.][k % 4] for k in range(args.channels)]
    cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=(cpr + 1) * C, device=dev,
                     raster=None if args.any_offsets else (offs[0], 9600.0), group=args.group)
    sh = ShardedChannelBank(cb, rank, world, cpr)
    pl = sh.plan
    n_loc = pl['lead'] + pl['n']
    xdev = synth_iq(n_loc, P.SRATE, offs[:4], modes[:4], seed=5, device=dev, n0=pl['first_sample'])
    hx = torch.empty(n_loc, dtype=torch.complex64, pin_memory=True)
    hx.copy_(xdev)
    raw16 = None
    if args.cs16:                                           # the int16 I/Q stream SDR hardware delivers (reference receiver.py:609-617)
        import ctypes
        from pysdr_b200._lib import check, load
        lib = load()
        h16 = torch.empty((n_loc, 2), dtype=torch.int16, pin_memory=True)
        h16.copy_((torch.view_as_real(xdev) * 2048.0).round().clamp_(-32768, 32767).to(torch.int16))
        raw16 = [torch.empty((n_loc, 2), dtype=torch.int16, device=dev) for _ in range(2 if args.overlap else 1)]
    # every step's samples come from pinned host memory inside the timed region.  Double-buffered (default): the copy of step
    # k + 1 runs on a copy stream while step k computes — what a streaming run over an hour-long shard does; --no-overlap: copy,
    # then compute, on one stream (the r02-early number)
    xbufs = [xdev, torch.empty_like(xdev)] if args.overlap else [xdev]
    copy_stream = torch.cuda.Stream(device=dev)
    ev_copied = [torch.cuda.Event() for _ in xbufs]
    ev_done = [torch.cuda.Event() for _ in xbufs]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[i])              # the step that last read this buffer has finished
            if raw16 is not None:                           # 4 bytes per sample over PCIe, scaled to complex64 on the device
                raw16[i].copy_(h16, non_blocking=True)
                check(lib.pysdr_cs16_to_cf32(ctypes.c_void_p(raw16[i].data_ptr()), ctypes.c_void_p(xbufs[i].data_ptr()), n_loc,
                                             1.0 / 2048.0, ctypes.c_void_p(copy_stream.cuda_stream)))
            else:
                xbufs[i].copy_(hx, non_blocking=True)       # this rank's shard (+ warm-up and halo) from pinned host memory
            ev_copied[i].record(copy_stream)

    def run(steps):
        am = None
        if args.overlap:
            issue_copy(0)
            for k in range(steps):
                i = k & 1
                if k + 1 < steps:
                    issue_copy(i ^ 1)
                torch.cuda.current_stream(dev).wait_event(ev_copied[i])
                am, _ = sh.step(xbufs[i])
                ev_done[i].record(torch.cuda.current_stream(dev))
        else:
            for _ in range(steps):
                if raw16 is not None:
                    raw16[0].copy_(h16, non_blocking=True)
                    check(lib.pysdr_cs16_to_cf32(ctypes.c_void_p(raw16[0].data_ptr()), ctypes.c_void_p(xdev.data_ptr()), n_loc,
                                                 1.0 / 2048.0, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
                else:
                    xdev.copy_(hx, non_blocking=True)
                am, _ = sh.step(xdev)
        return am

    for e in ev_done:
        e.record(torch.cuda.current_stream(dev))
    run(args.warmup)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    am = run(args.steps)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        sec = world * pl['n'] / P.SRATE
        print(json.dumps({"workload": "cfg5: %d channels %s, 10 MS/s -> 48 kHz (3/625), %.1f s of signal time-sharded "
                                      "over %d GPUs (%.1f s = %d chunks per rank, resident per rank), H2D of every shard inside the "
                                      "timed region (%s), AGC carry = one all-gather of 19 doubles per channel per rank"
                                      % (args.channels, "at any offsets (K1 = k1_chan, tensor cores)" if args.any_offsets else
                                         "on a 9.6 kHz raster (K1 = wola.cu)", sec, world, pl['n'] / P.SRATE, cpr,
                                         "double-buffered: the next step's copy runs under this step's kernels" if args.overlap else
                                         "copy, then compute, on one stream"),
                          "k1_last": cb.banks[0].k1_last,
                          "n_gpus": world, "ms_per_step": ms, "Msamples_per_s": world * pl['n'] / ms / 1e3,
                          "realtime_factor": sec / (ms / 1e3), "one_hour_capture_s": 3600.0 / (sec / (ms / 1e3)),
                          "h2d_bytes_per_rank_per_step": int(n_loc * (4 if args.cs16 else 8)),
                          "capture_format": "cs16 (int16 I/Q, scaled on the device)" if args.cs16 else "complex64", "collective_bytes_per_rank": int(args.channels * 19 * 8),
                          "audio_samples_per_channel_per_rank": int(am[0].numel())}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sharded", action="store_true", help="multi-GPU time-sharded run (launch with torchrun)")
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--block-chunks", type=int, default=188, help="IN_CHUNK_SIZE chunks per streaming block (188 = 4.0 s)")
    ap.add_argument("--any-offsets", action="store_true", help="sharded mode: no raster, every bank's K1 on the tensor cores (k1_chan)")
    ap.add_argument("--cs16", action="store_true", help="sharded mode: the host capture is int16 I/Q (4 bytes per sample over PCIe)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false", help="sharded mode: copy then compute on one stream")
    ap.add_argument("--group", type=int, default=128, help="receivers per bank (any-offsets mode): 128 = two column groups of 64 "
                    "channels per bank on the tensor-core K1, 96 = one group of 96")
    ap.add_argument("--k1", type=int, default=1, help="0: FP32 tap-stationary K1 (r01/r02-early path), 1: tensor-core K1s where they apply")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    if args.sharded:
        return run_sharded(args)
    import __graft_entry__ as ge
    ge.build()
    from pysdr_b200.channelizer import ChannelBank, raster_offsets
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.synth import synth_iq
    P = RUN_TIME_PARAMS(['-fs', '10', '-mode', 'USB', '-fc', '7000', '-af_bw', '2'])
    C = int(P.IN_CHUNK_SIZE)
    n = args.block_chunks * C
    offs = raster_offsets(args.channels, 9600.0, 0.0)
    modes = [['AM', 'NFM', 'USB', 'CW'][k % 4] for k in range(args.channels)]
    afs = [[5e3, 10e3, 2e3, 500.][k % 4] for k in range(args.channels)]
    x = synth_iq(n, P.SRATE, offs[:4], modes[:4], seed=5, device="cuda")
    cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=n, group=args.group)
    for b in cb.banks:
        b.set_k1_mma(args.k1)
    for _ in range(args.warmup):
        cb.process(x)
    torch.cuda.synchronize()
    l0 = cb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        cb.process(x)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.time() - t0) / args.steps
    ms = e0.elapsed_time(e1) / args.steps
    lp = (int(P.FILT_LEN) + int(P.UP) - 1) // int(P.UP)
    n_out = cb.n_out
    flops = 8.0 * lp * n_out * args.channels
    sec = n / P.SRATE
    # stage times of the same step (CUDA events inside the library, separate pass: the event records break the PDL chain)
    for b in cb.banks:
        b.set_timing(True)
    for _ in range(2):
        cb.process(x)
    tm = [b.get_timing() for b in cb.banks]
    stages = {"k1_ms": sum(t['k1_ms'] / t['calls'] for t in tm), "af_filter_ms": sum(t['front_rest_ms'] / t['calls'] for t in tm),
              "agc_back_ms": sum(t['back_ms'] / t['calls'] for t in tm)}
    cb_k1_last = cb.banks[0].k1_last
    k1_kernel = {0: "k1_generic", 1: "k1_fast (FP32 tap-stationary)", 2: "k1_mma", 3: "k1_chan (tcgen05 split-TF32, channels as columns)"}[cb.banks[0].k1_last]
    # the same channels' baseband IQ through the raster channelizer (wola.cu): K1 only
    from pysdr_b200.channelizer import RasterChannelizer
    rc = RasterChannelizer(P, offs[0], 9600.0, args.channels)
    launches_per_block, n_groups = (cb.launch_count() - l0) // args.steps, len(cb.banks)
    del cb                                                        # its 12 GB of per-bank buffers: the K1-only pass needs none of them
    torch.cuda.empty_cache()
    yw_buf = torch.empty((args.channels, n_out + 8), dtype=torch.complex64, device="cuda")   # allocated once, not per call
    for _ in range(args.warmup):
        rc.process(x, out=yw_buf)
    torch.cuda.synchronize()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    for _ in range(args.steps):
        yw = rc.process(x, out=yw_buf)
    w1.record()
    torch.cuda.synchronize()
    ms_w = w0.elapsed_time(w1) / args.steps
    wola = {"ms_per_block_k1_only": ms_w, "Msamples_per_s": n / ms_w / 1e3, "realtime_factor": sec / (ms_w / 1e3),
            "hbm_algorithmic_GBps(8 B in + 8 B per channel-output)": (8.0 * n + 8.0 * yw.numel()) / ms_w / 1e6}
    print(json.dumps({"wola_k1": wola}))
    del rc, yw, yw_buf
    # whole chain with the raster channelizer in front of the groups' audio-rate stages
    cbr = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=n, raster=(offs[0], 9600.0))
    for _ in range(args.warmup):
        cbr._n0 = 0
        cbr.process(x)
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(args.steps):
        cbr._n0 = 0
        cbr.process(x)
    r1.record()
    torch.cuda.synchronize()
    ms_r = r0.elapsed_time(r1) / args.steps
    print(json.dumps({"raster_mode_whole_chain": {"ms_per_block": ms_r, "Msamples_per_s": n / ms_r / 1e3,
                                                  "realtime_factor": sec / (ms_r / 1e3),
                                                  "one_hour_capture_s_on_1_gpu": 3600.0 / (sec / (ms_r / 1e3))}}))
    del cbr
    print(json.dumps({"workload": "cfg5 geometry: %d channels (any offsets: one K1 contraction per bank), 10 MS/s, 3/625, %.2f s blocks (%d chunks), modes AM/NFM/USB/CW"
                                  % (args.channels, sec, args.block_chunks),
                      "ms_per_block": ms, "wall_ms_per_block": wall * 1e3, "Msamples_per_s": n / ms / 1e3,
                      "realtime_factor": sec / (ms / 1e3), "k1_kernel": k1_kernel, "stage_ms_per_block": stages,
                      "k1_TFLOP_per_s(8 flop/tap)": flops / stages["k1_ms"] / 1e9,
                      "k1_tensor_TFLOP_per_s(3 split-TF32 products)": (3 * flops / stages["k1_ms"] / 1e9) if cb_k1_last == 3 else None,
                      "one_hour_capture_s_on_1_gpu": 3600.0 / (sec / (ms / 1e3)),
                      "gpu_launches_per_block": launches_per_block, "groups": n_groups, "receivers_per_bank": args.group}))


if __name__ == "__main__":
    main()
