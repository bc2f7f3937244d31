#!/usr/bin/env python
"""bench.py — IQ Msamples/s through the 4-RX demod chain (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            own arm (B200 path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port, MP_SCHEME 3
                                                           = one process per receiver, reference mp.py:146-175)

Workload (config.workload): BASELINE.json configs[1] — 4 independent receivers AM/NFM/USB/CW on a 60 s, 8 MS/s
synthetic complex64 capture (2812 whole IN_CHUNK_SIZE blocks = 479 912 792 samples per GPU).  One step = one pass of
the whole capture through all four receivers.  N>1: one process per GPU, the capture is N x 60 s long and sharded in
time (weak scaling); each rank warms its filter memories on the chunk preceding its shard, and the only collective
is the all-gather of per-block AGC peaks (n_rx x 2812 floats per rank).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FCS_KHZ = [-500.0, 700.0, 1400.0, 3100.0]           # offsets {-1.5,-0.3,+0.4,+2.1} MHz around the LO
MODES = ['AM', 'NFM', 'USB', 'CW']
AF_BW_KHZ = [5, 10, 2, 0.5]
SRATE_MHZ = 8
N_CHUNKS = 2812                                      # 60 s at 8 MS/s in whole 170 666-sample blocks
ALGO_BYTES_PER_SAMPLE = 8.0 + 4 * (3.0 / 500.0) * 4  # 8.096 B (SURVEY 8d / BASELINE.md section 3)
METRIC = "IQ Msamples/s through 4-RX demod chain"


def cfg_argv():
    return (['-fs', str(SRATE_MHZ), '-fc'] + [str(f) for f in FCS_KHZ] + ['-mode'] + MODES +
            ['-foffset', '100', '-af_bw'] + [str(b) for b in AF_BW_KHZ])


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per K1 launch from the committed ncu --set full capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.stop_flag = False
        self.sm, self.reasons = [], set()
        self.sm_max = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path = our oracle port (upstream sig_proc is not obtainable,
    see oracle/sig_proc_oracle.py) in the reference's MP_SCHEME 3 shape: one worker process per receiver, the same
    chunk sequence for each, joined per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import numpy as np
    sample_chunks = int(args.ref_chunks)
    ctx = mp.get_context("fork")
    from oracle import receiver_oracle as rxo
    Po = rxo.make_P(SRATE_MHZ * 1e6, [f * 1e3 for f in FCS_KHZ], MODES, foffset=100e3, af_bw=[b * 1e3 for b in AF_BW_KHZ])
    offs = [Po.FOFFSET + f - Po.FC[0] for f in Po.FC]
    from pysdr_b200.synth import synth_iq
    n = sample_chunks * Po.IN_CHUNK_SIZE
    x = synth_iq(n, Po.SRATE, offs, MODES, seed=1234).numpy()

    def worker(irx, conn):
        os.environ["OMP_NUM_THREADS"] = "1"
        from oracle import sig_proc_oracle as dsp
        P = rxo.make_P(SRATE_MHZ * 1e6, [f * 1e3 for f in FCS_KHZ], MODES, foffset=100e3,
                       af_bw=[b * 1e3 for b in AF_BW_KHZ])
        rx = dsp.Receiver(P, offs[irx], irx, str(irx + 1), dtype=np.complex64, fast=True)
        C = P.IN_CHUNK_SIZE
        while True:
            msg = conn.recv()
            if msg == 'quit':
                break
            acc = 0.0
            for c in range(sample_chunks):
                am = rx.demod_data(x[c * C:(c + 1) * C])
                acc += float(am[0])
            conn.send(acc)

    procs, conns = [], []
    for irx in range(4):
        a, b = ctx.Pipe()
        p = ctx.Process(target=worker, args=(irx, b), daemon=True)
        p.start()
        procs.append(p)
        conns.append(a)

    def step():
        for c in conns:
            c.send('go')
        for c in conns:
            c.recv()

    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for c in conns:
        c.send('quit')
    val = n * args.steps / dt / 1e6
    sample = "%d chunks (%.2f s of the 60 s capture, %d samples) per step, complex64 numpy/scipy oracle port" % (
        sample_chunks, n / Po.SRATE, n)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64/f32",
            "data": "synthetic",
            "config": {"workload": "cfg2: 4 RX AM/NFM/USB/CW, 8 MS/s -> 48 kHz, FILT_LEN 1001; bounded sample: " + sample,
                       "parallelism": "MP_SCHEME 3: one CPU process per receiver"},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": 4, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def cpu_baseline_single(seconds_target=12.0):
    """Oracle port, one process / one thread, all receivers sequentially per chunk (= MP_SCHEME 1)."""
    import numpy as np
    from oracle import receiver_oracle as rxo
    from oracle import sig_proc_oracle as dsp
    from pysdr_b200.synth import synth_iq
    P = rxo.make_P(SRATE_MHZ * 1e6, [f * 1e3 for f in FCS_KHZ], MODES, foffset=100e3, af_bw=[b * 1e3 for b in AF_BW_KHZ])
    offs = [P.FOFFSET + f - P.FC[0] for f in P.FC]
    chunks = 24
    n = chunks * P.IN_CHUNK_SIZE
    x = synth_iq(n, P.SRATE, offs, MODES, seed=1234).numpy()
    rx = [dsp.Receiver(P, offs[i], i, str(i + 1), dtype=np.complex64, fast=True) for i in range(4)]
    C = P.IN_CHUNK_SIZE
    done = 0
    t0 = time.perf_counter()
    while True:
        for c in range(chunks):
            for r in rx:
                r.demod_data(x[c * C:(c + 1) * C])
        done += n
        if time.perf_counter() - t0 > seconds_target:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": "%d samples (%.1f s of signal) of the same 4-RX workload, complex64 numpy/scipy oracle port, "
                      "1 process (MP_SCHEME 1)" % (done, done / P.SRATE)}


# ------------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.receiver import receiver_offsets
    from pysdr_b200.synth import synth_iq
    import __graft_entry__ as ge
    ge.build()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = RUN_TIME_PARAMS(cfg_argv())
    C = int(P.IN_CHUNK_SIZE)
    n_chunks = int(args.chunks)
    n = n_chunks * C                                   # samples per GPU per step
    offs = receiver_offsets(P)
    from pysdr_b200.dist import ShardedCapture
    bank = ReceiverBank(P, offs, max_in=n + (C if rank > 0 else 0), device=dev)
    shard = ShardedCapture(bank, P, rank, world, n_chunks)          # plan: warm-up chunk + K1 halo for rank > 0
    plan = shard.plan
    warm = plan['warm_chunks']
    xbuf = synth_iq(plan['lead'] + n, P.SRATE, offs, MODES, seed=1234, device=dev, n0=plan['first_sample'])
    x_main = xbuf[plan['lead']:]

    def step():
        shard.step(xbuf)                                            # front -> all-gather of AGC peaks -> back

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = bank.launches
    bank.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.result()
    ms = e0.elapsed_time(e1)
    tm = bank.get_timing()
    bank.set_timing(False)
    launches = bank.launches - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = world * n * args.steps / (ms * 1e-3) / 1e6

    # ---- e2e: host buffers through the same bank API, H2D and D2H inside the timed region ---------------
    e2e = None
    if not args.no_e2e:
        from pysdr_b200.receiver import ReplayStreamer
        streamer = ReplayStreamer(P, seg_chunks=64, device=dev)              # the public host-buffer API
        hx = torch.empty(n, dtype=torch.complex64, pin_memory=True)
        hx.copy_(x_main)                                                    # untimed: the capture lives on the host

        def e2e_step():
            h_am, _ = streamer.run(hx)                                      # each GPU replays its own host capture
            return float(h_am[0, 0, 0])                                     # host read of the step's result

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(3, min(args.steps, 10))
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        n_out_tot = (P.UP * n) // P.DOWN
        e2e = {"value": world * n * ksteps / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(n * 8),
               "d2h_bytes_per_step": int(4 * n_out_tot * 4), "steps": ksteps,
               "how": "pinned host complex64 capture -> 64-chunk segments double-buffered H2D on a copy stream -> "
                      "bank.process -> audio D2H to pinned host, per GPU"}
        del hx, streamer
        # the same capture as the hardware delivers it: CS16 (reference receiver.py:609-617), 4 bytes per sample over PCIe,
        # converted on the device.  Informational: the headline e2e above stays on the complex64 capture of the config.
        st16 = ReplayStreamer(P, seg_chunks=64, device=dev, fmt='cs16')
        h16 = torch.empty(2 * n, dtype=torch.int16, pin_memory=True)
        for s0 in range(0, n, 1 << 24):                                     # untimed quantisation, in pieces
            s1 = min(n, s0 + (1 << 24))
            q = torch.view_as_real(x_main[s0:s1]).mul(2048.0).round_().clamp_(-2048, 2047).to(torch.int16)
            h16[2 * s0:2 * s1].copy_(q.reshape(-1))
            del q
        for _ in range(2):
            st16.run(h16)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            float(st16.run(h16)[0][0, 0, 0])
        barrier()
        dt16 = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt16], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt16 = float(t.item())
        e2e["cs16_source"] = {"value": world * n * ksteps / dt16 / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(n * 4),
                              "note": "same path fed with an int16 I/Q capture (SDR hardware format), scaled on the device"}
        del h16, st16

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, which = measured_peak()
    k1_ms = tm["k1_ms"] / tm["calls"] if tm["calls"] else None      # rank 0 has no warm-up call: calls == steps
    achieved = ALGO_BYTES_PER_SAMPLE * n / (k1_ms * 1e-3) / 1e9 if k1_ms else None
    tr = ncu_traffic()
    roof = {"bound": "hbm", "kernel": "k1_fast_kernel<4,11> (fused mix + polyphase decimate, 4 RX)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "peak_source": which + ", burst figure",
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * n, "k1_ms_per_launch": k1_ms,
            "traffic": (tr or {}).get("dram_bytes_per_launch_scaled_to", {}).get(str(n)) if tr else None,
            "traffic_note": (tr or {}).get("note") if tr else "no ncu --set full capture committed yet",
            "stage_ms_per_step": {"k1": k1_ms, "front_rest(K2 detect+AF FIR+peaks+rolls)": tm["front_rest_ms"] / max(1, tm["calls"]),
                                  "back(AGC scan+apply)": tm["back_ms"] / max(1, tm["calls"])},
            "whole_chain_frac": (ALGO_BYTES_PER_SAMPLE * world * n * args.steps / (ms * 1e-3) / 1e9) / (peak * world)}
    cpu = None if args.no_cpu else cpu_baseline_single()
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64 samples; f64/u64 phase; f64 AGC)", "data": "synthetic",
            "config": {"workload": "cfg2: 4 independent receivers (AM, NFM, USB, CW) on a 60 s 8 MS/s synthetic IQ capture "
                                   "per GPU (%d blocks x %d = %d samples), 8 MS/s -> 48 kHz (3/500), FILT_LEN 1001, AF FIR 1001"
                                   % (n_chunks, C, n),
                       "parallelism": "time-sharded x%d (filter-memory warm-up chunk + all-gather of AGC block peaks)" % world
                       if world > 1 else "single GPU, all 4 receivers share one read",
                       "l2": "input per step (%.2f GB) exceeds L2; no flush needed" % (n * 8 / 1e9),
                       "timed_region": "inputs resident in HBM; CUDA events on the launch stream; max over ranks"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was diverted."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)                 # keep the driver-facing stdout for the JSON line only
    os.dup2(2, 1)                            # anything else printed to fd 1 (e.g. "NCCL version ...") goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--chunks", type=int, default=N_CHUNKS, help="IN_CHUNK_SIZE blocks per GPU per step")
    ap.add_argument("--ref-chunks", type=int, default=96, help="reference arm: chunks per step (bounded sample)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
