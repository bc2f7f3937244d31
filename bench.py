#!/usr/bin/env python
"""bench.py — IQ Msamples/s through the 4-RX demod chain (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            own arm (B200 path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port, MP_SCHEME 3
                                                           = one process per receiver, reference mp.py:146-175)

Workload (config.workload): BASELINE.json configs[1] — 4 independent receivers AM/NFM/USB/CW on a 60 s, 8 MS/s
synthetic complex64 capture (2812 whole IN_CHUNK_SIZE blocks = 479 912 792 samples per GPU).  One step = one pass of
the whole capture through all four receivers (three launches: tensor-core K1, AF filter, fused AGC back kernel).  N>1: one
process per GPU, the capture is N x 60 s long and sharded in time (weak scaling); each rank warms its filter memories on the
chunk preceding its shard and the AGC state crosses the shard boundaries as 19 doubles per receiver per rank, stored by our
own kernels into the later ranks' HBM over NVLink peer memory (--carry nccl: one all-gather of the same bytes instead).
Before anything is timed every rank checks its shard against a single-stream pass (parity_check in the JSON line).
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


def workload_config(n_chunks):
    """The `config` object — identical in the own arm and the reference arm (what differs between the arms lives under
    `arm`, `cpu_baseline.sample` and `e2e.how`)."""
    C = 170666
    return {"workload": "cfg2: 4 independent receivers (AM, NFM, USB, CW) on a 60 s 8 MS/s synthetic IQ capture "
                        "(%d blocks x %d = %d samples per GPU), 8 MS/s -> 48 kHz (3/500), FILT_LEN 1001, AF FIR 1001"
                        % (n_chunks, C, n_chunks * C),
            "l2": "input per step (%.2f GB) exceeds L2; no flush needed" % (n_chunks * C * 8 / 1e9),
            "offsets_khz": FCS_KHZ, "modes": MODES, "af_bw_khz": AF_BW_KHZ}


def host_info():
    model = "?"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    try:
        usable = len(os.sched_getaffinity(0))
    except Exception:
        usable = os.cpu_count() or 1
    return {"cpu_model": model, "nproc": os.cpu_count(), "usable_cores": usable}


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the cores NVML reports as local to its GPU, BEFORE any pinned allocation, so that
    the capture's pages and the staging buffers land on the GPU's own NUMA node (first touch).  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w in range(n_words) for b in range(64) if (int(mask[w]) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"bound": True, "cores": len(allowed), "first": allowed[0], "last": allowed[-1]}
        return {"bound": False, "why": "NVML affinity set empty or outside this process's cpuset"}
    except Exception as e:                     # no NVML / no permission: run unbound and say so
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


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
    """dram bytes per K1 launch from the committed ncu --set full captures, keyed by kernel name."""
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
    """The reference's own CPU implementation of the path = our oracle port (upstream sig_proc is not obtainable, see
    oracle/sig_proc_oracle.py) in the reference's MP_SCHEME 3 shape — one worker process per receiver fed with the same
    chunk sequence (reference mp.py:146-175, receiver.py:726-739) — with the LO taken from a table as the reference
    arranges (params.py:470-471).  To use all the host cores the box has, G = usable_cores // 4 such 4-process groups run
    side by side, each replaying its own segment of the capture (what several pySDR instances on one host would do).
    Each step is a bounded sample of the 2812-block workload, sized from a calibration step so that the whole
    --steps/--warmup run takes about a minute."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import numpy as np
    hi = host_info()
    # headline: ONE group = the parallelism the reference itself has (NUM_RX processes).  --ref-groups G > 1 runs G such
    # groups side by side; the all-cores figure is reported next to the headline under arm.all_cores.
    groups = max(1, min(int(args.ref_groups) if args.ref_groups else 1, 16))
    ctx = mp.get_context("fork")
    from oracle import receiver_oracle as rxo
    Po = rxo.make_P(SRATE_MHZ * 1e6, [f * 1e3 for f in FCS_KHZ], MODES, foffset=100e3, af_bw=[b * 1e3 for b in AF_BW_KHZ])
    offs = [Po.FOFFSET + f - Po.FC[0] for f in Po.FC]
    from pysdr_b200.synth import synth_iq
    C = Po.IN_CHUNK_SIZE
    max_chunks = 192                                                 # per group; segments of one shared synthetic capture
    x = synth_iq(max_chunks * C, Po.SRATE, offs, MODES, seed=1234).numpy()

    def worker(g, irx, conn):
        os.environ["OMP_NUM_THREADS"] = "1"
        from oracle import sig_proc_oracle as dsp
        P = rxo.make_P(SRATE_MHZ * 1e6, [f * 1e3 for f in FCS_KHZ], MODES, foffset=100e3,
                       af_bw=[b * 1e3 for b in AF_BW_KHZ])
        rx = dsp.Receiver(P, offs[irx], irx, str(irx + 1), dtype=np.complex64, fast=True)
        while True:
            msg = conn.recv()
            if msg == 'quit':
                break
            acc = 0.0
            for c in range(int(msg)):
                k = (c + 7 * g) % max_chunks                        # every group walks its own part of the capture
                am = rx.demod_data(x[k * C:(k + 1) * C])
                acc += float(am[0])
            conn.send(acc)

    procs, conns = [], []
    for g in range(groups):
        for irx in range(4):
            a, b = ctx.Pipe()
            p = ctx.Process(target=worker, args=(g, irx, b), daemon=True)
            p.start()
            procs.append(p)
            conns.append(a)

    def step(chunks):
        for c in conns:
            c.send(chunks)
        for c in conns:
            c.recv()

    t0 = time.perf_counter()
    step(8)                                                          # calibration (also warms caches / FFT plans)
    rate = 8 * C * groups / (time.perf_counter() - t0)              # samples/s over all groups
    total_steps = args.steps + max(1, args.warmup)
    chunks = int(args.ref_chunks) if args.ref_chunks else int(max(8, min(max_chunks, 60.0 * rate / (total_steps * C * groups))))
    for _ in range(max(1, args.warmup)):
        step(chunks)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(chunks)
    dt = time.perf_counter() - t0
    for c in conns:
        c.send('quit')
    n = chunks * C * groups
    val = n * args.steps / dt / 1e6
    all_cores = None
    if not args.ref_groups and hi["usable_cores"] >= 8 and not args.no_all_cores:
        # informational: G = usable_cores // 4 independent 4-process groups (several pySDR instances on one host)
        import subprocess
        G = min(hi["usable_cores"] // 4, 16)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-groups", str(G),
                                "--steps", "3", "--warmup", "1", "--no-all-cores"], capture_output=True, text=True, timeout=240)
            sub = json.loads(r.stdout.strip().splitlines()[-1])
            all_cores = {"value": sub["value"], "unit": "Msamples/s", "processes": 4 * G,
                         "note": "%d independent 4-process groups side by side, each on its own segment of the capture" % G}
        except Exception as e:
            all_cores = {"error": "%s: %s" % (type(e).__name__, e)}
    sample = ("%d chunks (%.2f s of the 60 s capture) per 4-process group per step, %d group(s) side by side = %d samples "
              "per step; complex64 numpy/scipy oracle port (upfirdn + fftconvolve), table LO" % (
                  chunks, chunks * C / Po.SRATE, groups, n))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64/f32",
            "data": "synthetic", "config": workload_config(int(args.chunks)),
            "arm": {"parallelism": "MP_SCHEME 3: one CPU process per receiver, x%d groups = %d processes" % (groups, 4 * groups),
                    "bounded_sample": sample, "host": hi, "all_cores": all_cores},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": 4 * groups, "kind": "port", "sample": sample,
                             "cpu_model": hi["cpu_model"], "nproc": hi["nproc"]},
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
    hi = host_info()
    return {"value": done / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port", "cpu_model": hi["cpu_model"],
            "nproc": hi["nproc"],
            "sample": "%d samples (%.1f s of signal) of the same 4-RX workload, complex64 numpy/scipy oracle port with a table "
                      "LO, 1 process (MP_SCHEME 1)" % (done, done / P.SRATE)}


# ------------------------------------------------------------------------------------------------------
def parity_check(P, offs, rank, world, dev, cpr=9, carry="peer"):
    """Correctness carried by the bench line itself: on a small capture of world x 9 blocks every rank compares the audio
    of ITS time shard (the timed code path: filter warm-up, halo in place, O(1) AGC carry over the collective) with the same
    blocks taken from a single-stream pass it runs locally.  N = 1: whole-capture call against chunk-at-a-time calls.
    Returns the worst max-abs relative error and difference SNR over ranks and receivers (gate 2e-5 / 90 dB: the two sides
    differ only by FFT block alignment and the re-associated AGC composition)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.dist import ShardedCapture
    from pysdr_b200.synth import synth_iq
    C = int(P.IN_CHUNK_SIZE)
    n_tot = world * cpr * C
    x = synth_iq(n_tot, P.SRATE, offs, MODES, seed=99, device=dev, block=1 << 18)
    env = torch.ones(n_tot, device=dev)
    env[2 * C:3 * C] = 4.0                                          # a burst whose AGC recovery crosses shard boundaries
    x = x * env
    single = ReceiverBank(P, offs, max_in=(rank + 1) * cpr * C, device=dev)
    single.set_k1_mma(0)                                            # reference side: the FP32 tap-stationary K1
    ref_am, _, _ = single.process(x[:(rank + 1) * cpr * C], want_dc=False)
    m0 = -((-int(P.UP) * rank * cpr * C) // int(P.DOWN))
    ref = [a[m0:].clone() for a in ref_am]
    if world == 1:
        b = ReceiverBank(P, offs, max_in=cpr * C, device=dev)
        b.set_k1_mma(2)                                             # the timed code path: tensor-core K1 (forced at this small size)
        am, _, _ = b.process(x[:cpr * C], want_dc=False)
        got = [a.clone() for a in am]
        k1_kernels = [b.k1_last]
        b2 = ReceiverBank(P, offs, max_in=C, device=dev)
        parts = [[] for _ in offs]
        for c in range(cpr):
            am, _, _ = b2.process(x[c * C:(c + 1) * C], want_dc=False)
            for r in range(len(offs)):
                parts[r].append(am[r].clone())
        ref = [torch.cat(p) for p in parts]
        k1_kernels.append(b2.k1_last)
        del b2
        how = ("whole-capture call (tensor-core K1, k1_last=%d) vs %d chunk-at-a-time calls (FP32 tap-stationary K1, k1_last=%d)"
               % (k1_kernels[0], cpr, k1_kernels[1]))
    else:
        b = ReceiverBank(P, offs, max_in=(cpr + 1) * C, device=dev)
        b.set_k1_mma(2)                                             # the timed code path: tensor-core K1 (forced at this small size)
        sh = ShardedCapture(b, P, rank, world, cpr, carry=carry)
        pl = sh.plan
        am, _, _ = sh.step(x[pl['first_sample']:pl['start'] + pl['n']])
        got = [a.clone() for a in am]
        how = ("each rank's time shard (%d blocks, tensor-core K1 k1_last=%d, O(1) AGC carry: %s) vs a local single-stream pass "
               "(FP32 tap-stationary K1)" % (cpr, b.k1_last, sh.carry_how))
    worst_rel, worst_snr = 0.0, 1e9
    for g, r in zip(got, ref):
        assert g.shape == r.shape, (g.shape, r.shape)
        d = (g.double() - r.double())
        rel = float(d.abs().max() / r.abs().max())
        snr = float(10 * torch.log10((r.double() ** 2).sum() / (d ** 2).sum().clamp_min(1e-300)))
        worst_rel, worst_snr = max(worst_rel, rel), min(worst_snr, snr)
    if world > 1:
        t = torch.tensor([worst_rel, -worst_snr], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst_rel, worst_snr = float(t[0]), -float(t[1])
    ok = worst_rel <= 2e-5 and worst_snr >= 90.0
    res = {"ok": bool(ok), "max_abs_rel_err": worst_rel, "min_diff_snr_db": worst_snr, "blocks_per_rank": cpr, "ranks": world,
           "compared": how, "gate": "rel <= 2e-5 and SNR >= 90 dB"}
    if not ok:
        raise SystemExit("bench.py: parity check failed before timing: %s" % json.dumps(res))
    del single, b
    return res


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
    numa = bind_to_gpu_numa_node(local) if not args.no_numa_bind else {"bound": False, "why": "--no-numa-bind"}
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
    shard = ShardedCapture(bank, P, rank, world, n_chunks, carry=args.carry)   # plan: warm-up chunk + K1 halo for rank > 0
    plan = shard.plan
    warm = plan['warm_chunks']
    xbuf = synth_iq(plan['lead'] + n, P.SRATE, offs, MODES, seed=1234, device=dev, n0=plan['first_sample'])
    x_main = xbuf[plan['lead']:]

    # sharded == single stream, before anything is timed (skipped only for profiler passes, whose lines are never bench values)
    parity = None if args.no_parity else parity_check(P, offs, rank, world, dev, carry=args.carry)

    def step():
        shard.step(xbuf)                                            # front -> AGC summary exchange (N > 1) -> back

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = bank.launches
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
    l1 = bank.launches
    # stage breakdown (roofline of the dominant kernel): a second, separate pass with CUDA events recorded between the
    # kernels on their stream — the event records would break the programmatic dependent launches of the headline loop.
    # The board reaches its 1 kW power cap after ~45 ms of back-to-back steps (sw_power_cap, SM clock 1965 -> ~1600 MHz), which
    # is about where the headline loop ends; a pause lets the power budget recover so that this pass measures the kernels in
    # the same (burst) regime as the headline loop and as MEASURED_PEAKS.json's burst copy figure it is compared with.
    time.sleep(1.0)
    stage_steps = min(args.steps, 10)
    bank.set_timing(True)
    for _ in range(stage_steps):
        step()
    tm = bank.get_timing()
    bank.set_timing(False)
    launches = l1 - l0
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

        # the ceiling of this box: the same pinned capture copied to the device and nothing else, all ranks at once
        dcap = torch.empty(64 * C, dtype=torch.complex64, device=dev)
        def bare_h2d():
            for s0 in range(0, n - 64 * C + 1, 64 * C):
                dcap.copy_(hx[s0:s0 + 64 * C], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        bare_h2d()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            bare_h2d()
        barrier()
        dt_h2d = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([dt_h2d], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_h2d = float(t.item())
        h2d_bytes = (n // (64 * C)) * 64 * C * 8
        del dcap
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
                      "bank.process -> audio D2H to pinned host, per GPU",
               "h2d_ceiling": {"GBps_per_gpu": h2d_bytes / dt_h2d / 1e9, "GBps_all_gpus": world * h2d_bytes / dt_h2d / 1e9,
                               "Msamples_per_s_all_gpus": world * h2d_bytes / 8 / dt_h2d / 1e6,
                               "how": "the same pinned capture copied host->device in the same 64-chunk pieces with no kernels, "
                                      "all ranks simultaneously (max over ranks)"}}
        e2e["frac_of_h2d_ceiling"] = e2e["value"] / e2e["h2d_ceiling"]["Msamples_per_s_all_gpus"]
        del streamer
        # the call the reference's own loop makes (receiver.py:724-725): one IN_CHUNK_SIZE chunk per call, host numpy in,
        # host numpy out, all four receivers of the chunk served from one upload (ReceiverBank.process_host)
        hx_np = hx.numpy()
        pcb = ReceiverBank(P, offs, max_in=C, device=dev)
        n_pc = 256
        for c in range(8):
            pcb.process_host(hx_np[c * C:(c + 1) * C], want_dc=False)
        barrier()
        t0 = time.perf_counter()
        for c in range(n_pc):
            am_pc, _, _ = pcb.process_host(hx_np[c * C:(c + 1) * C], want_dc=False)
        barrier()
        dt_pc = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_pc], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_pc = float(t.item())
        e2e["per_chunk"] = {"value": world * n_pc * C / dt_pc / 1e6, "unit": "Msamples/s", "ms_per_chunk": dt_pc / n_pc * 1e3,
                            "realtime_factor": (C / P.SRATE) / (dt_pc / n_pc), "chunks": n_pc,
                            "how": "ReceiverBank.process_host per 170666-sample chunk (21.3 ms of signal): numpy chunk -> pinned "
                                   "staging -> H2D -> kernels -> D2H -> numpy, synchronous, per GPU"}
        e2e["numa"] = numa
        del hx, pcb, hx_np
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
    k1_last = bank.k1_last
    tr = (tr or {}).get({2: "k1_mma_kernel", 1: "k1_fast_kernel<4,11>"}.get(k1_last, ""), None) if tr else None
    roof = {"bound": "hbm",
            "kernel": {2: "k1_mma_kernel (fused mix + polyphase decimate of 4 RX as a split-TF32 GEMM on tcgen05: TMA -> tensor "
                          "memory A operand, taps resident in shared memory; stream edges on one FP32 warp per CTA)",
                       1: "k1_fast_kernel<4,11> (fused mix + polyphase decimate, 4 RX, FP32 tap-stationary)"}.get(k1_last, "k1_generic_kernel"),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "peak_source": which + ", burst figure",
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * n, "k1_ms_per_launch": k1_ms,
            "traffic": (tr or {}).get("dram_bytes_per_launch_scaled_to", {}).get(str(n)) if tr else None,
            "traffic_note": (tr or {}).get("note") if tr else "no ncu --set full capture committed yet",
            "stage_pass": "%d steps with CUDA events between the kernels, after a 1 s pause (burst power regime, like the headline loop)"
                          % stage_steps,
            "stage_ms_per_step": {"k1": k1_ms, "front_rest(K2: detect + AF FIR)": tm["front_rest_ms"] / max(1, tm["calls"]),
                                  "back(fused: block peaks + state update, AGC scan, gain)": tm["back_ms"] / max(1, tm["calls"])},
            "whole_chain_frac": (ALGO_BYTES_PER_SAMPLE * world * n * args.steps / (ms * 1e-3) / 1e9) / (peak * world)}
    cpu = None if args.no_cpu else cpu_baseline_single()
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64 samples; f64/u64 phase; f64 AGC)", "data": "synthetic",
            "config": workload_config(n_chunks),
            "arm": {"parallelism": ("time-sharded x%d (filter-memory warm-up chunk + a 152-byte AGC summary per receiver per rank, %s)"
                                    % (world, shard.carry_how)) if world > 1 else "single GPU, all 4 receivers share one read",
                    "timed_region": "inputs resident in HBM; CUDA events on the launch stream; max over ranks"},
            "parity_check": parity,
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
    ap.add_argument("--ref-chunks", type=int, default=0, help="reference arm: chunks per group per step (0 = sized from a calibration step)")
    ap.add_argument("--ref-groups", type=int, default=0, help="reference arm: 4-process groups side by side (0 = usable cores // 4)")
    ap.add_argument("--no-all-cores", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="profiler passes only: skip the pre-timing parity check")
    ap.add_argument("--carry", default="peer", choices=["peer", "nccl"],
                    help="N>1: AGC carry over NVLink peer memory written by our own kernels (default) or one NCCL all-gather")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
