"""Time-sharding of one long capture across ranks (one process per GPU, torch.distributed).

The path shards naturally along time: K1 is stateless given the absolute sample index (LO phase
= exact u64 function of n) and a halo of ceil(FILT_LEN/UP)-1 preceding raw samples; the audio-rate FIR memories
(FILT_LEN+1 baseband samples) are rebuilt by processing the chunk(s) just before the shard; the only state
that depends on the whole past is the block AGC — its inputs are the per-block peaks, so ONE all-gather of
n_rx x n_blocks floats per rank lets every rank replay the (tiny, serial) AGC recursion over all earlier blocks.
No IQ samples are ever exchanged.

Host-side planning and the collective live here and are back-end agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
import math

import torch
import torch.distributed as dist


def pll_settle_chunks(P, rel=1.0e-6, bn=50.0, zeta=0.70710678118654752440):
    """Chunks after which the AM-Synch carrier loop (bank.cu: second-order PLL at the audio rate, noise bandwidth bn, damping
    zeta) has forgotten its start-up state to `rel`: the transient decays like exp(-zeta*wn*t), wn = bn / (zeta + 1/(4 zeta))
    per second (the kernel's theta = wn / FS_OUT per sample)."""
    fs_out = float(P.SRATE) * int(P.UP) / int(P.DOWN)
    wn = bn / (zeta + 1.0 / (4.0 * zeta))
    t = -math.log(rel) / (zeta * wn)                                    # seconds
    out_per_chunk = int(P.IN_CHUNK_SIZE) * int(P.UP) / int(P.DOWN)
    return int(math.ceil(t * fs_out / out_per_chunk))


def shard_plan(P, rank, world, chunks_per_rank, min_warm_chunks=0):
    """Where rank's shard starts and what it must read before it.  All quantities in input samples.  min_warm_chunks: warm-up
    chunks wanted beyond what the filter memories need (AM-Synch: the carrier loop's settling time)."""
    C = int(P.IN_CHUNK_SIZE)
    need = (int(P.FILT_LEN) + int(P.UP) - 1) // int(P.UP) - 1           # K1 halo (raw samples)
    hist_out = int(P.FILT_LEN) + 1                                      # AF memory (baseband samples)
    start = rank * chunks_per_rank * C
    if rank == 0:
        warm_chunks, halo = 0, 0
    else:
        warm_in = math.ceil(hist_out * int(P.DOWN) / int(P.UP))         # inputs that produce >= hist_out outputs
        warm_chunks = max(1, math.ceil(warm_in / C), int(min_warm_chunks))
        halo = need
        if warm_chunks * C >= start:                                    # the warm-up reaches back to the stream start:
            warm_chunks, halo = start // C, 0                           # x[<0] = 0, exactly what seek(0) gives
    lead = warm_chunks * C + halo                                       # samples to read before `start`
    return dict(start=start, n=chunks_per_rank * C, warm_chunks=warm_chunks, halo=halo, lead=lead,
                first_sample=start - lead, n_blocks=chunks_per_rank)


def exchange_agc_peaks(peaks, rank, world, group=None):
    """peaks: [n_rx, n_blocks] of this rank -> [n_rx, rank*n_blocks] peaks of all EARLIER blocks (None for rank 0).
    The one collective of the path."""
    if world == 1:
        return None
    n_rx, n_blocks = peaks.shape
    if peaks.is_cuda:
        allp = torch.empty((world, n_rx, n_blocks), dtype=peaks.dtype, device=peaks.device)
        dist.all_gather_into_tensor(allp, peaks.contiguous(), group=group)
    else:
        parts = [torch.empty_like(peaks) for _ in range(world)]
        dist.all_gather(parts, peaks.contiguous(), group=group)
        allp = torch.stack(parts)
    if rank == 0:
        return None
    return allp[:rank].permute(1, 0, 2).reshape(n_rx, rank * n_blocks).contiguous()


AGC_SUMMARY_LEN = 19          # include/pysdr_b200.h PYSDR_AGC_SUMMARY_LEN


def exchange_agc_summaries(own, out, world, group=None):
    """own: float64 [n_rx, AGC_SUMMARY_LEN] summary of this rank's shard (pysdr_bank_agc_summary); out: preallocated
    float64 [world, n_rx, AGC_SUMMARY_LEN].  THE collective of the time-sharded path: one all-gather of 152 bytes per receiver
    per rank, independent of the shard length (the r01 protocol gathered every block peak: O(blocks))."""
    if world == 1:
        out[0].copy_(own)
        return out
    if own.is_cuda:
        dist.all_gather_into_tensor(out, own, group=group)
    else:
        parts = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(parts, own, group=group)
        out.copy_(torch.stack(parts))
    return out


def agc_enter_reference(sums, n_before, ref=0.25, beta=0.1, nb=8, gmax=1.0e4, floor=1.0e-9):
    """Host restatement of agc_enter_kernel (bank.cu) for one receiver: sums = float64 [n_before, AGC_SUMMARY_LEN].
    Returns (gain, ring, k) entering the next shard.  Used by the gloo test and as documentation of the carry."""
    ring, k, gain = [0.0] * nb, 0, 1.0
    for q in range(n_before):
        o = [float(v) for v in sums[q]]
        n = int(o[18])
        for j in range(min(n, 7)):
            ring[k % nb] = o[3 + j]
            k += 1
            want = min(ref / max(max(ring), floor), gmax)
            gain = want if want < gain else beta * want + (1.0 - beta) * gain
        if n > 7:
            gain = min(o[0], o[1] + o[2] * gain)
            for t in range(8):
                ring[(k + (n - 7) - 1 - t) % nb] = o[17 - t]
            k += n - 7
    return gain, ring, k


def agc_summary_reference(peaks, ref=0.25, beta=0.1, gmax=1.0e4, floor=1.0e-9):
    """Host restatement of agc_summary_kernel for one receiver's block peaks (float32 array, n >= 1)."""
    import numpy as np
    p = np.asarray(peaks, np.float32)
    n = len(p)
    A, C, D = 1.0e300, 0.0, 1.0
    for b in range(7, n):
        w = min(ref / max(float(np.max(p[b - 7:b + 1])), floor), gmax)
        A, C, D = min(w, beta * w + (1.0 - beta) * A), (1.0 - beta) * C + beta * w, (1.0 - beta) * D
    o = np.zeros(AGC_SUMMARY_LEN, np.float64)
    o[0:3] = A, C, D
    o[3:3 + min(n, 7)] = p[:7]
    last = p[max(0, n - 8):]
    o[18 - len(last):18] = last
    o[18] = n
    return o


class PeerCarry:
    """The AGC carry between time shards over NVLink peer memory (include/pysdr_b200.h: pysdr_bank_agc_summary_push,
    pysdr_bank_process_back_xchg).  Owns one symmetric-memory buffer per rank (torch.distributed._symmetric_memory supplies
    the allocation and the peer mappings; every byte on it is written and read by OUR kernels).  All ranks must call
    push_and_back the same number of times (the step number is the protocol's sequence number)."""

    def __init__(self, bank, rank, world, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        self.bank, self.rank, self.world = bank, rank, world
        nbytes = bank.lib.pysdr_xchg_bytes(world, bank.n_rx)
        if nbytes <= 0:
            raise ValueError("peer carry: unsupported world size %d / receiver count %d" % (world, bank.n_rx))
        self.buf = symm.empty((nbytes + 7) // 8, dtype=torch.float64, device=bank.device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        torch.cuda.synchronize(bank.device)
        dist.barrier(group=group)                                          # every buffer is zero before anyone's first store
        self.bases = (ctypes.c_uint64 * world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.seq = 0

    def push_and_back(self, skip_blocks, want_dc=False):
        import ctypes
        from ._lib import check
        from .bank import _stream_ptr
        b = self.bank
        self.seq += 1
        if self.rank < self.world - 1:                                     # the last shard has no reader
            check(b.lib.pysdr_bank_agc_summary_push(b.h, int(skip_blocks), self.bases, self.world, self.rank, self.seq, _stream_ptr()))
        check(b.lib.pysdr_bank_process_back_xchg(b.h, self.bases, self.world, self.rank, self.seq, int(skip_blocks),
                                                 ctypes.c_void_p(b._am.data_ptr()),
                                                 ctypes.c_void_p(b._am_dc.data_ptr()) if want_dc else None, b.max_out,
                                                 _stream_ptr()))
        return b.views()

    def shard(self, x, skip_blocks, halo_in_place, want_dc=False):
        """front + push + back in the three launches of a single-GPU step (pysdr_bank_process_shard_xchg)."""
        import ctypes
        from ._lib import check
        from .bank import _stream_ptr
        b = self.bank
        b._check_input(x)
        b.sync_demod()
        self.seq += 1
        n_out = ctypes.c_int64(0)
        check(b.lib.pysdr_bank_process_shard_xchg(b.h, ctypes.c_void_p(x.data_ptr()), x.numel(), 1 if halo_in_place else 0,
                                                  b._iq_copy_ptr(), self.bases, self.world, self.rank, self.seq, int(skip_blocks),
                                                  ctypes.c_void_p(b._am.data_ptr()),
                                                  ctypes.c_void_p(b._am_dc.data_ptr()) if want_dc else None, b.max_out,
                                                  ctypes.byref(n_out), _stream_ptr()))
        b.n_out = n_out.value
        return b.views()


class ShardedCapture:
    """Per-rank driver of one time shard on the GPU bank (used by bench.py for N>1 and by receiver-level tools).

    The warm-up chunk(s) that rebuild the audio-rate filter memory are processed in the SAME call as the shard
    (one K1 launch over [start - warm*C, start + n)); their outputs are simply not returned."""

    def __init__(self, bank, P, rank, world, chunks_per_rank, carry="nccl"):
        """carry = "nccl": one all-gather of the AGC summaries per step; "peer": our own kernels store the summaries straight
        into the later ranks' memory over NVLink and the fused back kernel waits on their flags (PeerCarry) — no collective
        library call on the data path.  "peer" falls back to "nccl" where symmetric memory is not available (CPU/gloo)."""
        self.bank, self.P, self.rank, self.world = bank, P, rank, world
        # AM-Synch: the carrier loop is a nonlinear recurrence with no closed-form hand-off, but it FORGETS: a shard whose
        # warm-up covers the loop's settling time (rel 1e-6: 20 chunks at cfg2's rates) tracks the single-stream loop to the
        # parity tolerance (converged-loop assumption: the carrier is inside the loop's pull-in range, as it is whenever
        # AM-Synch is usable at all; the phase agrees modulo 2 pi, which the detector output does not see).
        self.pll_warm = pll_settle_chunks(P) if any(bank._mode_of(r) == 'AM-Synch' for r in range(bank.n_rx)) else 0
        self.plan = shard_plan(P, rank, world, chunks_per_rank, min_warm_chunks=self.pll_warm)
        dev = bank.device
        w = self.plan['warm_chunks']
        self.peaks_ext = torch.zeros((bank.n_rx, w + self.plan['n_blocks']), dtype=torch.float32, device=dev)
        self.own = torch.zeros((bank.n_rx, self.plan['n_blocks']), dtype=torch.float32, device=dev)
        C = int(P.IN_CHUNK_SIZE)
        up, down = int(P.UP), int(P.DOWN)
        s0 = self.plan['start']
        self.skip_out = (-((-up * s0) // down)) - (-((-up * (s0 - w * C)) // down))     # outputs of the warm-up blocks
        if bank.max_in < self.plan['n'] + w * C:
            raise ValueError("bank.max_in must cover the shard plus its %d warm-up chunk(s)" % w)
        # O(1) carry: needs at least 8 real blocks per shard (the summary's peak buffer); shorter shards gather the peaks
        self.o1 = self.plan['n_blocks'] >= 8
        self.summary = torch.zeros((bank.n_rx, AGC_SUMMARY_LEN), dtype=torch.float64, device=dev)
        self.all_sum = torch.zeros((world, bank.n_rx, AGC_SUMMARY_LEN), dtype=torch.float64, device=dev)
        self.peer = None
        self.carry_how = "ONE NCCL all-gather"
        if carry == "peer" and world > 1 and self.o1 and dev.type == "cuda":
            # every rank must take the same path: agree on whether the peer mapping came up everywhere
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            try:
                self.peer = PeerCarry(bank, rank, world)
            except Exception as e:                                           # no peer access / symmetric memory on this box
                self.peer_error = "%s: %s" % (type(e).__name__, e)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self.carry_how = ("stored by the summary kernel into the later ranks' HBM over NVLink peer memory, flag-waited "
                                  "inside the fused back kernel; no collective call")
            else:
                self.peer = None
                self.carry_how = "ONE NCCL all-gather (peer-memory mapping unavailable on this box)"

    def front(self, xbuf, copy_own=True):
        """K1 + audio-rate filters + block peaks of this shard (and its warm-up).  xbuf: device tensor holding samples
        [first_sample, start+n) of the capture.  Leaves this rank's own block peaks in self.own."""
        p, b = self.plan, self.bank
        C = int(self.P.IN_CHUNK_SIZE)
        w = p['warm_chunks']
        if w:
            b.seek(p['start'] - w * C)
            b.process_front(xbuf[p['halo']:], self.peaks_ext, halo_in_place=p['halo'] > 0)
            if copy_own:
                self.own.copy_(self.peaks_ext[:, w:])
        else:
            b.seek(0)
            b.process_front(xbuf[p['lead']:], self.peaks_ext)
            self.own = self.peaks_ext
        return self.own

    def back(self, prev, want_dc=False):
        """AGC replay over the peaks of ALL earlier blocks (prev, from exchange_agc_peaks; the warm-up blocks' own peaks
        are not valid and are skipped) and gain application.  Returns (am, iq, am_dc) views of this rank's shard."""
        am, iq, dc = self.bank.process_back(prev_peaks=prev, want_dc=want_dc, skip_blocks=self.plan['warm_chunks'])
        k = self.skip_out
        return [a[k:] for a in am], [a[k:] for a in iq], [a[k:] for a in dc]

    def step(self, xbuf, want_dc=False):
        import ctypes
        from ._lib import check
        from .bank import _stream_ptr
        if not self.o1:
            own = self.front(xbuf)
            prev = exchange_agc_peaks(own, self.rank, self.world)
            return self.back(prev, want_dc=want_dc)
        b = self.bank
        if self.world == 1:                                                  # nothing to exchange: the plain whole-capture call
            b.seek(0)
            return b.process(xbuf[self.plan['lead']:], want_dc=want_dc)
        w = self.plan['warm_chunks']
        if self.peer is not None:
            p = self.plan
            if w:
                b.seek(p['start'] - w * int(self.P.IN_CHUNK_SIZE))
                am, iq, dc = self.peer.shard(xbuf[p['halo']:], w, p['halo'] > 0, want_dc)
            else:
                b.seek(0)
                am, iq, dc = self.peer.shard(xbuf[p['lead']:], 0, False, want_dc)
            k = self.skip_out
            return [a[k:] for a in am], [a[k:] for a in iq], [a[k:] for a in dc]
        self.front(xbuf, copy_own=False)
        check(b.lib.pysdr_bank_agc_summary(b.h, w, ctypes.c_void_p(self.summary.data_ptr()), _stream_ptr()))
        exchange_agc_summaries(self.summary, self.all_sum, self.world)       # the one collective of the path
        am, iq, dc = b.process_back_carry(self.all_sum, self.rank, want_dc=want_dc, skip_blocks=w)
        k = self.skip_out
        return [a[k:] for a in am], [a[k:] for a in iq], [a[k:] for a in dc]


# ---- the other natural axis (SURVEY.md 8e axis 1): shard by receiver, no data-path collective -----------------------
def receiver_shard(n_rx, rank, world):
    """Receiver indices owned by `rank`: contiguous, sizes differing by at most one (what MP_SCHEME 3 does with one
    process per receiver, reference receiver.py:726-739, generalised to world ranks)."""
    base, extra = divmod(int(n_rx), int(world))
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def gather_audio(local, n_rx, rank, world, group=None):
    """Optional convenience: collect every receiver's audio on all ranks.  local: dict {irx: 1-D float32 tensor of equal
    length} for the receivers of receiver_shard(n_rx, rank, world).  Returns a [n_rx, n] tensor.  (The data path itself
    needs no collective: each rank can equally well write its own receivers' files.)"""
    per = -(-int(n_rx) // int(world))
    n = next(iter(local.values())).numel() if local else 0
    dev = next(iter(local.values())).device if local else torch.device("cpu")
    nt = torch.tensor([n], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nt, op=dist.ReduceOp.MAX, group=group)
    n = int(nt.item())
    mine = torch.zeros((per, n), dtype=torch.float32, device=dev)
    for j, irx in enumerate(receiver_shard(n_rx, rank, world)):
        mine[j].copy_(local[irx])
    if world == 1:
        return mine[:n_rx]
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    rows = []
    for r in range(world):
        rows.extend(parts[r][j] for j in range(len(receiver_shard(n_rx, r, world))))
    return torch.stack(rows)
