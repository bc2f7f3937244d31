"""Many-channel receiver bank (BASELINE config 5: 1024 simultaneous channel receivers on one 10 MS/s stream).

The reference's surface stops at MAX_RX = 6 (params.py:33); this is the north-star extension.  Channels are served in
groups of up to 128 receivers, one ReceiverBank (= one set of K1/K2 launches) per group, all groups reading the SAME
device-resident block of IQ.  Every number still comes from libpysdr_b200.so; this class is only the loop over
groups.  At 3/625 the contraction is compute-bound (13 kFLOP per input sample for 1024 channels, 474 flop/B): banks of 16
or more channels run it on the tensor cores (k1_chan.cu: split-TF32 GEMM on tcgen05, channels as columns, any offsets);
on a uniform raster wola.cu replaces it by one windowing pass + one inverse DFT per output instant."""
import copy

import numpy as np
import torch

from . import design
from .bank import ReceiverBank


class ChannelBank:
    GROUP = 128          # receivers per ReceiverBank (PYSDR_MAX_RX): 1024 channels = 8 sets of audio-rate launches per block

    def __init__(self, P, offsets_hz, modes, af_bw=0.0, bfo=0.0, max_in=None, device=None, raster=None, group=None):
        """raster=(f0_hz, df_hz): the offsets are f0 + c*df on a raster wola.cu serves (df/fs = a/3125): the baseband of all
        channels then comes from ONE RasterChannelizer pass per block, written straight into the groups' (shared) complex
        memory, and the groups only run their audio-rate stages."""
        n = len(offsets_hz)
        modes = list(modes) if isinstance(modes, (list, tuple)) else [modes] * n
        af_bw = list(af_bw) if isinstance(af_bw, (list, tuple, np.ndarray)) else [af_bw] * n
        bfo = list(bfo) if isinstance(bfo, (list, tuple, np.ndarray)) else [bfo] * n
        if not (len(modes) == len(af_bw) == len(bfo) == n):
            raise ValueError("one mode / AF bandwidth / BFO per channel")
        self.P, self.n_ch = P, n
        if group is not None:
            self.GROUP = int(group)
        self.banks, self.slices = [], []
        for g0 in range(0, n, self.GROUP):
            g1 = min(n, g0 + self.GROUP)
            Pg = copy.copy(P)
            Pg.MODE, Pg.AF_BW, Pg.BFO = modes[g0:g1], af_bw[g0:g1], bfo[g0:g1]
            Pg.AF_FILTER_NUM = None
            Pg.NUM_RX = g1 - g0
            self.banks.append(ReceiverBank(Pg, list(offsets_hz[g0:g1]), max_in=max_in, device=device))
            self.slices.append((g0, g1))
        self.device = self.banks[0].device
        self.n_out = 0
        self.raster = None
        if raster is not None:
            import ctypes
            from ._lib import check
            f0, df = raster
            if any(abs(offsets_hz[c] - (f0 + c * df)) > 1e-6 for c in range(n)):
                raise ValueError("offsets are not on the given raster")
            if 'AM-Synch' in modes:
                raise ValueError("raster mode does not serve AM-Synch channels (their rx.iq needs K1's separate copy)")
            self.raster = RasterChannelizer(P, f0, df, n, device=self.device)
            b0 = self.banks[0]
            self._hc = b0._hc
            stride = (self._hc + b0.max_out + 3) // 2 * 2
            self._C = torch.zeros((n, stride), dtype=torch.complex64, device=self.device)      # one memory for all groups
            for b, (g0, g1) in zip(self.banks, self.slices):
                check(b.lib.pysdr_bank_adopt_c_memory(b.h, ctypes.c_void_p(self._C[g0].data_ptr()), stride))
                check(b.lib.pysdr_bank_set_k1_external(b.h, 1))
                b._cmem = self._C[g0:g1]
            self._x_hist = torch.zeros(self.raster.lp - 1, dtype=torch.complex64, device=self.device)
            self._n0 = 0

    def process(self, x, want_dc=False):
        """x: device complex64 block (whole IN_CHUNK_SIZE chunks).  Returns (am, iq): lists of n_ch device views, valid
        until the next call."""
        am, iq = [], []
        self.banks[0]._check_input(x)                                 # dtype / device / contiguity / capacity, once per call
        if self.raster is not None:                                   # K1 of every channel in one pass, into the shared memory
            self.raster.process(x, n0=self._n0, hist=self._x_hist, out=self._C, out_col=self._hc)
        if want_dc:
            for b in self.banks:
                a, q, _ = b.process(x, want_dc=True)
                am.extend(a)
                iq.extend(q)
            self.n_out = self.banks[0].n_out
            self._advance(x)
            return am, iq
        # audio only: the banks are driven with prepared arguments (one ctypes call each) and the per-channel views are built
        # once per output length — with K1 on the tensor cores (k1_chan.cu, or wola.cu in raster mode) a block of 1024 channels
        # is ~5 ms of GPU time, and building 3 x 1024 tensor views per block costs more than that on the host
        import ctypes
        from ._lib import check
        if getattr(self, '_fast_args', None) is None:                   # (re)built after construction and after invalidate()
            self._fast_args = []
            for b in self.banks:
                b.sync_demod()
                self._fast_args.append((b.lib.pysdr_bank_process, b.h, b._iq_copy_ptr(), ctypes.c_void_p(b._am.data_ptr()), b.max_out))
            self._views_for = None
        xp, n_in = ctypes.c_void_p(x.data_ptr()), x.numel()
        n_out = ctypes.c_int64(0)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        for fn, h, iq_p, am_p, max_out in self._fast_args:
            check(fn(h, xp, n_in, 0, iq_p, am_p, None, max_out, ctypes.byref(n_out), st))
        self.n_out = n_out.value
        self._advance(x)                                              # only once every bank has taken the block
        if self._views_for != self.n_out:
            self._views = ([], [])
            for b in self.banks:
                b.n_out = self.n_out
                a, q, _ = b.views()
                self._views[0].extend(a)
                self._views[1].extend(q)
            self._views_for = self.n_out
        return list(self._views[0]), list(self._views[1])      # the cached views, in lists the caller may keep or edit

    def _advance(self, x):
        """Raster mode: the channelizer's own stream position and raw history move once all banks have processed the block, so
        a failing bank call leaves the two in step."""
        if self.raster is None:
            return
        keep = self._x_hist.numel()
        if x.numel() >= keep:
            self._x_hist.copy_(x[x.numel() - keep:])
        else:
            self._x_hist.copy_(torch.cat((self._x_hist[x.numel():], x)))
        self._n0 += x.numel()

    def invalidate(self):
        """Call after changing MODE / AF_BW / BFO on the banks' parameter objects: the audio-only path (want_dc=False) caches
        the per-bank call arguments and re-reads the parameters (ReceiverBank.sync_demod) only when they are rebuilt."""
        self._fast_args = None

    def launch_count(self):
        return sum(b.lib.pysdr_bank_launch_count(b.h) for b in self.banks)


def raster_offsets(n_ch, spacing_hz, centre_hz=0.0):
    """n_ch offsets on a uniform raster centred on centre_hz (config 5: 1024 channels, 9.6 kHz apart)."""
    return [centre_hz + spacing_hz * (k - (n_ch - 1) / 2.0) for k in range(n_ch)]


class ShardedChannelBank:
    """Config 5's multi-GPU shape: the capture is split in time across ranks, every rank runs all channel groups on its
    shard, and the AGC carry of ALL channels travels in ONE all-gather of [n_ch, n_blocks] block peaks."""

    def __init__(self, cb, rank, world, chunks_per_rank):
        from .dist import ShardedCapture
        self.cb, self.rank, self.world = cb, rank, world
        self.shards = [ShardedCapture(b, b.P, rank, world, chunks_per_rank) for b in cb.banks]
        self.plan = self.shards[0].plan

    def step(self, xbuf):
        """xbuf: device samples [first_sample, start+n) of this rank's shard.  Returns (am, iq) lists over channels."""
        import ctypes
        import torch.distributed as dist
        from ._lib import check
        from .bank import _stream_ptr
        from .dist import AGC_SUMMARY_LEN, exchange_agc_peaks
        cb = self.cb
        if cb.raster is not None:                                        # K1 of every channel for shard + warm-up in one pass
            p = self.plan
            C = int(cb.P.IN_CHUNK_SIZE)
            n0 = p['start'] - p['warm_chunks'] * C
            cb.raster.process(xbuf[p['halo']:], n0=n0, hist=xbuf[:p['halo']], out=cb._C, out_col=cb._hc)
        am, iq = [], []
        if not self.shards[0].o1:                                        # short shards: gather every block peak
            own = torch.cat([sh.front(xbuf) for sh in self.shards])      # [n_ch, n_blocks]
            prev = exchange_agc_peaks(own, self.rank, self.world)
            for sh, (g0, g1) in zip(self.shards, cb.slices):
                a, q, _ = sh.back(None if prev is None else prev[g0:g1].contiguous())
                am.extend(a)
                iq.extend(q)
            return am, iq
        # O(1) carry: one all-gather of 19 doubles per channel per rank (152 KB per rank at 1024 channels, whatever the shard
        # length; the peak gather was 691 MB at config-5 scale), then every bank enters from the summaries of the earlier ranks.
        # The banks are driven with prepared ctypes arguments and the per-channel views are built once per output length: the
        # generic per-bank Python path (parameter re-reads, 3 x 128 tensor views and as many slices per bank) cost ~16 ms per
        # step for 1024 channels, three times the GPU work.  After changing MODE / AF_BW / BFO call invalidate().
        G = cb.GROUP
        dev = cb.device
        w = self.plan['warm_chunks']
        p = self.plan
        C = int(cb.P.IN_CHUNK_SIZE)
        if getattr(self, '_fast', None) is None:
            nb = len(cb.banks)
            self._sum = torch.zeros((nb, G, AGC_SUMMARY_LEN), dtype=torch.float64, device=dev)
            self._all = torch.zeros((self.world, nb, G, AGC_SUMMARY_LEN), dtype=torch.float64, device=dev)
            self._per_bank = torch.zeros((nb, self.world, G, AGC_SUMMARY_LEN), dtype=torch.float64, device=dev)
            self._fast = []
            for k, sh in enumerate(self.shards):
                b = sh.bank
                b.sync_demod()
                compact = None if b.n_rx == G else torch.zeros((self.world, b.n_rx, AGC_SUMMARY_LEN), dtype=torch.float64, device=dev)
                self._fast.append((b, b.lib, b.h, b._iq_copy_ptr(), ctypes.c_void_p(sh.peaks_ext.data_ptr()),
                                   ctypes.c_void_p(self._sum[k].data_ptr()), compact,
                                   ctypes.c_void_p((compact if compact is not None else self._per_bank[k]).data_ptr()),
                                   ctypes.c_void_p(b._am.data_ptr()), b.max_out))
            self._views_for = None
        st = _stream_ptr()
        n_out = ctypes.c_int64(0)
        if w:
            seek_to, x_ptr, n_in, halo_flag = p['start'] - w * C, xbuf.data_ptr() + 8 * p['halo'], xbuf.numel() - p['halo'], 1 if p['halo'] > 0 else 0
        else:
            seek_to, x_ptr, n_in, halo_flag = 0, xbuf.data_ptr() + 8 * p['lead'], xbuf.numel() - p['lead'], 0
        cb.banks[0]._check_input(xbuf[p['halo'] if w else p['lead']:])     # dtype / device / contiguity / capacity, once per step
        xp = ctypes.c_void_p(x_ptr)
        for b, lib, h, iq_p, peaks_p, sum_p, compact, sums_p, am_p, max_out in self._fast:
            check(lib.pysdr_bank_seek(h, int(seek_to), st))
            check(lib.pysdr_bank_process_front(h, xp, n_in, halo_flag, iq_p, max_out, peaks_p, ctypes.byref(n_out), st))
            check(lib.pysdr_bank_agc_summary(h, w, sum_p, st))
        if self.world > 1:
            dist.all_gather_into_tensor(self._all, self._sum)            # the one collective, for every channel
            self._per_bank.copy_(self._all.permute(1, 0, 2, 3))          # [bank][rank][G][19]
        else:
            self._per_bank.copy_(self._sum.unsqueeze(1))
        for k, (b, lib, h, iq_p, peaks_p, sum_p, compact, sums_p, am_p, max_out) in enumerate(self._fast):
            if compact is not None:                                      # a bank with fewer receivers: rows [rank][n_rx][19]
                compact.copy_(self._per_bank[k][:, :b.n_rx, :])
            check(lib.pysdr_bank_process_back_carry(h, sums_p if self.rank else None, self.rank, w, am_p, None, max_out, st))
        if self._views_for != n_out.value:
            self._views = ([], [])
            for sh in self.shards:
                b = sh.bank
                b.n_out = n_out.value
                a, q, _ = b.views()
                ks = sh.skip_out
                self._views[0].extend(v[ks:] for v in a)
                self._views[1].extend(v[ks:] for v in q)
            self._views_for = n_out.value
        return list(self._views[0]), list(self._views[1])      # the cached views, in lists the caller may keep or edit

    def invalidate(self):
        """Call after changing MODE / AF_BW / BFO on the banks' parameter objects (see step)."""
        self._fast = None


def raster_tables(P, f0_hz, df_hz, n_ch, nd=3125):
    """Host tables of wola.cu: folded taps of channel 0 (complex64[UP][lp], angles from the quantised 64-bit increment as
    K1 folds its own), the transform position of every channel's bin (base-5 digit reversal of a*c mod 3125) and the
    per-channel 64-bit LO increments.  Pure numpy (checked on the CPU against the oracle's resampler)."""
    from math import gcd
    fs = int(round(P.SRATE))
    if abs(df_hz - round(df_hz)) > 1e-9 or abs(P.SRATE - fs) > 1e-9:
        raise ValueError("raster and sample rate must be whole numbers of Hz")
    g = gcd(int(round(df_hz)), fs)
    a, n_d = int(round(df_hz)) // g, fs // g
    if n_d != nd:
        raise ValueError("raster %g Hz at %g S/s is %d/%d of the sample rate; wola.cu serves x/%d" % (df_hz, P.SRATE, a, n_d, nd))
    up = int(P.UP)
    h = np.asarray(design.resampler_bank(P.SRATE, P.UP, P.DOWN, P.FILT_LEN, design.VIDEO_BWs, P.VIDEO_BW)[design.video_index(P)],
                   np.float64)
    lp = (len(h) + up - 1) // up
    if lp > 625:
        raise ValueError("at most 625 taps per polyphase branch (got %d)" % lp)
    hp = np.zeros(lp * up)
    hp[:len(h)] = h
    offsets = [f0_hz + c * df_hz for c in range(n_ch)]
    incs = [design.freq_to_phase_inc(f, P.SRATE) for f in offsets]
    j = np.arange(lp)
    ang = 2.0 * np.pi * np.array([(incs[0] * int(k)) % (1 << 64) for k in j], np.float64) / 2.0 ** 64
    g0 = np.stack([hp[p + up * j] * np.exp(1j * ang) for p in range(up)]).astype(np.complex64)

    def digitrev5(k):
        r = 0
        for _ in range(5):
            r = r * 5 + k % 5
            k //= 5
        return r
    pos = np.array([digitrev5((a * c) % nd) for c in range(n_ch)], np.int32)
    return g0, pos, incs, lp, offsets


class RasterChannelizer:
    """K1 for many channels on a UNIFORM raster (config 5: 1024 channels, 9.6 kHz apart, 10 MS/s): the baseband IQ of
    every channel from one shared windowing pass and one 3125-point inverse DFT per output instant (wola.cu) instead of
    one 334-tap complex FIR per channel — the same numbers as ReceiverBank's K1 to float32 round-off, ~7x fewer flops.
    Needs df/fs = a/3125 in lowest terms and at most 625 taps per polyphase branch."""

    ND = 3125

    def __init__(self, P, f0_hz, df_hz, n_ch, device=None):
        from . import _lib
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PysdrError("RasterChannelizer needs a CUDA device")
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        self.P, self.n_ch = P, int(n_ch)
        self.up, self.down = int(P.UP), int(P.DOWN)
        try:
            g0, pos, incs, self.lp, self.offsets = raster_tables(P, f0_hz, df_hz, self.n_ch)
        except ValueError as e:
            raise _lib.PysdrError(str(e))
        self.g0 = torch.from_numpy(np.ascontiguousarray(g0)).to(self.device)
        self.pos = torch.from_numpy(pos).to(self.device)
        self.inc = torch.from_numpy(np.array(incs, np.uint64).view(np.int64)).to(self.device)

    def process(self, x, n0=0, n_before=0, hist=None, out=None, out_col=0):
        """x: device complex64.  Without `hist`, element n_before of x is absolute sample n0 and the n_before samples in
        front of it are the filter history (a stream start passes 0); with `hist` (device complex64, the samples just
        before n0) x starts at n0.  Returns complex64[n_ch, n_out] baseband IQ at FS_OUT — written into
        out[:, out_col:out_col+n_out] when `out` (a row-contiguous [n_ch, stride] tensor) is given."""
        import ctypes
        from ._lib import check
        if hist is not None:
            n_before, x_ptr, h_ptr, n_in = hist.numel(), x.data_ptr(), ctypes.c_void_p(hist.data_ptr()), x.numel()
        else:
            x_ptr, h_ptr, n_in = x.data_ptr() + 8 * n_before, None, x.numel() - n_before
        m0 = design.n_out_total(n0, self.up, self.down)
        n_out = design.n_out_total(n0 + n_in, self.up, self.down) - m0
        if out is None:
            out, out_col = torch.empty((self.n_ch, max(n_out, 1)), dtype=torch.complex64, device=self.device), 0
        assert out.stride(1) == 1 and out.shape[0] >= self.n_ch and out_col + n_out <= out.shape[1]
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(self.lib.pysdr_wola_channelize(ctypes.c_void_p(x_ptr), h_ptr, int(n0), int(n_before), int(n_in), int(m0), int(n_out),
                                             self.up, self.down, self.lp, ctypes.c_void_p(self.g0.data_ptr()), self.n_ch,
                                             ctypes.c_void_p(self.pos.data_ptr()), ctypes.c_void_p(self.inc.data_ptr()),
                                             ctypes.c_void_p(out.data_ptr() + 8 * out_col), out.stride(0), st))
        self.n_out = n_out
        return out[:self.n_ch, out_col:out_col + n_out]
