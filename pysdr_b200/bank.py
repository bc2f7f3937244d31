"""ReceiverBank — all receivers fed by one IQ stream, served from ONE read of the samples.

Thin host wrapper over the C ABI (include/pysdr_b200.h).  PyTorch is used only for device memory,
streams and (in dist.py) torch.distributed; every number is produced by libpysdr_b200.so.

Mirrors what the reference does per chunk for every receiver, reference receiver.py:724-725 ->
demodulate_data :231-252 -> dsp.Receiver.demod_data :235.
"""
import ctypes

import numpy as np
import torch

from . import _lib, design
from ._lib import BankConfig, PysdrError, check


MAX_RX_PER_BANK = 128           # include/pysdr_b200.h PYSDR_MAX_RX


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _DeviceMemory:
    """float32 device memory owned by the library, exposed to torch without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class ReceiverBank:
    """n_rx receivers on one stream.

    P supplies SRATE, UP, DOWN, IN_CHUNK_SIZE, FILT_LEN, VIDEO_BW (construction-time) and, read
    at every call, MODE / AF_BW / AF_FILTER_NUM / BFO (scalars as in the reference, or per-RX lists).
    freqs: per-receiver LO offsets in Hz (reference receiver.py:829-835)."""

    def __init__(self, P, freqs, max_in=None, device=None, video_bws=design.VIDEO_BWs, af_bws=design.AF_BWs):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise PysdrError("pysdr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.P = P
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.n_rx = len(freqs)
        if self.n_rx < 1 or self.n_rx > MAX_RX_PER_BANK:
            raise PysdrError("1..%d receivers per bank (reference MAX_RX=6)" % MAX_RX_PER_BANK)
        self.af_len = int(P.FILT_LEN)
        self.max_in = int(max_in if max_in is not None else P.IN_CHUNK_SIZE)
        cfg = BankConfig(float(P.SRATE), int(P.UP), int(P.DOWN), int(P.IN_CHUNK_SIZE), self.n_rx, int(P.FILT_LEN),
                         self.af_len, self.max_in)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.pysdr_bank_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.h = h
        self.max_out = (int(P.UP) * self.max_in) // int(P.DOWN) + 2
        self.video_bws, self.af_bws = video_bws, af_bws
        self.filter_bank = design.resampler_bank(P.SRATE, P.UP, P.DOWN, P.FILT_LEN, video_bws, P.VIDEO_BW)
        self.filter_bank_real = design.af_bank_real(P.FS_OUT, self.af_len, af_bws)
        self.filter_bank_cmpx = design.af_bank_cmpx(P.FS_OUT, self.af_len, af_bws)
        self.filter_bank_lp = design.af_bank_lp(P.FS_OUT, self.af_len, af_bws)
        self.fo = [0.0] * self.n_rx
        self._demod_key = [None] * self.n_rx
        vidx = design.video_index(P, video_bws)
        for r in range(self.n_rx):
            self.set_freq(r, freqs[r])
            self.set_dec_taps(r, self.filter_bank[vidx])
        # output buffers (owned here, handed to the library per call)
        self._iq = None                                       # separate rx.iq copy: allocated only if AM-Synch needs it
        # rx.iq is read straight from the bank's complex memory (K1's output) — no second copy is written
        ptr, stride, hc = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
        check(self.lib.pysdr_bank_c_memory(self.h, ctypes.byref(ptr), ctypes.byref(stride), ctypes.byref(hc)))
        self._hc = hc.value
        mem = _DeviceMemory(ptr.value, self.n_rx * stride.value * 2)
        self._cmem = torch.view_as_complex(torch.as_tensor(mem, device=self.device).view(self.n_rx, stride.value, 2))
        self._am = torch.empty((self.n_rx, 2 * self.max_out), dtype=torch.float32, device=self.device)
        self._am_dc = torch.empty((self.n_rx, 2 * self.max_out), dtype=torch.float32, device=self.device)
        self.n_out = 0
        self._s1 = None

    def scratch1(self):
        """One float32 of device scratch (chunk power of the auto-mute detector)."""
        if self._s1 is None:
            self._s1 = torch.zeros(1, dtype=torch.float32, device=self.device)
        return self._s1

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.pysdr_bank_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- control plane --------------------------------------------------------------------------------
    def set_freq(self, rx, f_hz):
        """rx.lo.change_freq(f): returns the frequency actually applied (reference gui.py:1928,1938)."""
        inc = design.freq_to_phase_inc(f_hz, self.P.SRATE)
        check(self.lib.pysdr_bank_set_lo(self.h, rx, ctypes.c_uint64(inc)))
        self.fo[rx] = design.phase_inc_to_freq(inc, self.P.SRATE)
        return self.fo[rx]

    def set_dec_taps(self, rx, h):
        """rx.dec.h = rx.dec.filter_bank[idx] (reference gui.py:1713)."""
        h = np.ascontiguousarray(h, np.float32)
        check(self.lib.pysdr_bank_set_dec_taps(self.h, rx, _f32_ptr(h), len(h)))

    def _mode_of(self, rx):
        m = design.per_rx(self.P.MODE, rx)
        if m not in design.MODE_IDS:
            raise PysdrError("mode %r is not a ReceiverBank mode (WFM/WFM2 run through sig_proc.Receiver's WFM chain)" % (m,))
        return m

    def sync_demod(self, force=False):
        """Read P.MODE / AF_BW / AF_FILTER_NUM / BFO at call time (reference receiver.py:115,130-131)."""
        P = self.P
        for r in range(self.n_rx):
            mode = self._mode_of(r)
            idx = design.af_index(P, r)
            bfo = float(design.per_rx(getattr(P, 'BFO', 0), r))
            key = (mode, idx, bfo)
            if not force and key == self._demod_key[r]:
                continue
            mid = design.MODE_IDS[mode]
            if mode == 'RAW':
                check(self.lib.pysdr_bank_set_demod(self.h, r, mid, None, 0, 0, ctypes.c_uint64(0)))
                self._demod_key[r] = key
                continue
            if mode in ('USB', 'SSB', 'LSB'):
                g = self.filter_bank_cmpx[idx]
                if mode == 'LSB':
                    g = np.conj(g)
                taps = np.ascontiguousarray(g, np.complex64).view(np.float32)
                cplx = 1
            elif mode in ('CW', 'IQ', 'RTTY'):
                taps = np.ascontiguousarray(self.filter_bank_lp[idx], np.float32)
                cplx = 0
            else:
                taps = np.ascontiguousarray(self.filter_bank_real[idx], np.float32)
                cplx = 0
            bfo_inc = design.freq_to_phase_inc(bfo, P.FS_OUT) if mode == 'CW' else 0
            check(self.lib.pysdr_bank_set_demod(self.h, r, mid, _f32_ptr(taps), self.af_len, cplx,
                                                ctypes.c_uint64(bfo_inc)))
            self._demod_key[r] = key

    def agc_reset(self, rx):
        check(self.lib.pysdr_bank_agc_reset(self.h, rx))

    def pll_reset(self, rx):
        check(self.lib.pysdr_bank_pll_reset(self.h, rx))

    def pll_get(self, rx):
        out = (ctypes.c_double * 2)()
        check(self.lib.pysdr_bank_pll_get(self.h, rx, out, _stream_ptr()))
        return dict(phi=out[0], w=out[1])

    def agc_get(self, rx):
        out = (ctypes.c_double * 5)()
        check(self.lib.pysdr_bank_agc_get(self.h, rx, out, _stream_ptr()))
        return dict(agc=out[0], gain=out[1], maxbuf=out[2], ref=out[3], err=out[4])

    def agc_trace(self, max_blocks=None):
        """(peaks, gains) of the last call: float32 [n_rx, n_blocks] — per-block peak of the pre-AGC audio and applied gain."""
        cap = int(max_blocks if max_blocks is not None else self.max_in // int(self.P.IN_CHUNK_SIZE) + 2)
        pk = np.zeros((self.n_rx, cap), np.float32)
        gn = np.zeros((self.n_rx, cap), np.float32)
        nb = ctypes.c_int64(0)
        check(self.lib.pysdr_bank_agc_trace(self.h, pk.ctypes.data_as(ctypes.c_void_p), gn.ctypes.data_as(ctypes.c_void_p),
                                            cap, ctypes.byref(nb), _stream_ptr()))
        n = nb.value
        return pk.reshape(-1)[:self.n_rx * n].reshape(self.n_rx, n), gn.reshape(-1)[:self.n_rx * n].reshape(self.n_rx, n)

    def reset(self):
        check(self.lib.pysdr_bank_reset(self.h))

    def seek(self, n0):
        check(self.lib.pysdr_bank_seek(self.h, int(n0), _stream_ptr()))

    def set_timing(self, on=True):
        check(self.lib.pysdr_bank_set_timing(self.h, 1 if on else 0))

    def get_timing(self):
        out = (ctypes.c_double * 4)()
        check(self.lib.pysdr_bank_get_timing(self.h, out, _stream_ptr()))
        return dict(k1_ms=out[0], front_rest_ms=out[1], back_ms=out[2], calls=int(out[3]))

    def force_direct_fir(self, on=True):
        check(self.lib.pysdr_bank_force_direct_fir(self.h, 1 if on else 0))

    def force_generic(self, on=True):
        check(self.lib.pysdr_bank_force_generic(self.h, 1 if on else 0))

    def set_k1_mma(self, mode):
        """0: never use the tensor-core K1 kernels; 1 (default): for calls with >= 8192 interior super-periods (banks of 16 or more
        receivers: 2048, many-channel kernel); 2: whenever possible."""
        check(self.lib.pysdr_bank_set_k1_mma(self.h, int(mode)))

    @property
    def k1_mma_available(self):
        return bool(self.lib.pysdr_bank_k1_mma_available(self.h))

    @property
    def k1_last(self):
        """Kernel of the last call: 0 generic, 1 tap-stationary FP32, 2 tensor-core interior + tap-stationary edge tiles, 3 the
        many-channel tensor-core kernel (k1_chan.cu, banks of 16 or more receivers)."""
        return self.lib.pysdr_bank_k1_last(self.h)

    @property
    def k1_variant(self):
        return self.lib.pysdr_bank_k1_variant(self.h)

    @property
    def position(self):
        return self.lib.pysdr_bank_position(self.h)

    @property
    def launches(self):
        return self.lib.pysdr_bank_launch_count(self.h)

    def get_state(self):
        n = self.lib.pysdr_bank_state_size(self.h)
        buf = (ctypes.c_char * n)()
        check(self.lib.pysdr_bank_get_state(self.h, buf, n, _stream_ptr()))
        return bytes(buf)

    def set_state(self, blob):
        check(self.lib.pysdr_bank_set_state(self.h, blob, len(blob), _stream_ptr()))
        self._demod_key = [None] * self.n_rx

    # -- data plane -------------------------------------------------------------------------------------
    def _check_input(self, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.complex64 and x.dim() == 1
                and x.is_contiguous()):
            raise PysdrError("process() wants a contiguous 1-D complex64 CUDA tensor")
        if x.numel() > self.max_in:
            raise PysdrError("chunk of %d samples exceeds this bank's max_in=%d" % (x.numel(), self.max_in))

    def process(self, x, halo_in_place=False, want_dc=True):
        """All receivers' demod_data on a device-resident chunk; returns (am, iq, am_dc) lists of views
        valid until the next call.  am[r] is float32 (complex64 for IQ/RTTY mode)."""
        self._check_input(x)
        self.sync_demod()
        n_out = ctypes.c_int64(0)
        check(self.lib.pysdr_bank_process(self.h, ctypes.c_void_p(x.data_ptr()), x.numel(), 1 if halo_in_place else 0,
                                          self._iq_copy_ptr(), ctypes.c_void_p(self._am.data_ptr()),
                                          ctypes.c_void_p(self._am_dc.data_ptr()) if want_dc else None,
                                          self.max_out, ctypes.byref(n_out), _stream_ptr()))
        self.n_out = n_out.value
        return self.views()

    def process_front(self, x, peaks, halo_in_place=False):
        self._check_input(x)
        self.sync_demod()
        n_out = ctypes.c_int64(0)
        check(self.lib.pysdr_bank_process_front(self.h, ctypes.c_void_p(x.data_ptr()), x.numel(),
                                                1 if halo_in_place else 0, self._iq_copy_ptr(),
                                                self.max_out, ctypes.c_void_p(peaks.data_ptr()), ctypes.byref(n_out),
                                                _stream_ptr()))
        self.n_out = n_out.value
        return self.n_out

    def process_back(self, prev_peaks=None, want_dc=True, skip_blocks=0):
        pp = ctypes.c_void_p(prev_peaks.data_ptr()) if prev_peaks is not None and prev_peaks.numel() else None
        n_prev = 0 if pp is None else prev_peaks.shape[1]
        check(self.lib.pysdr_bank_process_back(self.h, pp, n_prev, int(skip_blocks), ctypes.c_void_p(self._am.data_ptr()),
                                               ctypes.c_void_p(self._am_dc.data_ptr()) if want_dc else None,
                                               self.max_out, _stream_ptr()))
        return self.views()

    def process_back_carry(self, summaries, n_before, want_dc=True, skip_blocks=0):
        """process_back whose AGC entry state comes from the summaries of the n_before earlier time shards (float64 device
        tensor [>= n_before, n_rx, 19], see dist.py) — entry state, scan and gain application in one launch."""
        check(self.lib.pysdr_bank_process_back_carry(self.h, ctypes.c_void_p(summaries.data_ptr()) if n_before else None,
                                                     int(n_before), int(skip_blocks), ctypes.c_void_p(self._am.data_ptr()),
                                                     ctypes.c_void_p(self._am_dc.data_ptr()) if want_dc else None,
                                                     self.max_out, _stream_ptr()))
        return self.views()

    def force_unfused(self, on=True):
        check(self.lib.pysdr_bank_force_unfused(self.h, 1 if on else 0))

    def n_blocks(self, n_in):
        return self.lib.pysdr_bank_n_blocks(self.h, int(n_in))

    def views(self):
        n = self.n_out
        am, dc, iq = [], [], []
        for r in range(self.n_rx):
            if self._mode_of(r) in ('IQ', 'RTTY'):
                am.append(torch.view_as_complex(self._am[r, :2 * n].view(n, 2)))
                dc.append(torch.view_as_complex(self._am_dc[r, :2 * n].view(n, 2)))
            else:
                am.append(self._am[r, :n])
                dc.append(self._am_dc[r, :n])
            iq.append(self.iq_row(r, n))
        return am, iq, dc

    def _iq_copy_ptr(self):
        """Device pointer for a separate rx.iq copy, or None when rx.iq can be read from the complex memory (every mode
        but AM-Synch, whose carrier loop de-rotates that memory in place)."""
        self._iq_separate = any(self._mode_of(r) == 'AM-Synch' for r in range(self.n_rx))
        if not self._iq_separate:
            return None
        if self._iq is None:
            self._iq = torch.empty((self.n_rx, self.max_out), dtype=torch.complex64, device=self.device)
        return ctypes.c_void_p(self._iq.data_ptr())

    def iq_row(self, r, n=None):
        """rx.iq of receiver r from the last call (device view, valid until the next call)."""
        n = self.n_out if n is None else n
        if getattr(self, '_iq_separate', False):
            return self._iq[r, :n]
        return self._cmem[r, self._hc:self._hc + n]

    def process_host(self, x_np, want_dc=True, want_iq=True):
        """Host buffers in, host buffers out (the reference-facing call): ONE C call does upload, kernels, download and the
        synchronisation (pysdr_bank_process_host); the arrays returned are copies (a few KB per receiver), so they stay valid
        as long as the caller keeps them."""
        if len(x_np) == 0:                                    # nothing in, nothing out (no state change)
            e = [np.zeros(0, np.complex64 if self._mode_of(r) in ('IQ', 'RTTY') else np.float32) for r in range(self.n_rx)]
            return e, [np.zeros(0, np.complex64) for _ in range(self.n_rx)], [v.copy() for v in e]
        x_np = np.asarray(x_np)
        if x_np.dtype != np.complex64 or not x_np.flags['C_CONTIGUOUS']:
            x_np = np.ascontiguousarray(x_np, np.complex64)
        n = len(x_np)
        if n > self.max_in:
            raise PysdrError("chunk of %d samples exceeds max_in=%d" % (n, self.max_in))
        self.sync_demod()
        p_am, p_iq, p_dc = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        row, n_out = ctypes.c_int64(0), ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            check(self.lib.pysdr_bank_process_host(self.h, x_np.ctypes.data_as(ctypes.c_void_p), n, 1 if want_iq else 0,
                                                   1 if want_dc else 0, ctypes.byref(p_am), ctypes.byref(p_iq), ctypes.byref(p_dc),
                                                   ctypes.byref(row), ctypes.byref(n_out), _stream_ptr()))
        no = self.n_out = n_out.value
        if getattr(self, '_hp', None) != (p_am.value, row.value):           # wrap the bank's pinned result block once
            nfl = row.value * self.n_rx
            as_f32 = lambda p: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(nfl,)).reshape(self.n_rx, row.value)
            self._h_am_np, self._h_iq_np, self._h_dc_np = as_f32(p_am), as_f32(p_iq), as_f32(p_dc)
            self._hp = (p_am.value, row.value)
        cplx = [self._mode_of(r) in ('IQ', 'RTTY') for r in range(self.n_rx)]

        def host(buf, r):
            return buf[r, :2 * no].view(np.complex64).copy() if cplx[r] else buf[r, :no].copy()
        am = [host(self._h_am_np, r) for r in range(self.n_rx)]
        iq = [self._h_iq_np[r, :2 * no].view(np.complex64).copy() for r in range(self.n_rx)] if want_iq else \
            [np.zeros(0, np.complex64) for _ in range(self.n_rx)]
        dc = [host(self._h_dc_np, r) for r in range(self.n_rx)] if want_dc else [a.copy() for a in am]
        return am, iq, dc
