"""Drop-in for the L1 surface pySDR imports as ``import sig_proc as dsp`` (reference receiver.py:45,
Plotting.py:31, params.py, srates.py:26) — same names, argument meaning and error behaviour, with the
arithmetic done by hand-written sm_100a kernels behind the C ABI (libpysdr_b200.so).

Surface:  up_dn, signal_generator, Receiver, spectrum, ring_buffer2, ring_buffer3, bpf, convolver.
Host numpy arrays in / out, exactly like the reference's callers expect; device-resident batch processing
lives in bank.ReceiverBank / receiver.py.
"""
import ctypes
import queue

import numpy as np
import torch

from . import _lib, design
from ._lib import PysdrError, check
from .bank import ReceiverBank, _stream_ptr
from .design import up_dn, bpf                      # noqa: F401  (re-exported reference names)


def _dev():
    if not torch.cuda.is_available():
        raise PysdrError("pysdr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda:%d" % torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------------
class signal_generator:
    """NCO (reference receiver.py:822 ``dsp.signal_generator(f, N, fs, True)``; ``.quad_mixer(x)``
    receiver.py:552-553; ``.change_freq(f)`` returns the applied frequency, gui.py:1928)."""

    def __init__(self, f, N, fs, cmplx=True):
        self.lib = _lib.load()
        self.N = int(N)
        self.fs = float(fs)
        self.complex = bool(cmplx)
        self.acc = 0
        self.change_freq(f)

    def change_freq(self, f):
        self.inc = design.freq_to_phase_inc(f, self.fs)
        self.fo = design.phase_inc_to_freq(self.inc, self.fs)
        return self.fo

    def quad_mixer(self, x):
        dev = _dev()
        host = not isinstance(x, torch.Tensor)
        xd = torch.from_numpy(np.ascontiguousarray(x, np.complex64)).to(dev) if host else x.contiguous()
        y = torch.empty_like(xd)
        n = xd.numel()
        check(self.lib.pysdr_quad_mixer(ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(y.data_ptr()), n,
                                        ctypes.c_uint64(self.acc), ctypes.c_uint64(self.inc), _stream_ptr()))
        self.acc = (self.acc + self.inc * n) & ((1 << 64) - 1)
        return y.cpu().numpy() if host else y


# ---------------------------------------------------------------------------------------------------
class _Lo:
    """``rx.lo`` (reference receiver.py:112,352, gui.py:1928,1938).  `_rx._bank` / `_rx._slot` name the bank row."""

    def __init__(self, rx):
        self._rx = rx

    @property
    def fo(self):
        return self._rx._bank.fo[self._rx._slot]

    def change_freq(self, f):
        if self._rx._wfm is not None:
            self._rx._wfm.set_freq(f)
        return self._rx._bank.set_freq(self._rx._slot, f)


class _Dec:
    """``rx.dec.h`` is assignable from ``rx.dec.filter_bank[idx]`` (reference gui.py:1713, receiver.py:127)."""

    def __init__(self, rx):
        self._rx = rx
        self.filter_bank = rx._bank.filter_bank
        self._h = self.filter_bank[design.video_index(rx.P, rx._bank.video_bws)]

    @property
    def h(self):
        return self._h

    @h.setter
    def h(self, taps):
        self._rx._bank.set_dec_taps(self._rx._slot, taps)          # takes effect at the next chunk boundary
        self._h = np.asarray(taps, np.float32)


class _Pll:
    """``rx.demod.am_pll`` — the AM-Synch carrier loop lives in the bank (am_pll_kernel, bank.cu)."""

    def __init__(self, rx):
        self._rx = rx

    def reset(self):                                     # reference receiver.py:649
        self._rx._bank.pll_reset(self._rx._slot)

    @property
    def phi(self):
        return self._rx._bank.pll_get(self._rx._slot)['phi']

    @property
    def w(self):
        return self._rx._bank.pll_get(self._rx._slot)['w']


class _Holder:
    pass


class _Demod:
    def __init__(self, rx):
        b = rx._bank
        self.filter_bank_real = b.filter_bank_real      # reference receiver.py:873
        self.filter_bank_cmpx = b.filter_bank_cmpx      # reference receiver.py:874
        self.am_pll = _Pll(rx)
        self.wfm_video = _WfmVideo(rx)                                   # reference gui.py:1704
        self.wfm_filter_bank = design.wfm_video_bank(rx.P.SRATE, rx.P.FILT_LEN, design.VIDEO_BWs, rx.P.VIDEO_BW)


class _Agc:
    """``rx.agc.reset()`` (reference receiver.py:648) and the read-outs of reference watchdog.py:298-302."""

    def __init__(self, rx):
        self._rx = rx

    def reset(self):
        self._rx._bank.agc_reset(self._rx._slot)

    def _get(self, k):
        return self._rx._bank.agc_get(self._rx._slot)[k]

    agc = property(lambda s: s._get('agc'))
    gain = property(lambda s: s._get('gain'))
    maxbuf = property(lambda s: s._get('maxbuf'))
    ref = property(lambda s: s._get('ref'))
    err = property(lambda s: s._get('err'))


class Receiver:
    """``dsp.Receiver(P, frq, irx, name, VIDEO_BWs, AF_BWs)`` (reference receiver.py:65,835).

    ``demod_data(x)`` takes the caller-owned complex64 chunk (copied to the device before returning, the
    caller may reuse it — reference receiver.py:445,588) and returns ``am``; side effects ``.am``, ``.iq``
    stay valid until the next call."""

    AUTO_MUTE_THRESH = 0.25

    def __init__(self, P, frq, irx, name, video_bws=design.VIDEO_BWs, af_bws=design.AF_BWs):
        self.P = P
        self.irx = irx
        self.name = name
        self.sub = 0
        self._slot = 0                                         # row of this receiver in its (one-row) bank
        self._view = _PView(P, irx)
        self._bank = ReceiverBank(self._view, [frq], max_in=int(P.IN_CHUNK_SIZE), video_bws=video_bws, af_bws=af_bws)
        self.lo = _Lo(self)
        self.dec = _Dec(self)
        self.demod = _Demod(self)
        self.agc = _Agc(self)
        self.am = np.zeros(0, np.float32)
        self.iq = np.zeros(0, np.complex64)
        self.am_dc = np.zeros(0, np.float32)
        self.mute_cnt = 0
        self._wfm = None
        self._pw = torch.zeros(1, dtype=torch.float32, device=self._bank.device)

    def _wfm_chain(self):
        stereo = design.per_rx(self.P.MODE, self.irx) == 'WFM2'
        if self._wfm is not None and self._wfm.stereo != stereo:      # WFM <-> WFM2: the stage after the discriminator
            self._wfm.stage2(stereo)                                  # (and its AGC) restarts, the video stage runs on
        if self._wfm is None:
            self._wfm = _WfmChain(self, stereo)
        return self._wfm

    def demod_data(self, x):
        if design.per_rx(self.P.MODE, self.irx) in ('WFM', 'WFM2'):      # demodulate first, then resample (gui.py:1703)
            if len(x) == 0:
                self.am, self.iq = np.zeros(0, np.float32), np.zeros(0, np.complex64)
                return self.am
            self.am, self.iq = self._wfm_chain().demod(x)
            self.am_dc = self.am
            return self.am
        am, iq, dc = self._bank.process_host(x)
        self.am, self.iq, self.am_dc = am[0], iq[0], dc[0]
        return self.am

    def auto_mute(self, x):
        """Mute while mean|x|^2 exceeds the threshold, held for MUTE_CHUNKS calls (reference receiver.py:238-245,
        params.py:447-450)."""
        xd = torch.from_numpy(np.ascontiguousarray(x, np.complex64)).to(self._bank.device)
        check(self._bank.lib.pysdr_mean_power(ctypes.c_void_p(xd.data_ptr()), xd.numel(),
                                              ctypes.c_void_p(self._pw.data_ptr()), _stream_ptr()))
        if float(self._pw.item()) > self.AUTO_MUTE_THRESH:
            self.mute_cnt = int(self.P.MUTE_CHUNKS)
        elif self.mute_cnt > 0:
            self.mute_cnt -= 1
        return self.mute_cnt > 0


class _PView:
    """What a single Receiver sees of P: per-RX entries of list-valued MODE/AF_BW/... are projected to its
    own index; everything is read through to P at call time."""

    def __init__(self, P, irx):
        object.__setattr__(self, '_P', P)
        object.__setattr__(self, '_irx', irx)

    def __getattr__(self, k):
        v = getattr(object.__getattribute__(self, '_P'), k)
        if k in ('MODE', 'AF_BW', 'AF_FILTER_NUM', 'BFO'):
            return design.per_rx(v, object.__getattribute__(self, '_irx'))
        return v


# ---------------------------------------------------------------------------------------------------
def _czt_tables(lib, win, chunk, nfft, M):
    """Bluestein tables for czt.cu (float64 on the host, stored complex64): wc[n] = w[n] conj(b[n]) and the spectrum of
    the wrapped chirp b[m] = exp(j pi m^2 / nfft), m in [-(chunk-1), nfft-1], divided by M, in [512][256] position order."""
    def chirp(m):
        m = np.asarray(m, np.int64)
        return np.exp(1j * np.pi * ((m * m) % (2 * nfft)).astype(np.float64) / nfft)
    wc = (np.asarray(win, np.float64) * np.conj(chirp(np.arange(chunk)))).astype(np.complex64)
    m = np.arange(-(chunk - 1), nfft)
    bpad = np.zeros(M, np.complex128)
    bpad[m % M] = chirp(m)
    B = np.fft.fft(bpad) / M
    k1 = np.array([lib.pysdr_fft_pos_to_freq(512, p) for p in range(512)], np.int64)
    k2 = np.array([lib.pysdr_fft_pos_to_freq(256, q) for q in range(256)], np.int64)
    bspec = B[k1[:, None] + 512 * k2[None, :]].astype(np.complex64)
    return np.ascontiguousarray(wc), np.ascontiguousarray(bspec)


class spectrum:
    """``dsp.spectrum(fs, chunk_size, NFFT, overlap, TAG=)`` (reference Plotting.py:376-377); attributes
    ``NFFT frq frq2 df fs chunk_size new_samps`` (Plotting.py:467,594,690; gui.py:1259,1289-1290,1369)."""

    def __init__(self, fs, chunk_size, NFFT, overlap, TAG=''):
        self.lib = _lib.load()
        self.fs = fs
        self.chunk_size = int(chunk_size)
        self.NFFT = int(NFFT)
        self.overlap = overlap
        self.TAG = TAG
        self.new_samps = int(self.chunk_size * (1 - overlap))
        from scipy import signal as _sig
        self.win = _sig.get_window('hann', self.chunk_size, fftbins=True).astype(np.float32)
        self.df = fs / float(self.NFFT)
        self.frq = (np.arange(self.NFFT) - self.NFFT // 2) * self.df
        self.frq2 = self.frq
        self.device = _dev()
        h = ctypes.c_void_p()
        self.czt = bool(self.NFFT & (self.NFFT - 1)) or self.NFFT > 16384 or self.NFFT < 64
        if self.czt:
            # e.g. the reference's RF panel: chunk 32818, NFFT 65636 = 4*61*269 (Plotting.py:370-375) -> chirp-z (czt.cu)
            M = 131072
            if self.NFFT + self.chunk_size - 1 > M:
                raise PysdrError("spectrum: NFFT=%d with chunk_size=%d exceeds the 2^17-point chirp-z transform"
                                 % (self.NFFT, self.chunk_size))
            wc, bspec = _czt_tables(self.lib, self.win, self.chunk_size, self.NFFT, M)
            check(self.lib.pysdr_czt_create(self.chunk_size, self.NFFT, max(1, self.new_samps),
                                            self.win.ctypes.data_as(ctypes.c_void_p), wc.ctypes.data_as(ctypes.c_void_p),
                                            bspec.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)))
        else:
            check(self.lib.pysdr_psd_create(self.chunk_size, self.NFFT, max(1, self.new_samps),
                                            self.win.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)))
        self.h = h
        self.buf = torch.zeros(self.chunk_size, dtype=torch.complex64, device=self.device)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                (self.lib.pysdr_czt_destroy if self.czt else self.lib.pysdr_psd_destroy)(self.h)
                self.h = None
        except Exception:
            pass

    def _to_dev(self, y):
        if isinstance(y, torch.Tensor):
            return y.to(self.device).to(torch.complex64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(y).astype(np.complex64))).to(self.device)

    def _lines(self, x, navg, dB):
        n = x.numel()
        nfr = 0 if n < self.chunk_size else 1 + (n - self.chunk_size) // max(1, self.new_samps)
        nl = nfr // navg
        out = torch.empty((max(nl, 1), self.NFFT), dtype=torch.float32, device=self.device)
        got = ctypes.c_int64(0)
        fn = self.lib.pysdr_czt_lines if self.czt else self.lib.pysdr_psd_lines
        check(fn(self.h, ctypes.c_void_p(x.data_ptr()), n, 1, int(navg), 1 if dB else 0,
                 ctypes.c_void_p(out.data_ptr()), ctypes.byref(got), _stream_ptr()))
        return out[:got.value]

    def periodogram(self, y, dB=True):
        """One frame; the newest len(y) samples are shifted into the chunk_size window.  Returns [] on
        error like the reference (Plotting.py:463-465)."""
        n = len(y)
        if n == 0 or n > self.chunk_size:
            return []
        yd = self._to_dev(y)
        self.buf = torch.cat((self.buf[n:], yd))
        return self._lines(self.buf, 1, dB)[0].cpu().numpy()

    def psd_est(self, x, dB=True):
        """Welch average over all full frames (reference sigs/iq.py:75-79)."""
        xd = self._to_dev(x)
        n = xd.numel()
        if n < self.chunk_size:
            return []
        nfr = 1 + (n - self.chunk_size) // max(1, self.new_samps)
        return self._lines(xd, nfr, dB)[0].cpu().numpy()

    def waterfall(self, x, navg, dB=True, to_host=True):
        out = self._lines(self._to_dev(x), navg, dB)
        return out.cpu().numpy() if to_host else out


# ---------------------------------------------------------------------------------------------------
class convolver:
    """``dsp.convolver(h, dtype).convolve_fast(x)`` (reference receiver.py:862,216): streaming FIR with
    carried history on the device (pysdr_fir_valid)."""

    def __init__(self, h, dtype=np.float32):
        self.lib = _lib.load()
        self.h = np.ascontiguousarray(h, np.float32)
        self.dtype = dtype
        self.device = _dev()
        self.hist = None

    def convolve_fast(self, x):
        x = np.asarray(x)
        cplx = np.iscomplexobj(x)
        tdt = torch.complex64 if cplx else torch.float32
        xd = torch.from_numpy(np.ascontiguousarray(x.astype(np.complex64 if cplx else np.float32))).to(self.device)
        L = len(self.h)
        if self.hist is None or self.hist.dtype != tdt:
            self.hist = torch.zeros(L - 1, dtype=tdt, device=self.device)
        src = torch.cat((self.hist, xd)).contiguous()
        out = torch.empty(xd.numel(), dtype=tdt, device=self.device)
        check(self.lib.pysdr_fir_valid(ctypes.c_void_p(src.data_ptr()), 1 if cplx else 0,
                                       self.h.ctypes.data_as(ctypes.c_void_p), L, xd.numel(),
                                       ctypes.c_void_p(out.data_ptr()), _stream_ptr()))
        self.hist = src[src.numel() - (L - 1):].clone()
        y = out.cpu().numpy()
        return y if cplx else y.astype(self.dtype)


# ---------------------------------------------------------------------------------------------------
class ring_buffer2:
    """Host-side sample FIFO (reference pySDR.py:103-112, watchdog.py:153-156,190,197, gui.py:1264-1266).
    Plumbing only — stays in Python like the reference."""

    def __init__(self, tag, size, PREVENT_OVERFLOW=False, BLOCK=False):
        self.tag = tag
        self.size = int(size)
        self.prevent_overflow = PREVENT_OVERFLOW
        self.buf = queue.Queue()
        self.data = None
        self.nsamps = 0

    def clear(self):
        self.data = None
        self.nsamps = 0

    def push(self, x):
        x = np.asarray(x)
        if self.data is None or self.nsamps == 0:
            self.data = x.copy()
        else:
            self.data = np.concatenate((self.data[-self.nsamps:], x))
        if len(self.data) > self.size:                  # overflow: keep the newest `size` samples
            if self.prevent_overflow:
                self.data = self.data[:self.size]
            else:
                self.data = self.data[len(self.data) - self.size:]
        self.nsamps = len(self.data)
        return self.nsamps

    def push_zeros(self, n):
        dt = self.data.dtype if self.data is not None else np.float32
        return self.push(np.zeros(int(n), dt))

    def ready(self, n):
        return self.nsamps >= n

    def pull(self, n, flush=False):
        """Oldest n samples; flush=True discards everything older than the newest n first (gui.py:1264)."""
        n = int(n)
        if self.nsamps < n:
            return np.zeros(0, np.float32)
        d = self.data[-self.nsamps:]
        if flush:
            out = d[len(d) - n:]
            self.data = None
            self.nsamps = 0
            return out
        out = d[:n]
        self.data = d[n:]
        self.nsamps = len(self.data)
        return out


class ring_buffer3(ring_buffer2):
    """Queue-backed variant used by mp.py (reference mp.py:90-95: ``*_psd_Q = rb.buf``)."""

    def __init__(self, tag, size):
        super().__init__(tag, size)

    def pull(self, n, flush=False):
        while not self.buf.empty():
            self.push(self.buf.get())
        return super().pull(n, flush)


# ---------------------------------------------------------------------------------------------------
class lfilter_stream:
    """``scipy.signal.lfilter(b, a, x, zi=z)`` with the state carried between calls, on the device
    (reference sigs/iir.py:90-105: ``y1,z1 = lfilter(b,a,x1,zi=zi); y2,z2 = lfilter(b,a,x2,zi=z1)``).
    Block-parallel linear scan in float64 (pysdr_lfilter); ``.z`` is scipy's ``zi`` vector."""

    def __init__(self, b, a, n_ch=1):
        self.lib = _lib.load()
        self.b = np.ascontiguousarray(np.atleast_1d(b), np.float64)
        self.a = np.ascontiguousarray(np.atleast_1d(a), np.float64)
        self.order = max(len(self.a), len(self.b)) - 1
        self.n_ch = int(n_ch)
        self.device = _dev()
        self.reset()

    def reset(self):
        self._z = torch.zeros((self.n_ch, max(1, self.order)), dtype=torch.float64, device=self.device)

    @property
    def z(self):
        return self._z.cpu().numpy()[:, :self.order]

    def run_dev(self, xd):
        """xd: float32 CUDA tensor [n] or [n_ch, n] (contiguous) -> same shape."""
        x2 = xd.reshape(self.n_ch, -1).contiguous()
        n = x2.shape[1]
        y = torch.empty_like(x2)
        if self.order == 0:
            return (x2 * float(self.b[0] / self.a[0])).reshape(xd.shape)
        check(self.lib.pysdr_lfilter(self.b.ctypes.data_as(ctypes.c_void_p), len(self.b),
                                     self.a.ctypes.data_as(ctypes.c_void_p), len(self.a),
                                     ctypes.c_void_p(x2.data_ptr()), ctypes.c_void_p(y.data_ptr()), n, self.n_ch, n,
                                     ctypes.c_void_p(self._z.data_ptr()), _stream_ptr()))
        return y.reshape(xd.shape)

    def run(self, x):
        xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(self.device)
        return self.run_dev(xd).cpu().numpy()


class squelch:
    """Noise squelch of reference sigs/squelch.m:92-145: the audio is split by an elliptic low-pass (3 kHz) and
    high-pass (4 kHz) (:103-105), each band's envelope is smoothed by ``filter(alpha,[1 alpha-1],|z|)``, alpha=.001
    (:125-128), and the squelch opens when the in-band / out-of-band ratio ``sq1./sq2`` (:141) exceeds ``thresh``.
    All four recursions are block-parallel scans with carried state."""

    def __init__(self, fs, alpha=0.001, thresh=2.0):
        from scipy import signal as _sig
        self.lib = _lib.load()
        B1, A1 = _sig.ellip(5, 5, 40, 3000 / (fs / 2.0))
        B2, A2 = _sig.ellip(5, 5, 40, 4000 / (fs / 2.0), 'high')
        self.f1, self.f2 = lfilter_stream(B1, A1), lfilter_stream(B2, A2)
        self.s1, self.s2 = lfilter_stream([alpha], [1, alpha - 1]), lfilter_stream([alpha], [1, alpha - 1])
        self.thresh = thresh
        self.device = _dev()

    def _env(self, filt, smooth, yd):
        z = filt.run_dev(yd)
        check(self.lib.pysdr_abs_f32(ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(z.data_ptr()), z.numel(), _stream_ptr()))
        return smooth.run_dev(z)

    def run(self, y):
        """y: real audio chunk (host) -> (ratio, open) arrays."""
        yd = torch.from_numpy(np.ascontiguousarray(y, np.float32)).to(self.device)
        sq1, sq2 = self._env(self.f1, self.s1, yd), self._env(self.f2, self.s2, yd)
        r = torch.empty_like(sq1)
        check(self.lib.pysdr_ratio_f32(ctypes.c_void_p(sq1.data_ptr()), ctypes.c_void_p(sq2.data_ptr()),
                                       ctypes.c_void_p(r.data_ptr()), 1e-30, r.numel(), _stream_ptr()))
        ratio = r.cpu().numpy()
        return ratio, ratio > self.thresh


# ---------------------------------------------------------------------------------------------------
class _NS:
    pass


class _WfmChain:
    """WFM / WFM2: video FIR at the RF rate -> FM discriminator at the RF rate -> resampler (AF low-pass) -> AGC
    ("BCB FM is wideband so we need to demodulate first before resampling", reference gui.py:1703,1759-1762).
    Built from two K1 launches (UP=DOWN=1 video stage, UP/DOWN resampler stage on the real discriminator output)
    and pysdr_fm_disc.  WFM is mono.  WFM2 is the stereo decoder (BASELINE config 4; no in-tree specification, the law
    is an open choice stated in DESIGN.md section 3): the resampler stage becomes a 3-row bank on the multiplex — LO 0
    (L+R), LO 38 kHz (L-R), LO 19 kHz (pilot, 500 Hz AF low-pass) — and pysdr_bank_set_stereo turns the rows into L/R
    with feed-forward pilot recovery (carrier = (pilot/|pilot|)^2); ``am`` is then complex64 L + 1j*R."""

    PILOT_AF = '500 Hz'

    def __init__(self, rx, stereo=False):
        P = rx.P
        self.rx = rx
        self.stereo = bool(stereo)
        self.lib = _lib.load()
        # P.WFM_MAX_CHUNKS (default 1 = the reference's chunk-at-a-time calls) lets a resident capture go through in one call
        self.max_in = int(P.IN_CHUNK_SIZE) * max(1, int(getattr(P, 'WFM_MAX_CHUNKS', 1)))
        self.filter_bank = design.wfm_video_bank(P.SRATE, P.FILT_LEN, design.VIDEO_BWs, P.VIDEO_BW)
        dev = rx._bank.device
        self.L = int(P.FILT_LEN)
        # video stage: LO + FIR + discriminator as ONE overlap-save FFT-convolution kernel (pysdr_wfm_video_disc); filters too
        # long for the 4096-point transform keep the r01 route (direct-form FIR through a K1-only bank, then pysdr_fm_disc)
        self.fast_video = 2 <= self.L <= 2045 and not getattr(P, 'WFM_DIRECT_VIDEO', False)
        self.vbank = None
        if self.fast_video:
            self.inc = design.freq_to_phase_inc(rx._bank.fo[0], P.SRATE)
            self.acc = 0
            self.vhist = torch.zeros(self.L + 1, dtype=torch.complex64, device=dev)
            self.vprev2 = torch.zeros((2, 2), dtype=torch.complex64, device=dev)
            self.vslot = 0
            self.vH = torch.empty(4096, dtype=torch.complex64, device=dev)
            self._vdirty = True
        else:
            vid = _NS()
            vid.SRATE, vid.UP, vid.DOWN, vid.FS_OUT = P.SRATE, 1, 1, int(P.SRATE)
            vid.IN_CHUNK_SIZE, vid.FILT_LEN, vid.VIDEO_BW = P.IN_CHUNK_SIZE, P.FILT_LEN, P.VIDEO_BW
            vid.MODE, vid.AF_BW, vid.AF_FILTER_NUM, vid.BFO, vid.VIDEO_FILTER_NUM = 'RAW', 0, 0, 0, None
            self.vbank = ReceiverBank(vid, [rx._bank.fo[0]], max_in=self.max_in)
            check(self.lib.pysdr_bank_set_k1_only(self.vbank.h, 1))
            self.prev2 = torch.zeros(2, dtype=torch.complex64, device=dev)
        self.set_video(self.filter_bank[design.video_index(P)])
        self.fm = torch.empty(self.max_in, dtype=torch.complex64, device=dev)
        self.stage2(stereo)

    def set_freq(self, f):
        """rx.lo.change_freq(f) for the video stage (phase continuous at the current sample)."""
        if self.vbank is not None:
            return self.vbank.set_freq(0, f)
        self.inc = design.freq_to_phase_inc(f, self.rx.P.SRATE)
        self._vdirty = True
        return design.phase_inc_to_freq(self.inc, self.rx.P.SRATE)

    def _refresh_video_spectrum(self):
        """G[j] = h[j] e^{+j 2 pi frac(inc j / 2^64)} (the LO folded into the taps in float64, as K1 does) -> spectrum."""
        j = np.arange(len(self.h), dtype=np.uint64)
        with np.errstate(over='ignore'):
            ph = (np.uint64(self.inc) * j).view(np.int64).astype(np.float64) / 18446744073709551616.0
        g = (self.h.astype(np.float64) * np.exp(2j * np.pi * ph)).astype(np.complex64)
        gd = torch.from_numpy(g).to(self.vH.device)
        check(self.lib.pysdr_fir_spectrum(ctypes.c_void_p(gd.data_ptr()), len(g), ctypes.c_void_p(self.vH.data_ptr()), _stream_ptr()))
        torch.cuda.current_stream(self.vH.device).synchronize()      # gd is a temporary
        self._vdirty = False

    def stage2(self, stereo):
        P = self.rx.P
        self.stereo = bool(stereo)
        res = _NS()
        res.SRATE, res.UP, res.DOWN, res.FS_OUT = P.SRATE, P.UP, P.DOWN, P.FS_OUT
        res.IN_CHUNK_SIZE, res.FILT_LEN, res.VIDEO_BW = P.IN_CHUNK_SIZE, P.FILT_LEN, P.VIDEO_BW
        res.MODE, res.AF_BW, res.AF_FILTER_NUM, res.BFO, res.VIDEO_FILTER_NUM = 'RAW', 0, 0, 0, None
        if self.stereo:
            res.MODE = ['IQ', 'IQ', 'IQ']
            res.AF_FILTER_NUM = [0, 0, design.AF_BWs.index(self.PILOT_AF)]
            self.rbank = ReceiverBank(res, [0.0, 38e3, 19e3], max_in=self.max_in)
            check(self.lib.pysdr_bank_set_stereo(self.rbank.h, 1, float(getattr(P, 'WFM_PILOT_MIN', 0.0))))
        else:
            self.rbank = ReceiverBank(res, [0.0], max_in=self.max_in)
        check(self.lib.pysdr_bank_set_real_input(self.rbank.h, 1))   # the discriminator output is real (stored as complex64)
        self._res_key = None
        self.deemph = None

    def set_video(self, h):
        self.h = np.asarray(h, np.float32)
        if self.vbank is not None:
            self.vbank.set_dec_taps(0, self.h)
        else:
            if len(self.h) != self.L:
                raise PysdrError("wfm_video.h must have FILT_LEN=%d taps" % self.L)
            self._vdirty = True

    def demod(self, x):
        """Host chunk in, host audio out (what rx.demod_data returns in WFM / WFM2 mode)."""
        xd = torch.from_numpy(np.ascontiguousarray(x, np.complex64)).to(self.fm.device)
        out, iq = self.demod_dev(xd)
        if self.stereo:
            a = out[0].cpu().numpy() + 1j * out[1].cpu().numpy()
            return a.astype(np.complex64), iq.cpu().numpy()
        return out[0].cpu().numpy(), iq.cpu().numpy()

    def demod_dev(self, xd):
        """Device samples (whole chunks, at most max_in) -> ([audio] or [L, R], baseband iq), device tensors."""
        P, rx = self.rx.P, self.rx
        n = xd.numel()
        fm = self.fm[:n]
        if self.vbank is None:
            if self._vdirty:
                self._refresh_video_spectrum()
            check(self.lib.pysdr_wfm_video_disc(ctypes.c_void_p(xd.data_ptr()), n, ctypes.c_void_p(self.vhist.data_ptr()),
                                                ctypes.c_void_p(self.vprev2.data_ptr()), self.vslot,
                                                ctypes.c_void_p(self.vH.data_ptr()), self.L, ctypes.c_uint64(self.acc),
                                                ctypes.c_uint64(self.inc), ctypes.c_void_p(fm.data_ptr()), _stream_ptr()))
            self.acc = (self.acc + self.inc * n) & ((1 << 64) - 1)
            if n > 0:
                self.vslot ^= 1
        else:
            _, y, _ = self.vbank.process(xd, want_dc=False)
            check(self.lib.pysdr_fm_disc(ctypes.c_void_p(y[0].data_ptr()), n, ctypes.c_void_p(self.prev2.data_ptr()),
                                         ctypes.c_void_p(fm.data_ptr()), _stream_ptr()))
        af_bw = float(design.per_rx(getattr(P, 'AF_BW', 0), rx.irx) or 0)
        if af_bw != self._res_key:
            taps = design.wfm_resampler_taps(P.SRATE, P.UP, P.FILT_LEN, af_bw)
            for r in range(self.rbank.n_rx):
                self.rbank.set_dec_taps(r, taps)
            self._res_key = af_bw
        am, iq, _ = self.rbank.process(fm, want_dc=False)
        tau = getattr(P, 'DEEMPH_US', 0) * 1e-6
        if tau > 0 and self.deemph is None:                     # one-pole de-emphasis, block-parallel scan
            al = 1.0 - np.exp(-1.0 / (P.FS_OUT * tau))
            self.deemph = [lfilter_stream([al], [1, al - 1]) for _ in range(2 if self.stereo else 1)]
        if self.stereo:
            n_out = self.rbank.n_out
            lr = [self.rbank._am[r, :n_out] for r in (0, 1)]    # float32 L, R rows after the common AGC
            if tau > 0:
                lr = [self.deemph[r].run_dev(lr[r].contiguous()) for r in (0, 1)]
            return lr, iq[0]
        a = am[0]
        if tau > 0:
            a = self.deemph[0].run_dev(a.contiguous())
        return [a], iq[0]


class _WfmVideo:
    """``rx.demod.wfm_video.h = rx.demod.wfm_filter_bank[idx]`` (reference gui.py:1704)."""

    def __init__(self, rx):
        self._rx = rx

    @property
    def h(self):
        return self._rx._wfm_chain().h

    @h.setter
    def h(self, taps):
        self._rx._wfm_chain().set_video(taps)
