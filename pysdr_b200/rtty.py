"""RTTY front end: the Kaiser-windowed quarter-symbol FFT filterbank of the reference's RTTY executive
(reference rtty.py:376-404 RTTY_Params, rtty.py:784-786 sizes, rtty.py:807 window, rtty.py:825-856 the loop).

The lines feed the mark/space detectors (rtty.py:849-853); the Baudot decoder itself is host logic and out of
this path.  Arithmetic runs in psd.cu (K3) with sub-step frame starts."""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import check


def nextpow2(n):
    """reference rtty.py:80-82."""
    return math.ceil(math.log(n, 2))


class RTTY_Params:
    """reference rtty.py:376-404 (sizes only; the decoder tables are not part of the filterbank)."""

    def __init__(self, FS_OUT, mark_bins=()):
        self.T = 22e-3
        self.FSK_SHIFT = 170
        self.SAMPS_PER_BIT = 4
        STOP_BITS = 1.5
        self.M = int(4 * (1 + 5 + STOP_BITS))
        self.N = int(round(self.T * FS_OUT))
        self.NFFT = int(2 ** nextpow2(self.N))
        NSTEP = self.N / 4.
        self.NSTART = [int(NSTEP * i + 0.5) for i in range(4)]
        bin_size = FS_OUT / float(self.NFFT)
        self.NBINS = int(round(self.FSK_SHIFT / bin_size))
        self.frq = np.fft.fftshift(np.fft.fftfreq(self.NFFT, d=1000. / FS_OUT)) + 0
        self.mark_bins = np.array(mark_bins)


class rtty_filterbank:
    """Streaming filterbank: ``push(iq)`` takes whole symbols (N samples each, as the executive pulls them from its
    ring buffer, rtty.py:825) and returns the 4 lines per symbol the reference computes for (prev, iq), i.e. none
    for the very first symbol (rtty.py:826-829).  Lines are float32[n, NFFT] = flipud(fftshift(10 log10 |X|^2))."""

    def __init__(self, FS_OUT, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PysdrError("rtty_filterbank needs a CUDA device")
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        self.RTTY = RTTY_Params(FS_OUT)
        self.N, self.NFFT, self.NSTART = self.RTTY.N, self.RTTY.NFFT, self.RTTY.NSTART
        self.window = np.kaiser(self.N, 8.6)                                   # rtty.py:807
        w32 = self.window.astype(np.float32)
        h = ctypes.c_void_p()
        check(self.lib.pysdr_psd_create(self.N, self.NFFT, self.N, w32.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h)))
        self.h = h
        offs = (ctypes.c_int32 * 4)(*self.NSTART)
        check(self.lib.pysdr_psd_configure(self.h, 4, offs, _lib.PSD_RAW | _lib.PSD_FLIP))
        self.prev = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.pysdr_psd_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self):
        self.prev = None

    def push(self, iq, to_host=True):
        if isinstance(iq, torch.Tensor):
            x = iq.to(self.device).to(torch.complex64).contiguous()
        else:
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(iq).astype(np.complex64))).to(self.device)
        if x.numel() % self.N:
            raise ValueError("rtty_filterbank.push: need whole symbols of N=%d samples" % self.N)
        if self.prev is not None:
            x = torch.cat((self.prev, x))
        n_sym = x.numel() // self.N
        out = torch.empty((max(4 * (n_sym - 1), 1), self.NFFT), dtype=torch.float32, device=self.device)
        got = ctypes.c_int64(0)
        if n_sym >= 2:
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            # n-1: the last symbol only closes the (prev, iq) pair, it does not start frames of its own
            check(self.lib.pysdr_psd_lines(self.h, ctypes.c_void_p(x.data_ptr()), x.numel() - 1, 1, 1, 1,
                                           ctypes.c_void_p(out.data_ptr()), ctypes.byref(got), st))
            assert got.value == 4 * (n_sym - 1)
        if n_sym >= 1:
            self.prev = x[(n_sym - 1) * self.N:].clone()
        out = out[:got.value]
        return out.cpu().numpy() if to_host else out
