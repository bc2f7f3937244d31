"""Compute part of the reference's ``three_box_plot`` (Plotting.py:312-434 ctor, :444-631 plot) without Qt:
PSD of the newest samples, waterfall shift-in / roll on retune, background level, peak picking and the
dynamic-range clipped image.  The PSD and the NFFT x 100 waterfall live on the device (K3, pysdr_waterfall_push);
only the peak picker runs on the host, on the NFFT-long row-mean vector, with the same scipy call as the reference
(Plotting.py:594)."""
import ctypes

import numpy as np
import torch

from . import _lib, sig_proc as dsp
from ._lib import check
from .bank import _stream_ptr


def jet(m=64):
    """The Matlab 'jet' colormap the reference tabulates (Tables.py:144-145, 64 RGBA rows), from its defining ramp."""
    n = int(np.ceil(m / 4.0))
    u = np.concatenate((np.arange(1, n + 1) / n, np.ones(n - 1), np.arange(n, 0, -1) / n))
    g = int(np.ceil(n / 2.0)) - (m % 4 == 1) + np.arange(1, len(u) + 1)
    r, b = g + n, g - n
    J = np.zeros((m, 3))
    gi = g[g <= m]
    J[gi - 1, 1] = u[:len(gi)]
    ri = r[r <= m]
    J[ri - 1, 0] = u[:len(ri)]
    bi = b[b >= 1]
    J[bi - 1, 2] = u[len(u) - len(bi):]
    out = np.full((m, 4), 255, np.uint8)
    out[:, :3] = np.floor(J * 255.0 + 0.5).astype(np.uint8)
    return out


def lookup_table(colors, npts=256):
    """ColorMap(pos=linspace(0,1,len), colors).getLookupTable(0, 1, npts): linear interpolation, ubyte
    (reference Plotting.py:139-141)."""
    colors = np.asarray(colors, np.float64)
    pos = np.linspace(0.0, 1.0, len(colors))
    x = np.linspace(0.0, 1.0, npts)
    lut = np.stack([np.interp(x, pos, colors[:, c]) for c in range(colors.shape[1])], axis=1)
    return lut.astype(np.uint8)


class three_box_compute:
    def __init__(self, P, fs, foff, chunk_size, Nfft, overlap, ncols=100):
        self.P = P
        self.foff = foff
        self.fc = 0
        if chunk_size > 65536:                                   # Plotting.py:370-375 (the reference's own workaround)
            chunk_size = int(65636 / 2)
            Nfft = 2 * chunk_size
        self.psd = dsp.spectrum(fs, chunk_size, Nfft, overlap)
        n = self.psd.NFFT
        self.lib = _lib.load()
        dev = self.psd.device
        self.ncols = ncols
        self.wf = torch.full((n, ncols), -1e38, dtype=torch.float32, device=dev)       # Plotting.py:385
        self.img = torch.empty((n, ncols), dtype=torch.float32, device=dev)
        self.bk = torch.zeros(1, dtype=torch.float32, device=dev)
        self.scratch = torch.empty(n * ncols + n + 8, dtype=torch.float32, device=dev)
        self.wf_cnt = 0
        self.wf_fc = 0
        self.pk_frqs = np.zeros(0)
        self.lut = torch.from_numpy(lookup_table(jet(64), 256)).to(dev)                # Plotting.py:139-141
        self.rgba = torch.empty((n, ncols, 4), dtype=torch.uint8, device=dev)
        self.pk_idx = torch.zeros(4096, dtype=torch.int32, device=dev)
        self.pk_cnt = torch.zeros(1, dtype=torch.int32, device=dev)

    def image_rgba(self, npsd=None):
        """RGBA8 waterfall image of the last plot() through the jet lookup table, on the device."""
        n = self.psd.NFFT if npsd is None else int(npsd)
        check(self.lib.pysdr_waterfall_rgba(ctypes.c_void_p(self.img.data_ptr()), n * self.ncols,
                                            ctypes.c_void_p(self.bk.data_ptr()), ctypes.c_void_p(self.scratch.data_ptr()),
                                            self.psd.NFFT, self.ncols, float(self.P.PAN_DR),
                                            ctypes.c_void_p(self.lut.data_ptr()), ctypes.c_void_p(self.rgba.data_ptr()),
                                            _stream_ptr()))
        return self.rgba[:n]

    def plot(self, y, fc):
        """One display frame (Plotting.py:444-631): returns dict(frq, PSD, image, bkgnd, peaks, pk_frqs)
        or None when the periodogram fails (Plotting.py:463-465)."""
        P = self.P
        PSD = self.psd.periodogram(y, True)                      # Plotting.py:462
        if len(PSD) == 0:
            return None
        frq = self.psd.frq - self.foff + fc                      # Plotting.py:467
        self.fc = fc - self.foff
        df = self.psd.frq[1] - self.psd.frq[0]                   # shift_waterfall, Plotting.py:689-695
        nbins = int(float(fc - self.wf_fc) / df + 0.5)
        if nbins != 0:
            self.wf_fc = fc
        if getattr(P, 'RIG_IF', 0) < 0:                          # Plotting.py:538-539
            PSD = np.flipud(PSD)
        if self.wf_cnt < self.ncols:
            self.wf_cnt += 1
        n = self.psd.NFFT
        line = torch.from_numpy(np.ascontiguousarray(PSD, np.float32)).to(self.wf.device)
        check(self.lib.pysdr_waterfall_push(ctypes.c_void_p(self.wf.data_ptr()), n, self.ncols, self.wf_cnt,
                                            ctypes.c_void_p(line.data_ptr()), len(PSD), nbins, float(P.PAN_DR),
                                            ctypes.c_void_p(self.img.data_ptr()), ctypes.c_void_p(self.bk.data_ptr()),
                                            ctypes.c_void_p(self.scratch.data_ptr()), _stream_ptr()))
        # Plotting.py:594 find_peaks(PSD2, distance=PEAK_DIST/df, height=bkgnd+10) on the device: PSD2 (row means over the last
        # wf_cnt lines) and the median never leave it; only the peak indices come back, with the background, in one read
        PSD2 = self.scratch[n * self.ncols:n * self.ncols + n]
        dist = P.PEAK_DIST / self.psd.df
        check(self.lib.pysdr_find_peaks(ctypes.c_void_p(PSD2.data_ptr()), n, ctypes.c_void_p(self.bk.data_ptr()), 10.0, 0.0,
                                        float(dist), ctypes.c_void_p(self.pk_idx.data_ptr()), ctypes.c_void_p(self.pk_cnt.data_ptr()),
                                        _stream_ptr()))
        cnt = int(self.pk_cnt.item())
        peaks = self.pk_idx[:cnt].cpu().numpy().astype(np.int64)
        bkgnd = float(self.bk.item())
        self.pk_frqs = frq[peaks]
        return dict(frq=frq, PSD=PSD, image=self.img[:len(PSD)], bkgnd=bkgnd, peaks=peaks, pk_frqs=self.pk_frqs)
