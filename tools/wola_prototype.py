#!/usr/bin/env python
"""Design aid for the next step of config 5 (NOT on the product path): for channels on a uniform raster the per-channel
fused mix + polyphase FIR

    Y[c, m] = e^{-j th_c(n_m)} * sum_j h[p_m + UP*j] * e^{+j w_c j} * x[n_m - j],      w_c = 2 pi (f0 + c*df) / fs

collapses to ONE shared windowing pass and ONE inverse DFT per output instant when df/fs = a/Nd in lowest terms
(10 MS/s, 9.6 kHz raster: a/Nd = 3/3125):

    v_m[j]   = h[p_m + UP*j] * e^{+j w_0 j} * x[n_m - j]                      (lp = 334 products, shared by all channels)
    Y[c, m]  = e^{-j th_c(n_m)} * sum_j v_m[j] * e^{+j 2 pi (a c mod Nd) j / Nd}   = Nd * IDFT_Nd(v_m zero-padded)[a c mod Nd]

i.e. 334 complex multiplies + a 3125-point (5^5) DFT instead of 1024 x 334 complex MACs per output instant — about 15x
fewer flops, after which config 5 is bound by its 27.66 B/sample of HBM traffic.  This script checks the identity in
float64 against the direct form that the product (ChannelBank / K1) and the oracle evaluate."""
import numpy as np
from math import gcd


def direct(x, h, up, down, fs, f0, df, n_ch, m_list):
    lp = (len(h) + up - 1) // up
    hp = np.zeros(lp * up)
    hp[:len(h)] = h
    out = np.zeros((n_ch, len(m_list)), complex)
    j = np.arange(lp)
    for k, m in enumerate(m_list):
        t = m * down
        nm, pm = t // up, t % up
        xs = x[nm - j]
        for c in range(n_ch):
            w = 2 * np.pi * (f0 + c * df) / fs
            out[c, k] = np.exp(-1j * w * nm) * np.sum(hp[pm + up * j] * np.exp(1j * w * j) * xs)
    return out


def wola(x, h, up, down, fs, f0, df, n_ch, m_list):
    g = gcd(int(round(df)), int(round(fs)))
    a, nd = int(round(df)) // g, int(round(fs)) // g
    lp = (len(h) + up - 1) // up
    assert lp <= nd
    hp = np.zeros(lp * up)
    hp[:len(h)] = h
    j = np.arange(lp)
    w0 = 2 * np.pi * f0 / fs
    out = np.zeros((n_ch, len(m_list)), complex)
    bins = (a * np.arange(n_ch)) % nd
    for k, m in enumerate(m_list):
        t = m * down
        nm, pm = t // up, t % up
        v = np.zeros(nd, complex)
        v[:lp] = hp[pm + up * j] * np.exp(1j * w0 * j) * x[nm - j]
        z = np.fft.ifft(v) * nd                                  # sum_j v[j] e^{+j 2 pi k j / nd}
        wc = 2 * np.pi * (f0 + np.arange(n_ch) * df) / fs
        out[:, k] = np.exp(-1j * wc * nm) * z[bins]
    return out, (a, nd)


if __name__ == "__main__":
    from scipy import signal
    rng = np.random.default_rng(0)
    fs, up, down = 10e6, 3, 625
    h = signal.firwin(1001, 5e3, window='hamming', fs=fs * up) * up
    x = rng.normal(size=40000) + 1j * rng.normal(size=40000)
    ms = [40, 41, 42, 100, 173]
    d = direct(x, h, up, down, fs, -4.9e6, 9600.0, 64, ms)
    w, (a, nd) = wola(x, h, up, down, fs, -4.9e6, 9600.0, 64, ms)
    print("raster %d/%d of fs; max |direct - wola| / max |direct| = %.2e" % (a, nd, np.max(np.abs(d - w)) / np.max(np.abs(d))))
