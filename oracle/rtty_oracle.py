"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference RTTY executive's FFT
filterbank.  Unlike the receiver chain this arithmetic is fully in-tree, so this oracle is PINNED: the fixture
tests/golden/rtty_fbank.npz was produced by the reference's own RTTY_Executive.run loop (rtty.py:780-856) executed
in the build container by tests/golden/make_golden_rtty.py; tests/test_oracle_pins.py checks this file against it.
"""
import math

import numpy as np


def nextpow2(n):
    """rtty.py:80-82."""
    return math.ceil(math.log(n, 2))


class RTTY_Params:
    """rtty.py:376-404."""

    def __init__(self, FS_OUT):
        self.T = 22e-3                                                  # rtty.py:380
        self.FSK_SHIFT = 170                                            # rtty.py:382
        self.N = int(round(self.T * FS_OUT))                            # rtty.py:389
        self.NFFT = int(2 ** nextpow2(self.N))                          # rtty.py:390
        NSTEP = self.N / 4.                                             # rtty.py:391
        self.NSTART = [int(NSTEP * i + 0.5) for i in range(4)]          # rtty.py:392-394
        bin_size = FS_OUT / float(self.NFFT)                            # rtty.py:398
        self.NBINS = int(round(self.FSK_SHIFT / bin_size))              # rtty.py:399
        self.frq = np.fft.fftshift(np.fft.fftfreq(self.NFFT, d=1000. / FS_OUT)) + 0   # rtty.py:402


def filterbank_lines(iq, FS_OUT):
    """Lines the executive computes for a stream of whole symbols (rtty.py:825-856): for every symbol after the
    first, x = [prev, iq] and four FFTs at the quarter-symbol starts; line = flipud(fftshift(10 log10 |X|^2))."""
    R = RTTY_Params(FS_OUT)
    window = np.kaiser(R.N, 8.6)                                        # rtty.py:807
    iq = np.asarray(iq)
    n_sym = len(iq) // R.N
    lines = []
    prev = None
    for s in range(n_sym):
        cur = iq[s * R.N:(s + 1) * R.N]                                 # rb.pull(N), rtty.py:825
        if prev is None:                                                # rtty.py:826-829
            prev = cur
            continue
        x = np.concatenate((prev, cur))                                 # rtty.py:837
        for i in range(4):                                              # rtty.py:837
            xx = x[R.NSTART[i]:(R.NSTART[i] + R.N)]                     # rtty.py:837
            X = np.fft.fftshift(np.fft.fft(xx * window, R.NFFT))        # rtty.py:839
            with np.errstate(divide='ignore'):
                XX = 10 * np.log10(np.square(X.real) + np.square(X.imag))   # rtty.py:841
            lines.append(np.flipud(XX))                                 # rtty.py:843
        prev = cur                                                      # rtty.py:856
    return np.array(lines).reshape(-1, R.NFFT)
