"""CPU oracle for the pySDR receive-DSP hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy/scipy *restatement* of the L1 DSP operators that
aa2il/pySDR's ``receiver.py`` / ``Plotting.py`` call through ``import sig_proc as
dsp`` (reference ``receiver.py:45``, ``Plotting.py:31``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; the product package ``pysdr_b200`` never does.

PARITY STATUS: **parity unpinned against upstream ``sig_proc``.**  That module
lives in the separate, un-vendored, un-pinned repo ``github.com/aa2il/libs``
(reference ``README.md:42-49,86-87``) which is absent from ``/root/reference``
and cannot be fetched (no network).  What *is* pinned, and checked in
``tests/test_oracle_pins.py``:

  * ``up_dn``                  -> 39-row golden table, reference ``srates.py:35-74``
  * ``IN_CHUNK_SIZE`` etc.     -> reference ``params.py:405-406,440-444``
  * ``adjust_foffset``         -> reference ``utils.py:277-289``
  * chunked ``lfilter`` carry  -> reference ``sigs/iir.py:83-125`` (+ notch coefficients)
  * AGC loop filter            -> reference ``sigs/agc.m:6-12``  (b=beta, a=[1 beta-1])
  * NFM discriminator          -> reference ``sigs/nfm.m:123-127``
  * squelch envelope smoother  -> reference ``sigs/squelch.m:125-128``
  * windowed-FFT idiom         -> reference ``rtty.py:807,839-841``
  * waterfall algebra          -> reference ``Plotting.py:536-548,583-626,689-695``

Every choice the reference tree does not determine is fixed ONCE here and
labelled "OPEN CHOICE" (see DESIGN.md section 3 for the list).

Numerics: inputs are complex64; the oracle evaluates every stage in float64 /
complex128 from float32-stored taps, and casts results to float32/complex64 at
the API surface, i.e. it is the "infinitely precise" evaluation of the same
float32 data + float32 coefficients the CUDA path consumes.  ``dtype=np.complex64``
switches the heavy stages to single precision (what numpy would do on complex64
arrays) and is used only for CPU-baseline timing.
"""
from __future__ import annotations

import math
from math import gcd

import numpy as np
from scipy import signal

# ----------------------------------------------------------------------------------------------
# Tables (restated from reference Tables.py:34-45; the module itself cannot be imported because it
# pulls rig_io / xlrd / unidecode).
# ----------------------------------------------------------------------------------------------
MODES = ["AM", "AM-Synch", "SSB", "USB", "LSB", "CW", "IQ", "WFM", "WFM2", "NFM", "RTTY"]
AF_BWs = ['Max', '50 Hz', '100 Hz', '500 Hz', '1 KHz', '2 KHz', '3 KHz',
          '4 KHz', '5 KHz', '8 KHz', '10 KHz', '15 KHz', '20 KHz', '45 KHz', '50 KHz', '100 KHz',
          '150 KHz', '200 KHz']
VIDEO_BWs = ['Max', '5 KHz', '10 KHz', '20 KHz', '25 KHz', '45 KHz', '50 KHz', '100 KHz',
             '150 KHz', '200 KHz', '300 KHz', '400 KHz', '500 KHz', '750 KHz', '1 MHz', 'Other']
RTLsrates = [0.25, 1.024, 1.536, 1.792, 1.92, 2.048, 2.16, 2.56, 2.88, 3.2]
SDRplaysrates = [0.25, 0.5, 1, 2, 2.048, 3, 4, 5, 6, 7, 8, 9, 10]
MAX_RX = 6                                                   # reference params.py:33


def bw_label_to_hz(label):
    """'10 KHz' -> 10e3, '1 MHz' -> 1e6, '50 Hz' -> 50; 'Max'/'Other' -> None
    (same parsing as reference Tables.py:48-62 / gui.py:1741-1751)."""
    if label in ('Max', 'Other'):
        return None
    a = label.split(" ")
    b = float(int(a[0]))
    if a[1] == "KHz":
        b *= 1e3
    elif a[1] == "MHz":
        b *= 1e6
    return b


def find_filter(max_bw, bw_list):
    """Widest table entry <= max_bw (reference Tables.py:48-62)."""
    best = None
    for bw in bw_list:
        b = bw_label_to_hz(bw)
        if b is not None and b <= max_bw:
            best = bw
    return best


# ----------------------------------------------------------------------------------------------
# a1: up_dn                                                           (golden: reference srates.py:35-74)
# ----------------------------------------------------------------------------------------------
def up_dn(fs1, fs2):
    """Rate-change factors: fs2/fs1 == UP/DOWN in lowest terms."""
    f1 = int(round(fs1))
    f2 = int(round(fs2))
    g = gcd(f1, f2)
    return f2 // g, f1 // g


def derived_rates(srate, fs_out_req, out_chunk=1024):
    """UP, DOWN, FS_OUT, IN_CHUNK_SIZE exactly as reference params.py:405-406,440-444."""
    up, down = up_dn(srate, fs_out_req)
    fs_out = int(srate * up / down)
    in_chunk = int(out_chunk * down / float(up) + 0 * 0.5)
    return up, down, fs_out, in_chunk


def rb_size(num_rx, fs_out, sdr_type='sdrplay', out_chunk=1024):
    """RB_SIZE rule, reference params.py:456-468."""
    rb = 32 * out_chunk
    if num_rx > 2:
        rb *= 4
    if sdr_type == 'rtlsdr':
        rb *= 2
    if fs_out > 100e3:
        rb *= 4
    elif fs_out > 50e3:
        rb *= 2
    return rb


# ----------------------------------------------------------------------------------------------
# a2: adjust_foffset                                                        (reference utils.py:277-289)
# ----------------------------------------------------------------------------------------------
def adjust_foffset(foffset, srate, rb):
    M = round(rb * foffset / srate)
    return M * srate / rb


# ----------------------------------------------------------------------------------------------
# a3: NCO / signal_generator
# OPEN CHOICE: the NCO phase is an exact 64-bit fixed-point accumulator, phase[n] = acc0 + n*inc
# (mod 2^64 cycles/2^64), inc = trunc(frac(f/fs) * 2^64).  It is a pure function of the absolute
# sample index (chunk- and shard-invariant).  quad_mixer multiplies by exp(-j*2*pi*phase): a signal
# at +f lands on 0 Hz ("Shift by tuning offset", reference receiver.py:551-553).
# ----------------------------------------------------------------------------------------------
TWO64 = 2.0 ** 64
MASK64 = (1 << 64) - 1


def freq_to_phase_inc(f, fs):
    r = float(f) / float(fs)
    r = r - math.floor(r)
    return int(r * TWO64) & MASK64


def phase_inc_to_freq(inc, fs):
    """Frequency actually applied (what change_freq returns, cf. reference gui.py:1928)."""
    inc = int(inc) & MASK64
    if inc >= (1 << 63):
        inc -= (1 << 64)
    return inc / TWO64 * float(fs)


def nco_phase_cycles(acc0, inc, n):
    """phase (in cycles, in [-0.5,0.5)) of samples acc0 + inc*k for k in n (int array); exact mod-2^64."""
    k = np.asarray(n).astype(np.int64).astype(np.uint64)     # negative indices wrap mod 2^64
    with np.errstate(over='ignore'):
        ph = np.uint64(acc0 & MASK64) + np.uint64(inc & MASK64) * k          # wraps mod 2^64
    # OPEN CHOICE: only the top 32 bits feed sin/cos (2^-32 cycle = 1.5e-9 rad resolution).
    top = (ph >> np.uint64(32)).astype(np.uint32).view(np.int32).astype(np.float64)
    return top * (2.0 ** -32)


class signal_generator:
    """NCO. Surface used by the reference: ctor ``(f,N,fs,complex)`` (receiver.py:822), ``.fo``,
    ``.quad_mixer(x)`` (receiver.py:552-553), ``.change_freq(f)`` -> applied f (gui.py:1928,1938)."""

    def __init__(self, f, N, fs, cmplx=True):
        self.N = int(N)
        self.fs = float(fs)
        self.complex = bool(cmplx)
        self.acc = 0                       # phase accumulator (uint64 cycles*2^64) at the next sample
        self.change_freq(f)

    def change_freq(self, f):
        self.inc = freq_to_phase_inc(f, self.fs)
        self.fo = phase_inc_to_freq(self.inc, self.fs)
        return self.fo

    def advance(self, n):
        self.acc = (self.acc + self.inc * int(n)) & MASK64

    def lo(self, n):
        """n LO samples exp(+j*2*pi*phase) from the current accumulator; advances it."""
        ph = nco_phase_cycles(self.acc, self.inc, np.arange(n))
        self.advance(n)
        z = np.exp(2j * np.pi * ph)
        return z if self.complex else z.real

    def quad_mixer(self, x):
        x = np.asarray(x)
        z = self.lo(len(x))
        return (x.astype(np.complex128) * np.conj(z)).astype(np.complex64)


# ----------------------------------------------------------------------------------------------
# a5: polyphase rational resampler ("dec"), taps + streaming arithmetic.
# Indexing contract (bit-exact gate): output m uses phase p=(m*DOWN)%UP, newest input
# n_m=(m*DOWN)//UP:      y[m] = sum_j h[p + j*UP] * xmix[n_m - j]          (x[<0] = 0)
# After n inputs have been consumed exactly ceil(UP*n/DOWN) outputs exist.
# OPEN CHOICE (taps): scipy firwin, Hamming window, FILT_LEN taps designed at rate SRATE*UP,
# cutoff VIDEO_BW/2 (two-sided video bandwidth), DC gain UP, stored as float32.
# ----------------------------------------------------------------------------------------------
def design_lowpass(ntaps, cutoff_hz, fs_hz, gain=1.0):
    cutoff_hz = min(float(cutoff_hz), 0.45 * fs_hz)
    h = signal.firwin(int(ntaps), cutoff_hz, window='hamming', fs=float(fs_hz)) * gain
    return h.astype(np.float32)


def video_cutoff_hz(label, srate, fs_out, video_bw_other):
    bw = bw_label_to_hz(label)
    lim = 0.45 * min(srate, fs_out)
    if label == 'Max':
        return lim
    if label == 'Other':
        bw = video_bw_other
    return min(0.5 * bw, 0.45 * srate)


def design_resampler_bank(srate, up, down, filt_len, video_bws=VIDEO_BWs, video_bw_other=10e3):
    fs_out = srate * up / down
    fs_v = srate * up
    return [design_lowpass(filt_len, video_cutoff_hz(lb, srate, fs_out, video_bw_other), fs_v, gain=up)
            for lb in video_bws]


def n_out_total(n_in, up, down):
    """# outputs that exist once n_in inputs were consumed: ceil(UP*n/DOWN)."""
    return -((-up * n_in) // down)


class decimator:
    """Streaming polyphase resampler; ``.h`` is assignable from ``.filter_bank[idx]``
    (reference gui.py:1713, receiver.py:127).

    OPEN CHOICE: the filter memory holds RAW input samples and the LO is applied at filter time
    (``resamp(x, lo)``): with a fixed LO this is identical to mix-then-filter; after ``change_freq``
    the new LO is also applied to the (lp-1)-sample memory, phase-continuous at the chunk boundary."""

    def __init__(self, srate, up, down, filt_len, video_bws=VIDEO_BWs, video_bw_other=10e3,
                 dtype=np.complex128):
        self.up, self.down = int(up), int(down)
        self.filter_bank = design_resampler_bank(srate, up, down, filt_len, video_bws, video_bw_other)
        self.h = self.filter_bank[0]
        self.dtype = dtype
        self.reset()

    def reset(self):
        self.n0 = 0                                  # absolute index of the next input sample
        self.hist = np.zeros(0, self.dtype)          # raw (un-mixed) input memory

    def _lp(self):
        return -(-len(self.h) // self.up)            # taps per polyphase branch (ceil)

    def _extended(self, x, lo, H):
        """[H raw history samples ; x], mixed with the LO whose accumulator sits at sample n0."""
        x = np.asarray(x).astype(self.dtype)
        hist = self.hist
        if len(hist) < H:
            hist = np.concatenate((np.zeros(H - len(hist), self.dtype), hist))
        raw = np.concatenate((hist[len(hist) - H:], x))
        if lo is None:
            return raw, raw
        if getattr(self, 'lo_table', False):
            # CPU-baseline path: the LO as a two-level table instead of one complex exponential per sample — the reference
            # arranges its tuning offset "so we don't have to compute sines/cosines over and over in the local osc"
            # (params.py:470-471, utils.py:277-289).  exp(-j th(k)) with th(k) = base + inc*k, k = q*B + r, is
            # T2[q] * T1[r]: T1 (B entries) is cached per LO increment, T2 has n/B entries per chunk; phases stay exact u64.
            B = 4096
            n = H + len(x)
            if getattr(self, '_t1_inc', None) != lo.inc:
                self._t1 = np.exp(-2j * np.pi * nco_phase_cycles(0, lo.inc, np.arange(B))).astype(self.dtype)
                self._t1_inc = lo.inc
            base = (lo.acc - lo.inc * H) & MASK64
            q = np.arange(-(-n // B), dtype=np.int64)
            t2 = np.exp(-2j * np.pi * nco_phase_cycles(base, (lo.inc * B) & MASK64, q)).astype(self.dtype)
            z = (t2[:, None] * self._t1[None, :]).reshape(-1)[:n]              # in self.dtype: no widening pass
            return raw, raw * z
        k = np.arange(-H, len(x), dtype=np.int64)
        ph = nco_phase_cycles(lo.acc, lo.inc, k)      # negative k wraps mod 2^64 = phase run backwards
        z = np.exp(-2j * np.pi * ph)
        return raw, (raw * z).astype(self.dtype)

    def resamp(self, x, lo=None):
        """Definition-level evaluation (gather + dot per phase). Advances lo by len(x)."""
        up, down = self.up, self.down
        wide = np.float64 if self.dtype == np.complex128 else np.float32
        h = np.asarray(self.h, wide)
        lp = self._lp()
        hp = np.zeros(lp * up, wide)
        hp[:len(h)] = h
        n0, n1 = self.n0, self.n0 + len(x)
        m0, m1 = n_out_total(n0, up, down), n_out_total(n1, up, down)
        need = lp - 1
        raw, xx = self._extended(x, lo, need)        # xx[k] <-> absolute index n0 - need + k
        m = np.arange(m0, m1, dtype=np.int64)
        t = m * down
        nm = t // up
        pm = t % up
        y = np.zeros(len(m), self.dtype)
        j = np.arange(lp, dtype=np.int64)
        for p in range(up):
            sel = np.nonzero(pm == p)[0]
            if len(sel) == 0:
                continue
            idx = (nm[sel] - (n0 - need))[:, None] - j[None, :]
            y[sel] = xx[idx] @ hp[p::up].astype(self.dtype)
        self.hist = raw[max(0, len(raw) - (need + down)):]
        self.n0 = n1
        if lo is not None:
            lo.advance(len(x))
        return y

    def resamp_fast(self, x, lo=None):
        """Same numbers through scipy.signal.upfirdn (used for CPU-baseline timing)."""
        up, down = self.up, self.down
        wide = np.float64 if self.dtype == np.complex128 else np.float32
        h = np.asarray(self.h, wide)
        lp = self._lp()
        n0, n1 = self.n0, self.n0 + len(x)
        m0, m1 = n_out_total(n0, up, down), n_out_total(n1, up, down)
        H = (n0 % down)
        while H < lp - 1:
            H += down                                 # (n0-H) % DOWN == 0: upfirdn's grid == ours
        raw, xx = self._extended(x, lo, H)
        y = signal.upfirdn(h, xx, up, down)
        q0 = m0 - ((n0 - H) // down) * up             # local index of absolute output m0
        y = y[q0:q0 + (m1 - m0)]
        keep = lp - 1 + down
        self.hist = raw[max(0, len(raw) - keep):]
        self.n0 = n1
        if lo is not None:
            lo.advance(len(x))
        return y.astype(self.dtype)


# ----------------------------------------------------------------------------------------------
# a6: demodulator filter banks + demod arithmetic (all at FS_OUT).
# OPEN CHOICES: FIR length = FILT_LEN (the author uses 1001 taps at audio rate too, reference
# receiver.py:861), Hamming firwin;  real bank cutoff = AF_BW, 'Max' = pass-through (delta);
# complex bank = low-pass(AF_BW/2) shifted up by AF_BW/2 (pass-band 0..AF_BW, USB; LSB conjugates);
# CW = low-pass(AF_BW/2) on the complex baseband then BFO re-insertion Re{. * e^{+j 2pi BFO n/fs}};
# NFM = 3-point discriminator of reference sigs/nfm.m:123-127 (no limiter), one-sample latency.
# ----------------------------------------------------------------------------------------------
def af_bw_hz(label, fs_out):
    bw = bw_label_to_hz(label)
    if bw is None:
        return 0.0
    return bw


def design_af_bank_real(fs_out, ntaps, af_bws=AF_BWs):
    bank = []
    for lb in af_bws:
        bw = af_bw_hz(lb, fs_out)
        if bw <= 0:
            h = np.zeros(ntaps, np.float32)
            h[(ntaps - 1) // 2] = 1.0                 # 'Max': pure delay, same latency as the others
        else:
            h = design_lowpass(ntaps, bw, fs_out)
        bank.append(h)
    return bank


def design_af_bank_cmpx(fs_out, ntaps, af_bws=AF_BWs):
    bank = []
    c = (ntaps - 1) / 2.0
    j = np.arange(ntaps)
    for lb in af_bws:
        bw = af_bw_hz(lb, fs_out)
        if bw <= 0:
            bw = 0.9 * fs_out / 2
        bw = min(bw, 0.9 * fs_out / 2)
        h = design_lowpass(ntaps, bw / 2, fs_out).astype(np.float64)
        g = h * np.exp(2j * np.pi * (bw / 2) * (j - c) / fs_out)
        bank.append(g.astype(np.complex64))
    return bank


def design_af_bank_cw(fs_out, ntaps, af_bws=AF_BWs):
    """Low-pass(AF_BW/2) prototypes used by CW (pre-BFO) and IQ."""
    bank = []
    for lb in af_bws:
        bw = af_bw_hz(lb, fs_out)
        if bw <= 0:
            h = np.zeros(ntaps, np.float32)
            h[(ntaps - 1) // 2] = 1.0
        else:
            h = design_lowpass(ntaps, bw / 2, fs_out)
        bank.append(h)
    return bank


PLL_BN_HZ = 50.0
PLL_ZETA = 0.70710678118654752440


class am_pll:
    """AM-Synch carrier loop (``demod.am_pll.reset()``, reference receiver.py:649).  OPEN CHOICE (no in-tree law):
    second-order PLL, atan2 phase detector, noise bandwidth 50 Hz, damping 1/sqrt(2), run at the audio rate:
        v = z e^{-j phi};  e = atan2(Im v, Re v);  w += K2 e;  phi += w + K1 e  (wrapped to [-pi, pi))
    ``run`` returns v (the detector uses Re v)."""

    def __init__(self, fs):
        th = PLL_BN_HZ / float(fs) / (PLL_ZETA + 1.0 / (4.0 * PLL_ZETA))
        d = 1.0 + 2.0 * PLL_ZETA * th + th * th
        self.k1 = 4.0 * PLL_ZETA * th / d
        self.k2 = 4.0 * th * th / d
        self.reset()

    def reset(self):
        self.phi = 0.0
        self.w = 0.0

    def run(self, z):
        z = np.asarray(z, np.complex128)
        v = np.empty(len(z), np.complex128)
        phi, w, k1, k2 = self.phi, self.w, self.k1, self.k2
        for i in range(len(z)):
            vi = z[i] * complex(math.cos(phi), -math.sin(phi))
            v[i] = vi
            e = math.atan2(vi.imag, vi.real)
            w += k2 * e
            phi += w + k1 * e
            if phi >= math.pi:
                phi -= 2.0 * math.pi
            elif phi < -math.pi:
                phi += 2.0 * math.pi
        self.phi, self.w = phi, w
        return v


class _holder:
    pass


class demodulator:
    """All modes share ONE carried memory: the last L+1 complex baseband samples (L-1 for the FIR,
    +2 for the 3-point NFM discriminator).  Detection (|.| or discriminator) is recomputed over that
    memory each call, so a mode change re-interprets the memory under the new mode."""

    def __init__(self, fs_out, filt_len, af_bws=AF_BWs, dtype=np.complex128, exact=True):
        self.fs = float(fs_out)
        self.L = int(filt_len)
        self.filter_bank_real = design_af_bank_real(fs_out, filt_len, af_bws)
        self.filter_bank_cmpx = design_af_bank_cmpx(fs_out, filt_len, af_bws)
        self.filter_bank_lp = design_af_bank_cw(fs_out, filt_len, af_bws)
        self.am_pll = am_pll(fs_out)
        self.wfm_video = _holder()
        self.wfm_video.h = None
        self.wfm_filter_bank = []
        self.dtype = dtype
        self.rdtype = np.float64 if dtype == np.complex128 else np.float32
        self.exact = exact
        self.reset()

    def reset(self):
        self.hist_c = np.zeros(self.L + 1, self.dtype)
        self.m0 = 0                                    # absolute output index (BFO phase)

    def _fir(self, g, src, n):
        """valid convolution: src has len(g)-1+n samples -> n outputs."""
        g = np.asarray(g)
        gw = g.astype(self.dtype if np.iscomplexobj(g) else self.rdtype)
        if n == 0:
            return np.zeros(0, np.result_type(gw.dtype, src.dtype))
        src = src[len(src) - (len(g) - 1 + n):]
        if self.exact or n < 256:
            return np.convolve(src, gw, mode='valid')
        return signal.fftconvolve(src, gw, mode='valid')

    def demod(self, iq, mode, af_idx, bfo_hz):
        """iq: complex baseband chunk @FS_OUT -> pre-AGC audio (real; complex for IQ/RTTY)."""
        iq = np.asarray(iq).astype(self.dtype)
        n = len(iq)
        m = self.m0 + np.arange(n)
        if mode == 'AM-Synch':                           # the carried memory then holds DE-ROTATED samples
            iq = self.am_pll.run(iq).astype(self.dtype)
        xx = np.concatenate((self.hist_c, iq))           # xx[k] <-> output index m0-(L+1)+k
        if mode == 'AM-Synch':
            a = self._fir(self.filter_bank_real[af_idx], xx[2:].real, n)
        elif mode == 'AM':
            a = self._fir(self.filter_bank_real[af_idx], np.abs(xx[2:]), n)
        elif mode in ('USB', 'SSB', 'LSB'):
            g = self.filter_bank_cmpx[af_idx]
            if mode == 'LSB':
                g = np.conj(g)
            a = self._fir(g, xx[2:], n).real
        elif mode == 'CW':
            z = self._fir(self.filter_bank_lp[af_idx], xx[2:], n)
            inc = freq_to_phase_inc(bfo_hz, self.fs)
            ph = nco_phase_cycles(0, inc, m)
            a = (z * np.exp(2j * np.pi * ph)).real
        elif mode in ('IQ', 'RTTY'):
            a = self._fir(self.filter_bank_lp[af_idx], xx[2:], n)
        elif mode == 'NFM':
            d = xx[2:] - xx[:-2]                         # nfm.m:124  d = IQ - y(1:end-2)
            y1 = xx[1:-1]                                # nfm.m:125
            fm = y1.real * d.imag - y1.imag * d.real     # nfm.m:126  (one sample of latency)
            a = self._fir(self.filter_bank_real[af_idx], fm, n)
        else:
            raise ValueError('mode %s not supported by the oracle' % mode)
        self.hist_c = xx[len(xx) - (self.L + 1):]
        self.m0 += n
        return a


# ----------------------------------------------------------------------------------------------
# a7: AGC.  Pinned: loop filter y = beta*x + (1-beta)*y_1 (reference sigs/agc.m:6-12, beta=.1) and the
# attribute names .agc .gain .maxbuf .ref .err (reference watchdog.py:298-302: MAX printed as a scalar
# and GAIN*MAX compared with the slider).
# OPEN CHOICE (the law around the loop filter): block AGC, one update per demod_data call
# (= per IN_CHUNK_SIZE input block): pk=max|a|; maxbuf=max(last NB block peaks); want=min(ref/maxbuf,
# GMAX); err=want-gain; attack (want<gain): gain=want at once; decay: gain=beta*want+(1-beta)*gain.
# ----------------------------------------------------------------------------------------------
AGC_NB = 8
AGC_REF = 0.25
AGC_BETA = 0.1
AGC_GMAX = 1.0e4
AGC_FLOOR = 1.0e-9


class agc:
    def __init__(self, ref=AGC_REF, beta=AGC_BETA, nb=AGC_NB):
        self.ref = float(ref)
        self.beta = float(beta)
        self.nb = int(nb)
        self.reset()

    def reset(self):
        self.ring = [0.0] * self.nb
        self.k = 0
        self.gain = 1.0
        self.agc = 1.0
        self.maxbuf = 0.0
        self.err = 0.0

    def update(self, pk):
        """One block: returns the gain applied to that block."""
        self.ring[self.k % self.nb] = float(pk)
        self.k += 1
        self.maxbuf = max(self.ring)
        want = min(self.ref / max(self.maxbuf, AGC_FLOOR), AGC_GMAX)
        self.err = want - self.gain
        if want < self.gain:
            self.gain = want
        else:
            self.gain = self.beta * want + (1.0 - self.beta) * self.gain
        self.agc = want                      # the gain the loop asks for; .gain is what it applies
        return self.gain

    def run(self, a):
        if len(a) == 0:
            return a
        g = self.update(np.max(np.abs(a)))
        return a * g


# ----------------------------------------------------------------------------------------------
# a4: Receiver
# ----------------------------------------------------------------------------------------------
def per_rx(v, irx):
    """P.MODE / P.AF_BW / P.BFO ... may be one value for all receivers (the reference) or a per-RX list
    (extension used by the 4-RX AM/NFM/USB/CW config; in MP_SCHEME 3 every RX process owns its P anyway)."""
    return v[irx] if isinstance(v, (list, tuple, np.ndarray)) else v


def _af_index(P, irx=0):
    """AF filter index from P at call time; None/-1 -> look AF_BW up in the table, miss -> 0 'Max'
    (reference gui.py:1720-1731, am.py:57, mp.py:76, params.py:202)."""
    idx = per_rx(getattr(P, 'AF_FILTER_NUM', None), irx)
    if idx is None or idx < 0:
        bw = per_rx(getattr(P, 'AF_BW', 0), irx)
        idx = 0
        for i, lb in enumerate(AF_BWs):
            b = bw_label_to_hz(lb)
            if b is not None and b == bw:
                idx = i
                break
    return idx


def _video_index(P, labels=VIDEO_BWs):
    """Start-up video filter: VIDEO_BW looked up in the table else last entry 'Other'
    (reference gui.py:1675-1685)."""
    idx = getattr(P, 'VIDEO_FILTER_NUM', None)
    if idx is not None and idx >= 0:
        return idx
    bw = P.VIDEO_BW
    lab = (str(int(bw * 1e-6)) + ' MHz') if bw > 1e6 - 1 else (str(int(bw * 1e-3)) + ' KHz')
    return labels.index(lab) if lab in labels else len(labels) - 1


class Receiver:
    """``dsp.Receiver(P, frq, irx, name, VIDEO_BWs, AF_BWs)`` (reference receiver.py:65,835);
    ``.demod_data(x)`` -> am (receiver.py:235), side effects ``.am`` ``.iq``."""

    def __init__(self, P, frq, irx, name, video_bws=VIDEO_BWs, af_bws=AF_BWs,
                 dtype=np.complex128, fast=False):
        self.P = P
        self.irx = irx
        self.name = name
        self.sub = 0
        self.dtype = dtype
        self.fast = fast
        self.lo = signal_generator(frq, P.IN_CHUNK_SIZE, P.SRATE, True)
        self.dec = decimator(P.SRATE, P.UP, P.DOWN, P.FILT_LEN, video_bws, P.VIDEO_BW, dtype)
        self.dec.h = self.dec.filter_bank[_video_index(P, video_bws)]
        self.dec.lo_table = bool(fast)               # CPU-baseline runs: table LO (same numbers to ~2e-16)
        self.demod = demodulator(P.FS_OUT, P.FILT_LEN, af_bws, dtype, exact=not fast)
        self.agc = agc()
        self.mute_cnt = 0
        self.am = np.zeros(0, np.float32)
        self.iq = np.zeros(0, np.complex64)

    def _mode(self):
        return per_rx(self.P.MODE, self.irx)

    # ---- WFM / WFM2: "BCB FM is wideband so we need to demodulate first before resampling" (reference
    # gui.py:1703): video FIR at the RF rate (demod.wfm_video.h = demod.wfm_filter_bank[idx], gui.py:1704) ->
    # FM discriminator at the RF rate -> resampler whose filter is chosen by the AF bandwidth ("the audio
    # filtering is done in the resampler", gui.py:1759-1762) -> AGC.
    # OPEN CHOICES: video FIR = firwin(FILT_LEN, VIDEO_BW/2, fs=SRATE); discriminator = the 3-point form of
    # sigs/nfm.m:123-127; resampler low-pass cutoff = AF_BW (15 kHz when AF_BW is 0 / 'Max'), gain UP; optional
    # one-pole de-emphasis P.DEEMPH_US (0 = off).  Mono only: stereo pilot recovery has no in-tree specification.
    def _wfm_setup(self):
        P = self.P
        self.demod.wfm_filter_bank = [design_lowpass(P.FILT_LEN, min(0.5 * (bw_label_to_hz(lb) or P.VIDEO_BW), 0.45 * P.SRATE),
                                                     P.SRATE) for lb in VIDEO_BWs]
        self.demod.wfm_video.h = self.demod.wfm_filter_bank[_video_index(P)]
        self.wfm_vid = decimator(P.SRATE, 1, 1, P.FILT_LEN, VIDEO_BWs, P.VIDEO_BW, self.dtype)
        self.wfm_prev2 = np.zeros(2, self.dtype)
        self._wfm_stage2()

    def _wfm_stage2(self):
        """(Re)build everything after the discriminator; a WFM <-> WFM2 switch restarts this stage and the AGC, the
        video stage keeps running."""
        P = self.P
        self.agc.reset()
        self.wfm_res = decimator(P.SRATE, P.UP, P.DOWN, P.FILT_LEN, VIDEO_BWs, P.VIDEO_BW, self.dtype)
        self.wfm_deemph = None
        self.wfm_stereo = self._mode() == 'WFM2'
        self._wfm_res_key = None
        if self.wfm_stereo:
            # OPEN CHOICE (no in-tree stereo decoder): three resamplers on the real multiplex — LO 0 (L+R), 38 kHz (L-R),
            # 19 kHz (pilot) — equal-delay AF stage ('Max' delta for the audio rows, 500 Hz low-pass for the pilot),
            # feed-forward carrier (pilot/|pilot|)^2, L = S + D, R = S - D, one block AGC on max(|L|,|R|).
            self.wfm_res_lo = [None, signal_generator(38e3, P.IN_CHUNK_SIZE, P.SRATE, True),
                               signal_generator(19e3, P.IN_CHUNK_SIZE, P.SRATE, True)]
            self.wfm_res3 = [self.wfm_res] + [decimator(P.SRATE, P.UP, P.DOWN, P.FILT_LEN, VIDEO_BWs, P.VIDEO_BW, self.dtype)
                                              for _ in range(2)]
            self.wfm_af3 = [demodulator(P.FS_OUT, P.FILT_LEN, AF_BWs, self.dtype, exact=not self.fast) for _ in range(3)]
            self.wfm_deemph = None

    def _demod_wfm(self, x):
        P = self.P
        if not hasattr(self, 'wfm_vid'):
            self._wfm_setup()
        elif self.wfm_stereo != (self._mode() == 'WFM2'):
            self._wfm_stage2()
        self.wfm_vid.h = self.demod.wfm_video.h
        y = self.wfm_vid.resamp(x, self.lo)                          # video-filtered baseband at SRATE
        yy = np.concatenate((self.wfm_prev2, y))
        d = yy[2:] - yy[:-2]
        y1 = yy[1:-1]
        fm = y1.real * d.imag - y1.imag * d.real                     # sigs/nfm.m:124-126, one sample of latency
        self.wfm_prev2 = yy[len(yy) - 2:]
        af_bw = per_rx(getattr(P, 'AF_BW', 0), self.irx) or 15e3
        key = float(af_bw)
        if getattr(self, '_wfm_res_key', None) != key:
            self.wfm_res.h = design_lowpass(P.FILT_LEN, af_bw, P.SRATE * P.UP, gain=P.UP)
            if self.wfm_stereo:
                self.wfm_res3[1].h = self.wfm_res3[2].h = self.wfm_res.h
            self._wfm_res_key = key
        tau = getattr(P, 'DEEMPH_US', 0) * 1e-6
        if self.wfm_stereo:
            mpx = fm.astype(self.dtype)
            z = [self.wfm_res3[r].resamp(mpx, self.wfm_res_lo[r]) for r in range(3)]
            idx = [0, 0, AF_BWs.index('500 Hz')]
            f = [self.wfm_af3[r].demod(z[r], 'IQ', idx[r], 0) for r in range(3)]
            self.iq = np.asarray(z[0], np.complex64)
            m2 = f[2].real ** 2 + f[2].imag ** 2
            pmin = float(getattr(P, 'WFM_PILOT_MIN', 0.0))
            ok = (m2 > pmin * pmin) & (m2 > 0)
            u2 = np.where(ok, f[2] * f[2] / np.where(ok, m2, 1.0), 0.0)
            D = 2.0 * (f[1] * np.conj(u2)).real
            S = f[0].real
            L, R = S + D, S - D
            if len(L):
                g = self.agc.update(max(np.max(np.abs(L)), np.max(np.abs(R))))
                L, R = L * g, R * g
            if tau > 0:
                if self.wfm_deemph is None:
                    al = 1.0 - math.exp(-1.0 / (P.FS_OUT * tau))
                    self.wfm_deemph = [iir_stream([al], [1, al - 1]) for _ in range(2)]
                L = self.wfm_deemph[0].run(np.asarray(L, np.float32))
                R = self.wfm_deemph[1].run(np.asarray(R, np.float32))
            self.am = (np.asarray(L, np.float32) + 1j * np.asarray(R, np.float32)).astype(np.complex64)
            return self.am
        z = self.wfm_res.resamp(fm.astype(self.dtype), None)
        self.iq = np.asarray(z, np.complex64)
        a = self.agc.run(z.real)
        if tau > 0:                                                  # OPEN CHOICE: de-emphasis after the block AGC
            if self.wfm_deemph is None:
                al = 1.0 - math.exp(-1.0 / (P.FS_OUT * tau))
                self.wfm_deemph = iir_stream([al], [1, al - 1])
            a = self.wfm_deemph.run(np.asarray(a, np.float32))       # the AGC output is float32 at the API surface
        self.am = np.asarray(a, np.float32)
        return self.am

    def demod_data(self, x):
        P = self.P
        x = np.asarray(x)
        if self._mode() in ('WFM', 'WFM2'):
            return self._demod_wfm(x)
        iq = self.dec.resamp_fast(x, self.lo) if self.fast else self.dec.resamp(x, self.lo)
        mode = self._mode()
        a = self.demod.demod(iq, mode, _af_index(P, self.irx), per_rx(getattr(P, 'BFO', 0), self.irx))
        if mode not in ('IQ', 'RTTY'):
            a = self.agc.run(a)
            self.am = np.asarray(a, np.float32)
        else:
            self.am = np.asarray(a, np.complex64)
        self.iq = np.asarray(iq, np.complex64)
        return self.am

    # OPEN CHOICE: auto-mute = mean |x|^2 of the raw chunk above AUTO_MUTE_THRESH holds the mute for
    # MUTE_CHUNKS calls (only MUTE_TIME/MUTE_CHUNKS are in-tree, reference params.py:447-450).
    AUTO_MUTE_THRESH = 0.25

    def auto_mute(self, x):
        x = np.asarray(x)
        pw = float(np.mean(x.real.astype(np.float64) ** 2 + x.imag.astype(np.float64) ** 2))
        if pw > self.AUTO_MUTE_THRESH:
            self.mute_cnt = int(self.P.MUTE_CHUNKS)
        elif self.mute_cnt > 0:
            self.mute_cnt -= 1
        return self.mute_cnt > 0


# ----------------------------------------------------------------------------------------------
# a13: IIR with carried state (reference sigs/iir.py:83-105) + the squelch detector sketched in
# reference sigs/squelch.m:92-145.
# ----------------------------------------------------------------------------------------------
def iir_designs():
    """The four designs exercised by reference sigs/iir.py (:45-46, :57, :134-135, :198)."""
    return {
        'notch50': signal.iirnotch(50.0, 20.0, 1000),
        'cheby2_band': signal.iirfilter(15, [50, 200], rp=3, rs=100, btype='band', analog=False,
                                        ftype='cheby2', fs=1000),
        'ellip7_lp': signal.iirfilter(7, 200, rp=3, rs=100, btype='low', analog=False,
                                      ftype='ellip', fs=8000),
        'butter3': signal.butter(3, 0.05),
    }


class iir_stream:
    """lfilter(b,a,x,zi=z) with z carried between chunks (reference sigs/iir.py:90-93)."""

    def __init__(self, b, a):
        self.b = np.atleast_1d(np.asarray(b, np.float64))
        self.a = np.atleast_1d(np.asarray(a, np.float64))
        self.reset()

    def reset(self):
        self.z = np.zeros(max(len(self.a), len(self.b)) - 1)

    def run(self, x):
        x = np.asarray(x, np.float64)
        if len(self.z) == 0:
            return x * (self.b[0] / self.a[0])
        y, self.z = signal.lfilter(self.b, self.a, x, zi=self.z)
        return y


def squelch_designs(fs):
    """reference sigs/squelch.m:103-105: ellip(5,5,40,3000/(fs/2)) low and ellip(5,5,40,4000/(fs/2),'high')."""
    B1, A1 = signal.ellip(5, 5, 40, 3000 / (fs / 2.0))
    B2, A2 = signal.ellip(5, 5, 40, 4000 / (fs / 2.0), 'high')
    return (B1, A1), (B2, A2)


class squelch:
    """Noise squelch: in-band / out-of-band envelope ratio (reference sigs/squelch.m:121-145),
    envelopes smoothed by filter(alpha,[1 alpha-1],|z|), alpha=.001 (:125-128).
    OPEN CHOICE: the open/closed decision is ratio > thresh."""

    def __init__(self, fs, alpha=0.001, thresh=2.0):
        (B1, A1), (B2, A2) = squelch_designs(fs)
        self.f1, self.f2 = iir_stream(B1, A1), iir_stream(B2, A2)
        self.s1, self.s2 = iir_stream([alpha], [1, alpha - 1]), iir_stream([alpha], [1, alpha - 1])
        self.thresh = thresh

    def run(self, y):
        sq1 = self.s1.run(np.abs(self.f1.run(y)))
        sq2 = self.s2.run(np.abs(self.f2.run(y)))
        ratio = sq1 / np.maximum(sq2, 1e-30)
        return ratio, ratio > self.thresh


# ----------------------------------------------------------------------------------------------
# a11: spectrum (PSD).  Pinned idiom: fftshift(fft(x*w, NFFT)), 10*log10(re^2+im^2)
# (reference rtty.py:839-841).   OPEN CHOICES: periodic Hann window (BASELINE.json says Hann),
# |X|^2 / sum(w^2) scaling, averaging = arithmetic mean of |X|^2 over segments, floor 1e-30.
# ----------------------------------------------------------------------------------------------
PSD_FLOOR = 1.0e-30


class spectrum:
    def __init__(self, fs, chunk_size, NFFT, overlap, TAG=''):
        self.fs = fs
        self.chunk_size = int(chunk_size)
        self.NFFT = int(NFFT)
        self.overlap = overlap
        self.TAG = TAG
        self.new_samps = int(self.chunk_size * (1 - overlap))
        self.win = signal.get_window('hann', self.chunk_size, fftbins=True).astype(np.float32)
        self.wsum2 = float(np.sum(self.win.astype(np.float64) ** 2))
        self.df = fs / float(self.NFFT)
        self.frq = (np.arange(self.NFFT) - self.NFFT // 2) * self.df       # fftshift(fftfreq)
        self.frq2 = self.frq
        self.buf = np.zeros(self.chunk_size, np.complex128)

    def _pwr(self, seg):
        X = np.fft.fft(seg * self.win.astype(np.float64), self.NFFT)
        return (np.square(X.real) + np.square(X.imag)) / self.wsum2

    def _db(self, p, dB):
        p = np.fft.fftshift(p)
        return 10 * np.log10(np.maximum(p, PSD_FLOOR)) if dB else p

    def periodogram(self, y, dB=True):
        """One frame: the newest len(y) samples are shifted into the chunk_size window."""
        y = np.asarray(y)
        if len(y) == 0 or len(y) > self.chunk_size:
            return []                                   # failure signalled by an empty sequence (Plotting.py:463)
        self.buf = np.concatenate((self.buf[len(y):], y.astype(np.complex128)))
        return self._db(self._pwr(self.buf), dB)

    def frames(self, x):
        """|X|^2 (unshifted) of every full frame of x: start k*new_samps, length chunk_size."""
        x = np.asarray(x).astype(np.complex128)
        nfr = 0 if len(x) < self.chunk_size else 1 + (len(x) - self.chunk_size) // self.new_samps
        out = np.empty((nfr, self.NFFT))
        for k in range(nfr):
            out[k] = self._pwr(x[k * self.new_samps:k * self.new_samps + self.chunk_size])
        return out

    def psd_est(self, x, dB=True):
        """Welch average over all full frames (reference sigs/iq.py:75-79 usage)."""
        fr = self.frames(x)
        if len(fr) == 0:
            return []
        return self._db(fr.mean(axis=0), dB)

    def waterfall(self, x, navg, dB=True):
        """Lines of navg consecutive frames each -> array (nlines, NFFT), fftshifted."""
        fr = self.frames(x)
        nl = len(fr) // navg
        fr = fr[:nl * navg].reshape(nl, navg, self.NFFT).mean(axis=1)
        p = np.fft.fftshift(fr, axes=1)
        return 10 * np.log10(np.maximum(p, PSD_FLOOR)) if dB else p


# ----------------------------------------------------------------------------------------------
# a12: three_box_plot compute part, restated 1:1 from reference Plotting.py (:385-388 init,
# :536-548 line insert, :583-587 background, :594 peaks, :618-626 clip, :689-695 roll).
# ----------------------------------------------------------------------------------------------
class waterfall_state:
    def __init__(self, nfft, df, ncols=100, pan_dr=60.0, peak_dist=10e3, rig_if=0):
        self.wf = -1e38 * np.ones((nfft, ncols))
        self.wf_cnt = 0
        self.wf_fc = 0
        self.line = -1e38 * np.ones((nfft, 1))
        self.df = df
        self.pan_dr = pan_dr
        self.peak_dist = peak_dist
        self.rig_if = rig_if

    def shift_waterfall(self, frq):
        nbins = int(float(frq - self.wf_fc) / self.df + 0.5)
        if nbins != 0:
            self.wf = np.roll(self.wf, -nbins, axis=0)
            self.wf_fc = frq
        return nbins

    def push(self, PSD, fc=0):
        self.shift_waterfall(fc)
        npsd = len(PSD)
        if self.rig_if < 0:
            PSD = np.flipud(PSD)
        self.line[0:npsd, 0] = PSD
        self.line[npsd:, 0] = -1e38
        self.wf = np.concatenate((self.wf[:, 1:self.wf.shape[1]], self.line), axis=1)
        if self.wf_cnt < self.wf.shape[1]:
            self.wf_cnt += 1
        PSD2 = np.mean(self.wf[:, -self.wf_cnt:], 1)
        bkgnd = np.median(PSD2)
        dist = self.peak_dist / self.df
        peaks, _ = signal.find_peaks(PSD2, distance=dist, height=bkgnd + 10)
        zz = self.wf[0:npsd, :] - bkgnd
        zmax = np.nanmax(zz)
        image = np.maximum(zz, zmax - self.pan_dr)
        return image, bkgnd, peaks
