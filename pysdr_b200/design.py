"""Host-side control plane of the receive path: tables, rate arithmetic, tuning-offset quantiser and FIR
design.  Runs once per (re)configuration, never per sample — the per-sample work is in csrc/*.cu.

Reference contracts: Tables.py:34-62 (mode / bandwidth tables, find_filter), params.py:405-406,440-468
(UP/DOWN, FS_OUT, IN_CHUNK_SIZE, RB_SIZE), utils.py:277-289 (adjust_foffset), srates.py:35-74 (up_dn table).
Filter design choices (not fixed by the reference tree) are listed in DESIGN.md section 3.
"""
import math
from math import gcd

import numpy as np
from scipy import signal

MODES = ["AM", "AM-Synch", "SSB", "USB", "LSB", 'CW', "IQ", "WFM", "WFM2", "NFM", "RTTY"]      # Tables.py:34
AF_BWs = ['Max', '50 Hz', '100 Hz', '500 Hz', '1 KHz', '2 KHz', '3 KHz',
          '4 KHz', '5 KHz', '8 KHz', '10 KHz', '15 KHz', '20 KHz', '45 KHz', '50 KHz', '100 KHz', '150 KHz',
          '200 KHz']                                                                          # Tables.py:36-37
PAN_BWs = ['1 KHz', '3 KHz', '5 KHz', '10 KHz', '20 KHz', '40 KHz', '50 KHz', '100 KHz', '150 KHz', 'All']
VIDEO_BWs = ['Max', '5 KHz', '10 KHz', '20 KHz', '25 KHz', '45 KHz', '50 KHz', '100 KHz', '150 KHz', '200 KHz',
             '300 KHz', '400 KHz', '500 KHz', '750 KHz', '1 MHz', 'Other']                     # Tables.py:41-42
RTLsrates = [0.25, 1.024, 1.536, 1.792, 1.92, 2.048, 2.16, 2.56, 2.88, 3.2]                   # Tables.py:44
SDRplaysrates = [0.25, 0.5, 1, 2, 2.048, 3, 4, 5, 6, 7, 8, 9, 10]                              # Tables.py:45
MAX_RX = 6                                                                                    # params.py:33

MODE_IDS = {"AM": 0, "AM-Synch": 7, "USB": 1, "SSB": 1, "LSB": 2, "CW": 3, "IQ": 4, "RTTY": 4, "NFM": 5,
            "RAW": 6}       # RAW: internal second stage of the WFM chain (sig_proc._WfmChain)


def bw_hz(label):
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
    """Tables.py:48-62."""
    best = None
    for bw in bw_list:
        b = bw_hz(bw)
        if b is not None and b <= max_bw:
            best = bw
    return best


def up_dn(fs1, fs2):
    f1, f2 = int(round(fs1)), int(round(fs2))
    g = gcd(f1, f2)
    return f2 // g, f1 // g


def rb_size(num_rx, fs_out, sdr_type='sdrplay', out_chunk=1024):
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


def adjust_foffset(P):
    """utils.py:277-289 (mutates P.FOFFSET)."""
    M = round(P.RB_SIZE * P.FOFFSET / P.SRATE)
    P.FOFFSET = M * P.SRATE / P.RB_SIZE


def n_out_total(n_in, up, down):
    return -((-up * n_in) // down)


def lowpass(ntaps, cutoff_hz, fs_hz, gain=1.0):
    cutoff_hz = min(float(cutoff_hz), 0.45 * fs_hz)
    return (signal.firwin(int(ntaps), cutoff_hz, window='hamming', fs=float(fs_hz)) * gain).astype(np.float32)


def resampler_bank(srate, up, down, filt_len, video_bws=VIDEO_BWs, video_bw_other=10e3):
    """dec.filter_bank: one prototype per VIDEO_BWs entry, designed at SRATE*UP, DC gain UP."""
    fs_out = srate * up / down
    bank = []
    for lb in video_bws:
        if lb == 'Max':
            fc = 0.45 * min(srate, fs_out)
        else:
            bw = video_bw_other if lb == 'Other' else bw_hz(lb)
            fc = min(0.5 * bw, 0.45 * srate)
        bank.append(lowpass(filt_len, fc, srate * up, gain=up))
    return bank


def _delta(ntaps):
    h = np.zeros(ntaps, np.float32)
    h[(ntaps - 1) // 2] = 1.0
    return h


def af_bank_real(fs_out, ntaps, af_bws=AF_BWs):
    return [_delta(ntaps) if not bw_hz(lb) else lowpass(ntaps, bw_hz(lb), fs_out) for lb in af_bws]


def af_bank_lp(fs_out, ntaps, af_bws=AF_BWs):
    return [_delta(ntaps) if not bw_hz(lb) else lowpass(ntaps, bw_hz(lb) / 2, fs_out) for lb in af_bws]


def af_bank_cmpx(fs_out, ntaps, af_bws=AF_BWs):
    bank = []
    c = (ntaps - 1) / 2.0
    j = np.arange(ntaps)
    for lb in af_bws:
        bw = bw_hz(lb) or 0.9 * fs_out / 2
        bw = min(bw, 0.9 * fs_out / 2)
        h = lowpass(ntaps, bw / 2, fs_out).astype(np.float64)
        bank.append((h * np.exp(2j * np.pi * (bw / 2) * (j - c) / fs_out)).astype(np.complex64))
    return bank


def bpf(f1, f2, fs, ntaps):
    """dsp.bpf(f1,f2,fs,ntaps) (reference receiver.py:861): real band-pass FIR."""
    return signal.firwin(int(ntaps), [float(f1), float(f2)], pass_zero=False, window='hamming', fs=float(fs)).astype(np.float32)


def per_rx(v, irx):
    return v[irx] if isinstance(v, (list, tuple, np.ndarray)) else v


def af_index(P, irx=0):
    """gui.py:1720-1731 semantics for AF_FILTER_NUM None/-1: look AF_BW up, miss -> 0 ('Max')."""
    idx = per_rx(getattr(P, 'AF_FILTER_NUM', None), irx)
    if idx is None or idx < 0:
        bw = per_rx(getattr(P, 'AF_BW', 0), irx)
        idx = 0
        for i, lb in enumerate(AF_BWs):
            b = bw_hz(lb)
            if b is not None and b == bw:
                idx = i
                break
    return idx


def video_index(P, labels=VIDEO_BWs):
    """gui.py:1675-1685."""
    idx = getattr(P, 'VIDEO_FILTER_NUM', None)
    if idx is not None and idx >= 0:
        return idx
    bw = P.VIDEO_BW
    lab = (str(int(bw * 1e-6)) + ' MHz') if bw > 1e6 - 1 else (str(int(bw * 1e-3)) + ' KHz')
    return labels.index(lab) if lab in labels else len(labels) - 1


def freq_to_phase_inc(f, fs):
    r = float(f) / float(fs)
    r = r - math.floor(r)
    return int(r * 2.0 ** 64) & ((1 << 64) - 1)


def phase_inc_to_freq(inc, fs):
    inc = int(inc) & ((1 << 64) - 1)
    if inc >= (1 << 63):
        inc -= (1 << 64)
    return inc / 2.0 ** 64 * float(fs)


def wfm_video_bank(srate, filt_len, video_bws=VIDEO_BWs, video_bw_other=200e3):
    """demod.wfm_filter_bank (reference gui.py:1704): video FIRs at the RF rate, cutoff VIDEO_BW/2 ('Max'/'Other'
    use P.VIDEO_BW)."""
    return [lowpass(filt_len, min(0.5 * (bw_hz(lb) or video_bw_other), 0.45 * srate), srate) for lb in video_bws]


def wfm_resampler_taps(srate, up, filt_len, af_bw):
    """For WFM the audio filtering is done in the resampler (reference gui.py:1759-1762): low-pass at AF_BW
    (15 kHz when AF_BW is 0 / 'Max'), designed at SRATE*UP with gain UP."""
    return lowpass(filt_len, af_bw or 15e3, srate * up, gain=up)
