"""CPU oracle for the L2 executive semantics of pySDR's receiver.py (TEST INFRASTRUCTURE ONLY).

Restates, over the numpy oracle operators in ``sig_proc_oracle``:

  * the parameter bag derivations of reference ``params.py:199-472`` (``make_P``),
  * ``demodulate_data``   reference ``receiver.py:231-297``  (per-chunk DC removal for AM/USB),
  * ``audio_out`` gain    reference ``receiver.py:192-225``  (af_gain = 10**AF_GAIN - 1, mute -> 0),
  * the replay chunk loop reference ``receiver.py:538-559, 684-782`` including its quirks: strict ``<``
    termination and the stale last chunk being demodulated once more when EOF is hit.

Parity status: the loop/parameter arithmetic is pinned by the reference tree itself; the operators
underneath are "parity unpinned" (see sig_proc_oracle.py header).
"""
from __future__ import annotations

import numpy as np

from . import sig_proc_oracle as dsp


class Params:
    """Attribute bag 'P' (reference params.py RUN_TIME_PARAMS)."""
    pass


def make_P(srate, fc_hz, mode, fs_out=48e3, foffset=0.0, vid_bw=0.0, af_bw=0.0, nfilt=1001, bfo=0,
           sdr_type='sdrplay', duration=1e38, af_filter_num=None, video_filter_num=None):
    """fc_hz: list of receiver centre frequencies in Hz. mode/af_bw/bfo: scalar or per-RX list."""
    P = Params()
    fc = np.atleast_1d(np.asarray(fc_hz, float))
    P.SDR_TYPE = sdr_type
    P.SRATE = float(srate)
    P.NUM_RX = len(fc)
    if P.NUM_RX > dsp.MAX_RX:                                  # params.py:270-276
        fc = fc[:dsp.MAX_RX]
        P.NUM_RX = len(fc)
    P.FC = fc
    P.MODE = mode
    P.FOFFSET = float(foffset)
    P.SOURCE = np.array([-1] * P.NUM_RX)                       # params.py:289-294
    P.rx = P.NUM_RX * [None]
    if P.FOFFSET == 0:                                         # params.py:309-314
        fo = 0.5 * (max(fc) + min(fc))
        P.FOFFSET = fo - max(fc)
    P.BFO = bfo
    first_mode = dsp.per_rx(mode, 0)
    if not isinstance(mode, (list, tuple)):
        if mode == 'CW' and bfo == 0:                          # params.py:316-318
            P.BFO = 700
    else:
        b = list(bfo) if isinstance(bfo, (list, tuple)) else [bfo] * P.NUM_RX
        P.BFO = [700 if (m == 'CW' and bb == 0) else bb for m, bb in zip(mode, b)]
    P.DURATION = duration
    P.VIDEO_BW = float(vid_bw)
    if P.VIDEO_BW == 0:                                        # params.py:322-327
        P.VIDEO_BW = 200e3 if first_mode == 'WFM' else 10e3
    P.FILT_LEN = int(nfilt)
    P.AF_BW = af_bw
    P.AF_FILTER_NUM = af_filter_num
    P.VIDEO_FILTER_NUM = video_filter_num
    P.FS_OUT = float(fs_out)
    P.UP, P.DOWN = dsp.up_dn(P.SRATE, P.FS_OUT)                # params.py:405
    P.FS_OUT = int(P.SRATE * P.UP / P.DOWN)                    # params.py:406
    P.AF_GAIN = 0.5                                            # params.py:425
    P.MUTED = dsp.MAX_RX * [False]
    P.AUTO_MUTED = False
    P.OUT_CHUNK_SIZE = 1024                                    # params.py:440
    P.IN_CHUNK_SIZE = int(P.OUT_CHUNK_SIZE * P.DOWN / float(P.UP) + 0 * 0.5)   # params.py:444
    P.ENABLE_AUTO_MUTE = False
    P.MUTE_TIME = .25
    P.MUTE_CHUNKS = int(P.MUTE_TIME * P.FS_OUT / P.OUT_CHUNK_SIZE)             # params.py:449
    P.RB_SIZE = dsp.rb_size(P.NUM_RX, P.FS_OUT, sdr_type, P.OUT_CHUNK_SIZE)    # params.py:456-468
    P.FOFFSET = dsp.adjust_foffset(P.FOFFSET, P.SRATE, P.RB_SIZE)              # params.py:472
    P.MP_SCHEME = 1
    P.audio_playback = False                                   # params.py:203
    P.SHOW_AF_PSD = False
    P.SHOW_BASEBAND_PSD = False
    P.PANADAPTOR = False
    P.PLOT_RX = 0
    P.RX_DONE = False
    P.nchunks = 0
    return P


def create_receivers(P, dtype=np.complex128, fast=False):
    """reference receiver.py:826-835."""
    foff = P.FOFFSET
    for irx in range(P.NUM_RX):
        if P.SOURCE[irx] >= 0:
            frq = P.FC[irx] - P.FC[P.SOURCE[irx]]
        else:
            frq = foff + P.FC[irx] - P.FC[0]
        P.rx[irx] = dsp.Receiver(P, frq, irx, str(irx + 1), dsp.VIDEO_BWs, dsp.AF_BWs, dtype=dtype, fast=fast)
    return P.rx


def demodulate_data(P, x, irx):
    """reference receiver.py:231-252: returns (am_for_psd_and_file, rx) — rx.am (audio) is NOT DC-removed."""
    rx = P.rx[irx]
    am = rx.demod_data(x)
    if P.ENABLE_AUTO_MUTE:
        P.AUTO_MUTED = rx.auto_mute(x)
    mode = dsp.per_rx(P.MODE, irx)
    if mode == 'AM' or mode == 'USB':
        am = am - np.mean(am)                                   # receiver.py:250-252
    return am


def af_gain(P, irx=0):
    """reference receiver.py:197-200."""
    if P.MUTED[irx] or P.AUTO_MUTED:
        return 0.
    return pow(10., P.AF_GAIN) - 1


def audio_payloads(P):
    """What reference receiver.py:153-225 pushes to each player for the chunk just demodulated: AUDIO_SCHEME 2 packs
    receiver i and its partner i + (NUM_RX+1)//2 as ``am1*g1 + 1j*am2*g2`` (:158-188, mute only — auto-mute is not
    looked at in this branch); otherwise one player per receiver gets ``am*gain`` with gain 0 when muted or auto-muted."""
    g = pow(10., P.AF_GAIN) - 1
    if getattr(P, 'AUDIO_SCHEME', 1) == 2:
        n2 = int((P.NUM_RX + 1) / 2)
        out = []
        for irx in range(n2):
            a1 = P.rx[irx].am.real
            g1 = 0. if P.MUTED[irx] else g
            j = irx + n2
            if j < P.NUM_RX:
                a2, g2 = P.rx[j].am.real, (0. if P.MUTED[j] else g)
            else:
                a2, g2 = 0, 0.
            out.append(a1 * g1 + 1j * a2 * g2)
        return out
    return [P.rx[irx].am * af_gain(P, irx) for irx in range(P.NUM_RX)]


def mode_freq_change(P):
    """reference receiver.py:634-650: a pending mode change takes effect between chunks; 'FM' means 'NFM'; receiver 0's
    AGC and carrier PLL restart."""
    if getattr(P, 'MODE_CHANGE', False):
        if P.NEW_MODE == 'FM':
            P.NEW_MODE = 'NFM'
        if P.MODE != P.NEW_MODE:
            P.MODE = P.NEW_MODE
            P.rx[0].agc.reset()
            P.rx[0].demod.am_pll.reset()
        P.MODE_CHANGE = False


def run_replay(P, raw, collect=('am', 'iq', 'am_dc'), on_iteration=None):
    """Replay loop, reference receiver.py:684-782 + read_chunk :538-559 (MP_SCHEME 1, no pre-mixer).
    Returns dict of per-RX lists of per-chunk arrays, and the number of loop iterations.  on_iteration(k) runs at the
    end of iteration k (where the reference saves the raw chunk, :755-758); 'audio' collects the player payloads."""
    out = {k: [[] for _ in range(P.NUM_RX)] for k in collect}
    if 'auto_muted' in out:
        out['auto_muted'] = []
    praw = 0
    x = np.zeros(P.IN_CHUNK_SIZE, np.complex64)                 # receiver.py:445
    dt = float(P.IN_CHUNK_SIZE) / P.SRATE
    t = 0.
    P.RX_DONE = False
    iters = 0
    while not P.RX_DONE:
        t += dt
        P.nchunks += 1
        iters += 1
        if praw + P.IN_CHUNK_SIZE < len(raw):                   # strict '<', receiver.py:544
            x = raw[praw:praw + P.IN_CHUNK_SIZE]
            praw += P.IN_CHUNK_SIZE
        else:
            P.RX_DONE = True                                    # ... and the stale x is processed again
        mode_freq_change(P)
        for irx in range(P.NUM_RX):
            am_dc = demodulate_data(P, x, irx)
            if 'am' in out:
                out['am'][irx].append(P.rx[irx].am.copy())
            if 'iq' in out:
                out['iq'][irx].append(P.rx[irx].iq.copy())
            if 'am_dc' in out:
                out['am_dc'][irx].append(np.asarray(am_dc))
        if 'audio' in out:
            for i, a in enumerate(audio_payloads(P)):
                out['audio'][i].append(np.asarray(a))
        if 'auto_muted' in out:
            out['auto_muted'].append(bool(P.AUTO_MUTED))
        if on_iteration is not None:
            on_iteration(iters)
        P.RX_DONE = P.RX_DONE or t >= P.DURATION
    return out, iters
