"""RUN_TIME_PARAMS — the parameter bag 'P' of the receive path (reference params.py:38-486), restricted to
the flags that reach the hot path and keeping their names, units (kHz/MHz on the CLI -> Hz in P) and
derivation order.  Hardware discovery, rig control, GUI geometry, hopping, UDP ... are out of scope.

Extension used by the 4-RX benchmark config: ``-mode``, ``-af_bw`` and ``-bfo`` accept one value per
receiver (the reference keeps one global value; in MP_SCHEME 3 every RX process owns its own P).
"""
import argparse

import numpy as np

from . import design
from .design import MAX_RX, MODES, RTLsrates, SDRplaysrates


class RUN_TIME_PARAMS:
    def __init__(self, argv=None, **overrides):
        ap = argparse.ArgumentParser(description='pysdr_b200 receive path')
        ap.add_argument('-fc', help='RF centre frequencies (KHz), one per receiver', type=float, default=[1000.], nargs='*')
        ap.add_argument('-mode', help='Demod mode(s)', type=str, default=['AM'], nargs='*', choices=MODES)
        ap.add_argument('-fs', help='RF sampling rate (MHz)', type=float, default=0)
        ap.add_argument('-fsout', help='Audio sampling rate (KHz)', type=float, default=48)
        ap.add_argument('-foffset', help='Tuning offset (KHz)', type=float, default=100)     # params.py:80-81
        ap.add_argument('-vid_bw', help='Video bandwidth (KHz)', type=float, default=0)
        ap.add_argument('-af_bw', help='Audio bandwidth(s) (KHz)', type=float, default=[0], nargs='*')
        ap.add_argument('-nfilt', help='Decimation filter length', type=int, default=1001)
        ap.add_argument('-bfo', help='BFO (Hz)', type=float, default=[0], nargs='*')
        ap.add_argument('-t', help='Duration (s)', type=float, default=1e38)
        ap.add_argument('-replay', help='Replay file [tskip]', type=str, default=None, nargs='*')
        ap.add_argument('-rtl', action='store_true', help='RTL rate table')
        ap.add_argument('-auto_mute', action='store_true')
        ap.add_argument('-pan_dr', help='Waterfall dynamic range (dB)', type=float, default=60)
        ap.add_argument('-pan_bw', help='Pan bandwidth (KHz)', type=float, default=0)
        ap.add_argument('-src', type=int, default=[-1], nargs='*')
        ap.add_argument('-audio', help='Audio scheme for routing RXs', type=int, default=1)   # params.py:71-72
        ap.add_argument('-delay', help='Audio buffer delay', type=int, default=16)            # params.py:73-74
        ap.add_argument('-mute', action='store_true')
        args = ap.parse_args(argv if argv is not None else [])
        for k, v in overrides.items():
            setattr(args, k, v)

        self.MP_SCHEME = 1                                    # params.py:200
        self.threads = []
        self.AF_FILTER_NUM = None                             # params.py:202
        self.VIDEO_FILTER_NUM = None
        self.audio_playback = False                           # params.py:203
        self.REPLAY_MODE = bool(args.replay)
        self.SDR_TYPE = 'replay' if self.REPLAY_MODE else ('rtlsdr' if args.rtl else 'sdrplay')   # utils.py:462-471
        self.REPLAY = args.replay[0] if args.replay else None

        fs = args.fs                                          # params.py:218-236: snap to the device's rate table
        if self.SDR_TYPE == 'rtlsdr':
            if fs == 0:
                fs = 2
            self.SRATE = 1e6 * RTLsrates[int(np.argmin(np.abs(np.array(RTLsrates) - fs)))]
        else:
            if fs == 0:
                fs = 1
            self.SRATE = 1e6 * SDRplaysrates[int(np.argmin(np.abs(np.array(SDRplaysrates) - fs)))]
        if overrides.get('srate_hz'):                         # replay files carry their own rate (receiver.py:811)
            self.SRATE = float(overrides['srate_hz'])

        fc = np.atleast_1d(np.array(args.fc, float)) * 1e3    # params.py:246-250
        self.NUM_RX = len(fc)
        self.MAX_RX = MAX_RX
        if self.NUM_RX > MAX_RX:                              # params.py:270-276
            fc = fc[0:MAX_RX]
            self.NUM_RX = len(fc)
        self.FC = fc
        self.VFO = MAX_RX * ['A']
        mode = list(args.mode) if isinstance(args.mode, (list, tuple)) else [args.mode]
        self.MODE = mode[0] if len(mode) == 1 else (mode + [mode[-1]] * self.NUM_RX)[:self.NUM_RX]
        self.FOFFSET = args.foffset * 1e3
        self.AUDIO_SCHEME = args.audio                        # params.py:287
        if self.AUDIO_SCHEME == 1:                            # params.py:296-303
            self.NUM_PLAYERS = int(self.NUM_RX)
        elif self.AUDIO_SCHEME == 2:
            self.NUM_PLAYERS = int((self.NUM_RX + 1) / 2)
        else:
            raise SystemExit('ERROR - Invalid audio playback scheme')
        self.LOOPBACK = False
        self.AUX_AUDIO = False
        src = np.atleast_1d(np.array(args.src) * 1)           # params.py:289-294
        while len(src) < self.NUM_RX:
            src = np.append(src, [-1])
        self.SOURCE = src
        self.rx = self.NUM_RX * [None]
        if self.FOFFSET == 0:                                 # params.py:309-314
            fo = 0.5 * (max(fc) + min(fc))
            self.FOFFSET = fo - max(fc)
        bfo = list(args.bfo) if isinstance(args.bfo, (list, tuple)) else [args.bfo]
        if isinstance(self.MODE, list):
            bfo = (bfo + [bfo[-1]] * self.NUM_RX)[:self.NUM_RX]
            self.BFO = [700 if (m == 'CW' and b == 0) else b for m, b in zip(self.MODE, bfo)]
        else:
            self.BFO = bfo[0]
            if self.MODE == 'CW' and self.BFO == 0:           # params.py:316-318
                self.BFO = 700
        self.DURATION = args.t
        self.VIDEO_BW = args.vid_bw * 1e3
        if self.VIDEO_BW == 0:                                # params.py:322-327
            self.VIDEO_BW = 200e3 if design.per_rx(self.MODE, 0) == 'WFM' else 10e3
        self.PAN_BW = args.pan_bw * 1e3
        self.PAN_DR = args.pan_dr
        self.FS_OUT = args.fsout * 1e3
        self.FILT_LEN = args.nfilt                            # params.py:345
        afbw = list(args.af_bw) if isinstance(args.af_bw, (list, tuple)) else [args.af_bw]
        afbw = [a * 1e3 for a in afbw]
        self.AF_BW = afbw[0] if len(afbw) == 1 else (afbw + [afbw[-1]] * self.NUM_RX)[:self.NUM_RX]
        self.PEAK_DIST = 10e3
        self.RIG_IF = 0
        if self.FS_OUT < 1 or self.FS_OUT > 192e3:            # params.py:400-404
            raise SystemExit('*** ERROR in RUN_TIME_PARAMS - Invalid output sampling rate ***')
        self.UP, self.DOWN = design.up_dn(self.SRATE, self.FS_OUT)          # params.py:405
        self.FS_OUT = int(self.SRATE * self.UP / self.DOWN)                 # params.py:406
        self.SHOW_RF_PSD = False
        self.SHOW_BASEBAND_PSD = False
        self.SHOW_AF_PSD = False
        self.PANADAPTOR = False
        self.PLOT_RX = 0
        self.SAVE_IQ = self.SAVE_BASEBAND = self.SAVE_DEMOD = False
        self.MODE_CHANGE = False
        self.NEW_MODE = self.MODE
        self.FREQ_CHANGE = False
        self.AF_GAIN = 0.5                                    # params.py:425
        self.MUTED = MAX_RX * [args.mute]                     # params.py:431
        self.OUT_CHUNK_SIZE = 1024                            # params.py:440
        self.IN_CHUNK_SIZE = int(self.OUT_CHUNK_SIZE * self.DOWN / float(self.UP) + 0 * 0.5)   # params.py:444
        self.ENABLE_AUTO_MUTE = args.auto_mute
        self.MUTE_TIME = .25
        self.MUTE_CHUNKS = int(self.MUTE_TIME * self.FS_OUT / self.OUT_CHUNK_SIZE)             # params.py:449
        self.AUTO_MUTED = False
        self.RB_SIZE = design.rb_size(self.NUM_RX, self.FS_OUT, self.SDR_TYPE, self.OUT_CHUNK_SIZE)  # :456-468
        self.DELAY = min(32, max(1, args.delay)) * self.OUT_CHUNK_SIZE                         # params.py:457
        design.adjust_foffset(self)                           # params.py:472
        self.SHUT_DOWN = False
        self.raw_iq_io = self.baseband_iq_io = self.demod_io = None
        self.gui = None
        self.RX_DONE = False
        self.nchunks = 0
        self.Stopper = None
        self.evt = None
        self.players = []
