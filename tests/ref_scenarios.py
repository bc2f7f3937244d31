"""Replay scenarios shared by the golden generator (tests/golden/make_golden_ref_callers.py, which runs the REFERENCE's
own receiver.py / params.py code in the build container) and by the tests that replay the same scenarios through the
restated oracle loop (CPU) and through the B200 executive (GPU).

A scenario = a pySDR command line + a synthetic capture (regenerated from a seed, never stored) + a list of control
events applied at the END of loop iteration k (1-based; the place where the reference's Run() saves the raw chunk,
receiver.py:755-758, i.e. after audio_out and before the RX_DONE update):

    ('mode', 'FM')            service_commands 'setMode'        receiver.py:356-359  (P.NEW_MODE / P.MODE_CHANGE)
    ('freq', irx, f_hz)       rx.lo.change_freq(f)              receiver.py:112,352, gui.py:1938
    ('mute', irx, flag)       P.MUTED[irx]                      receiver.py:168,197
    ('af_gain', v)            P.AF_GAIN slider                  receiver.py:173,200
    ('af_filter', bw, idx)    'setAudioFilter'                  receiver.py:374-375
    ('video_filter', idx)     rx[0].dec.h = filter_bank[idx]    receiver.py:371
"""
import numpy as np

from tests.util import golden_input

SCENARIOS = {
    # two AM receivers, own players, AF-PSD / demod-file / baseband-file taps on; 5 whole chunks + a ragged tail
    'am2': dict(argv=['-replay', 'baseband_iq_am2.dat', '-fs', '2.048', '-fc', '1000', '1030', '-mode', 'AM',
                      '-foffset', '100', '-af_bw', '5'],
                n_chunks=5, tail=1234, seed=11, events={}),
    # three USB receivers on AUDIO_SCHEME 2 (two receivers per player, odd one out), one muted, slider moved
    'usb3_scheme2': dict(argv=['-replay', 'baseband_iq_usb3.dat', '-fs', '2.048', '-fc', '1000', '1020', '980',
                               '-mode', 'USB', '-foffset', '100', '-af_bw', '2', '-audio', '2'],
                         n_chunks=4, tail=0, seed=12, events={2: [('mute', 2, True), ('af_gain', 0.8)]}),
    # CW -> NFM mode change requested as 'FM' (receiver.py:640-641), retune, AF and video filter swaps, duration limit
    'cw_events': dict(argv=['-replay', 'baseband_iq_cw.dat', '-fs', '2.048', '-fc', '1000', '1015', '-mode', 'CW',
                            '-foffset', '100', '-af_bw', '0.5', '-t', '0.14'],
                      n_chunks=9, tail=77, seed=13,
                      events={1: [('freq', 1, 116500.0)], 2: [('af_filter', 1000.0, 4)], 3: [('mode', 'FM')],
                              4: [('video_filter', 3)], 5: [('mute', 0, True)]}),
    # auto-mute on a loud burst (chunk 2 scaled x40): held for MUTE_CHUNKS chunks
    'nfm_automute': dict(argv=['-replay', 'baseband_iq_nfm.dat', '-fs', '2.048', '-fc', '1000', '-mode', 'NFM',
                               '-foffset', '100', '-af_bw', '10', '-auto_mute'],
                         n_chunks=15, tail=5, seed=14, events={}, burst=(2, 40.0)),
}

# command lines whose derived parameter bag is pinned (reference params.py:199-472 + utils.py:277-289)
PARAM_CASES = [
    ['-replay', 'x.dat', '-fs', '8', '-fc', '-500', '700', '1400', '3100', '-mode', 'AM', '-foffset', '100', '-af_bw', '5'],
    ['-replay', 'x.dat', '-fs', '2.048', '-fc', '1000', '-mode', 'USB', '-foffset', '100'],
    ['-replay', 'x.dat', '-fs', '10', '-fc', '7030', '7040', '-mode', 'CW'],
    ['-replay', 'x.dat', '-fs', '2', '-fc', '94100', '-mode', 'WFM', '-foffset', '0'],
    ['-replay', 'x.dat', '-fs', '1', '-fc', '600', '-mode', 'AM', '-fsout', '24', '-vid_bw', '20', '-af_bw', '3'],
    ['-replay', 'x.dat', '-fs', '6', '-fc', '14074', '14080', '14100', '-mode', 'USB', '-foffset', '250', '-bfo', '600'],
    ['-replay', 'x.dat', '-fs', '0.5', '-fc', '3573', '-mode', 'LSB', '-foffset', '-33', '-fsout', '12'],
    ['-replay', 'x.dat', '-fs', '9', '-fc', '1', '2', '3', '4', '5', '6', '7', '8', '-mode', 'NFM', '-foffset', '123.456'],
    ['-replay', 'x.dat', '-fs', '4', '-fc', '10000', '-mode', 'IQ', '-fsout', '96', '-foffset', '77'],
    ['-replay', 'x.dat', '-fs', '3', '-fc', '10000', '-mode', 'RTTY', '-fsout', '192', '-audio', '2', '-delay', '3'],
]
PARAM_FIELDS = ['SRATE', 'UP', 'DOWN', 'FS_OUT', 'IN_CHUNK_SIZE', 'OUT_CHUNK_SIZE', 'RB_SIZE', 'DELAY', 'FOFFSET', 'BFO',
                'VIDEO_BW', 'AF_BW', 'MUTE_CHUNKS', 'NUM_RX', 'NUM_PLAYERS', 'FILT_LEN', 'DURATION', 'AF_GAIN', 'AUDIO_SCHEME',
                'PAN_DR', 'PEAK_DIST', 'MODE', 'ENABLE_AUTO_MUTE', 'REPLAY_MODE', 'SDR_TYPE']


def scenario_offsets(P):
    """Receiver offsets the reference computes (receiver.py:826-835), from any P with FOFFSET/FC/SOURCE."""
    out = []
    for irx in range(P.NUM_RX):
        if P.SOURCE[irx] >= 0:
            out.append(float(P.FC[irx] - P.FC[P.SOURCE[irx]]))
        else:
            out.append(float(P.FOFFSET + P.FC[irx] - P.FC[0]))
    return out


def scenario_input(sc, P):
    """complex64 capture of a scenario: n_chunks * IN_CHUNK_SIZE + tail samples (LCG noise + one carrier per receiver)."""
    C = int(P.IN_CHUNK_SIZE)
    n = sc['n_chunks'] * C + sc['tail']
    x = golden_input(n, P.SRATE, scenario_offsets(P), sc['seed'])
    if 'burst' in sc:
        k, s = sc['burst']
        x[k * C:(k + 1) * C] *= np.float32(s)
    return x


class Recorder(object):
    """save_data sink standing in for fileio.sdr_fileio writers (receiver.py:293-296,757)."""

    def __init__(self, hook=None):
        self.saved, self.hook = [], hook

    def save_data(self, x, VERBOSITY=0):
        self.saved.append(np.array(x, copy=True))
        if self.hook:
            self.hook(len(self.saved))

    def close(self):
        pass


class Player(object):
    """audio_io.AudioIO stand-in: never active, keeps the ring buffer it is given (receiver.py:840-849)."""

    def __init__(self, P, fs, rb, device, ch='B', Tag=''):
        self.rb, self.fs, self.active, self.Start_Time, self.starts = rb, fs, False, 0, 0

    def start_playback(self, delay, flag):
        self.starts += 1
        return False

    def stop(self):
        pass


class PushLog(object):
    """Ring-buffer stand-in that keeps every pushed block."""

    def __init__(self, tag='', size=0):
        self.tag, self.size, self.pushed, self.nsamps = tag, size, [], 0

    def push(self, x):
        self.pushed.append(np.array(x, copy=True))
        self.nsamps += len(x)


class ReplayFile(object):
    """fileio.sdr_fileio stand-in for the replay source (receiver.py:810-813, :531): .srate .fc .read_data()."""

    def __init__(self, raw, srate, fc):
        self.raw, self.srate, self.fc = raw, srate, fc

    def read_data(self):
        return self.raw


def apply_event(P, ev):
    kind = ev[0]
    if kind == 'mode':                                   # what service_commands 'setMode' does (receiver.py:356-359)
        P.NEW_MODE = ev[1]
        P.MODE_CHANGE = (P.MODE != ev[1])
    elif kind == 'freq':
        P.rx[ev[1]].lo.change_freq(ev[2])
    elif kind == 'mute':
        P.MUTED[ev[1]] = ev[2]
    elif kind == 'af_gain':
        P.AF_GAIN = ev[1]
    elif kind == 'af_filter':
        P.AF_BW, P.AF_FILTER_NUM = ev[1], ev[2]
    elif kind == 'video_filter':
        P.rx[0].dec.h = P.rx[0].dec.filter_bank[ev[1]]
    else:
        raise ValueError(ev)
