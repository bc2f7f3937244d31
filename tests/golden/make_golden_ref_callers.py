"""Generates tests/golden/ref_callers.npz + ref_params.json by EXECUTING the reference's own code in the build container:

    python tests/golden/make_golden_ref_callers.py

What runs is /root/reference's unmodified  params.RUN_TIME_PARAMS (params.py:38-472), utils.adjust_foffset
(utils.py:277-289), receiver.SDR_EXECUTIVE.__init__/Run/read_chunk/mode_freq_change (receiver.py:408-782),
receiver.demodulate_data (:231-297), receiver.audio_out (:153-225) and Plotting.three_box_plot.plot/shift_waterfall
(Plotting.py:444-631, 689-695), loaded through tests/golden/ref_harness.py (stub modules for Qt / Soapy / rig_io ...).
The module the reference imports as `sig_proc` is the numpy oracle (oracle/sig_proc_oracle.py): upstream's own
`sig_proc` is not obtainable, so the arithmetic UNDER the seam stays the oracle's, while everything ABOVE the seam —
chunking and its quirks, DC removal, audio routing / gains / muting, mode and filter changes, the parameter bag, the
waterfall algebra — is the reference's code, and its outputs are the fixtures.

Only numeric outputs are saved; the reference tree is read, never written or copied."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import ref_scenarios as rs                 # noqa: E402
from tests.golden import ref_harness as rh            # noqa: E402
from oracle import sig_proc_oracle as dsp             # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


Recorder, Player, ReplayFile, apply_event = rs.Recorder, rs.Player, rs.ReplayFile, rs.apply_event


def run_scenario(mods, name, sc):
    rx_mod = mods['receiver']
    P = rh.run_time_params(mods, sc['argv'])
    x = rs.scenario_input(sc, P)
    srate0, fc0 = P.SRATE, float(P.FC[0])
    holder = {}

    class file_io(object):                               # receiver.py:810 says `file_io` although :41 imports `fileio`
        @staticmethod
        def sdr_fileio(fname, mode, PP):
            holder['replay_name'] = fname
            return ReplayFile(x, srate0, fc0)

    rx_mod.file_io = file_io
    rx_mod.AudioIO = Player
    # what pySDR.py's main sets up around the executive (pySDR.py:95-130), reduced to the data path
    P.evt = None
    P.Stopper = None
    P.gui = rh._Anything()
    P.audio_playback = True
    P.SHOW_AF_PSD = True
    P.PLOT_RX = 0
    P.rb_af = rs.PushLog('AF', P.RB_SIZE)
    P.SHOW_BASEBAND_PSD = True
    P.rb_baseband = rs.PushLog('BB', P.RB_SIZE)
    P.SAVE_DEMOD = P.SAVE_BASEBAND = P.SAVE_IQ = True
    P.demod_io, P.baseband_iq_io = Recorder(), Recorder()
    auto_muted, am_all = [], []

    def end_of_iteration(k):
        auto_muted.append(bool(P.AUTO_MUTED))
        am_all.append([np.array(P.rx[i].am, copy=True) for i in range(P.NUM_RX)])
        for ev in sc['events'].get(k, []):
            apply_event(P, ev)

    P.raw_iq_io = Recorder(hook=end_of_iteration)
    ex = rx_mod.SDR_EXECUTIVE(P, False)
    ex.Run()
    it = len(P.raw_iq_io.saved)
    out = {
        'iters': it, 'nchunks': int(P.nchunks), 'mode_final': str(P.MODE), 'in_chunk': int(P.IN_CHUNK_SIZE),
        'fs_out': int(P.FS_OUT), 'up': int(P.UP), 'down': int(P.DOWN), 'n_players': len(P.players),
        'auto_muted': np.array(auto_muted), 'shut_down': bool(P.SHUT_DOWN),
        'rb_af': np.array([np.asarray(v) for v in P.rb_af.pushed]),
        'baseband_io': np.array(P.baseband_iq_io.saved),
        # the file taps see the same arrays as the PSD taps (receiver.py:254-296): stored once, equality recorded
        'demod_io_equals_rb_af': bool(np.array_equal(np.array(P.demod_io.saved), np.array(P.rb_af.pushed))),
        'rb_baseband_equals_baseband_io': bool(np.array_equal(np.array(P.rb_baseband.pushed), np.array(P.baseband_iq_io.saved))),
        'raw_first': np.array([v[0] for v in P.raw_iq_io.saved]), 'raw_last': np.array([v[-1] for v in P.raw_iq_io.saved]),
        'am': np.array(am_all),
    }
    for i, pl in enumerate(P.players):
        out['player%d' % i] = np.array([np.asarray(v) for v in pl.rb.pushed])
        out['player%d_starts' % i] = pl.starts
    print("%-14s iters %d nchunks %d mode %s players %d rb_af %s taps equal %s %s" % (
        name, it, P.nchunks, P.MODE, len(P.players), out['rb_af'].shape, out['demod_io_equals_rb_af'],
        out['rb_baseband_equals_baseband_io']))
    return out


# ---- Plotting.three_box_plot.plot on a stub widget -------------------------------------------------------------------
WF_NFFT, WF_CHUNK, WF_FS = 256, 128, 48.0
WF_FRAMES = 130
WF_KEEP = [1, 2, 57, 100, 101, 129]                      # frames whose image is stored (before / after the 100-column fill)
WF_RETUNE = {60: 1.5, 110: -2.25}                        # frame -> new fc (KHz): shift_waterfall rolls the rows


def waterfall_input(frame):
    from tests.util import lcg_iq
    n = WF_CHUNK
    t = (frame * n + np.arange(n)) / (WF_FS * 1e3)
    x = lcg_iq(n, 9000 + frame, scale=0.02).astype(np.complex128)
    x += 0.3 * np.exp(2j * np.pi * 6000.0 * t) + 0.1 * np.exp(-2j * np.pi * 11000.0 * t) * (1 + 0.5 * np.sin(2 * np.pi * 40 * t))
    return x.astype(np.complex64)


def run_waterfall(mods):
    tb = mods['Plotting'].three_box_plot
    images = {}

    class Imager(object):
        def getXRange(self):
            return [0.0, 1.0]

        def imagesc(self, z, **kw):
            self.last = np.array(z, copy=True)

    class Widget(rh._Anything):
        pass

    class PP(object):
        PAN_BW, PAN_DIR, RIG_IF, NUM_RX, MAIN_RX, FC, frqArx, frqAtx = 0, 'Up/Down', 0, 1, 0, [0.0], None, None
        AF_BW, VIDEO_BW, PAN_DR, PEAK_DIST = 0, 10e3, 60.0, 2.0       # PEAK_DIST in the units of psd.df (KHz here)

    w = Widget()
    w.P = PP()
    w.psd = dsp.spectrum(WF_FS, WF_CHUNK, WF_NFFT, 0.0)
    w.foff, w.TRANSPOSE = 0.0, False
    w.imager = Imager()
    w.wf = -1e38 * np.ones((WF_NFFT, 100))                            # Plotting.py:385-388
    w.wf_cnt, w.wf_fc = 0, 0
    w.line = -1e38 * np.ones((WF_NFFT, 1))
    w.shift_waterfall = lambda frq: tb.shift_waterfall(w, frq)
    fc = 0.0
    bk, pk, fcs = [], [], []
    for f in range(WF_FRAMES):
        fc = WF_RETUNE.get(f, fc)
        y = waterfall_input(f)
        tb.plot(w, np.arange(len(y)), y, fc, False, True)
        z = w.imager.last
        bk.append(float(np.nanmax(w.wf[:, -w.wf_cnt:].mean(1)) * 0 + np.median(np.mean(w.wf[:, -w.wf_cnt:], 1))))
        pk.append(np.asarray(w.pk_frqs, np.float64))
        fcs.append(fc)
        if f in WF_KEEP:
            images[f] = z.astype(np.float32)
    out = {'wf_bkgnd': np.array(bk), 'wf_fc': np.array(fcs), 'wf_cnt_final': w.wf_cnt,
           'wf_npk': np.array([len(p) for p in pk]), 'wf_pk': np.concatenate(pk) if pk else np.zeros(0)}
    for f, z in images.items():
        out['wf_img_%d' % f] = z
    print("waterfall: %d frames, peaks/frame %s..., images %s" % (WF_FRAMES, out['wf_npk'][:8], sorted(images)))
    return out


def main():
    mods = rh.load(dsp)
    # ---- parameter bags -----------------------------------------------------------------------------------------
    cases = []
    for argv in rs.PARAM_CASES:
        P = rh.run_time_params(mods, argv)
        d = {}
        for k in rs.PARAM_FIELDS:
            v = getattr(P, k)
            d[k] = v.tolist() if isinstance(v, np.ndarray) else (float(v) if isinstance(v, (np.floating,)) else v)
        d['FC'] = [float(f) for f in P.FC]
        d['SOURCE'] = [int(s) for s in P.SOURCE]
        d['MUTED'] = list(P.MUTED)
        cases.append({'argv': argv, 'P': d})
    # ---- adjust_foffset on a grid (utils.py:277-289) --------------------------------------------------------------
    class Bag(object):
        pass
    grid = []
    for rb in (32768, 65536, 131072, 262144, 524288):
        for srate in (250e3, 1.024e6, 2.048e6, 2.4e6, 8e6, 10e6):
            for fo in (0.0, 100e3, -100e3, 12345.678, 250e3, -33e3, 1.0, 0.49 * srate):
                b = Bag()
                b.RB_SIZE, b.SRATE, b.FOFFSET = rb, srate, fo
                mods['utils'].adjust_foffset(b)
                grid.append([rb, srate, fo, float(b.FOFFSET)])
    json.dump({'param_cases': cases, 'adjust_foffset': grid,
               'MODES': mods['Tables'].MODES, 'AF_BWs': mods['Tables'].AF_BWs, 'VIDEO_BWs': mods['Tables'].VIDEO_BWs,
               'RTLsrates': mods['Tables'].RTLsrates, 'SDRplaysrates': mods['Tables'].SDRplaysrates,
               'find_filter': [[bw, mods['Tables'].find_filter(bw, mods['Tables'].AF_BWs)] for bw in (24e3, 48e3, 96e3, 192e3, 2048.0)]},
              open(os.path.join(HERE, 'ref_params.json'), 'w'), indent=0, sort_keys=True)
    # ---- replay scenarios + waterfall -------------------------------------------------------------------------------
    blob = {}
    for name, sc in rs.SCENARIOS.items():
        for k, v in run_scenario(mods, name, sc).items():
            a = np.asarray(v)
            if a.dtype == np.float64 and a.ndim >= 2:
                a = a.astype(np.float32)
            elif a.dtype == np.complex128:
                a = a.astype(np.complex64)
            blob['%s/%s' % (name, k)] = a
    for k, v in run_waterfall(mods).items():
        blob['waterfall/%s' % k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'ref_callers.npz'), **blob)
    print("wrote ref_callers.npz (%d arrays, %.0f KB) and ref_params.json" % (
        len(blob), os.path.getsize(os.path.join(HERE, 'ref_callers.npz')) / 1e3))


if __name__ == "__main__":
    main()
