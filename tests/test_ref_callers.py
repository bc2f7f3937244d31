"""Pins against fixtures written by the REFERENCE's own code (tests/golden/make_golden_ref_callers.py executes
/root/reference's params.py, utils.py, receiver.py and Plotting.py through stub modules, with the numpy oracle standing
behind `import sig_proc`):

  CPU  the product's RUN_TIME_PARAMS / adjust_foffset / tables equal the reference's outputs;
       the restated oracle loop (oracle/receiver_oracle.py) reproduces the reference's SDR_EXECUTIVE.Run /
       demodulate_data / audio_out outputs EXACTLY (same arithmetic under the seam) — iteration counts, the stale last
       chunk, DC removal, player payloads, mutes, mode / filter / frequency changes;
       the oracle's waterfall algebra reproduces three_box_plot.plot's images, background and peaks.
  GPU  pysdr_b200.receiver.SDR_EXECUTIVE and pysdr_b200.plotting.three_box_compute reproduce the same fixtures within the
       north_star tolerance (max-abs rel err <= 1e-4, difference SNR >= 80 dB).

A third test drives the product through the reference's REAL callers when both a GPU and /root/reference are present
(never the case for the driver's runs; it documents the INTEGRATION.md binding for a maintainer's box)."""
import json
import os
import sys

import numpy as np
import pytest

from tests import ref_scenarios as rs
from tests.util import assert_parity

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "ref_callers.npz"), allow_pickle=False)
GP = json.load(open(os.path.join(HERE, "golden", "ref_params.json")))


def g(name, key):
    return G['%s/%s' % (name, key)]


# ---- parameter surface ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GP['param_cases'], ids=lambda c: ' '.join(c['argv'][2:8]))
def test_run_time_params_match_reference(case):
    from pysdr_b200.params import RUN_TIME_PARAMS
    P = RUN_TIME_PARAMS(case['argv'])
    ref = case['P']
    for k in rs.PARAM_FIELDS:
        got = getattr(P, k)
        if isinstance(got, np.ndarray):
            got = got.tolist()
        assert got == ref[k], (k, got, ref[k])
        assert type(got) is type(ref[k]) or isinstance(got, (int, float)) and isinstance(ref[k], (int, float)), (k, type(got))
    assert [float(f) for f in P.FC] == ref['FC']
    assert [int(s) for s in P.SOURCE] == ref['SOURCE']
    assert list(P.MUTED) == ref['MUTED']
    for k in ('UP', 'DOWN', 'FS_OUT', 'IN_CHUNK_SIZE', 'RB_SIZE', 'MUTE_CHUNKS', 'DELAY'):
        assert isinstance(getattr(P, k), int), k                 # integer arithmetic stays integer (params.py:405-406,444)


def test_adjust_foffset_matches_reference_grid():
    from pysdr_b200 import design
    from oracle import sig_proc_oracle as odsp

    class Bag(object):
        pass
    for rb, srate, fo, want in GP['adjust_foffset']:
        b = Bag()
        b.RB_SIZE, b.SRATE, b.FOFFSET = rb, srate, fo
        design.adjust_foffset(b)
        assert b.FOFFSET == want, (rb, srate, fo)
        assert odsp.adjust_foffset(fo, srate, rb) == want


def test_tables_match_reference():
    from pysdr_b200 import design
    from oracle import sig_proc_oracle as odsp
    for mod in (design, odsp):
        assert list(mod.MODES) == GP['MODES']
        assert list(mod.AF_BWs) == GP['AF_BWs']
        assert list(mod.VIDEO_BWs) == GP['VIDEO_BWs']
        assert list(mod.RTLsrates) == GP['RTLsrates'] and list(mod.SDRplaysrates) == GP['SDRplaysrates']
        for bw, lab in GP['find_filter']:
            assert mod.find_filter(bw, mod.AF_BWs) == lab


# ---- the replay loop ---------------------------------------------------------------------------------------------------
def _product_P(sc):
    from pysdr_b200.params import RUN_TIME_PARAMS
    return RUN_TIME_PARAMS(sc['argv'])


def _wire(P, sc, x, ring):
    """What pySDR.py's main sets up around the executive, as in the generator: players, PSD taps, file taps, and the
    control events riding the raw-IQ tap (fired at the end of every loop iteration)."""
    P.sdr = rs.ReplayFile(x, P.SRATE, float(P.FC[0]))
    P.audio_playback = True
    P.players = [rs.Player(P, P.FS_OUT, ring('Audio%d' % (i + 1), P.RB_SIZE), None) for i in range(P.NUM_PLAYERS)]
    P.SHOW_AF_PSD, P.PLOT_RX, P.rb_af = True, 0, ring('AF', P.RB_SIZE)
    P.SHOW_BASEBAND_PSD, P.rb_baseband = True, ring('BB', P.RB_SIZE)
    P.SAVE_DEMOD = P.SAVE_BASEBAND = P.SAVE_IQ = True
    P.demod_io, P.baseband_iq_io = rs.Recorder(), rs.Recorder()
    log = dict(auto_muted=[], am=[])

    def end_of_iteration(k):
        log['auto_muted'].append(bool(P.AUTO_MUTED))
        log['am'].append([np.array(P.rx[i].am, copy=True) for i in range(P.NUM_RX)])
        for ev in sc['events'].get(k, []):
            rs.apply_event(P, ev)
    P.raw_iq_io = rs.Recorder(hook=end_of_iteration)
    return log


@pytest.mark.parametrize("name", list(rs.SCENARIOS))
def test_oracle_loop_reproduces_reference_loop_exactly(name):
    """oracle/receiver_oracle.run_replay (the restated loop) == the reference's own Run()/demodulate_data/audio_out,
    bit for bit: both stand on the same numpy operators, so any difference is a difference in the loop logic."""
    from oracle import receiver_oracle as rxo
    sc = rs.SCENARIOS[name]
    P = _product_P(sc)                                           # parameter bag (pinned above); operators: the oracle's
    P.AUDIO_SCHEME = int(P.AUDIO_SCHEME)
    x = rs.scenario_input(sc, P)
    rxo.create_receivers(P)

    def on_iter(k):
        for ev in sc['events'].get(k, []):
            rs.apply_event(P, ev)
    out, iters = rxo.run_replay(P, x, collect=('am', 'iq', 'am_dc', 'audio', 'auto_muted'), on_iteration=on_iter)
    assert iters == int(g(name, 'iters')) and P.nchunks == int(g(name, 'nchunks'))
    assert str(P.MODE) == str(g(name, 'mode_final'))
    assert int(P.IN_CHUNK_SIZE) == int(g(name, 'in_chunk')) and int(P.FS_OUT) == int(g(name, 'fs_out'))
    assert np.array_equal(np.array(out['auto_muted']), g(name, 'auto_muted'))
    am = np.array([[out['am'][i][c] for i in range(P.NUM_RX)] for c in range(iters)])
    assert np.array_equal(am, g(name, 'am'))
    assert np.array_equal(np.array(out['am_dc'][0]).astype(np.float32), g(name, 'rb_af'))
    assert np.array_equal(np.array(out['iq'][0]), g(name, 'baseband_io'))
    for i in range(int(g(name, 'n_players'))):
        ref = g(name, 'player%d' % i)
        got = np.array(out['audio'][i]).astype(ref.dtype)
        assert np.array_equal(got, ref), "player %d" % i


def test_oracle_waterfall_reproduces_reference_plot():
    from oracle import sig_proc_oracle as odsp
    from tests.golden import make_golden_ref_callers as mk        # frame synthesis only (no reference access)
    psd = odsp.spectrum(mk.WF_FS, mk.WF_CHUNK, mk.WF_NFFT, 0.0)
    wf = odsp.waterfall_state(mk.WF_NFFT, psd.df, ncols=100, pan_dr=60.0, peak_dist=2.0)
    fc = 0.0
    k = 0
    pk = G['waterfall/wf_pk']
    for f in range(mk.WF_FRAMES):
        fc = mk.WF_RETUNE.get(f, fc)
        line = psd.periodogram(mk.waterfall_input(f), True)
        image, bkgnd, peaks = wf.push(line, fc)
        assert bkgnd == G['waterfall/wf_bkgnd'][f]
        n = int(G['waterfall/wf_npk'][f])
        assert np.array_equal(psd.frq[peaks] + fc, pk[k:k + n]), f
        k += n
        if f in mk.WF_KEEP:
            assert np.array_equal(image.astype(np.float32), G['waterfall/wf_img_%d' % f]), f
    assert wf.wf_cnt == int(G['waterfall/wf_cnt_final'])


# ---- the B200 executive against the reference's outputs -------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(rs.SCENARIOS))
def test_b200_executive_reproduces_reference_loop(name):
    from pysdr_b200 import sig_proc as dsp
    from pysdr_b200.receiver import SDR_EXECUTIVE
    sc = rs.SCENARIOS[name]
    P = _product_P(sc)
    x = rs.scenario_input(sc, P)
    log = _wire(P, sc, x, dsp.ring_buffer2)
    pushed = {}
    for rb in [p.rb for p in P.players] + [P.rb_af, P.rb_baseband]:
        pushed[rb.tag] = []
        rb.push = (lambda t: (lambda v: pushed[t].append(np.array(v, copy=True))))(rb.tag)   # keep every block
    ex = SDR_EXECUTIVE(P)
    iters = ex.Run()
    assert iters == int(g(name, 'iters')) and P.nchunks == int(g(name, 'nchunks'))
    assert str(P.MODE) == str(g(name, 'mode_final')) and P.SHUT_DOWN == bool(g(name, 'shut_down'))
    assert int(P.IN_CHUNK_SIZE) == int(g(name, 'in_chunk')) and int(P.FS_OUT) == int(g(name, 'fs_out'))
    assert np.array_equal(np.array(log['auto_muted']), g(name, 'auto_muted'))
    raw = P.raw_iq_io.saved
    assert np.array_equal(np.array([v[0] for v in raw]), g(name, 'raw_first'))       # which chunk each iteration saw
    assert np.array_equal(np.array([v[-1] for v in raw]), g(name, 'raw_last'))
    ref_am = g(name, 'am')
    for c in range(iters):
        for i in range(P.NUM_RX):
            assert_parity(log['am'][c][i], ref_am[c, i], "%s am chunk %d rx %d" % (name, c, i))
    for c in range(iters):
        assert_parity(pushed['AF'][c], g(name, 'rb_af')[c], "%s AF-PSD tap chunk %d" % (name, c))
        assert_parity(P.demod_io.saved[c], g(name, 'rb_af')[c], "%s demod file chunk %d" % (name, c))
        assert_parity(pushed['BB'][c], g(name, 'baseband_io')[c], "%s baseband tap chunk %d" % (name, c))
        assert_parity(P.baseband_iq_io.saved[c], g(name, 'baseband_io')[c], "%s baseband file chunk %d" % (name, c))
    for i in range(int(g(name, 'n_players'))):
        ref = g(name, 'player%d' % i)
        assert P.players[i].starts == int(g(name, 'player%d_starts' % i))
        for c in range(iters):
            got = pushed['Audio%d' % (i + 1)][c]
            if np.max(np.abs(ref[c])) == 0:
                assert np.max(np.abs(got)) == 0, "%s player %d chunk %d must be silent (muted)" % (name, i, c)
            else:
                assert_parity(got, ref[c], "%s player %d chunk %d" % (name, i, c))


@pytest.mark.gpu
def test_b200_waterfall_reproduces_reference_plot():
    from pysdr_b200.plotting import three_box_compute
    from tests.golden import make_golden_ref_callers as mk

    class PP(object):
        PAN_DR, PEAK_DIST, RIG_IF = 60.0, 2.0, 0
    tb = three_box_compute(PP(), mk.WF_FS, 0.0, mk.WF_CHUNK, mk.WF_NFFT, 0.0)
    fc = 0.0
    k = 0
    pk = G['waterfall/wf_pk']
    for f in range(mk.WF_FRAMES):
        fc = mk.WF_RETUNE.get(f, fc)
        r = tb.plot(mk.waterfall_input(f), fc)
        assert abs(r['bkgnd'] - G['waterfall/wf_bkgnd'][f]) <= 2e-3, (f, r['bkgnd'], G['waterfall/wf_bkgnd'][f])   # dB
        n = int(G['waterfall/wf_npk'][f])
        assert len(r['pk_frqs']) == n and np.allclose(r['pk_frqs'], pk[k:k + n], rtol=0, atol=1e-9), f
        k += n
        if f in mk.WF_KEEP:
            ref = G['waterfall/wf_img_%d' % f]
            got = r['image'].cpu().numpy()
            assert got.shape == ref.shape
            assert np.max(np.abs(got - ref)) <= 2e-3, (f, float(np.max(np.abs(got - ref))))        # dB, on a 60 dB range
    assert tb.wf_cnt == int(G['waterfall/wf_cnt_final'])


@pytest.mark.gpu
def test_reference_callers_drive_b200_receivers():
    """The INTEGRATION.md binding, executed: the reference's unmodified receiver.py (SDR_EXECUTIVE.Run, demodulate_data,
    audio_out) with `sig_proc` = pysdr_b200.sig_proc.  Needs the reference tree AND a GPU, so it is skipped on the driver's
    GPU box (no /root/reference there) and in the CPU container (no GPU); tools/run_reference_callers.py is the same
    thing as a script."""
    from tests.golden import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present on this box")
    from tools.run_reference_callers import run
    res = run('am2')
    assert res['iters'] == int(g('am2', 'iters'))
    for c in range(res['iters']):
        for i in range(2):
            assert_parity(res['am'][c][i], g('am2', 'am')[c, i], "reference-driven am chunk %d rx %d" % (c, i))
