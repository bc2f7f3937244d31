"""Edge cases of the C-ABI path: ragged / tiny inputs, error codes, maximum receiver count, mid-stream retune."""
import os

import numpy as np
import pytest
import torch

from oracle import receiver_oracle as rxo
from oracle import sig_proc_oracle as odsp
from tests.util import assert_parity, make_both

pytestmark = pytest.mark.gpu


def _noise(n, seed, scale=0.1):
    rng = np.random.default_rng(seed)
    return ((rng.normal(size=n) + 1j * rng.normal(size=n)) * scale).astype(np.complex64)


def _bank(P, max_in):
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.receiver import receiver_offsets
    return ReceiverBank(P, receiver_offsets(P), max_in=max_in)


@pytest.mark.parametrize("tail", [1, 2, 167, 499, 500, 501, 100000])
def test_ragged_last_call(tail):
    """A capture that does not end on an IN_CHUNK_SIZE boundary: whole chunks, then a short final call."""
    P, Po = make_both(8, [1000, 1300], ['USB', 'AM'], af_bw_khz=[2, 5])
    C = P.IN_CHUNK_SIZE
    x = _noise(2 * C + tail, 3)
    bank = _bank(P, C)
    rxo.create_receivers(Po)
    cuts = [0, C, 2 * C, 2 * C + tail]
    for a, b in zip(cuts[:-1], cuts[1:]):
        am, iq, _ = bank.process(torch.from_numpy(x[a:b]).cuda())
        exp = odsp.n_out_total(b, P.UP, P.DOWN) - odsp.n_out_total(a, P.UP, P.DOWN)
        assert bank.n_out == exp
        for r in range(2):
            ref = Po.rx[r].demod_data(x[a:b])
            assert len(ref) == exp
            if exp:
                assert_parity(iq[r].cpu().numpy(), Po.rx[r].iq, "iq rx%d [%d,%d)" % (r, a, b))
                assert_parity(am[r].cpu().numpy(), ref, "am rx%d [%d,%d)" % (r, a, b))
    # the stream is now off the block grid: a further call must be refused loudly, not mis-segmented
    from pysdr_b200._lib import PysdrError
    if (2 * C + tail) % C:
        with pytest.raises(PysdrError, match="boundary"):
            bank.process(torch.from_numpy(x[:C]).cuda())


def test_tiny_whole_stream_calls():
    """Streams shorter than the filters (everything is start-up transient, x[<0] = 0)."""
    P, Po = make_both(2.048, [1000], ['USB'], af_bw_khz=[2], nfilt=101)
    for n in (1, 2, 42, 43, 127, 128, 129, 1000):
        x = _noise(n, n)
        bank = _bank(P, P.IN_CHUNK_SIZE)
        am, iq, _ = bank.process(torch.from_numpy(x).cuda())
        rxo.create_receivers(Po)
        ref = Po.rx[0].demod_data(x)
        assert bank.n_out == len(ref) == odsp.n_out_total(n, P.UP, P.DOWN)
        assert_parity(iq[0].cpu().numpy(), Po.rx[0].iq, "iq n=%d" % n)
        assert_parity(am[0].cpu().numpy(), ref, "am n=%d" % n)


def test_error_codes_are_loud():
    from pysdr_b200._lib import PysdrError
    from pysdr_b200.bank import ReceiverBank
    P, _ = make_both(8, [1000], ['USB'])
    bank = _bank(P, P.IN_CHUNK_SIZE)
    with pytest.raises(PysdrError, match="exceeds"):
        bank.process(torch.zeros(P.IN_CHUNK_SIZE + 1, dtype=torch.complex64, device="cuda"))
    with pytest.raises(PysdrError):
        bank.process(torch.zeros(16, dtype=torch.complex64))                       # host tensor
    with pytest.raises(PysdrError):
        bank.process(torch.zeros(16, dtype=torch.float32, device="cuda"))          # wrong dtype
    with pytest.raises(PysdrError, match="multiple"):
        bank.seek(12345)
    with pytest.raises(PysdrError, match="1..128"):
        ReceiverBank(P, [0.0] * 129)
    P.MODE = 'WFM'
    with pytest.raises(PysdrError, match="WFM"):
        bank.process(torch.zeros(P.IN_CHUNK_SIZE, dtype=torch.complex64, device="cuda"))
    import pysdr_b200.sig_proc as dsp
    with pytest.raises(PysdrError, match="chirp-z"):
        dsp.spectrum(8000., 70000, 140000, 0.)                                     # beyond the 2^17-point chirp-z transform
    sp = dsp.spectrum(48., 4096, 8192, 0.5)
    assert len(sp.psd_est(np.zeros(100, np.complex64), True)) == 0                 # shorter than one frame -> []
    rx = dsp.Receiver(make_both(2.048, [1000], ['USB'])[0], 1e5, 0, '1')
    assert len(rx.demod_data(np.zeros(0, np.complex64))) == 0                      # empty chunk -> empty audio


def test_eight_receivers_all_modes():
    modes = ['AM', 'NFM', 'USB', 'LSB', 'CW', 'IQ', 'AM', 'CW']
    fcs = [1000 + 150 * i for i in range(8)]
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.bank import ReceiverBank
    P = RUN_TIME_PARAMS(['-fs', '8', '-fc'] + [str(f) for f in fcs[:6]] + ['-mode'] + modes[:6] + ['-foffset', '100',
                        '-af_bw', '5', '10', '2', '3', '0.5', '20'])
    # the reference caps receivers at MAX_RX=6 (params.py:270-276); the bank itself takes 8
    assert P.NUM_RX == 6
    P.MODE = modes
    P.AF_BW = [5e3, 10e3, 2e3, 3e3, 500., 20e3, 5e3, 500.]
    P.BFO = [0, 0, 0, 0, 700, 0, 0, 600]
    offs = [P.FOFFSET + (f - fcs[0]) * 1e3 for f in fcs]
    n = 2 * P.IN_CHUNK_SIZE
    x = _noise(n, 17)
    bank = ReceiverBank(P, offs, max_in=n)
    am, iq, _ = bank.process(torch.from_numpy(x).cuda())
    Po = rxo.make_P(8e6, [f * 1e3 for f in fcs[:6]], modes[:6], foffset=100e3)
    Po.MODE, Po.AF_BW, Po.BFO = modes, P.AF_BW, P.BFO
    for r in range(8):
        orx = odsp.Receiver(Po, offs[r], r, str(r + 1))
        ref = np.concatenate([orx.demod_data(x[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE]) for c in range(2)])
        got = am[r].cpu().numpy()
        assert got.dtype == (np.complex64 if modes[r] == 'IQ' else np.float32)
        assert_parity(got, ref, "rx%d %s" % (r, modes[r]))


def test_retune_and_filter_swap_mid_batch_stream():
    """lo.change_freq / dec.h swaps between device-resident calls keep phase continuity (oracle semantics:
    the new LO and taps also apply to the carried raw filter memory)."""
    P, Po = make_both(8, [1000, 1250], ['USB', 'CW'], af_bw_khz=[2, .5])
    C = P.IN_CHUNK_SIZE
    x = _noise(4 * C, 23)
    bank = _bank(P, 2 * C)
    rxo.create_receivers(Po)
    am, _, _ = bank.process(torch.from_numpy(x[:2 * C]).cuda())
    got = [[a.cpu().numpy().copy()] for a in am]
    ref = [[Po.rx[r].demod_data(x[c * C:(c + 1) * C]).copy() for c in range(2)] for r in range(2)]
    f_new = bank.set_freq(1, 271828.1828)
    assert f_new == Po.rx[1].lo.change_freq(271828.1828)
    bank.set_dec_taps(0, bank.filter_bank[4])
    Po.rx[0].dec.h = Po.rx[0].dec.filter_bank[4]
    am, _, _ = bank.process(torch.from_numpy(x[2 * C:]).cuda())
    for r in range(2):
        got[r].append(am[r].cpu().numpy().copy())
        ref[r] += [Po.rx[r].demod_data(x[c * C:(c + 1) * C]).copy() for c in (2, 3)]
        assert_parity(np.concatenate(got[r]), np.concatenate(ref[r]), "retuned stream rx%d" % r)


def test_lfilter_stream_and_squelch_classes():
    import pysdr_b200.sig_proc as dsp
    from scipy import signal
    rng = np.random.default_rng(9)
    b, a = signal.butter(3, 0.05)
    x = rng.normal(size=30000).astype(np.float32)
    f = dsp.lfilter_stream(b, a)
    y = np.concatenate([f.run(x[:7000]), f.run(x[7000:7001]), f.run(x[7001:])])
    ref, zref = signal.lfilter(b, a, x.astype(np.float64), zi=np.zeros(3))
    assert_parity(y, ref, "lfilter_stream")
    np.testing.assert_allclose(f.z[0], zref, rtol=1e-6, atol=1e-9)
    n = np.arange(48000)
    for tone, expect_open in ((1000.0, True), (8000.0, False)):
        sig = np.sin(2 * np.pi * tone * n / 48000).astype(np.float32)
        sq, so = dsp.squelch(48000), odsp.squelch(48000)
        r1, o1 = [], []
        for c in range(0, len(sig), 12000):
            r, o = sq.run(sig[c:c + 12000])
            r1.append(r); o1.append(o)
        rr, oo = so.run(sig.astype(np.float64))
        r1 = np.concatenate(r1)
        assert bool(np.concatenate(o1)[-1]) == bool(oo[-1]) == expect_open
        np.testing.assert_allclose(r1[2000:], rr[2000:], rtol=2e-3)      # ratio of two small envelopes: float32 I/O


def test_three_box_compute_matches_plotting_py_algebra():
    from pysdr_b200.plotting import three_box_compute
    from pysdr_b200.params import RUN_TIME_PARAMS
    P = RUN_TIME_PARAMS(['-fs', '2.048', '-fc', '1000', '-mode', 'USB'])
    P.PEAK_DIST = 2.0                                             # kHz units of the AF panel (fs given in kHz)
    tb = three_box_compute(P, 48., 0, 1024, 2048, 0.5)
    so = odsp.spectrum(48., 1024, 2048, 0.5)
    ws = odsp.waterfall_state(2048, so.df, 100, pan_dr=P.PAN_DR, peak_dist=P.PEAK_DIST)
    rng = np.random.default_rng(10)
    t = np.arange(512 * 12)
    x = (0.2 * np.exp(2j * np.pi * 0.11 * t) + 0.01 * (rng.normal(size=len(t)) + 1j * rng.normal(size=len(t)))).astype(np.complex64)
    for k in range(12):
        fc = 0.0 if k < 8 else 3.0                                # retune -> waterfall roll (Plotting.py:689-695)
        seg = x[k * 512:(k + 1) * 512]
        out = tb.plot(seg, fc)
        PSDo = so.periodogram(seg, True)
        img_ref, bk_ref, pk_ref = ws.push(PSDo, fc)
        top = PSDo > PSDo.max() - 60
        assert np.max(np.abs(out['PSD'] - PSDo)[top]) < 2e-3
        assert abs(out['bkgnd'] - bk_ref) <= 2e-3 + 1e-5 * abs(bk_ref)
        if k >= 2:
            assert len(out['peaks']) >= 1 and abs(int(out['peaks'][np.argmax(out['PSD'][out['peaks']])]) -
                                                  int(pk_ref[np.argmax(PSDo[pk_ref])])) <= 1
        np.testing.assert_allclose(out['image'].cpu().numpy(), img_ref, rtol=0, atol=5e-2)
    # RGBA through the jet lookup table (Plotting.py:139-142): same index law on the host
    from pysdr_b200.plotting import jet, lookup_table
    rgba = tb.image_rgba().cpu().numpy()
    img = out['image'].cpu().numpy()
    hi = img.max()
    idx = np.clip(np.floor((img - (hi - P.PAN_DR)) * (255.0 / P.PAN_DR) + 0.5), 0, 255).astype(int)
    want = lookup_table(jet(64), 256)[idx]
    assert rgba.shape == want.shape == (2048, 100, 4)
    assert np.mean(np.any(rgba != want, axis=-1)) < 1e-3         # LUT index ties at float rounding only
    assert np.abs(rgba.astype(int) - want.astype(int)).max() <= 8


@pytest.mark.parametrize("deemph,direct", [(0, False), (75, False), (0, True)])
def test_wfm_demod_first_then_resample(deemph, direct):
    """BASELINE config 4 geometry (2.4 MS/s replay rate, 1/50): WFM chain = video FIR @RF rate -> discriminator ->
    resampler (AF low-pass) -> AGC (-> de-emphasis), reference gui.py:1703-1704,1759-1762.  Mono."""
    import pysdr_b200.sig_proc as dsp
    P, Po = make_both(2.4, [100000], ['WFM'], foffset_khz=100, srate_hz=2.4e6, af_bw_khz=[15])
    assert (P.UP, P.DOWN, P.IN_CHUNK_SIZE, P.VIDEO_BW) == (1, 50, 51200, 200e3) == (Po.UP, Po.DOWN, Po.IN_CHUNK_SIZE, Po.VIDEO_BW)
    P.DEEMPH_US = Po.DEEMPH_US = deemph
    P.WFM_DIRECT_VIDEO = direct                     # True: the r01 route (direct-form video FIR through K1 + pysdr_fm_disc)
    C = P.IN_CHUNK_SIZE
    n = np.arange(5 * C)
    ph = 2 * np.pi * P.FOFFSET * n / P.SRATE + (75e3 / 1e3) * np.sin(2 * np.pi * 1e3 * n / P.SRATE) \
        + (20e3 / 5e3) * np.sin(2 * np.pi * 5e3 * n / P.SRATE)
    x = (0.3 * np.exp(1j * ph) + _noise(len(n), 4, 0.003)).astype(np.complex64)
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1')
    orx = odsp.Receiver(Po, Po.FOFFSET, 0, '1')
    for c in range(5):
        if c == 3:                                                  # video filter swap (gui.py:1704) and AF change
            rx.demod.wfm_video.h = rx.demod.wfm_filter_bank[8]
            orx._demod_wfm(np.zeros(0, np.complex64)) if not hasattr(orx, 'wfm_vid') else None
            orx.demod.wfm_video.h = orx.demod.wfm_filter_bank[8]
            P.AF_BW = Po.AF_BW = 10e3
        if c == 2:                                                  # retune: phase continuous, applied to the filter memory too
            assert rx.lo.change_freq(P.FOFFSET + 2.5e3) == orx.lo.change_freq(Po.FOFFSET + 2.5e3)
        am = rx.demod_data(x[c * C:(c + 1) * C])
        ref = orx.demod_data(x[c * C:(c + 1) * C])
        assert (rx._wfm.vbank is None) == (not direct)
        assert am.dtype == np.float32 and len(am) == len(ref) == 1024
        assert_parity(rx.iq, orx.iq, "wfm resampled chunk %d" % c)
        assert_parity(am, ref, "wfm audio chunk %d" % c)
    assert np.max(np.abs(am)) > 0.05                                # a real demodulated tone, not silence


def test_rtty_filterbank_vs_reference_run():
    """The one PINNED parity target: device lines vs the fixture the reference's own
    RTTY_Executive.run wrote (tests/golden/make_golden_rtty.py), fed in uneven pushes of whole symbols."""
    from pysdr_b200.rtty import rtty_filterbank
    from tests.util import rtty_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rtty_fbank.npz"))
    fb = rtty_filterbank(48000)
    assert (fb.N, fb.NFFT, fb.NSTART) == (int(g['N']), int(g['NFFT']), list(g['NSTART']))
    x = rtty_input(int(g['n_sym']), fb.N)
    parts, pos = [], 0
    for k in (1, 1, 3, 2, 5):                                   # first symbol alone yields nothing (rtty.py:826-829)
        parts.append(fb.push(x[pos * fb.N:(pos + k) * fb.N]))
        pos += k
    assert [len(p) for p in parts] == [0, 4, 12, 8, 20]
    lines = np.concatenate(parts)
    ref = g['lines'].astype(np.float64)
    lin, rlin = 10.0 ** (lines / 10.0), 10.0 ** (ref / 10.0)
    assert_parity(lin, rlin, "rtty filterbank power", rel_tol=1e-4, snr_min=80)
    top = ref > ref.max() - 60.0
    assert np.max(np.abs(lines[top] - ref[top])) < 2e-3         # dB, within 60 dB of the strongest bin
    assert np.array_equal(np.argmax(lines, axis=1), np.argmax(ref, axis=1))


def test_am_synch_pll_chain():
    """AM-Synch: carrier PLL (open-choice law, oracle am_pll) -> in-phase arm -> AF FIR -> AGC; carrier 17 Hz off the
    receiver centre so the loop has to pull in; mid-stream am_pll.reset() (reference receiver.py:649)."""
    import pysdr_b200.sig_proc as dsp
    P, Po = make_both(2.048, [1000], ['AM-Synch'], foffset_khz=100, af_bw_khz=[5])
    C = P.IN_CHUNK_SIZE
    n = np.arange(7 * C)
    env = 1.0 + 0.6 * np.sin(2 * np.pi * 1e3 * n / P.SRATE) + 0.2 * np.sin(2 * np.pi * 2.3e3 * n / P.SRATE)
    x = (0.2 * env * np.exp(2j * np.pi * (P.FOFFSET + 17.0) * n / P.SRATE + 0.7j) + _noise(len(n), 9, 0.002)).astype(np.complex64)
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1')
    orx = odsp.Receiver(Po, Po.FOFFSET, 0, '1')
    for c in range(7):
        if c == 5:
            rx.demod.am_pll.reset(); orx.demod.am_pll.reset()
        am = rx.demod_data(x[c * C:(c + 1) * C])
        ref = orx.demod_data(x[c * C:(c + 1) * C])
        assert len(am) == len(ref)
        assert_parity(am, ref, "am-synch chunk %d" % c)
        assert abs(rx.demod.am_pll.phi - orx.demod.am_pll.phi) < 1e-4 and abs(rx.demod.am_pll.w - orx.demod.am_pll.w) < 1e-7
    # locked: loop frequency = carrier offset, 17 Hz at 48 kHz
    assert abs(rx.demod.am_pll.w * P.FS_OUT / (2 * np.pi) - 17.0) < 0.5


def test_am_synch_in_bank_with_other_modes_and_direct_fir():
    """AM-Synch next to AM/USB in one bank; FFT and direct-form AF filters both match the oracle; the state blob
    carries the loop state across a bank hand-over."""
    from pysdr_b200.receiver import receiver_offsets
    P, Po = make_both(2.048, [1000, 1020, 1045], ['AM-Synch', 'AM', 'USB'], foffset_khz=100, af_bw_khz=[5, 5, 2])
    C = P.IN_CHUNK_SIZE
    n = np.arange(3 * C)
    x = _noise(len(n), 21, 0.002).astype(np.complex128)
    for k, off in enumerate(receiver_offsets(P)):
        env = 1.0 + 0.5 * np.sin(2 * np.pi * (700.0 + 300 * k) * n / P.SRATE)
        x = x + 0.1 * env * np.exp(2j * np.pi * (off + 11.0) * n / P.SRATE)
    x = x.astype(np.complex64)
    rxo.create_receivers(Po)
    ref = [[], [], []]
    for k in range(3):
        for r in range(3):
            ref[r].append(np.array(Po.rx[r].demod_data(x[k * C:(k + 1) * C])))
    for direct in (0, 1):
        bank = _bank(P, C)
        bank.lib.pysdr_bank_force_direct_fir(bank.h, direct)
        outs = [bank.process_host(x[k * C:(k + 1) * C], want_dc=False)[0] for k in range(2)]
        blob = bank.get_state()
        bank2 = _bank(P, C)
        bank2.lib.pysdr_bank_force_direct_fir(bank2.h, direct)
        bank2.set_state(blob)
        outs.append(bank2.process_host(x[2 * C:3 * C], want_dc=False)[0])
        for r in range(3):
            got = np.concatenate([o[r] for o in outs])
            assert_parity(got, np.concatenate(ref[r]), "rx%d direct=%d" % (r, direct))


def _fm_stereo_iq(P, n, fl=1e3, fr=3e3, pilot=True, seed=5):
    """Stereo-multiplex FM at offset P.FOFFSET: L/R tones, 19 kHz pilot (10 %), 38 kHz DSB-SC difference channel."""
    t = np.arange(n) / P.SRATE
    L, R = 0.4 * np.sin(2 * np.pi * fl * t), 0.4 * np.sin(2 * np.pi * fr * t + 0.5)
    mpx = 0.45 * (L + R) + 0.45 * (L - R) * np.cos(2 * np.pi * 38e3 * t + 2 * 0.3)
    if pilot:
        mpx = mpx + 0.1 * np.cos(2 * np.pi * 19e3 * t + 0.3)
    ph = 2 * np.pi * P.FOFFSET * t + 2 * np.pi * 75e3 * np.cumsum(mpx) / P.SRATE
    return (0.3 * np.exp(1j * ph) + _noise(n, seed, 0.001)).astype(np.complex64)


def test_wfm2_stereo_pilot_recovery():
    """BASELINE config 4: wideband stereo FM at 2.4 MS/s, 1/50, pilot recovery and 75 us de-emphasis: parity with the
    oracle's stereo law, and a functional check that L and R really separate."""
    import pysdr_b200.sig_proc as dsp
    P, Po = make_both(2.4, [100000], ['WFM2'], foffset_khz=100, srate_hz=2.4e6, af_bw_khz=[15])
    P.VIDEO_BW = Po.VIDEO_BW = 300e3                # params.py:326 widens the default for 'WFM' only; -vid_bw 300
    P.DEEMPH_US = Po.DEEMPH_US = 75
    C = P.IN_CHUNK_SIZE
    nchunk = 6
    x = _fm_stereo_iq(P, nchunk * C)
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1')
    orx = odsp.Receiver(Po, Po.FOFFSET, 0, '1')
    outs = []
    for c in range(nchunk):
        am = rx.demod_data(x[c * C:(c + 1) * C])
        ref = orx.demod_data(x[c * C:(c + 1) * C])
        assert am.dtype == np.complex64 and len(am) == len(ref) == 1024
        if c >= 1:                                  # chunk 0 is all start-up transient (pilot filter still empty)
            assert_parity(am.real, ref.real, "wfm2 L chunk %d" % c)
            assert_parity(am.imag, ref.imag, "wfm2 R chunk %d" % c)
        outs.append(am)
    a = np.concatenate(outs[3:])                    # settled part
    w = np.hanning(len(a))
    SL, SR = np.abs(np.fft.rfft(a.real * w)), np.abs(np.fft.rfft(a.imag * w))
    k1, k3 = int(round(1e3 * len(a) / P.FS_OUT)), int(round(3e3 * len(a) / P.FS_OUT))
    pk = lambda S, k: S[k - 2:k + 3].max()
    assert pk(SL, k1) > 20 * pk(SL, k3)             # left carries the 1 kHz tone, > 26 dB separation
    assert pk(SR, k3) > 20 * pk(SR, k1)             # right carries the 3 kHz tone


def test_wfm_to_wfm2_switch_restarts_chain():
    import pysdr_b200.sig_proc as dsp
    P, Po = make_both(2.4, [100000], ['WFM'], foffset_khz=100, srate_hz=2.4e6, af_bw_khz=[15])
    P.VIDEO_BW = Po.VIDEO_BW = 300e3
    C = P.IN_CHUNK_SIZE
    x = _fm_stereo_iq(P, 4 * C, seed=6)
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1')
    orx = odsp.Receiver(Po, Po.FOFFSET, 0, '1')
    for c in range(4):
        if c == 2:
            P.MODE = Po.MODE = 'WFM2'
        am = rx.demod_data(x[c * C:(c + 1) * C])
        ref = orx.demod_data(x[c * C:(c + 1) * C])
        assert am.dtype == ref.dtype and len(am) == len(ref)
        if c != 2:                                  # chunk 2: stage-2 start-up, the pilot filter is still empty
            assert_parity(am.real, ref.real, "chunk %d" % c)
            assert_parity(am.imag, ref.imag, "chunk %d R" % c) if c > 2 else None


def test_many_channel_bank_cfg5_geometry():
    """BASELINE config 5 geometry (10 MS/s, 3/625, channels on a 9.6 kHz raster) at a testable size: 20 channels in
    three groups, mixed modes, two chunks; every channel against its own oracle receiver."""
    from pysdr_b200.channelizer import ChannelBank, raster_offsets
    P, Po = make_both(10, [7000], ['USB'], af_bw_khz=[2])
    assert (P.UP, P.DOWN, P.IN_CHUNK_SIZE) == (3, 625, 213333)
    n_ch, C = 20, P.IN_CHUNK_SIZE
    offs = raster_offsets(n_ch, 9600.0, 150e3)
    modes = [['AM', 'NFM', 'USB', 'CW', 'LSB'][k % 5] for k in range(n_ch)]
    afs = [[5e3, 10e3, 2e3, 500., 3e3][k % 5] for k in range(n_ch)]
    n = np.arange(2 * C)
    x = _noise(len(n), 77, 0.01).astype(np.complex128)
    for k, f in enumerate(offs):
        x = x + 0.02 * (1 + 0.5 * np.sin(2 * np.pi * (300.0 + 40 * k) * n / P.SRATE)) * np.exp(2j * np.pi * (f + 700.0) * n / P.SRATE)
    x = x.astype(np.complex64)
    cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=2 * C, group=8)
    am, iq = cb.process(torch.from_numpy(x).cuda())
    assert len(am) == n_ch and cb.n_out == odsp.n_out_total(2 * C, P.UP, P.DOWN)
    for k in range(n_ch):
        Pk = rxo.make_P(P.SRATE, [7000e3], modes[k], foffset=100e3, af_bw=afs[k], bfo=700.0)
        orx = odsp.Receiver(Pk, offs[k], 0, str(k), fast=True)
        ref = np.concatenate([np.array(orx.demod_data(x[c * C:(c + 1) * C])) for c in range(2)])
        assert_parity(am[k].cpu().numpy(), ref, "channel %d (%s)" % (k, modes[k]))


def test_replay_streamer_equals_resident_processing():
    """Host capture through the double-buffered streaming executive == the same capture processed resident."""
    from pysdr_b200.receiver import ReplayStreamer, receiver_offsets
    P, _ = make_both(8, [-500, 700, 1400], ['AM', 'IQ', 'USB'], af_bw_khz=[5, 45, 2])
    C = P.IN_CHUNK_SIZE
    n_chunks = 11
    x = _noise(n_chunks * C, 8, 0.05)
    bank = _bank(P, n_chunks * C)
    am, iq, _ = bank.process(torch.from_numpy(x).cuda(), want_dc=False)
    st = ReplayStreamer(P, seg_chunks=4, want_iq=True)           # segments of 4, 4 and 3 chunks
    hx = st.pin(x)
    for _ in range(2):                                           # second run reuses buffers and restarts the stream
        h_am, n_seg = st.run(hx)
        assert len(n_seg) == 3 and sum(n_seg) == bank.n_out
        for r in range(3):
            got = st.audio(r)
            ref = am[r].cpu().numpy()
            assert got.dtype == ref.dtype and len(got) == len(ref)
            assert_parity(got, ref, "streamed rx%d" % r, rel_tol=2e-5, snr_min=90)
    pos = 0
    for i, no in enumerate(n_seg):
        assert np.array_equal(st.h_iq[i, 0, :no].numpy(), iq[0][pos:pos + no].cpu().numpy())     # K1 bit-exact
        pos += no


@pytest.mark.parametrize("chunk,nfft,overlap", [(32818, 65636, 0.0), (1000, 3000, 0.5), (32768, 65536, 0.5), (30, 45, 0.0)])
def test_spectrum_non_power_of_two_nfft_chirp_z(chunk, nfft, overlap):
    """The reference's RF panel geometry (Plotting.py:370-375: chunk 32818, NFFT 65636 = 4*61*269) and other lengths the
    radix-16 transforms do not cover, through Bluestein's identity on the 2^17-point four-step FFT; oracle = numpy fft."""
    import pysdr_b200.sig_proc as dsp
    sp = dsp.spectrum(8000., chunk, nfft, overlap)
    so = odsp.spectrum(8000., chunk, nfft, overlap)
    assert sp.czt and np.array_equal(sp.frq, so.frq)
    hop = max(1, int(chunk * (1 - overlap)))
    n = chunk + 5 * hop + 7
    t = np.arange(n)
    x = (_noise(n, 12, 0.02).astype(np.complex128) + 0.3 * np.exp(2j * np.pi * 0.1234 * t)
         + 0.05 * np.exp(-2j * np.pi * 0.37 * t)).astype(np.complex64)
    lin, ref = sp.psd_est(x, False), so.psd_est(x, False)
    assert lin.shape == ref.shape == (nfft,)
    assert_parity(lin, ref, "czt psd %d/%d" % (chunk, nfft), rel_tol=1e-4, snr_min=80)
    wf, wref = sp.waterfall(x, 2, False), so.waterfall(x, 2, False)
    assert wf.shape == wref.shape == (3, nfft)
    assert_parity(wf, wref, "czt waterfall", rel_tol=1e-4, snr_min=80)
    one, oref = sp.periodogram(x[:chunk], True), so.periodogram(x[:chunk], True)
    top = oref > oref.max() - 60
    assert np.max(np.abs(one - oref)[top]) < 5e-3                                  # dB, within 60 dB of the peak


def test_replay_from_capture_file_through_streamer(tmp_path):
    """-replay end to end: capture file -> pinned read -> rates from the file header (receiver.py:808-822) ->
    streaming executive -> demod file; against the oracle loop on the same samples."""
    from pysdr_b200.fileio import sdr_fileio, open_replay
    from pysdr_b200.receiver import ReplayStreamer
    Pw, _ = make_both(2.048, [1000], ['USB'], af_bw_khz=[2])
    Pw.SAVE_DIR = str(tmp_path)
    C = Pw.IN_CHUNK_SIZE
    x = _noise(6 * C, 14, 0.05)
    w = sdr_fileio('raw_iq', 'w', Pw, 2, 'RAW_IQ')
    w.save_data(x); w.close()
    P, Po = make_both(8, [1000], ['USB'], af_bw_khz=[2])         # started with another rate: the file decides
    open_replay(P, w.fname)
    assert P.IN_CHUNK_SIZE == C and P.SRATE == 2.048e6
    st = ReplayStreamer(P, seg_chunks=4)
    hx = P.sdr.read_data(pinned=True)
    assert hx.is_pinned() and hx.numel() == 6 * C
    st.run(hx)
    got = st.audio(0)
    from pysdr_b200.receiver import receiver_offsets
    Po2 = rxo.make_P(2.048e6, [1000e3], 'USB', foffset=100e3, af_bw=2e3)
    # FOFFSET stays as quantised for the start-up rate (params.py:472 runs before the file is opened, receiver.py:811)
    orx = odsp.Receiver(Po2, receiver_offsets(P)[0], 0, '1', fast=True)
    ref = np.concatenate([orx.demod_data(x[c * C:(c + 1) * C]) for c in range(6)])
    assert_parity(got, ref, "replayed file")
    P.SAVE_DIR = str(tmp_path)
    d = sdr_fileio('demod', 'w', P, 1, 'USB')
    d.save_data(got); d.close()
    assert np.array_equal(sdr_fileio(d.fname, 'r', None).read_data(), got)


def test_cs16_source_streams_at_half_the_pcie_bytes():
    """CS16 capture (reference receiver.py:609-617: sc = 1/2048) converted on the device == the same samples fed as
    complex64; the conversion is exact in float32."""
    from pysdr_b200.receiver import ReplayStreamer
    P, _ = make_both(2.048, [1000, 1030], ['USB', 'AM'], af_bw_khz=[2, 5])
    C = P.IN_CHUNK_SIZE
    rng = np.random.default_rng(3)
    raw = rng.integers(-2048, 2048, size=2 * 7 * C + 2, dtype=np.int16)[:2 * 7 * C]
    x = (raw[0::2].astype(np.float64) / 2048.0 + 1j * raw[1::2].astype(np.float64) / 2048.0).astype(np.complex64)
    a = ReplayStreamer(P, seg_chunks=3, want_iq=True)
    a.run(a.pin(x))
    b = ReplayStreamer(P, seg_chunks=3, want_iq=True, fmt='cs16')
    b.run(b.pin(raw))
    for r in range(2):
        assert np.array_equal(a.audio(r), b.audio(r))
    assert torch.equal(a.h_iq[:3], b.h_iq[:3])
    with pytest.raises(ValueError):
        b.run(a.pin(x))


def test_paired_fft_keeps_a_weak_signal_accurate_next_to_a_strong_one():
    """AM and NFM receivers share one complex FFT in the AF filter (k2_fftconv.cu).  A carrier-less NFM channel (its
    discriminator output ~1e-7) must not inherit the rounding error of a strong AM envelope (~0.3): the pair is brought
    to the same binade by an exact power-of-two scale.  Without it this case fails the gate by two orders of magnitude."""
    P, Po = make_both(8, [1000, 1400], ['AM', 'NFM'], af_bw_khz=[5, 10])
    C = P.IN_CHUNK_SIZE
    n = np.arange(3 * C)
    from pysdr_b200.receiver import receiver_offsets
    offs = receiver_offsets(P)
    x = (0.3 * (1 + 0.5 * np.sin(2 * np.pi * 1e3 * n / P.SRATE)) * np.exp(2j * np.pi * offs[0] * n / P.SRATE)
         + _noise(len(n), 31, 3e-4)).astype(np.complex64)
    bank = _bank(P, 3 * C)
    am, _, _ = bank.process(torch.from_numpy(x).cuda(), want_dc=False)
    rxo.create_receivers(Po)
    for r in range(2):
        ref = np.concatenate([Po.rx[r].demod_data(x[c * C:(c + 1) * C]) for c in range(3)])
        assert_parity(am[r].cpu().numpy(), ref, "rx%d" % r)


def test_raster_channelizer_equals_per_channel_k1():
    """wola.cu: 96 channels on the 9.6 kHz raster of config 5 (3/3125 of 10 MS/s) from one windowing pass + one 3125-point
    inverse DFT per output instant == the per-receiver K1 of ChannelBank (and the oracle's resampler), two calls with
    carried history, ragged second call."""
    from pysdr_b200.channelizer import ChannelBank, RasterChannelizer
    P, Po = make_both(10, [7000], ['IQ'])
    n_ch, f0, df = 96, -450e3, 9600.0
    C = P.IN_CHUNK_SIZE
    n1, n2 = 2 * C, C + 12345
    x = _noise(n1 + n2, 41, 0.05)
    t = np.arange(len(x))
    x = (x + 0.2 * np.exp(2j * np.pi * (f0 + 17 * df + 300.0) * t / P.SRATE)).astype(np.complex64)
    rc = RasterChannelizer(P, f0, df, n_ch)
    xd = torch.from_numpy(x).cuda()
    y1 = rc.process(xd[:n1], n0=0, n_before=0)
    keep = rc.lp - 1
    y2 = rc.process(xd[n1 - keep:], n0=n1, n_before=keep)
    got = torch.cat((y1, y2), dim=1).cpu().numpy()
    assert got.shape == (n_ch, odsp.n_out_total(n1 + n2, P.UP, P.DOWN))
    cb = ChannelBank(P, rc.offsets, 'IQ', max_in=n1 + n2)
    _, iq = cb.process(xd[:3 * C])                                # whole chunks only for the bank
    k = cb.n_out
    for c in (0, 1, 17, 50, 95):
        assert_parity(got[c, :k], iq[c].cpu().numpy(), "wola vs K1, channel %d" % c, rel_tol=2e-5, snr_min=90)
    for c in (17, 95):
        dec = odsp.decimator(Po.SRATE, Po.UP, Po.DOWN, Po.FILT_LEN, odsp.VIDEO_BWs, Po.VIDEO_BW)
        dec.h = dec.filter_bank[odsp._video_index(Po)]
        lo = odsp.signal_generator(rc.offsets[c], Po.IN_CHUNK_SIZE, Po.SRATE, True)
        ref = dec.resamp_fast(x, lo)
        assert_parity(got[c], ref, "wola vs oracle, channel %d" % c, rel_tol=2e-5, snr_min=90)


def test_many_channel_bank_raster_mode_two_blocks():
    """Config 5 with the raster channelizer in front: ONE wola.cu pass per block writes every channel's baseband into the
    groups' shared complex memory, the groups run only their audio-rate stages; two blocks (carried input history, AF
    memories, AGC), every channel against its own oracle receiver."""
    from pysdr_b200.channelizer import ChannelBank, raster_offsets
    P, Po = make_both(10, [7000], ['USB'], af_bw_khz=[2])
    n_ch, C = 20, P.IN_CHUNK_SIZE
    offs = raster_offsets(n_ch, 9600.0, 150e3)
    modes = [['AM', 'NFM', 'USB', 'CW', 'LSB'][k % 5] for k in range(n_ch)]
    afs = [[5e3, 10e3, 2e3, 500., 3e3][k % 5] for k in range(n_ch)]
    n = np.arange(3 * C)
    x = _noise(len(n), 78, 0.01).astype(np.complex128)
    for k, f in enumerate(offs):
        x = x + 0.02 * (1 + 0.5 * np.sin(2 * np.pi * (300.0 + 40 * k) * n / P.SRATE)) * np.exp(2j * np.pi * (f + 700.0) * n / P.SRATE)
    x = x.astype(np.complex64)
    cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=2 * C, raster=(offs[0], 9600.0))
    xd = torch.from_numpy(x).cuda()
    outs, iqs = [], []
    for a, b in ((0, 2 * C), (2 * C, 3 * C)):
        am, iq = cb.process(xd[a:b])
        outs.append([v.cpu().numpy().copy() for v in am])
        iqs.append([v.cpu().numpy().copy() for v in iq])
    for k in range(n_ch):
        Pk = rxo.make_P(P.SRATE, [7000e3], modes[k], foffset=100e3, af_bw=afs[k], bfo=700.0)
        orx = odsp.Receiver(Pk, offs[k], 0, str(k), fast=True)
        ref, refq = [], []
        for c in range(3):
            ref.append(np.array(orx.demod_data(x[c * C:(c + 1) * C])))
            refq.append(orx.iq.copy())
        got = np.concatenate([o[k] for o in outs])
        assert_parity(np.concatenate([q[k] for q in iqs]), np.concatenate(refq), "iq channel %d" % k, rel_tol=2e-5, snr_min=90)
        assert_parity(got, np.concatenate(ref), "channel %d (%s)" % (k, modes[k]))
