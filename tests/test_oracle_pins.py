"""Pins the oracle against every known-answer item the reference tree holds for the path
(SURVEY.md section 8c).  Reference line numbers are cited per test."""
import os

import numpy as np
import pytest
from scipy import signal

from oracle import sig_proc_oracle as dsp
from oracle import receiver_oracle as rxo

# reference srates.py:35-74 (expected output pasted by the author as comments): fs1 MHz, fs2 kHz, UP, DOWN
SRATES_TABLE = """
0.25 48 24 125
0.25 96 48 125
0.25 192 96 125
0.5 48 12 125
0.5 96 24 125
0.5 192 48 125
1.0 48 6 125
1.0 96 12 125
1.0 192 24 125
2.0 48 3 125
2.0 96 6 125
2.0 192 12 125
2.048 48 3 128
2.048 96 3 64
2.048 192 3 32
3.0 48 2 125
3.0 96 4 125
3.0 192 8 125
4.0 48 3 250
4.0 96 3 125
4.0 192 6 125
5.0 48 6 625
5.0 96 12 625
5.0 192 24 625
6.0 48 1 125
6.0 96 2 125
6.0 192 4 125
7.0 48 6 875
7.0 96 12 875
7.0 192 24 875
8.0 48 3 500
8.0 96 3 250
8.0 192 3 125
9.0 48 2 375
9.0 96 4 375
9.0 192 8 375
10.0 48 3 625
10.0 96 6 625
10.0 192 12 625
"""


def test_up_dn_golden_table():
    rows = [r.split() for r in SRATES_TABLE.strip().splitlines()]
    assert len(rows) == 39
    for f1, f2, up, dn in rows:
        fs1 = int(float(f1) * 1e6)                 # srates.py:29
        fs2 = int(float(f2) * 1e3)
        assert dsp.up_dn(fs1, fs2) == (int(up), int(dn)), (f1, f2)


@pytest.mark.parametrize("srate,up,down,chunk", [
    (2.048e6, 3, 128, 43690), (8e6, 3, 500, 170666), (10e6, 3, 625, 213333), (2.4e6, 1, 50, 51200)])
def test_chunk_sizes(srate, up, down, chunk):
    # reference params.py:405-406,444 ; SURVEY section 8 derived-size table
    u, d, fs_out, ic = dsp.derived_rates(srate, 48e3)
    assert (u, d, fs_out, ic) == (up, down, 48000, chunk)


def test_adjust_foffset_examples():
    # reference utils.py:277-289 ; SURVEY 8a row a2: 100 kHz @ 8 MS/s, RB_SIZE 131072 -> 99975.5859375
    assert dsp.rb_size(4, 48000) == 131072
    assert dsp.adjust_foffset(100e3, 8e6, 131072) == 99975.5859375
    assert dsp.rb_size(1, 48000) == 32768
    assert dsp.adjust_foffset(100e3, 2.048e6, 32768) == 100e3


def test_make_P_cfg2():
    P = rxo.make_P(8e6, [1e6 - 1.5e6, 1e6 - .3e6, 1e6 + .4e6, 1e6 + 2.1e6], ['AM', 'NFM', 'USB', 'CW'],
                   foffset=100e3)
    assert (P.UP, P.DOWN, P.FS_OUT, P.IN_CHUNK_SIZE, P.RB_SIZE) == (3, 500, 48000, 170666, 131072)
    assert P.FOFFSET == 99975.5859375
    assert P.MUTE_CHUNKS == 11                     # int(.25*48000/1024), params.py:449
    assert P.BFO == [0, 0, 0, 700]                 # params.py:316-318
    assert P.VIDEO_BW == 10e3                      # params.py:322-327


def test_iir_notch_coefficients_and_chunk_identity():
    # reference sigs/iir.py:57 (coefficients quoted in SURVEY a13) and :83-125 (chunked == whole, exact)
    b, a = dsp.iir_designs()['notch50']
    np.testing.assert_allclose(b, [0.99220706, -1.88728999, 0.99220706], atol=5e-9)
    np.testing.assert_allclose(a, [1, -1.88728999, 0.98441413], atol=5e-9)
    rng = np.random.default_rng(0)
    fs, T = 1000, 2
    t = np.linspace(0, T, T * fs)
    x = np.sin(2 * np.pi * 15 * t) + np.sin(2 * np.pi * 50 * t) + rng.normal(0, .1, T * fs) * 0.03
    for name, (b, a) in dsp.iir_designs().items():
        y = signal.lfilter(b, a, x)
        for nchunk in (2, 3):
            f = dsp.iir_stream(b, a)
            n3 = len(x) // nchunk
            parts = [x[i * n3:(i + 1) * n3] for i in range(nchunk - 1)] + [x[(nchunk - 1) * n3:]]
            yy = np.concatenate([f.run(p) for p in parts])
            assert np.max(np.abs(y - yy)) == 0.0, name


def test_agc_loop_filter_is_agc_m():
    # reference sigs/agc.m:6-12: b=beta, a=[1 beta-1], beta=.1 on a step
    beta = .1
    x = np.concatenate((np.zeros(100), np.ones(100)))
    y = signal.lfilter([beta], [1, beta - 1], x)
    g = dsp.agc(ref=1.0, beta=beta, nb=1)
    g.gain = 0.0
    # decay branch of the oracle AGC (want > gain) is exactly that loop filter driven by 'want'
    out = []
    for v in x:
        want = v
        g.gain = g.beta * want + (1 - g.beta) * g.gain
        out.append(g.gain)
    np.testing.assert_allclose(out, y, rtol=0, atol=1e-15)
    # and through the public update(): constant peak 0.5, ref 1 -> want 2, gain walks 1 -> 2 by the loop filter
    g = dsp.agc(ref=1.0, beta=beta, nb=1)
    gains = [g.update(0.5) for _ in range(50)]
    ref = 1 + signal.lfilter([beta], [1, beta - 1], np.ones(50))
    np.testing.assert_allclose(gains, ref, rtol=1e-12)
    assert g.maxbuf == 0.5 and abs(g.err - (2 - gains[-2])) < 1e-12


def test_agc_attack_and_maxbuf():
    g = dsp.agc()
    g.update(0.01)
    assert g.gain > 1
    gg = g.update(10.0)                            # loud block: immediate attack to ref/peak
    assert gg == dsp.AGC_REF / 10.0
    for _ in range(dsp.AGC_NB - 1):
        g.update(0.01)
    assert g.maxbuf == 10.0                        # still remembered
    g.update(0.01)
    assert g.maxbuf == 0.01                        # aged out after NB blocks


def test_nfm_discriminator_is_nfm_m():
    # reference sigs/nfm.m:123-127
    rng = np.random.default_rng(1)
    y = (rng.normal(size=300) + 1j * rng.normal(size=300))
    IQ = y[2:]
    d = IQ - y[:-2]
    y1 = y[1:-1]
    fm_ref = y1.real * d.imag - y1.imag * d.real
    dm = dsp.demodulator(48000, 1)
    dm.filter_bank_real = [np.array([1.0], np.float32)]      # no AF filtering: isolate the discriminator
    fm = np.concatenate([dm.demod(y[:100], 'NFM', 0, 0), dm.demod(y[100:], 'NFM', 0, 0)])
    # streaming version has one sample of latency and a zero-history start-up
    np.testing.assert_allclose(fm[2:], fm_ref, rtol=1e-12, atol=1e-12)
    # pure tone of frequency f: discriminator = Im(conj(y1)*d) = 2*A^2*sin(2*pi*f/fs)
    n = np.arange(1000)
    z = 0.5 * np.exp(2j * np.pi * 1000 * n / 48000)
    dm.reset()
    fm = dm.demod(z, 'NFM', 0, 0)
    np.testing.assert_allclose(fm[2:], 2 * 0.25 * np.sin(2 * np.pi * 1000 / 48000), rtol=1e-9)


def test_squelch_smoother_is_squelch_m():
    # reference sigs/squelch.m:125-128: sq = filter(alpha,[1 alpha-1],abs(z)), alpha=.001
    rng = np.random.default_rng(2)
    z = rng.normal(size=5000)
    alpha = 0.001
    ref = signal.lfilter([alpha], [1, alpha - 1], np.abs(z))
    s = dsp.iir_stream([alpha], [1, alpha - 1])
    got = np.concatenate([s.run(np.abs(z[:1234])), s.run(np.abs(z[1234:]))])
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    sq = dsp.squelch(48000)
    n = np.arange(48000)
    tone = np.sin(2 * np.pi * 1000 * n / 48000)               # in-band energy only -> open
    ratio, op = sq.run(tone)
    assert op[-1] and ratio[-1] > 10
    sq = dsp.squelch(48000)
    hiss = np.sin(2 * np.pi * 8000 * n / 48000)               # out-of-band only -> closed
    ratio, op = sq.run(hiss)
    assert (not op[-1]) and ratio[-1] < 0.1


def test_spectrum_follows_rtty_fft_idiom():
    # reference rtty.py:839-841:  X = fftshift(fft(xx*window, NFFT)); 10*log10(re^2+im^2)
    rng = np.random.default_rng(3)
    N, NFFT = 256, 512
    x = (rng.normal(size=N) + 1j * rng.normal(size=N)).astype(np.complex64)
    sp = dsp.spectrum(48., N, NFFT, 0.0)
    got = sp.periodogram(x, True)
    w = sp.win.astype(np.float64)
    X = np.fft.fftshift(np.fft.fft(x.astype(np.complex128) * w, NFFT))
    ref = 10 * np.log10((np.square(X.real) + np.square(X.imag)) / np.sum(w ** 2))
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    assert sp.NFFT == NFFT and abs(sp.df - 48. / NFFT) < 1e-15 and sp.new_samps == N
    assert sp.frq[NFFT // 2] == 0 and abs(sp.frq[0] + 24.) < 1e-12
    assert len(sp.periodogram(np.zeros(0), True)) == 0        # failure -> empty (Plotting.py:463-465)


def test_spectrum_overlap_and_average():
    rng = np.random.default_rng(4)
    N, NFFT = 128, 256
    x = (rng.normal(size=N * 8) + 1j * rng.normal(size=N * 8))
    sp = dsp.spectrum(48., N, NFFT, 0.5)
    assert sp.new_samps == N // 2
    fr = sp.frames(x)
    assert len(fr) == 1 + (len(x) - N) // (N // 2)
    # streaming periodogram with new_samps pushes == frame k once the window is full
    sp2 = dsp.spectrum(48., N, NFFT, 0.5)
    outs = [sp2.periodogram(x[i * 64:(i + 1) * 64], False) for i in range(16)]
    np.testing.assert_allclose(np.fft.ifftshift(outs[1]), fr[0], rtol=1e-12)
    np.testing.assert_allclose(np.fft.ifftshift(outs[5]), fr[4], rtol=1e-12)
    np.testing.assert_allclose(sp.psd_est(x, False), np.fft.fftshift(fr.mean(0)), rtol=1e-12)


def test_waterfall_algebra_matches_plotting_py():
    # reference Plotting.py:385-388,536-548,583-587,594,618-626,689-695 restated literally here
    rng = np.random.default_rng(5)
    nfft, ncols, df = 64, 100, 0.5
    ws = dsp.waterfall_state(nfft, df, ncols, pan_dr=60.0, peak_dist=4.0)
    wf = -1e38 * np.ones((nfft, ncols))
    cnt = 0
    wf_fc = 0
    for it in range(7):
        PSD = rng.normal(size=nfft) * 3 - 80
        PSD[20] += 40
        fc = 0 if it < 4 else 2.0                 # retune by 2.0 -> roll by int(2/.5+.5)=4 bins
        nb = int(float(fc - wf_fc) / df + 0.5)
        if nb != 0:
            wf = np.roll(wf, -nb, axis=0)
            wf_fc = fc
        line = PSD.reshape(-1, 1)
        wf = np.concatenate((wf[:, 1:], line), axis=1)
        cnt = min(cnt + 1, ncols)
        PSD2 = np.mean(wf[:, -cnt:], 1)
        bk = np.median(PSD2)
        zz = wf - bk
        img_ref = np.maximum(zz, np.nanmax(zz) - 60.0)
        img, bkgnd, peaks = ws.push(PSD, fc)
        assert bkgnd == bk
        np.testing.assert_array_equal(img, img_ref)
        pk_ref, _ = signal.find_peaks(PSD2, distance=4.0 / df, height=bk + 10)
        np.testing.assert_array_equal(peaks, pk_ref)


def test_nco_is_exact_function_of_index():
    inc = dsp.freq_to_phase_inc(99975.5859375, 8e6)
    assert inc == (1638 << 47)                      # M/RB_SIZE = 1638/131072 exactly representable
    lo = dsp.signal_generator(99975.5859375, 1000, 8e6)
    a = lo.lo(1000)
    b = lo.lo(1000)
    n = np.arange(2000)
    ref = np.exp(2j * np.pi * ((1638 * n) % 131072) / 131072.)
    np.testing.assert_allclose(np.concatenate((a, b)), ref, atol=2e-9)
    assert lo.change_freq(-1.5e6) == pytest.approx(-1.5e6, abs=1e-9)


def test_decimator_definition_vs_upfirdn_and_chunking():
    rng = np.random.default_rng(6)
    for (srate, up, down, L) in [(8e6, 3, 500, 1001), (2.048e6, 3, 128, 1001), (2.4e6, 1, 50, 301),
                                 (0.25e6, 24, 125, 1001), (1e6, 6, 125, 200)]:
        n = 7 * down + 123
        x = (rng.normal(size=n) + 1j * rng.normal(size=n))
        d0 = dsp.decimator(srate, up, down, L)
        d0.h = d0.filter_bank[2]
        whole = d0.resamp(x)
        assert len(whole) == dsp.n_out_total(n, up, down)
        ref = signal.upfirdn(d0.h.astype(np.float64), x, up, down)[:len(whole)]
        np.testing.assert_allclose(whole, ref, rtol=1e-10, atol=1e-12)
        for fn in ('resamp', 'resamp_fast'):
            d1 = dsp.decimator(srate, up, down, L)
            d1.h = d1.filter_bank[2]
            cuts = [0, 100, 100 + down, 3 * down + 7, n]
            parts = [getattr(d1, fn)(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
            lens = [dsp.n_out_total(b, up, down) - dsp.n_out_total(a, up, down) for a, b in zip(cuts[:-1], cuts[1:])]
            assert [len(p) for p in parts] == lens
            np.testing.assert_allclose(np.concatenate(parts), whole, rtol=1e-9, atol=1e-11)


def test_replay_loop_quirks():
    # strict '<' and the stale last chunk (reference receiver.py:544, 715-725)
    P = rxo.make_P(2.048e6, [1e6], 'USB', foffset=100e3, nfilt=101)
    rxo.create_receivers(P)
    rng = np.random.default_rng(7)
    raw = (rng.normal(size=3 * P.IN_CHUNK_SIZE) + 1j * rng.normal(size=3 * P.IN_CHUNK_SIZE)).astype(np.complex64) * .1
    out, iters = rxo.run_replay(P, raw)
    # 3 full chunks present but '<' only admits 2; third iteration hits EOF and re-processes chunk 2
    assert iters == 3
    assert len(out['am'][0]) == 3
    assert all(len(a) in (1023, 1024, 1025) for a in out['am'][0])
    # DC-removed copy has zero mean per chunk; audio copy does not (receiver.py:250-252 vs :194)
    assert abs(np.mean(out['am_dc'][0][1])) < 1e-9


def test_rtty_filterbank_matches_reference_run():
    """PINNED: tests/golden/rtty_fbank.npz holds the lines the reference's own RTTY_Executive.run produced
    (tests/golden/make_golden_rtty.py); the restated oracle must reproduce sizes and values."""
    from oracle import rtty_oracle as ro
    from tests.util import rtty_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rtty_fbank.npz"))
    R = ro.RTTY_Params(48000)
    assert (R.N, R.NFFT, R.NSTART, R.NBINS) == (int(g['N']), int(g['NFFT']), list(g['NSTART']), int(g['NBINS'])) == (1056, 2048, [0, 264, 528, 792], 7)
    assert np.array_equal(R.frq, g['frq'])
    assert np.allclose(np.kaiser(R.N, 8.6), g['window'], rtol=0, atol=0)
    lines = ro.filterbank_lines(rtty_input(int(g['n_sym']), R.N), 48000)
    assert lines.shape == g['lines'].shape == (44, 2048)
    assert np.max(np.abs(lines - g['lines'])) < 1e-5            # float32 storage of values up to 74 dB
