"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Gates (BASELINE.json north_star): decimation sample indexing bit-exact (n_out per chunk, sample<->input
mapping), audio / baseband / PSD max-abs relative error <= 1e-4 and difference SNR >= 80 dB.
"""
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


# ------------------------------------------------------------------------------------------------ K1
GEOMS = [
    # srate_mhz, srate_hz(replay), n_rx, nfilt, note
    (8, None, 4, 1001, "cfg2 3/500"),
    (2.048, None, 1, 1001, "cfg1 3/128"),
    (8, None, 1, 1001, "3/500 1rx"),
    (8, None, 2, 1001, "3/500 2rx"),
    (8, None, 3, 1001, "3/500 3rx (padded to 4)"),
    (8, None, 6, 1001, "3/500 6rx (4+2)"),
    (10, None, 4, 1001, "cfg5 rate 3/625 (odd DOWN)"),
    (2, None, 2, 1001, "3/125"),
    (0.25, None, 2, 1001, "24/125 (many phases)"),
    (6, None, 1, 1001, "1/125 (UP=1, 1001 taps/phase)"),
    (2.4, 2.4e6, 1, 301, "cfg4 rate 1/50"),
    (1, None, 4, 200, "6/125 short filter"),
    (7, None, 2, 1001, "6/875"),
]


@pytest.mark.parametrize("srate_mhz,srate_hz,n_rx,nfilt,note", GEOMS)
@pytest.mark.parametrize("variant", ["fast", "generic"])
def test_k1_mix_decimate_parity(srate_mhz, srate_hz, n_rx, nfilt, note, variant):
    fcs = [1000 + 137.5 * i * (1 if i % 2 else -1) for i in range(n_rx)]
    P, Po = make_both(srate_mhz, fcs, ['IQ'], foffset_khz=100 if srate_mhz >= 1 else 20, nfilt=nfilt, srate_hz=srate_hz)
    n_chunks = 3
    n = n_chunks * P.IN_CHUNK_SIZE
    x = _noise(n, 11)
    bank = _bank(P, n)
    if variant == "generic":
        bank.force_generic(True)
        assert bank.k1_variant == 0
    else:
        assert bank.k1_variant == 1, "fast K1 should support " + note
    xd = torch.from_numpy(x).cuda()
    # whole capture in one call
    _, iq, _ = bank.process(xd)
    got_whole = [v.cpu().numpy() for v in iq]
    # chunked through a second bank
    bank2 = _bank(P, P.IN_CHUNK_SIZE)
    if variant == "generic":
        bank2.force_generic(True)
    parts = [[] for _ in range(n_rx)]
    for c in range(n_chunks):
        _, iqc, _ = bank2.process(xd[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE])
        # bit-exact indexing: number of outputs per chunk
        exp = odsp.n_out_total((c + 1) * P.IN_CHUNK_SIZE, P.UP, P.DOWN) - odsp.n_out_total(c * P.IN_CHUNK_SIZE, P.UP, P.DOWN)
        assert bank2.n_out == exp
        for r in range(n_rx):
            parts[r].append(iqc[r].cpu().numpy())
    rxo.create_receivers(Po)
    for r in range(n_rx):
        ref = np.concatenate([Po.rx[r].dec.resamp(x[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE], Po.rx[r].lo)
                              for c in range(n_chunks)])
        assert len(got_whole[r]) == len(ref) == odsp.n_out_total(n, P.UP, P.DOWN)
        assert_parity(got_whole[r], ref, "K1 %s %s rx%d" % (variant, note, r))
        # chunked == whole, bit for bit (same per-output arithmetic order)
        np.testing.assert_array_equal(np.concatenate(parts[r]), got_whole[r])


def test_k1_indexing_is_bit_exact_on_impulses():
    """An impulse at input k must land on exactly the outputs/taps the contract names:
    y[m] = h[(m*DOWN)%UP + (n_m - k)*UP] with n_m=(m*DOWN)//UP (LO at 0 Hz)."""
    P, Po = make_both(8, [1000], ['IQ'], foffset_khz=0)
    P.FOFFSET = 0.0
    n = 2 * P.IN_CHUNK_SIZE
    bank = _bank(P, n)
    bank.set_freq(0, 0.0)
    h = bank.filter_bank[2]
    for k in (0, 1, 499, 500, 170665, 170666, 250001):
        bank.reset()
        x = np.zeros(n, np.complex64)
        x[k] = 1.0
        _, iq, _ = bank.process(torch.from_numpy(x).cuda())
        y = iq[0].cpu().numpy()
        m = np.arange(len(y), dtype=np.int64)
        t = m * P.DOWN
        nm, pm = t // P.UP, t % P.UP
        idx = pm + (nm - k) * P.UP
        ok = (idx >= 0) & (idx < len(h))
        exp = np.zeros(len(y), np.float32)
        exp[ok] = h[idx[ok]]
        np.testing.assert_array_equal(y.imag, 0)
        np.testing.assert_array_equal(y.real, exp)                  # bit-exact: single non-zero product


def test_k1_unaligned_input_pointer_and_halo_in_place():
    P, Po = make_both(8, [1000, 1400], ['IQ'])
    n = 2 * P.IN_CHUNK_SIZE
    x = _noise(n + 4001, 5)
    xd = torch.from_numpy(x).cuda()
    ref_bank = _bank(P, n)
    _, iq, _ = ref_bank.process(xd[:n].clone())
    ref = [v.cpu().numpy().copy() for v in iq]
    # 8-byte (odd element) aligned view of the same samples
    shifted = torch.empty(n + 1, dtype=torch.complex64, device="cuda")
    shifted[1:] = xd[:n]
    b2 = _bank(P, n)
    _, iq2, _ = b2.process(shifted[1:])
    for r in range(2):
        np.testing.assert_array_equal(iq2[r].cpu().numpy(), ref[r])
    # time shard with the halo in place: second chunk processed alone equals the second half of the whole
    b3 = _bank(P, P.IN_CHUNK_SIZE)
    b3.seek(P.IN_CHUNK_SIZE)
    _, iq3, _ = b3.process(xd[P.IN_CHUNK_SIZE:n], halo_in_place=True)
    m_split = odsp.n_out_total(P.IN_CHUNK_SIZE, P.UP, P.DOWN)
    for r in range(2):
        np.testing.assert_array_equal(iq3[r].cpu().numpy(), ref[r][m_split:])


# ------------------------------------------------------------------------------------------------ chain
@pytest.mark.parametrize("mode,af_khz", [("AM", 5), ("NFM", 10), ("USB", 2), ("LSB", 3), ("CW", 0.5), ("IQ", 45),
                                         ("AM", 0), ("USB", 0)])
def test_full_chain_parity_per_mode(mode, af_khz):
    from pysdr_b200.synth import synth_iq
    P, Po = make_both(8, [1000], [mode], af_bw_khz=[af_khz])
    n_chunks = 4
    n = n_chunks * P.IN_CHUNK_SIZE
    x = synth_iq(n, P.SRATE, [P.FOFFSET], [mode], seed=3).numpy()
    x[:P.IN_CHUNK_SIZE] *= 0.2                                    # level step -> exercises AGC attack/decay
    bank = _bank(P, n)
    am, iq, dc = bank.process(torch.from_numpy(x).cuda())
    rxo.create_receivers(Po)
    ref_am, ref_dc = [], []
    for c in range(n_chunks):
        ref_dc.append(rxo.demodulate_data(Po, x[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE], 0))
        ref_am.append(Po.rx[0].am.copy())
    ref_am, ref_dc = np.concatenate(ref_am), np.concatenate(ref_dc)
    assert_parity(am[0].cpu().numpy(), ref_am, "am %s" % mode)
    assert_parity(dc[0].cpu().numpy(), ref_dc, "am_dc %s" % mode)   # mean subtraction cancels signal
    st = bank.agc_get(0)
    if mode != 'IQ':
        assert abs(st['gain'] - Po.rx[0].agc.gain) <= 1e-5 * Po.rx[0].agc.gain
        assert abs(st['maxbuf'] - Po.rx[0].agc.maxbuf) <= 1e-4 * Po.rx[0].agc.maxbuf
        assert st['ref'] == Po.rx[0].agc.ref


def test_cfg2_four_receivers_chunked_equals_whole_and_oracle():
    from pysdr_b200.synth import synth_iq
    fcs = [-500, 700, 1400, 3100]
    modes = ['AM', 'NFM', 'USB', 'CW']
    P, Po = make_both(8, fcs, modes, af_bw_khz=[5, 10, 2, .5])
    offs = [P.FOFFSET + f - P.FC[0] for f in P.FC]
    n_chunks = 5
    n = n_chunks * P.IN_CHUNK_SIZE
    xd = synth_iq(n, P.SRATE, offs, modes, seed=9, device="cuda")
    bank = _bank(P, n)
    am, iq, dc = bank.process(xd)
    whole = [a.cpu().numpy().copy() for a in am]
    b2 = _bank(P, P.IN_CHUNK_SIZE)
    parts = [[] for _ in range(4)]
    for c in range(n_chunks):
        a2, _, _ = b2.process(xd[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE])
        for r in range(4):
            parts[r].append(a2[r].cpu().numpy().copy())
    x = xd.cpu().numpy()
    rxo.create_receivers(Po)
    for r in range(4):
        assert_parity(np.concatenate(parts[r]), whole[r], "chunked vs whole rx%d" % r, rel_tol=2e-5, snr_min=90)
        ref = np.concatenate([Po.rx[r].demod_data(x[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE]) for c in range(n_chunks)])
        assert_parity(whole[r], ref, "cfg2 rx%d %s" % (r, modes[r]))


def test_af_filter_fft_path_equals_direct_form():
    """K2's overlap-save FFT convolution against the direct-form FIR kernel (GPU vs GPU) and chunk invariance."""
    from pysdr_b200.synth import synth_iq
    fcs = [-500, 700, 1400, 3100, 1000, 1200]
    modes = ['AM', 'NFM', 'USB', 'CW', 'LSB', 'IQ']
    P, Po = make_both(8, fcs, modes, af_bw_khz=[5, 10, 2, .5, 3, 20])
    offs = [P.FOFFSET + f - P.FC[0] for f in P.FC]
    n = 7 * P.IN_CHUNK_SIZE                                        # 7168 outputs: 3 overlap-save blocks of 3096
    xd = synth_iq(n, P.SRATE, offs, modes, seed=5, device="cuda")
    a = _bank(P, n)
    am_fft = [v.cpu().numpy().copy() for v in a.process(xd)[0]]
    b = _bank(P, n)
    b.force_direct_fir(True)
    am_dir = [v.cpu().numpy().copy() for v in b.process(xd)[0]]
    for r in range(6):
        assert_parity(am_fft[r], am_dir[r], "fft vs direct rx%d %s" % (r, modes[r]), rel_tol=2e-5, snr_min=90)
    c = _bank(P, P.IN_CHUNK_SIZE)
    parts = [[] for _ in range(6)]
    for k in range(7):
        am, _, _ = c.process(xd[k * P.IN_CHUNK_SIZE:(k + 1) * P.IN_CHUNK_SIZE])
        for r in range(6):
            parts[r].append(am[r].cpu().numpy().copy())
    for r in range(6):
        # block boundaries of the overlap-save differ between the two runs -> equal to FFT round-off, not bitwise
        assert_parity(np.concatenate(parts[r]), am_fft[r], "chunked vs whole rx%d" % r, rel_tol=2e-5, snr_min=90)


def test_state_checkpoint_roundtrip():
    P, Po = make_both(8, [1000, 1400], ['USB', 'AM'], af_bw_khz=[2, 5])
    n = 4 * P.IN_CHUNK_SIZE
    xd = torch.from_numpy(_noise(n, 21)).cuda()
    a = _bank(P, P.IN_CHUNK_SIZE)
    outs = []
    for c in range(4):
        am, _, _ = a.process(xd[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE])
        outs.append([v.cpu().numpy().copy() for v in am])
        if c == 1:
            blob = a.get_state()
    b = _bank(P, P.IN_CHUNK_SIZE)
    b.set_state(blob)
    assert b.position == 2 * P.IN_CHUNK_SIZE
    for c in (2, 3):
        am, _, _ = b.process(xd[c * P.IN_CHUNK_SIZE:(c + 1) * P.IN_CHUNK_SIZE])
        for r in range(2):
            np.testing.assert_array_equal(am[r].cpu().numpy(), outs[c][r])


def test_front_back_split_with_replayed_peaks_equals_streaming():
    """Time-shard contract: a shard that (i) warms its filter memories on the preceding chunk and (ii) replays
    the earlier blocks' AGC peaks reproduces the single-stream result exactly."""
    P, Po = make_both(8, [1000, 1400], ['USB', 'NFM'], af_bw_khz=[2, 10])
    C = P.IN_CHUNK_SIZE
    n = 6 * C
    xd = torch.from_numpy(_noise(n, 33)).cuda()
    xd[:2 * C] *= 0.3
    ref_bank = _bank(P, n)
    pk_all = torch.zeros((2, 6), dtype=torch.float32, device="cuda")
    ref_bank.process_front(xd, pk_all)
    am, _, _ = ref_bank.process_back()
    ref = [a.cpu().numpy().copy() for a in am]
    m3 = odsp.n_out_total(3 * C, P.UP, P.DOWN)
    # shard = chunks 3..5, warm-up on chunk 2
    sh = _bank(P, 3 * C)
    sh.seek(2 * C)
    warm = torch.zeros((2, 1), dtype=torch.float32, device="cuda")
    sh.process_front(xd[2 * C:3 * C], warm, halo_in_place=True)
    sh.process_back()                                              # discard
    pk = torch.zeros((2, 3), dtype=torch.float32, device="cuda")
    sh.process_front(xd[3 * C:], pk, halo_in_place=True)
    am2, _, _ = sh.process_back(prev_peaks=pk_all[:, :3].contiguous())
    torch.testing.assert_close(pk, pk_all[:, 3:], rtol=1e-5, atol=0)   # overlap-save block boundaries differ: FFT round-off
    for r in range(2):
        # earlier blocks' AGC is replayed as a parallel composition (re-associated float64): equal to ~1e-7
        assert_parity(am2[r].cpu().numpy(), ref[r][m3:], "shard vs single stream rx%d" % r, rel_tol=2e-5, snr_min=90)


# ------------------------------------------------------------------------------------------------ L1 surface
def test_reference_surface_receiver_object():
    import pysdr_b200.sig_proc as dsp
    P, Po = make_both(2.048, [1000], ['USB'], af_bw_khz=[2])
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1', dsp.design.VIDEO_BWs, dsp.design.AF_BWs)
    orx = odsp.Receiver(Po, Po.FOFFSET, 0, '1')
    x = _noise(6 * P.IN_CHUNK_SIZE, 2)
    C = P.IN_CHUNK_SIZE
    for c in range(6):
        if c == 2:                                                 # retune (gui.py:1938) -> applied frequency returned
            assert rx.lo.change_freq(123456.789) == orx.lo.change_freq(123456.789) == rx.lo.fo
        if c == 3:                                                 # video filter swap (gui.py:1713)
            rx.dec.h = rx.dec.filter_bank[5]
            orx.dec.h = orx.dec.filter_bank[5]
        if c == 4:                                                 # mode + AF filter change read at call time
            P.MODE = Po.MODE = 'AM'
            P.AF_BW = Po.AF_BW = 5e3
            rx.agc.reset(); orx.agc.reset()                        # receiver.py:648
        chunk = x[c * C:(c + 1) * C].copy()
        am = rx.demod_data(chunk)
        chunk[:] = 0                                               # caller may reuse its buffer (receiver.py:445)
        ref = orx.demod_data(x[c * C:(c + 1) * C])
        assert am is rx.am and am.dtype == np.float32 and rx.iq.dtype == np.complex64
        assert_parity(rx.iq, orx.iq, "iq chunk %d" % c)
        assert_parity(am, ref, "am chunk %d" % c)
    assert abs(rx.agc.gain - orx.agc.gain) < 1e-5 * orx.agc.gain and rx.agc.ref == orx.agc.ref
    assert abs(rx.agc.maxbuf - orx.agc.maxbuf) < 1e-4 * orx.agc.maxbuf
    # auto-mute (receiver.py:238-245)
    big = (np.ones(C) * (1 + 1j)).astype(np.complex64)
    assert rx.auto_mute(big) and orx.auto_mute(big)
    small = x[:C]
    for _ in range(P.MUTE_CHUNKS):
        assert rx.auto_mute(small) == orx.auto_mute(small)
    assert not rx.auto_mute(small)


def test_signal_generator_quad_mixer():
    import pysdr_b200.sig_proc as dsp
    x = _noise(50000, 8)
    a = dsp.signal_generator(99975.5859375, 1000, 8e6, True)
    b = odsp.signal_generator(99975.5859375, 1000, 8e6, True)
    y = np.concatenate([a.quad_mixer(x[:20000]), a.quad_mixer(x[20000:])])
    yr = np.concatenate([b.quad_mixer(x[:20000]), b.quad_mixer(x[20000:])])
    assert a.fo == b.fo
    assert_parity(y, yr, "quad_mixer")


def test_executive_replay_loop_matches_oracle_loop():
    from pysdr_b200.receiver import SDR_EXECUTIVE
    P, Po = make_both(2.048, [1000, 1050], ['USB', 'AM'], af_bw_khz=[2, 5], nfilt=301)
    raw = _noise(3 * P.IN_CHUNK_SIZE, 4)
    rxo.create_receivers(Po)
    ref, it_ref = rxo.run_replay(Po, raw)
    got = {0: [], 1: []}
    ex = SDR_EXECUTIVE(P)
    iters = ex.Run(raw, sink=lambda irx, audio, dc, iq: got[irx].append((audio, dc, iq)))
    assert iters == it_ref == 3                                    # strict '<' + stale last chunk (receiver.py:544,715-725)
    g = pow(10., .5) - 1
    for irx in range(2):
        for c in range(3):
            assert_parity(got[irx][c][0], ref['am'][irx][c] * g, "audio rx%d c%d" % (irx, c))
            assert_parity(got[irx][c][2], ref['iq'][irx][c], "iq rx%d c%d" % (irx, c))
        assert_parity(got[1][1][1], ref['am_dc'][1][1], "am_dc")


def test_convolver_streaming():
    import pysdr_b200.sig_proc as dsp
    h = dsp.bpf(800., 1300., 48000, 1001)
    cv = dsp.convolver(h, np.float32)
    rng = np.random.default_rng(1)
    x = rng.normal(size=5000).astype(np.float32)
    y = np.concatenate([cv.convolve_fast(x[:1024]), cv.convolve_fast(x[1024:])])
    ref = np.convolve(np.concatenate((np.zeros(1000), x.astype(np.float64))), h.astype(np.float64), mode='valid')
    assert_parity(y, ref, "convolver")


# ------------------------------------------------------------------------------------------------ IIR
def _run_lfilter(b, a, x, cuts, mode=0, n_ch=1):
    import ctypes
    from pysdr_b200 import _lib
    lib = _lib.load()
    n = x.shape[-1]
    order = max(len(a), len(b)) - 1
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda().reshape(n_ch, n)
    yd = torch.empty_like(xd)
    zi = torch.zeros((n_ch, order), dtype=torch.float64, device="cuda")
    bb = np.ascontiguousarray(b, np.float64)
    aa = np.ascontiguousarray(a, np.float64)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pysdr_lfilter_set_mode(mode))
    try:
        for lo, hi in zip(cuts[:-1], cuts[1:]):                      # chunked with carried zi (sigs/iir.py:90-105)
            _lib.check(lib.pysdr_lfilter(bb.ctypes.data_as(ctypes.c_void_p), len(bb), aa.ctypes.data_as(ctypes.c_void_p), len(aa),
                                         ctypes.c_void_p(xd[:, lo:].data_ptr()), ctypes.c_void_p(yd[:, lo:].data_ptr()), hi - lo,
                                         n_ch, n, ctypes.c_void_p(zi.data_ptr()), st))
    finally:
        lib.pysdr_lfilter_set_mode(0)
    return yd.cpu().numpy(), zi.cpu().numpy()


def _iir_test_signal(n=20000):
    rng = np.random.default_rng(5)
    t = np.arange(n) / 1000.
    return (np.sin(2 * np.pi * 15 * t) + np.sin(2 * np.pi * 50 * t) + rng.normal(0, .1, n)).astype(np.float32)


@pytest.mark.parametrize("name", ["notch50", "butter3", "squelch_lo", "squelch_hi", "onepole"])
def test_lfilter_block_scan_matches_scipy(name):
    """Well-conditioned designs take the block-parallel scan (forced, mode 1) and match scipy within 1e-4."""
    from scipy import signal
    designs = dict(odsp.iir_designs())
    (designs["squelch_lo"], designs["squelch_hi"]) = odsp.squelch_designs(48000)
    designs["onepole"] = ([0.001], [1, 0.001 - 1])                  # squelch.m:125-128 / agc.m:6-12 form
    b, a = designs[name]
    x = _iir_test_signal()
    n = len(x)
    order = max(len(a), len(b)) - 1
    ref, zref = signal.lfilter(b, a, x.astype(np.float64), zi=np.zeros(order))
    for cuts in ([0, n], [0, 7000, 7001, 12345, n]):
        y, zi = _run_lfilter(b, a, x, cuts, mode=1)
        assert_parity(y[0], ref, "lfilter scan " + name)
        np.testing.assert_allclose(zi[0], zref, rtol=1e-6, atol=1e-9 * max(1.0, np.max(np.abs(zref))))
    # two channels at once, different data
    x2 = np.stack((x, x[::-1].copy()))
    y2, _ = _run_lfilter(b, a, x2, [0, 5000, n], mode=1, n_ch=2)
    assert_parity(y2[1], signal.lfilter(b, a, x2[1].astype(np.float64)), "lfilter scan ch1 " + name)


@pytest.mark.parametrize("name", ["notch50", "ellip7_lp", "butter3", "cheby2_band"])
def test_lfilter_auto_mode_tracks_scipy_on_all_reference_designs(name):
    """All four designs of reference sigs/iir.py.  The narrow-band high-order direct forms (ellip-7, cheby2-15)
    are ill-conditioned — scipy's own float64 result is rounding-dominated — so the library evaluates them
    sequentially in scipy's operation order and is compared tightly against scipy itself."""
    from scipy import signal
    b, a = odsp.iir_designs()[name]
    x = _iir_test_signal()
    n = len(x)
    order = max(len(a), len(b)) - 1
    ref, zref = signal.lfilter(b, a, x.astype(np.float64), zi=np.zeros(order))
    y, zi = _run_lfilter(b, a, x, [0, 7000, 7001, 12345, n], mode=0)
    ok = np.isfinite(ref)
    assert ok.all() or name == "cheby2_band"
    assert_parity(y[0][ok], ref[ok], "lfilter auto " + name)


# ------------------------------------------------------------------------------------------------ PSD
@pytest.mark.parametrize("chunk,nfft,overlap", [(4096, 8192, 0.5), (8192, 8192, 0.5), (1024, 2048, 0.5), (512, 512, 0.0)])
def test_spectrum_parity(chunk, nfft, overlap):
    import pysdr_b200.sig_proc as dsp
    rng = np.random.default_rng(6)
    n = chunk * 12
    tt = np.arange(n)
    x = (0.3 * np.exp(2j * np.pi * 0.123 * tt) + 0.01 * (rng.normal(size=n) + 1j * rng.normal(size=n))).astype(np.complex64)
    sp = dsp.spectrum(48., chunk, nfft, overlap)
    so = odsp.spectrum(48., chunk, nfft, overlap)
    assert (sp.NFFT, sp.new_samps, sp.chunk_size, sp.df) == (so.NFFT, so.new_samps, so.chunk_size, so.df)
    np.testing.assert_array_equal(sp.frq, so.frq)
    # linear power: strict gate; dB: compare where the oracle is within 100 dB of its peak
    est, ref = sp.psd_est(x, False), so.psd_est(x, False)
    assert_parity(est, ref, "psd_est linear")
    est_db, ref_db = sp.psd_est(x, True), so.psd_est(x, True)
    top = ref_db > ref_db.max() - 60                               # dB is a relative measure: gate it within 60 dB of the peak
    assert np.max(np.abs(est_db - ref_db)[top]) < 1e-3
    wf, wref = sp.waterfall(x, 4, False), so.waterfall(x, 4, False)
    assert wf.shape == wref.shape
    assert_parity(wf, wref, "waterfall lines")
    # streaming periodogram protocol (gui.py:1289-1313): new_samps per call
    ns = sp.new_samps
    for k in range(5):
        g, r = sp.periodogram(x[k * ns:(k + 1) * ns], True), so.periodogram(x[k * ns:(k + 1) * ns], True)
    top = r > r.max() - 60
    assert np.max(np.abs(g - r)[top]) < 2e-3
    assert_parity(10 ** (g / 10), 10 ** (r / 10), "streaming periodogram (linear)")
    assert len(sp.periodogram(np.zeros(0), True)) == 0


def test_real_input_psd_and_af_panel_geometry():
    import pysdr_b200.sig_proc as dsp
    rng = np.random.default_rng(7)
    x = rng.normal(size=4096 * 6).astype(np.float32)
    sp = dsp.spectrum(48., 4 * 1024, 8 * 1024, 0.5)                # AF panel, gui.py:618-621
    so = odsp.spectrum(48., 4 * 1024, 8 * 1024, 0.5)
    assert_parity(sp.psd_est(x, False), so.psd_est(x, False), "AF psd")


def test_waterfall_push_matches_plotting_py():
    import ctypes
    from pysdr_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(8)
    nfft, ncols, df = 1024, 100, 0.5
    ws = odsp.waterfall_state(nfft, df, ncols, pan_dr=60.0, peak_dist=4.0)
    wf = torch.full((nfft, ncols), -1e38, dtype=torch.float32, device="cuda")
    img = torch.empty((nfft, ncols), dtype=torch.float32, device="cuda")
    bk = torch.empty(1, dtype=torch.float32, device="cuda")
    scratch = torch.empty(nfft * ncols + nfft + 8, dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cnt = 0
    for it in range(6):
        PSD = (rng.normal(size=nfft) * 3 - 80).astype(np.float32)
        PSD[300] += 40
        fc = 0 if it < 3 else 2.0
        roll = int(float(fc - ws.wf_fc) / df + 0.5)                # Plotting.py:690-691 (push() applies it)
        img_ref, bk_ref, _ = ws.push(PSD.astype(np.float64), fc)
        cnt = min(cnt + 1, ncols)
        line = torch.from_numpy(PSD).cuda()
        _lib.check(lib.pysdr_waterfall_push(ctypes.c_void_p(wf.data_ptr()), nfft, ncols, cnt, ctypes.c_void_p(line.data_ptr()),
                                            nfft, int(roll), 60.0, ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(bk.data_ptr()),
                                            ctypes.c_void_p(scratch.data_ptr()), st))
        np.testing.assert_allclose(wf.cpu().numpy(), ws.wf.astype(np.float32), rtol=0, atol=0)
        assert abs(bk.item() - bk_ref) <= 1e-5 * abs(bk_ref)
        np.testing.assert_allclose(img.cpu().numpy(), img_ref, rtol=1e-5, atol=1e-3)
