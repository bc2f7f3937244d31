"""Round-2 robustness cases: even AF filter lengths through the multi-block FFT path (16-byte aligned staging), AGC
read-outs, K1 on streams cut into pieces shorter than one tile (every tile an edge tile), two devices in one process."""
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


def _sig(n, P, offs, seed):
    t = np.arange(n)
    x = _noise(n, seed, 0.01).astype(np.complex128)
    for k, f in enumerate(offs):
        x += 0.05 * (1 + 0.4 * np.sin(2 * np.pi * (400.0 + 90 * k) * t / P.SRATE)) * np.exp(2j * np.pi * (f + 300.0) * t / P.SRATE)
    return x.astype(np.complex64)


def _bank(P, max_in, device=None):
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.receiver import receiver_offsets
    return ReceiverBank(P, receiver_offsets(P), max_in=max_in, device=device)


@pytest.mark.parametrize("nfilt", [200, 100, 1000])
def test_even_af_length_many_blocks_per_call(nfilt):
    """-nfilt even => V = N-(L-1) odd: every other overlap-save block starts at an odd sample of the complex memory.  With
    >= 8 chunks per call the FFT path runs several blocks per receiver (misaligned 16-byte cp.async)."""
    from pysdr_b200.receiver import receiver_offsets
    P, Po = make_both(2.048, [1000, 1020, 1045, 990], ['AM', 'NFM', 'USB', 'CW'], af_bw_khz=[5, 10, 2, 0.5], nfilt=nfilt)
    C = P.IN_CHUNK_SIZE
    k = 12
    x = _sig(k * C, P, receiver_offsets(P), 77)
    bank = _bank(P, k * C)
    am, iq, _ = bank.process(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    rxo.create_receivers(Po)
    for r in range(4):
        ref = np.concatenate([Po.rx[r].demod_data(x[c * C:(c + 1) * C]) for c in range(k)])
        assert_parity(am[r].cpu().numpy(), ref, "nfilt %d rx%d" % (nfilt, r))


def test_agc_readouts_match_oracle():
    """rx.agc.{agc,gain,maxbuf,ref,err} (reference watchdog.py:298-302): agc = the gain the loop asks for, gain = applied."""
    from pysdr_b200 import sig_proc as dsp
    P, Po = make_both(2.048, [1000], ['AM'], af_bw_khz=[5], nfilt=301)
    C = P.IN_CHUNK_SIZE
    x = _sig(12 * C, P, [P.FOFFSET], 5)
    x[1 * C:2 * C] *= 4.0                                           # attack; 8 blocks later the burst leaves the peak buffer
    rx = dsp.Receiver(P, P.FOFFSET, 0, '1')
    orx = odsp.Receiver(Po, Po.FOFFSET, 0, '1')
    for c in range(12):
        rx.demod_data(x[c * C:(c + 1) * C])
        orx.demod_data(x[c * C:(c + 1) * C])
        for k in ('agc', 'gain', 'maxbuf', 'ref', 'err'):
            got, ref = getattr(rx.agc, k), getattr(orx.agc, k)
            scale = abs(orx.agc.gain) if k == 'err' else max(abs(ref), 1e-6)      # err = agc - gain: a difference
            assert abs(got - ref) <= 1e-4 * scale, (c, k, got, ref)
    assert rx.agc.agc != rx.agc.gain                                # decaying: the loop filter lags the wanted gain
    rx.agc.reset()
    assert rx.agc.gain == 1.0 and rx.agc.agc == 1.0 and rx.agc.maxbuf == 0.0


@pytest.mark.parametrize("geom", [(8, 1001), (2.048, 1001), (10, 301)])
def test_k1_every_tile_an_edge_tile(geom):
    """Calls far shorter than one K1 tile (a tile is ~9000 samples): every tile of every call is filled by produce()'s
    edge path — carried history before x[0], plain stores, zero fill past the end, odd alignments — never by the
    interior bulk-copy path.  Baseband must equal the one-call result BIT FOR BIT (K1's arithmetic does not depend on how
    the stream is cut) and match the oracle."""
    srate, nfilt = geom
    P, Po = make_both(srate, [1000, 1300, 870], ['IQ', 'IQ', 'IQ'], nfilt=nfilt)
    C = P.IN_CHUNK_SIZE
    x = _noise(C, 123)
    whole = _bank(P, C)
    _, iq_w, _ = whole.process(torch.from_numpy(x).cuda())
    iq_w = [v.cpu().numpy().copy() for v in iq_w]
    rng = np.random.default_rng(9)
    for trial in range(3):
        # one stream of C samples cut at random points: pieces of 1 .. 3000 samples, odd and even lengths
        cuts = [0]
        while cuts[-1] < C:
            cuts.append(min(C, cuts[-1] + int(rng.integers(1, 3000))))
        from pysdr_b200._lib import check
        import ctypes
        b = _bank(P, C)
        xd = torch.from_numpy(x).cuda()
        parts = [[] for _ in range(3)]
        for a0, a1 in zip(cuts[:-1], cuts[1:]):
            # K1 only, at arbitrary stream positions: the bank's k1_only mode moves the input memory along without the
            # block-aligned audio stages
            check(b.lib.pysdr_bank_set_k1_only(b.h, 1))
            n_out = ctypes.c_int64(0)
            check(b.lib.pysdr_bank_process_front(b.h, ctypes.c_void_p(xd[a0:a1].data_ptr()), a1 - a0, 0, None, b.max_out,
                                                 ctypes.c_void_p(b._am.data_ptr()), ctypes.byref(n_out),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.n_out = n_out.value
            for r in range(3):
                parts[r].append(b.iq_row(r).cpu().numpy().copy())
        for r in range(3):
            got = np.concatenate(parts[r])
            assert got.shape == iq_w[r].shape
            assert np.array_equal(got, iq_w[r]), "trial %d rx%d: cut stream differs from the one-call result" % (trial, r)
    rxo.create_receivers(Po)
    for r in range(3):
        Po.rx[r].demod_data(x)
        assert_parity(iq_w[r], Po.rx[r].iq, "iq rx%d" % r)


def test_two_devices_in_one_process():
    """Per-device one-time kernel set-up (cudaFuncSetAttribute is per device): a bank on cuda:1 after one on cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    P, Po = make_both(8, [1000, 1300], ['USB', 'AM'], af_bw_khz=[2, 5])
    C = P.IN_CHUNK_SIZE
    x = _noise(2 * C, 3)
    outs = []
    for d in (0, 1):
        with torch.cuda.device(d):
            b = _bank(P, 2 * C, device="cuda:%d" % d)
            am, iq, _ = b.process(torch.from_numpy(x).to("cuda:%d" % d))
            torch.cuda.synchronize()
            outs.append([v.cpu().numpy().copy() for v in am])
    for r in range(2):
        assert np.array_equal(outs[0][r], outs[1][r])


def test_o1_agc_carry_three_shards_emulated_on_one_gpu():
    """The O(1) AGC hand-over of time-sharded runs (pysdr_bank_agc_summary / pysdr_bank_agc_enter: 19 doubles per
    receiver per shard instead of every block peak), with three shards played one after the other on ONE device — the
    NCCL version of the same thing needs 2 GPUs (test_gpu_multi.py).  Sharded audio == single-stream audio."""
    import ctypes
    from pysdr_b200._lib import check
    from pysdr_b200.bank import _stream_ptr
    from pysdr_b200.dist import AGC_SUMMARY_LEN, ShardedCapture, agc_enter_reference
    from pysdr_b200.receiver import receiver_offsets
    world, cpr = 3, 9
    P, _ = make_both(2.048, [1000, 1020, 1045, 990], ['AM', 'NFM', 'USB', 'CW'], af_bw_khz=[5, 10, 2, 0.5], nfilt=301)
    C = P.IN_CHUNK_SIZE
    n = world * cpr * C
    x = _sig(n, P, receiver_offsets(P), 321)
    env = np.ones(n, np.float32)
    env[4 * C:6 * C] = 5.0                                          # a burst in shard 0 whose decay crosses into shard 1
    env[20 * C:] = 0.2                                              # a fade in shard 2
    xd = torch.from_numpy(x * env).cuda()
    single = _bank(P, n)
    am, _, _ = single.process(xd)
    ref = [a.cpu().numpy().copy() for a in am]
    pk_ref, gn_ref = single.agc_trace()
    shards, all_sum = [], torch.zeros((world, 4, AGC_SUMMARY_LEN), dtype=torch.float64, device="cuda")
    for r in range(world):
        b = _bank(P, (cpr + 1) * C)
        sh = ShardedCapture(b, P, r, world, cpr)
        assert sh.o1
        pl = sh.plan
        sh.front(xd[pl['first_sample']:pl['start'] + pl['n']], copy_own=False)
        check(b.lib.pysdr_bank_agc_summary(b.h, pl['warm_chunks'], ctypes.c_void_p(all_sum[r].data_ptr()), _stream_ptr()))
        shards.append(sh)
    torch.cuda.synchronize()
    sums = all_sum.cpu().numpy()
    for r in range(world):
        sh, b = shards[r], shards[r].bank
        check(b.lib.pysdr_bank_agc_enter(b.h, ctypes.c_void_p(all_sum.data_ptr()), r, _stream_ptr()))
        st = b.agc_get(0)
        g_host, ring_host, _ = agc_enter_reference(sums[:, 0], r)   # the host restatement of the same carry
        assert abs(st['gain'] - g_host) <= 1e-12 * g_host
        if r:
            assert abs(st['gain'] - gn_ref[0, r * cpr - 1]) <= 1e-6 * st['gain']          # = the single stream's gain there
        got, _, _ = sh.back(None)
        m0 = odsp.n_out_total(r * cpr * C, P.UP, P.DOWN)
        for k in range(4):
            g = got[k].cpu().numpy()
            assert_parity(g, ref[k][m0:m0 + len(g)], "shard %d rx%d vs single stream" % (r, k), rel_tol=2e-5, snr_min=90)


@pytest.mark.parametrize("modes", [['AM', 'NFM', 'USB', 'CW'], ['IQ', 'AM', 'LSB'], ['RTTY'], ['USB', 'IQ', 'AM', 'NFM', 'CW', 'LSB']])
def test_fused_back_kernel_equals_standalone_tail_kernels(modes):
    """agc_back_fused_kernel (block peaks + state update -> grid barrier -> scan -> grid barrier -> gain / DC removal, with
    seek() and the AGC restart folded in) against the stand-alone kernels it replaces: bit-identical audio, DC-removed
    audio, baseband, AGC state and carried memories, over whole-capture calls, chunked calls, re-seeks and a ragged tail."""
    n_rx = len(modes)
    P, _ = make_both(2.048, [1000 + 17 * k for k in range(n_rx)], modes, af_bw_khz=[2] * n_rx, nfilt=301)
    C = P.IN_CHUNK_SIZE
    from pysdr_b200.receiver import receiver_offsets
    x = torch.from_numpy(_sig(9 * C + 777, P, receiver_offsets(P), 11)).cuda()
    banks = [_bank(P, 9 * C + 777), _bank(P, 9 * C + 777)]
    banks[1].force_unfused(True)
    plans = [[(0, 9 * C + 777)],                                      # one ragged call
             [(0, 4 * C), (4 * C, 5 * C), (5 * C, 9 * C)],            # chunked, carried state between calls
             [(0, 2 * C), (2 * C, 9 * C + 777)]]
    for plan in plans:
        outs = []
        for b in banks:
            b.seek(0)                                                 # folded into the next call on the fused bank
            got = []
            for a0, a1 in plan:
                am, iq, dc = b.process(x[a0:a1], want_dc=True)
                got.append(([v.clone() for v in am], [v.clone() for v in iq], [v.clone() for v in dc]))
            got.append(b.agc_get(0))
            outs.append(got)
        for (amA, iqA, dcA), (amB, iqB, dcB) in zip(outs[0][:-1], outs[1][:-1]):
            for r in range(n_rx):
                assert torch.equal(amA[r], amB[r]) and torch.equal(iqA[r], iqB[r]) and torch.equal(dcA[r], dcB[r]), (plan, r)
        assert outs[0][-1] == outs[1][-1]
    for b in banks:                                                   # a seek to a later block: memories cleared, AGC kept
        b.seek(3 * C)
        am, _, _ = b.process(x[3 * C:6 * C])
        b._keep = [v.clone() for v in am]
    for r in range(n_rx):
        assert torch.equal(banks[0]._keep[r], banks[1]._keep[r])
    assert banks[0].get_state() == banks[1].get_state()


@pytest.mark.parametrize("nfft,chunk,overlap,cplx", [(8192, 8192, 0.5, True), (8192, 4096, 0.5, False), (4096, 4096, 0.5, True),
                                                      (2048, 1500, 0.25, True), (1024, 1024, 0.0, True), (512, 512, 0.75, False)])
def test_psd_fast_path_many_lines(nfft, chunk, overlap, cplx):
    """The waterfall of a long capture (>= 296 lines: one persistent CTA per line, psd_fast.cu: first radix-16 pass on the
    loaded registers, table twiddles, last radix fused with |X|^2) against numpy (reference Plotting.py:376-377,462 /
    sigs/iq.py:75-79: periodic Hann, |FFT|^2 / sum w^2, mean over the line's frames, 10 log10, fftshift) and against the
    generic kernel."""
    import os
    from pysdr_b200 import sig_proc as dsp
    navg, lines = 3, 320
    sp = dsp.spectrum(48., chunk, nfft, overlap)
    hop = sp.new_samps
    n = chunk + hop * (navg * lines - 1)
    rng = np.random.default_rng(nfft + chunk)
    t = np.arange(n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.05 + np.exp(2j * np.pi * 0.1234 * t) + 0.3 * np.exp(-2j * np.pi * 0.31 * t)
    if not cplx:
        x = x.real
    xd = torch.from_numpy(x.astype(np.complex64)).cuda()
    got = sp.waterfall(xd, navg, dB=False)
    assert got.shape == (lines, nfft)
    os.environ["PYSDR_PSD_GENERIC"] = "1"
    try:
        gen = sp.waterfall(xd, navg, dB=False)
    finally:
        del os.environ["PYSDR_PSD_GENERIC"]
    w = sp.win.astype(np.float64)
    xs = x.astype(np.complex64).astype(np.complex128)
    for line in (0, 1, lines // 2, lines - 1):
        ref = np.zeros(nfft)
        for f in range(navg):
            s0 = (line * navg + f) * hop
            ref += np.abs(np.fft.fft(xs[s0:s0 + chunk] * w, nfft)) ** 2
        ref = np.fft.fftshift(ref / (navg * np.sum(w * w)))
        assert_parity(got[line], ref.astype(np.float32), "fast psd line %d" % line)
        assert_parity(got[line], gen[line], "fast vs generic psd line %d" % line, rel_tol=2e-5, snr_min=90)


def _dev_find_peaks(x, height, distance):
    import ctypes
    from pysdr_b200 import _lib
    from pysdr_b200._lib import check
    lib = _lib.load()
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    idx = torch.zeros(4096, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(lib.pysdr_find_peaks(ctypes.c_void_p(xd.data_ptr()), xd.numel(), None, 0.0, float(height), float(distance),
                               ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(cnt.data_ptr()), None))
    torch.cuda.synchronize()
    return idx[:int(cnt.item())].cpu().numpy()


@pytest.mark.parametrize("seed,n,dist", [(1, 8192, 41.7), (2, 2048, 3.0), (3, 8192, 1.0), (4, 500, 120.2), (5, 16384, 17.0)])
def test_device_peak_picker_equals_scipy_find_peaks(seed, n, dist):
    """pysdr_find_peaks against scipy.signal.find_peaks(x, distance=, height=) — the call of reference Plotting.py:594 — on
    PSD-like lines: noise floor + carriers, flat-topped plateaus (clipped display lines) and distinct heights (with equal
    heights scipy's own argsort order is unspecified, so the fixture avoids exact ties between different peaks)."""
    from scipy import signal
    rng = np.random.default_rng(seed)
    x = (-100.0 + 6.0 * rng.normal(size=n)).astype(np.float32)
    for k in range(25):
        c = int(rng.integers(5, n - 5))
        x[max(0, c - 3):c + 4] += np.float32(20.0 + 30.0 * rng.random()) * np.hanning(9)[1:8].astype(np.float32)[:len(x[max(0, c - 3):c + 4])]
    for k in range(6):                                        # plateaus of width 2..5
        c = int(rng.integers(10, n - 10))
        x[c:c + 2 + k % 4] = np.float32(-40.0 + k)
    h = float(np.median(x) + 10.0)
    ref, _ = signal.find_peaks(x, distance=dist, height=h)
    got = _dev_find_peaks(x, h, dist)
    np.testing.assert_array_equal(got, ref)


def test_psd_more_than_65535_lines_in_one_call():
    """A whole-capture call may ask for more lines than one launch's grid covers (r01: a silent capacity cliff): 70 000 lines
    of a 256-point spectrum come back in batches and equal the same lines computed piecewise."""
    from pysdr_b200 import sig_proc as dsp
    sp = dsp.spectrum(48., 256, 256, 0.0)
    n = 256 * 70000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda", dtype=torch.float32))
    full = sp.waterfall(x, 1, dB=True, to_host=False)
    assert full.shape == (70000, 256)
    for l0 in (0, 65534, 65535, 69990):
        part = sp.waterfall(x[l0 * 256:(l0 + 5) * 256], 1, dB=True, to_host=False)
        assert torch.allclose(part, full[l0:l0 + 5], rtol=0, atol=1e-4)


def test_am_synch_time_shards_after_loop_settling_warmup():
    """AM-Synch under time sharding (r01 refused it): the carrier loop has no closed-form hand-off but it forgets its start-up
    state — a shard warmed up over the loop's settling time (dist.pll_settle_chunks: 20 chunks for rel 1e-6) tracks the
    single-stream loop to the parity gate.  Two shards of 32 chunks played on one device (the O(1) AGC summaries carried as in
    the multi-GPU run); carrier 11 Hz off the receiver centre, inside the 50 Hz loop's pull-in range."""
    import ctypes
    from pysdr_b200._lib import check
    from pysdr_b200.bank import _stream_ptr
    from pysdr_b200.dist import AGC_SUMMARY_LEN, ShardedCapture, pll_settle_chunks
    from pysdr_b200.receiver import receiver_offsets
    world, cpr = 2, 32
    P, _ = make_both(2.048, [1000, 1020], ['AM-Synch', 'AM'], foffset_khz=100, af_bw_khz=[5, 5])
    C = P.IN_CHUNK_SIZE
    warm = pll_settle_chunks(P)
    assert 10 <= warm < cpr
    n = world * cpr * C
    t = np.arange(n)
    x = _noise(n, 33, 0.002).astype(np.complex128)
    for k, off in enumerate(receiver_offsets(P)):
        env = 1.0 + 0.5 * np.sin(2 * np.pi * (700.0 + 300 * k) * t / P.SRATE)
        x = x + 0.1 * env * np.exp(2j * np.pi * (off + 11.0) * t / P.SRATE)
    xd = torch.from_numpy(x.astype(np.complex64)).cuda()
    single = _bank(P, n)
    am, _, _ = single.process(xd)
    ref = [a.cpu().numpy().copy() for a in am]
    all_sum = torch.zeros((world, 2, AGC_SUMMARY_LEN), dtype=torch.float64, device="cuda")
    shards = []
    for r in range(world):
        b = _bank(P, (cpr + warm) * C)
        sh = ShardedCapture(b, P, r, world, cpr)
        pl = sh.plan
        assert pl['warm_chunks'] == (warm if r else 0)
        sh.front(xd[pl['first_sample']:pl['start'] + pl['n']], copy_own=False)
        check(b.lib.pysdr_bank_agc_summary(b.h, pl['warm_chunks'], ctypes.c_void_p(all_sum[r].data_ptr()), _stream_ptr()))
        shards.append(sh)
    for r in range(world):
        sh, b = shards[r], shards[r].bank
        check(b.lib.pysdr_bank_agc_enter(b.h, ctypes.c_void_p(all_sum.data_ptr()), r, _stream_ptr()))
        got, _, _ = sh.back(None)
        m0 = odsp.n_out_total(r * cpr * C, P.UP, P.DOWN)
        for k in range(2):
            g = got[k].cpu().numpy()
            assert_parity(g, ref[k][m0:m0 + len(g)], "AM-Synch shard %d rx%d vs single stream" % (r, k))
