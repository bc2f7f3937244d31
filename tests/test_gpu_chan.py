"""Many-channel tensor-core K1 (pysdr_b200/csrc/k1_chan.cu): the NCO mix + polyphase decimation (reference receiver.py:235,822,866)
of a whole bank of channel receivers as a dense split-TF32 contraction on tcgen05 — BASELINE config 5 with ANY set of channel
offsets.  Checked against the oracle's resampler and against the FP32 tap-stationary kernel on the same inputs.  Mode 2 forces
the tensor-core kernel on every call that has a few interior super-periods; mode 0 pins the FP32 kernel."""
import numpy as np
import pytest
import torch

from oracle import receiver_oracle as rxo
from oracle import sig_proc_oracle as odsp
from tests.util import assert_parity, make_both

pytestmark = pytest.mark.gpu


def _sig(n, srate, offs, seed, amp=0.02):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = ((rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.01).astype(np.complex128)
    for k, f in enumerate(offs):
        x += amp * (1 + 0.4 * np.sin(2 * np.pi * (300.0 + 23 * k) * t / srate)) * np.exp(2j * np.pi * (f + 500.0) * t / srate)
    return x.astype(np.complex64)


def _offsets(n_ch, span_hz, seed):
    """Irregular offsets (no raster): what wola.cu cannot serve."""
    rng = np.random.default_rng(seed)
    return sorted((rng.uniform(-span_hz, span_hz, n_ch)).round(1).tolist())


def _oracle_iq(Po, off, x):
    dec = odsp.decimator(Po.SRATE, Po.UP, Po.DOWN, Po.FILT_LEN, odsp.VIDEO_BWs, Po.VIDEO_BW)
    dec.h = dec.filter_bank[odsp._video_index(Po)]
    lo = odsp.signal_generator(off, Po.IN_CHUNK_SIZE, Po.SRATE, True)
    return dec.resamp_fast(x, lo)


def _channel_bank(P, offs, max_in, mode, group):
    from pysdr_b200.channelizer import ChannelBank
    cb = ChannelBank(P, offs, 'IQ', max_in=max_in, group=group)
    for b in cb.banks:
        b.set_k1_mma(mode)
    return cb


@pytest.mark.parametrize("n_ch,shift", [(24, 0), (24, 1), (96, 0), (128, 1), (100, 0), (16, 0)])
def test_chan_k1_cfg5_geometry_matches_oracle_and_fp32_kernel(n_ch, shift):
    """10 MS/s -> 48 kHz (3/625: DOWN odd, so rows of a class are two super-periods apart and there are six classes), irregular
    offsets, 4 chunks in one call.  96 channels = one column group of N = 192, 128 = two groups of 64, 100 = two groups of 56
    (50 live), 16/24 = narrow tiles.  shift = 1: the capture starts at an odd sample of its allocation, so every class's 16-byte
    alignment falls the other way (the other tap image)."""
    P, Po = make_both(10, [7000], ['IQ'])
    assert (P.UP, P.DOWN) == (3, 625)
    C, k = P.IN_CHUNK_SIZE, 4
    offs = _offsets(n_ch, 2.0e6, 100 + n_ch)
    xh = _sig(k * C + 1, P.SRATE, offs[::7], 5 + n_ch)
    x = xh[shift:shift + k * C]
    xd = torch.from_numpy(xh).cuda()[shift:shift + k * C]
    cb = _channel_bank(P, offs, k * C, 2, 128)
    _, iq = cb.process(xd)
    assert all(b.k1_last == 3 for b in cb.banks), [b.k1_last for b in cb.banks]
    got = [v.cpu().numpy().copy() for v in iq]
    cb0 = _channel_bank(P, offs, k * C, 0, 128)
    _, iq0 = cb0.process(xd)
    assert all(b.k1_last == 1 for b in cb0.banks)
    for c in range(n_ch):
        assert_parity(got[c], iq0[c].cpu().numpy(), "chan K1 vs fp32 K1, channel %d" % c, rel_tol=2e-5, snr_min=90)
    for c in sorted({0, 1, n_ch // 2, n_ch - 9, n_ch - 1}):
        assert_parity(got[c], _oracle_iq(Po, offs[c], x), "chan K1 vs oracle, channel %d" % c)


@pytest.mark.parametrize("srate_mhz,updown,span_hz", [(8, (3, 500), 1.5e6), (2.048, (3, 128), 0.4e6)])
def test_chan_k1_even_down_geometry(srate_mhz, updown, span_hz):
    """DOWN even (one parity, three classes), 32 channels, 5 chunks: 8 MS/s -> 48 kHz (3/500), and 2.048 MS/s -> 48 kHz (3/128),
    where the 334-tap window is longer than two super-periods — k1_mma.cu's row-per-super-period layout does not cover that, a
    row per output instant does."""
    P, Po = make_both(srate_mhz, [7000], ['IQ'])
    assert (P.UP, P.DOWN) == updown
    C, k, n_ch = P.IN_CHUNK_SIZE, 5, 32
    offs = _offsets(n_ch, span_hz, 9)
    x = _sig(k * C, P.SRATE, offs[::5], 17)
    xd = torch.from_numpy(x).cuda()
    cb = _channel_bank(P, offs, k * C, 2, 128)
    _, iq = cb.process(xd)
    assert cb.banks[0].k1_last == 3
    got = [v.cpu().numpy().copy() for v in iq]
    cb0 = _channel_bank(P, offs, k * C, 0, 128)
    _, iq0 = cb0.process(xd)
    for c in range(n_ch):
        assert_parity(got[c], iq0[c].cpu().numpy(), "chan K1 vs fp32 K1 (%d/%d), channel %d" % (updown + (c,)), rel_tol=2e-5, snr_min=90)
    for c in (0, 13, 31):
        assert_parity(got[c], _oracle_iq(Po, offs[c], x), "chan K1 vs oracle (%d/%d), channel %d" % (updown + (c,)))


def test_chan_k1_streaming_calls_full_chain():
    """Two consecutive calls (3 + 2 chunks: the second call's first outputs read the carried raw history on the edge warp)
    through the whole AM/NFM/USB/CW/LSB chain of 20 channels in ONE bank on the tensor-core kernel: every channel against its own
    oracle receiver, chunk at a time."""
    from pysdr_b200.channelizer import ChannelBank
    P, Po = make_both(10, [7000], ['USB'], af_bw_khz=[2])
    n_ch, C = 20, P.IN_CHUNK_SIZE
    offs = _offsets(n_ch, 1.0e6, 3)
    modes = [['AM', 'NFM', 'USB', 'CW', 'LSB'][k % 5] for k in range(n_ch)]
    afs = [[5e3, 10e3, 2e3, 500., 3e3][k % 5] for k in range(n_ch)]
    n = np.arange(5 * C)
    rng = np.random.default_rng(78)
    x = ((rng.normal(size=len(n)) + 1j * rng.normal(size=len(n))) * 0.01).astype(np.complex128)
    for k, f in enumerate(offs):
        x = x + 0.02 * (1 + 0.5 * np.sin(2 * np.pi * (300.0 + 40 * k) * n / P.SRATE)) * np.exp(2j * np.pi * (f + 700.0) * n / P.SRATE)
    x = x.astype(np.complex64)
    cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=3 * C, group=32)
    cb.banks[0].set_k1_mma(2)
    xd = torch.from_numpy(x).cuda()
    outs = []
    for a, b in ((0, 3 * C), (3 * C, 5 * C)):
        am, _ = cb.process(xd[a:b])
        assert cb.banks[0].k1_last == 3
        outs.append([v.cpu().numpy().copy() for v in am])
    for k in range(n_ch):
        Pk = rxo.make_P(P.SRATE, [7000e3], modes[k], foffset=100e3, af_bw=afs[k], bfo=700.0)
        orx = odsp.Receiver(Pk, offs[k], 0, str(k), fast=True)
        ref = np.concatenate([np.array(orx.demod_data(x[c * C:(c + 1) * C])) for c in range(5)])
        assert_parity(np.concatenate([o[k] for o in outs]), ref, "channel %d (%s)" % (k, modes[k]))


def test_chan_k1_impulse_indexing():
    """A unit impulse at a known sample: every tap of the polyphase response lands at the same output index as with the FP32
    kernel, for an impulse in each of the six classes' windows."""
    P, _ = make_both(10, [7000], ['IQ'])
    C = P.IN_CHUNK_SIZE
    n = 3 * C
    offs = _offsets(16, 1.0e6, 21)
    for pos in (70001, 2 * C - 7, C + 625 * 11 + 208, C + 625 * 12 + 416):
        x = torch.zeros(n, dtype=torch.complex64, device="cuda")
        x[pos] = 1.0 + 0.5j
        outs = []
        for mode in (2, 0):
            cb = _channel_bank(P, offs, n, mode, 128)
            _, iq = cb.process(x)
            assert cb.banks[0].k1_last == (3 if mode else 1)
            outs.append(torch.stack(list(iq)).cpu().numpy().copy())
        for c in (0, 7, 15):
            nz2, nz0 = np.nonzero(outs[0][c])[0], np.nonzero(outs[1][c])[0]
            assert nz0.size > 0 and nz2.min() == nz0.min() and nz2.max() == nz0.max(), (pos, c, nz2.min(), nz0.min(), nz2.max(), nz0.max())
            assert np.max(np.abs(outs[0][c] - outs[1][c])) <= 2e-6 * np.max(np.abs(outs[1][c]))


def test_chan_k1_many_tiles_per_cta_is_stable_and_reproducible():
    """One 4 s block of config 5 (188 chunks) through a bank of 96 channels: 1506 tiles on 148 persistent CTAs, the sample ring
    wraps ~50 times and the tap ring ~70 times per CTA.  Three runs must agree BIT FOR BIT (one issuing thread, fixed summation
    order) and match the FP32 kernel."""
    P, _ = make_both(10, [7000], ['IQ'])
    C = P.IN_CHUNK_SIZE
    n = 188 * C
    offs = _offsets(96, 2.0e6, 33)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda", dtype=torch.float32) * 0.05)
    cb0 = _channel_bank(P, offs, n, 0, 128)
    _, iq0 = cb0.process(x)
    ref = torch.stack([iq0[c] for c in (0, 40, 95)]).clone()
    del cb0, iq0
    cb = _channel_bank(P, offs, n, 1, 128)                # default mode: a call this long takes the tensor-core kernel by itself
    first = None
    for rep in range(3):
        cb.banks[0].reset()
        _, iq = cb.process(x)
        assert cb.banks[0].k1_last == 3
        got = torch.stack([iq[c] for c in (0, 40, 95)])
        e = ((got - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)).max().item()
        assert e < 2e-5, (rep, e)
        if first is None:
            first = got.clone()
        else:
            assert torch.equal(torch.view_as_real(got), torch.view_as_real(first)), "run %d differs from run 0" % rep


def test_chan_k1_retune_mid_stream_rebuilds_the_tap_images():
    """set_freq between two calls: the folded taps of the retuned channels change, so the host rebuilds the class images
    (k1_chan_upload_taps behind g_dirty); both calls and every channel equal the FP32 kernel driven the same way (whose retune
    is checked against the oracle in test_gpu_edges.py::test_retune_and_filter_swap_mid_batch_stream)."""
    P, _ = make_both(10, [7000], ['IQ'])
    C, n_ch = P.IN_CHUNK_SIZE, 16
    offs = _offsets(n_ch, 1.5e6, 77)
    new = {3: offs[3] + 12345.0, 11: -offs[11]}
    x = _sig(6 * C, P.SRATE, [offs[3], new[3], offs[11], new[11]], 19)
    xd = torch.from_numpy(x).cuda()
    outs = []
    for mode in (2, 0):
        cb = _channel_bank(P, offs, 3 * C, mode, 128)
        b = cb.banks[0]
        _, iq = cb.process(xd[:3 * C])
        first = torch.stack(list(iq)).cpu().numpy().copy()
        for r, f in new.items():
            b.set_freq(r, f)
        _, iq = cb.process(xd[3 * C:])
        assert b.k1_last == (3 if mode else 1)
        outs.append((first, torch.stack(list(iq)).cpu().numpy().copy()))
    for c in range(n_ch):
        for part in (0, 1):
            assert_parity(outs[0][part][c], outs[1][part][c], "retune, call %d, channel %d: chan vs fp32 K1" % (part, c), rel_tol=2e-5, snr_min=90)
    # the retune did something: channel 3 moved onto the second test carrier (x holds one at offs[3] + 500 Hz and one at
    # new[3] + 500 Hz), so its second call still sees a strong in-band tone
    p1 = np.mean(np.abs(outs[0][1][3]) ** 2)
    assert p1 > 1e-5, p1


def test_chan_k1_time_shards_emulated_on_one_gpu():
    """Two time shards of a 20-channel bank played one after the other on ONE device (the 2-GPU version is test_gpu_multi.py):
    each shard's K1 call starts at a non-zero absolute position with its filter history in place in front of x (the edge warp
    reads it there) and one warm-up chunk; sharded audio through the O(1) AGC carry equals the single-stream FP32 run."""
    import ctypes
    from pysdr_b200._lib import check
    from pysdr_b200.bank import _stream_ptr
    from pysdr_b200.channelizer import ChannelBank
    from pysdr_b200.dist import AGC_SUMMARY_LEN, ShardedCapture
    P, _ = make_both(10, [7000], ['USB'], af_bw_khz=[2])
    C, n_ch, world, cpr = P.IN_CHUNK_SIZE, 20, 2, 9
    offs = _offsets(n_ch, 1.0e6, 5)
    modes = [['AM', 'NFM', 'USB', 'CW', 'LSB'][k % 5] for k in range(n_ch)]
    afs = [[5e3, 10e3, 2e3, 500., 3e3][k % 5] for k in range(n_ch)]
    n = world * cpr * C
    x = _sig(n, P.SRATE, offs[::3], 23)
    env = np.ones(n, np.float32)
    env[4 * C:6 * C] = 4.0                                          # a burst in shard 0 whose AGC decay crosses into shard 1
    xd = torch.from_numpy(x * env).cuda()
    single = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=n, group=32)
    single.banks[0].set_k1_mma(0)
    am, _ = single.process(xd)
    ref = [a.cpu().numpy().copy() for a in am]
    shards, all_sum = [], torch.zeros((world, n_ch, AGC_SUMMARY_LEN), dtype=torch.float64, device="cuda")
    for r in range(world):
        cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=(cpr + 1) * C, group=32)
        b = cb.banks[0]
        b.set_k1_mma(2)
        sh = ShardedCapture(b, cb.banks[0].P, r, world, cpr)
        assert sh.o1
        pl = sh.plan
        sh.front(xd[pl['first_sample']:pl['start'] + pl['n']], copy_own=False)
        assert b.k1_last == 3
        check(b.lib.pysdr_bank_agc_summary(b.h, pl['warm_chunks'], ctypes.c_void_p(all_sum[r].data_ptr()), _stream_ptr()))
        shards.append(sh)
    for r in range(world):
        sh, b = shards[r], shards[r].bank
        got, _, _ = b.process_back_carry(all_sum, r, want_dc=False, skip_blocks=sh.plan['warm_chunks'])
        m0 = odsp.n_out_total(r * cpr * C, P.UP, P.DOWN)
        for k in range(n_ch):
            g = got[k][sh.skip_out:].cpu().numpy()
            # tensor-core K1 on the sharded side, FP32 K1 on the single stream: the north-star gate (same-kernel runs agree to 2e-5)
            assert_parity(g, ref[k][m0:m0 + len(g)], "shard %d channel %d (%s) vs single stream" % (r, k, modes[k]))
