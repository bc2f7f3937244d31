"""Full-size checks (BASELINE.json configs at their real sizes) through size-independent properties: total output
count, determinism, exact linearity under power-of-two scaling, segments ≡ whole, spot parity against the oracle deep
inside the capture, and Parseval for the panadapter lines.  The oracle only ever sees a few chunks."""
import numpy as np
import pytest
import torch

from oracle import receiver_oracle as rxo
from oracle import sig_proc_oracle as odsp
from tests.util import assert_parity, make_both

pytestmark = pytest.mark.gpu

FCS = [-500, 700, 1400, 3100]
MODES = ['AM', 'NFM', 'USB', 'CW']
N_CHUNKS = 2812                                            # 60 s of 8 MS/s in whole IN_CHUNK_SIZE blocks


@pytest.fixture(scope="module")
def cfg2():
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.receiver import receiver_offsets
    from pysdr_b200.synth import synth_iq
    P, Po = make_both(8, FCS, MODES, af_bw_khz=[5, 10, 2, .5])
    n = N_CHUNKS * P.IN_CHUNK_SIZE
    offs = receiver_offsets(P)
    x = synth_iq(n, P.SRATE, offs, MODES, seed=31, device="cuda")
    bank = ReceiverBank(P, offs, max_in=n)
    am, iq, _ = bank.process(x, want_dc=False)
    assert bank.k1_last == 2, "the whole-capture call should run the tensor-core K1"
    out = dict(P=P, Po=Po, x=x, n=n, offs=offs, n_out=bank.n_out, am=[a.clone() for a in am], iq=[q.clone() for q in iq],
               trace=bank.agc_trace())
    del bank
    return out


def test_cfg2_full_size_output_count_and_determinism(cfg2):
    from pysdr_b200.bank import ReceiverBank
    P = cfg2['P']
    assert cfg2['n'] == 479912792 and cfg2['n_out'] == odsp.n_out_total(cfg2['n'], P.UP, P.DOWN) == 2879477
    bank = ReceiverBank(P, cfg2['offs'], max_in=cfg2['n'])
    am, iq, _ = bank.process(cfg2['x'], want_dc=False)
    for r in range(4):                                     # a second run is bit-identical (no atomics, fixed reduction order)
        assert torch.equal(am[r], cfg2['am'][r]) and torch.equal(iq[r], cfg2['iq'][r])
        assert torch.isfinite(am[r]).all()


def test_cfg2_full_size_k1_is_exactly_linear_under_power_of_two_scaling(cfg2):
    from pysdr_b200.bank import ReceiverBank
    P = cfg2['P']
    bank = ReceiverBank(P, cfg2['offs'], max_in=cfg2['n'])
    x4 = cfg2['x'] * 4.0
    _, iq, _ = bank.process(x4, want_dc=False)
    for r in range(4):
        assert torch.equal(iq[r], cfg2['iq'][r] * 4.0)     # scaling by 2^k commutes with every rounding in K1


@pytest.mark.parametrize("k1_mma", [0, 1])
def test_cfg2_full_size_segments_equal_whole(cfg2, k1_mma):
    """k1_mma = 0 (tap-stationary FP32 K1 pinned): the baseband of uneven segments equals the whole-capture call BIT FOR BIT
    (the per-output arithmetic does not depend on how the stream is cut).  k1_mma = 1 (default: tensor-core K1 on calls of
    >= 8192 interior super-periods, so the long segments take it and the single-chunk segment does not): the two K1 kernels
    and the two row alignments round differently, agreement is 5e-6 of peak (split-TF32 products, fp32 accumulation)."""
    from pysdr_b200.bank import ReceiverBank
    P = cfg2['P']
    C = P.IN_CHUNK_SIZE
    cuts = [0, 700, 701, 1999, N_CHUNKS]                   # uneven segments, one of them a single chunk
    if k1_mma == 0:
        whole = ReceiverBank(P, cfg2['offs'], max_in=cfg2['n'])
        whole.set_k1_mma(0)
        w_am, w_iq, _ = whole.process(cfg2['x'], want_dc=False)
        assert whole.k1_last == 1
        ref_am, ref_iq = [a.clone() for a in w_am], [q.clone() for q in w_iq]
        del whole
    else:
        ref_am, ref_iq = cfg2['am'], cfg2['iq']
    bank = ReceiverBank(P, cfg2['offs'], max_in=1300 * C)
    bank.set_k1_mma(k1_mma)
    pos = 0
    used = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        am, iq, _ = bank.process(cfg2['x'][a * C:b * C], want_dc=False)
        used.append(bank.k1_last)
        k = bank.n_out
        for r in range(4):
            if k1_mma == 0:
                assert torch.equal(iq[r], ref_iq[r][pos:pos + k])                        # K1: bit-exact
            else:
                e = (iq[r] - ref_iq[r][pos:pos + k]).abs().max().item() / ref_iq[r].abs().max().item()
                assert e < 5e-6, (r, a, b, e)
            ref = ref_am[r][pos:pos + k]
            err = (am[r] - ref).abs().max().item() / ref.abs().max().item()
            assert err < 2e-5, (r, a, b, err)                                            # K2: FFT block alignment differs
        pos += k
    assert pos == cfg2['n_out']
    assert used == ([1, 1, 1, 1] if k1_mma == 0 else [2, 1, 2, 2]), used


@pytest.mark.parametrize("blk", [1, 1406, 2805])
def test_cfg2_full_size_spot_parity_with_oracle(cfg2, blk):
    """Baseband IQ of 3 chunks starting at block `blk`, against an oracle that is started 2 chunks earlier with its
    sample counter and LO phase advanced to that point (K1 only needs lp-1 samples of history)."""
    P, Po = cfg2['P'], cfg2['Po']
    C = P.IN_CHUNK_SIZE
    warm = 1
    s0 = (blk - warm) * C
    xs = cfg2['x'][s0:(blk + 3) * C].cpu().numpy()
    m_lo = odsp.n_out_total(blk * C, P.UP, P.DOWN)
    m_hi = odsp.n_out_total((blk + 3) * C, P.UP, P.DOWN)
    for r in range(4):
        orx = odsp.Receiver(Po, cfg2['offs'][r], r, str(r), fast=True)
        orx.lo.advance(s0)
        orx.dec.n0 = s0
        got = []
        for c in range(warm + 3):
            orx.demod_data(xs[c * C:(c + 1) * C])
            if c >= warm:
                got.append(orx.iq.copy())
        ref = np.concatenate(got)
        assert len(ref) == m_hi - m_lo
        assert_parity(cfg2['iq'][r][m_lo:m_hi].cpu().numpy(), ref, "iq rx%d blocks %d..%d" % (r, blk, blk + 3))


def test_cfg2_full_size_agc_trajectory_and_audio(cfg2):
    """The AUDIO of the full 60 s capture against the oracle (r01 only spot-checked the baseband):
      (1) the oracle's AGC law run over all 2812 blocks on the device's own block peaks reproduces the device's gain
          trajectory (the float64 ordered scan vs the serial recursion);
      (2) for blocks 1, 1406 and 2805 an oracle receiver started two chunks earlier yields the pre-AGC audio: its block
          peak equals the device's, and  pre-AGC audio x oracle-trajectory gain  equals the device's audio at the gate."""
    P, Po = cfg2['P'], cfg2['Po']
    C = P.IN_CHUNK_SIZE
    pk, gn = cfg2['trace']
    assert pk.shape == gn.shape == (4, N_CHUNKS)
    traj = np.zeros_like(gn, dtype=np.float64)
    for r in range(4):
        g = odsp.agc()
        traj[r] = [g.update(p) for p in pk[r]]
        assert np.max(np.abs(traj[r] - gn[r]) / traj[r]) <= 2e-7, r        # float32 storage of a float64 recursion
    for blk in (1, 1406, 2805):
        # three consecutive blocks: a keyed CW carrier is silent for whole blocks, and the float32 FFT filter's rounding
        # error scales with the loudest sample in its 4096-point window — errors are measured against the receiver's level
        # over the span (and the capture, for the peaks), not against a silent block's own residue
        warm, span = min(2, blk), 3
        s0 = (blk - warm) * C
        xs = cfg2['x'][s0:(blk + span) * C].cpu().numpy()
        for r in range(4):
            orx = odsp.Receiver(Po, cfg2['offs'][r], r, str(r), fast=True)
            orx.lo.advance(s0)
            orx.dec.n0 = s0
            orx.demod.m0 = odsp.n_out_total(s0, P.UP, P.DOWN)
            ref = []
            for c in range(warm + span):
                iq = orx.dec.resamp_fast(xs[c * C:(c + 1) * C], orx.lo)
                a = orx.demod.demod(iq, MODES[r], odsp._af_index(Po, r), odsp.per_rx(Po.BFO, r))
                if c >= warm:
                    b = blk + c - warm
                    assert abs(np.max(np.abs(a)) - pk[r, b]) <= 1e-5 * np.max(pk[r]), (b, r)
                    ref.append(np.asarray(a) * traj[r, b])
            m_lo = odsp.n_out_total(blk * C, P.UP, P.DOWN)
            m_hi = odsp.n_out_total((blk + span) * C, P.UP, P.DOWN)
            ref = np.concatenate(ref)
            assert len(ref) == m_hi - m_lo
            assert_parity(cfg2['am'][r][m_lo:m_hi].cpu().numpy(), ref, "audio rx%d blocks %d..%d" % (r, blk, blk + span - 1))


def test_cfg3_full_size_psd_parseval(cfg2):
    """Panadapter lines over the whole capture (8192-point Hann, 50 % overlap, 16 frames per line): for every checked
    line  sum_k PSD[k] * sum(w^2) / NFFT  equals the mean windowed frame energy (Parseval), and line count is exact."""
    import pysdr_b200.sig_proc as dsp
    x = cfg2['x']
    sp = dsp.spectrum(8000., 8192, 8192, 0.5)
    lines = sp.waterfall(x, 16, dB=False, to_host=False)
    n_frames = 1 + (x.numel() - 8192) // 4096
    assert lines.shape == (n_frames // 16, 8192) and n_frames == 117165
    w = torch.from_numpy(sp.win).to(x.device).double()
    wsum2 = (w * w).sum()
    for ln in (0, 3661, lines.shape[0] - 1):
        e = 0.0
        for f in range(16):
            s = (ln * 16 + f) * 4096
            fr = x[s:s + 8192]
            e += ((fr.real.double() * w) ** 2 + (fr.imag.double() * w) ** 2).sum().item()
        want = e / 16 / wsum2.item()                       # = sum_k |X_k|^2 / (NFFT * sum w^2), averaged
        got = lines[ln].double().sum().item() / 8192
        assert abs(got - want) <= 2e-5 * want, (ln, got, want)
