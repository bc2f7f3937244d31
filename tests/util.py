import json
import os

import numpy as np

REL_TOL = 1e-4          # north_star: max-abs relative error <= 1e-4
SNR_MIN_DB = 80.0       # north_star: difference SNR >= 80 dB


def err_metrics(got, ref):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    d = got.astype(np.complex128) - ref.astype(np.complex128)
    peak = np.max(np.abs(ref)) if ref.size else 1.0
    rel = float(np.max(np.abs(d)) / max(peak, 1e-300)) if ref.size else 0.0
    pr = float(np.sum(np.abs(ref.astype(np.complex128)) ** 2))
    pd = float(np.sum(np.abs(d) ** 2))
    snr = 10 * np.log10(pr / pd) if pd > 0 else np.inf
    return rel, snr


def assert_parity(got, ref, what="", rel_tol=REL_TOL, snr_min=SNR_MIN_DB):
    rel, snr = err_metrics(got, ref)
    log = os.environ.get("PYSDR_PARITY_LOG")                 # margins of every comparison, for profiles/parity_report_*.txt
    if log:
        with open(log, "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "what": what, "rel": rel,
                                "snr_db": None if snr == np.inf else snr, "rel_tol": rel_tol, "snr_min": snr_min}) + "\n")
    assert rel <= rel_tol and snr >= snr_min, "%s: max-abs rel err %.3e (tol %.1e), diff SNR %.1f dB (min %.0f)" % (
        what, rel, rel_tol, snr, snr_min)
    return rel, snr


def cfg_args(srate_mhz, fcs_khz, modes, foffset_khz=100, af_bw_khz=None, nfilt=1001, bfo=None):
    a = ['-fs', str(srate_mhz), '-fc'] + [str(f) for f in fcs_khz] + ['-mode'] + list(modes) + \
        ['-foffset', str(foffset_khz), '-nfilt', str(nfilt)]
    if af_bw_khz is not None:
        a += ['-af_bw'] + [str(b) for b in af_bw_khz]
    if bfo is not None:
        a += ['-bfo'] + [str(b) for b in bfo]
    return a


def make_both(srate_mhz, fcs_khz, modes, foffset_khz=100, af_bw_khz=None, nfilt=1001, bfo=None, srate_hz=None):
    """(product P, oracle P) for the same configuration."""
    from pysdr_b200.params import RUN_TIME_PARAMS
    from oracle import receiver_oracle as rxo
    kw = {}
    if srate_hz:
        kw['srate_hz'] = srate_hz
    P = RUN_TIME_PARAMS(cfg_args(srate_mhz, fcs_khz, modes, foffset_khz, af_bw_khz, nfilt, bfo), **kw)
    mode = list(modes) if len(modes) > 1 else modes[0]
    af = 0.0 if af_bw_khz is None else ([b * 1e3 for b in af_bw_khz] if len(af_bw_khz) > 1 else af_bw_khz[0] * 1e3)
    b = 0 if bfo is None else (list(bfo) if len(bfo) > 1 else bfo[0])
    Po = rxo.make_P(srate_hz or P.SRATE, [f * 1e3 for f in fcs_khz], mode, foffset=foffset_khz * 1e3, af_bw=af,
                    nfilt=nfilt, bfo=b)
    return P, Po


def lcg_iq(n, seed, scale=0.05):
    """Portable deterministic complex64 noise (integer LCG, no library RNG) for golden fixtures."""
    k = np.arange(n, dtype=np.uint64)
    with np.errstate(over='ignore'):
        h = (k + np.uint64(seed)) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
        h ^= h >> np.uint64(29)
        h = h * np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
    re = ((h & np.uint64(0xFFFF)).astype(np.float64) / 32768.0 - 1.0)
    im = (((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.float64) / 32768.0 - 1.0)
    return ((re + 1j * im) * scale).astype(np.complex64)


def golden_input(n, srate, offsets_hz, seed):
    """LCG noise + one AM-ish carrier per receiver offset (float64 phase, cast once)."""
    x = lcg_iq(n, seed).astype(np.complex128)
    t = np.arange(n, dtype=np.float64) / srate
    for i, f in enumerate(offsets_hz):
        x += 0.1 * (1 + 0.5 * np.sin(2 * np.pi * (700.0 + 300 * i) * t)) * np.exp(2j * np.pi * (f + 900.0) * t)
    return x.astype(np.complex64)


def rtty_input(n_sym, N, FS_OUT=48000, seed=505):
    """Noise + a 45.45 baud FSK pair (mark 915 Hz / space 1085 Hz) + a weak carrier; complex64."""
    n = n_sym * N
    t = np.arange(n)
    bits = ((np.arange(n_sym) * 7 + 3) % 5 < 2).astype(np.float64)
    f = np.repeat(np.where(bits > 0, 915.0, 1085.0), N)
    ph = 2 * np.pi * np.cumsum(f) / FS_OUT
    x = lcg_iq(n, seed, scale=0.01).astype(np.complex128) + 0.2 * np.exp(1j * ph) + 0.02 * np.exp(-2j * np.pi * 3000.0 * t / FS_OUT)
    return x.astype(np.complex64)
