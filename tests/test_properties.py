"""Property tests (hypothesis, CPU only) of the host-side arithmetic every kernel launch depends on: rate reduction,
output counting, phase increments (three implementations, bit-exact), offset quantisation and the shard planners."""
import ctypes
from fractions import Fraction

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import sig_proc_oracle as odsp
from pysdr_b200 import _lib, design
from pysdr_b200.dist import receiver_shard, shard_plan

RATES = st.sampled_from([250e3, 1e6, 1.024e6, 2e6, 2.048e6, 2.4e6, 3e6, 6e6, 8e6, 10e6])
FS_OUT = st.sampled_from([8000, 12000, 24000, 48000, 96000])


@given(RATES, FS_OUT)
def test_up_dn_is_the_reduced_fraction(fs, fs_out):
    up, down = design.up_dn(fs, fs_out)
    assert (up, down) == odsp.up_dn(fs, fs_out)
    assert Fraction(up, down) == Fraction(int(fs_out), int(fs)) and np.gcd(up, down) == 1


@given(st.integers(1, 64), st.integers(1, 1000), st.lists(st.integers(0, 500000), min_size=1, max_size=8))
def test_output_count_is_additive_over_any_chunking(up, down, chunks):
    """ceil(UP*n/DOWN) outputs after n inputs, so a call covering inputs [a, b) emits n_out(b) - n_out(a): the counts of
    any chunking telescope to the whole (the bit-exact indexing gate)."""
    pos, total = 0, 0
    for c in chunks:
        total += design.n_out_total(pos + c, up, down) - design.n_out_total(pos, up, down)
        pos += c
    assert total == design.n_out_total(pos, up, down) == -((-up * pos) // down) == odsp.n_out_total(pos, up, down)


@settings(max_examples=300)
@given(st.floats(-6e6, 6e6, allow_nan=False), RATES)
def test_phase_increment_three_implementations_bit_exact(f, fs):
    lib = _lib.load()
    a = design.freq_to_phase_inc(f, fs)
    b = odsp.freq_to_phase_inc(f, fs)
    c = int(lib.pysdr_freq_to_phase_inc(ctypes.c_double(f), ctypes.c_double(fs)))
    assert a == b == c and 0 <= a < 2 ** 64
    back = design.phase_inc_to_freq(a, fs)
    assert back == lib.pysdr_phase_inc_to_freq(ctypes.c_uint64(a), ctypes.c_double(fs))
    alias = (f + fs / 2) % fs - fs / 2                       # the increment lives on the circle: f modulo fs
    assert abs(back - alias) <= fs * 2.0 ** -52 + 1e-9 or abs(abs(back - alias) - fs) <= 1e-6


@given(RATES, st.floats(-3e6, 3e6, allow_nan=False), st.sampled_from([32768, 65536, 131072, 262144]))
def test_adjust_foffset_is_idempotent_and_periodic_in_the_ring_buffer(fs, foff, rb):
    class P:
        pass
    P.SRATE, P.FOFFSET, P.RB_SIZE = fs, foff, rb
    design.adjust_foffset(P)
    f1 = P.FOFFSET
    cycles = f1 * rb / fs                                     # whole LO periods per ring buffer (utils.py:277-289)
    assert abs(cycles - round(cycles)) < 1e-6 and abs(f1 - foff) <= fs / rb / 2 + 1e-6
    design.adjust_foffset(P)
    assert P.FOFFSET == f1


@given(st.integers(1, 2048), st.integers(1, 16))
def test_receiver_shards_partition_the_receivers(n_rx, world):
    parts = [receiver_shard(n_rx, r, world) for r in range(world)]
    assert sum(parts, []) == list(range(n_rx))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@given(st.sampled_from([(2.048e6, 48000), (8e6, 48000), (10e6, 48000), (2.4e6, 48000), (250e3, 48000)]),
       st.integers(1, 8), st.integers(1, 400), st.sampled_from([101, 301, 1001]))
def test_time_shards_tile_the_capture_and_carry_enough_history(rates, world, chunks_per_rank, nfilt):
    class P:
        pass
    fs, fs_out = rates
    P.UP, P.DOWN = design.up_dn(fs, fs_out)
    P.IN_CHUNK_SIZE = int(1024 * P.DOWN / float(P.UP))
    P.FILT_LEN = nfilt
    C = P.IN_CHUNK_SIZE
    end = 0
    for r in range(world):
        p = shard_plan(P, r, world, chunks_per_rank)
        assert p['start'] == end and p['start'] % C == 0 and p['n'] == chunks_per_rank * C
        end = p['start'] + p['n']
        if r == 0:
            assert p['lead'] == 0
        else:
            warm_out = design.n_out_total(p['start'], P.UP, P.DOWN) - design.n_out_total(p['start'] - p['warm_chunks'] * C, P.UP, P.DOWN)
            assert p['first_sample'] == p['start'] - p['lead'] >= 0
            if p['first_sample'] > 0 or p['halo']:
                assert warm_out >= nfilt + 1                  # the AF memory (FILT_LEN+1 baseband samples) is rebuilt
                assert p['halo'] == (nfilt + P.UP - 1) // P.UP - 1
            else:                                             # warm-up from the very first sample: nothing precedes it
                assert p['halo'] == 0 and p['start'] - p['warm_chunks'] * C == 0
    assert end == world * chunks_per_rank * C
