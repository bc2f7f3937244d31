"""World-size-2 gloo test (CPU) of the time-shard protocol in pysdr_b200/dist.py: shard planning, filter-memory
warm-up, and the single all-gather of AGC block peaks.  The per-shard arithmetic is done by the oracle here
(no GPU in this container); the same protocol drives the CUDA bank in bench.py / tests -m gpu."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import receiver_oracle as rxo
from oracle import sig_proc_oracle as odsp
from pysdr_b200.dist import (AGC_SUMMARY_LEN, agc_enter_reference, agc_summary_reference, exchange_agc_peaks,
                             exchange_agc_summaries, shard_plan)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mkP():
    return rxo.make_P(2.048e6, [1000e3, 1040e3], ['USB', 'AM'], foffset=100e3, af_bw=[2e3, 5e3], nfilt=101)


def _capture(P, n_chunks):
    rng = np.random.default_rng(42)
    n = n_chunks * P.IN_CHUNK_SIZE
    x = ((rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.05).astype(np.complex64)
    x[: 2 * P.IN_CHUNK_SIZE] *= 0.2
    return x


def _pre_agc(rx, P, x):
    """demod_data without the AGC stage (what process_front produces)."""
    iq = rx.dec.resamp(x, rx.lo)
    return rx.demod.demod(iq, odsp.per_rx(P.MODE, rx.irx), odsp._af_index(P, rx.irx), odsp.per_rx(P.BFO, rx.irx))


def _worker(rank, world, port, chunks_per_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = _mkP()
        C = P.IN_CHUNK_SIZE
        x = _capture(P, world * chunks_per_rank)
        plan = shard_plan(P, rank, world, chunks_per_rank)
        assert plan['start'] == rank * chunks_per_rank * C and plan['n'] == chunks_per_rank * C
        offs = [P.FOFFSET + f - P.FC[0] for f in P.FC]
        pre, peaks = [], np.zeros((2, chunks_per_rank), np.float32)
        for irx in range(2):
            rx = odsp.Receiver(P, offs[irx], irx, str(irx + 1))
            # seek: absolute indices, LO phase, raw halo
            n0 = plan['start'] - plan['warm_chunks'] * C
            rx.dec.n0 = n0
            rx.lo.acc = (rx.lo.inc * n0) & odsp.MASK64
            rx.demod.m0 = odsp.n_out_total(n0, P.UP, P.DOWN)
            if plan['halo']:
                rx.dec.hist = x[n0 - plan['halo']:n0].astype(np.complex128)
            for w in range(plan['warm_chunks']):
                _pre_agc(rx, P, x[n0 + w * C:n0 + (w + 1) * C])           # discard: rebuilds the AF memory
            a = []
            for c in range(chunks_per_rank):
                s = plan['start'] + c * C
                a.append(_pre_agc(rx, P, x[s:s + C]))
                peaks[irx, c] = np.max(np.abs(a[-1]))
            pre.append(a)
        out = []
        if chunks_per_rank >= 8:
            # O(1) carry: 19 doubles per receiver per rank cross the wire, whatever the shard length
            own = torch.from_numpy(np.array([agc_summary_reference(peaks[irx]) for irx in range(2)]))
            allsum = exchange_agc_summaries(own, torch.zeros((world, 2, AGC_SUMMARY_LEN), dtype=torch.float64), world).numpy()
            for irx in range(2):
                g = odsp.agc()
                g.gain, g.ring, g.k = agc_enter_reference(allsum[:, irx], rank)
                out.append(np.concatenate([pre[irx][c] * g.update(peaks[irx, c]) for c in range(chunks_per_rank)]))
            q.put((rank, out))
            return
        prev = exchange_agc_peaks(torch.from_numpy(peaks), rank, world)
        assert (prev is None) == (rank == 0)
        for irx in range(2):
            g = odsp.agc()
            if prev is not None:
                assert tuple(prev.shape) == (2, rank * chunks_per_rank)
                for pk in prev[irx].numpy():
                    g.update(np.float32(pk))
            out.append(np.concatenate([pre[irx][c] * g.update(peaks[irx, c]) for c in range(chunks_per_rank)]))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("cpr", [3, 8])
def test_two_rank_time_shard_equals_single_stream(cpr):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cpr, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-stream reference
    P = _mkP()
    C = P.IN_CHUNK_SIZE
    x = _capture(P, world * cpr)
    rxo.create_receivers(P)
    for irx in range(2):
        ref = np.concatenate([P.rx[irx].demod_data(x[c * C:(c + 1) * C]).astype(np.float64) for c in range(world * cpr)])
        got = np.concatenate([res[r][irx] for r in range(world)])
        assert got.shape == ref.shape
        # float32 peaks crossing the wire vs float64 peaks in the single stream: gains agree to ~1e-7
        assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))


def test_shard_plan_geometry():
    P = rxo.make_P(8e6, [1e6], 'USB', foffset=100e3)
    p0 = shard_plan(P, 0, 8, 2812)
    p3 = shard_plan(P, 3, 8, 2812)
    assert p0['lead'] == 0 and p0['warm_chunks'] == 0
    assert p3['start'] == 3 * 2812 * 170666 and p3['halo'] == 333 and p3['warm_chunks'] == 1
    assert p3['first_sample'] == p3['start'] - 170666 - 333
    # 1002 baseband samples of AF memory need ceil(1002*500/3)=167000 inputs < one 170666-sample chunk
    Pw = rxo.make_P(0.25e6, [1e6], 'USB', foffset=20e3)               # 24/125: chunk 5333 inputs -> 1024 outputs
    assert shard_plan(Pw, 1, 2, 10)['warm_chunks'] == 1


def _rx_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pysdr_b200.dist import gather_audio, receiver_shard
        P = rxo.make_P(2.048e6, [1000e3, 1040e3, 1075e3], ['USB', 'AM', 'CW'], foffset=100e3, af_bw=[2e3, 5e3, 500.], nfilt=101)
        x = _capture(P, 3)
        offs = [P.FOFFSET + f - P.FC[0] for f in P.FC]
        mine = receiver_shard(3, rank, world)
        local = {}
        for irx in mine:
            rx = odsp.Receiver(P, offs[irx], irx, str(irx + 1))
            C = P.IN_CHUNK_SIZE
            local[irx] = torch.from_numpy(np.concatenate([rx.demod_data(x[c * C:(c + 1) * C]) for c in range(3)]).astype(np.float32))
        allrx = gather_audio(local, 3, rank, world)
        q.put((rank, mine, allrx.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_receiver_shard_needs_no_exchange():
    from pysdr_b200.dist import receiver_shard
    assert [receiver_shard(3, r, 2) for r in range(2)] == [[0, 1], [2]]
    assert [receiver_shard(4, r, 8) for r in range(8)] == [[0], [1], [2], [3], [], [], [], []]
    assert sum((receiver_shard(1024, r, 8) for r in range(8)), []) == list(range(1024))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rx_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == [0, 1] and res[1][1] == [2]
    assert np.array_equal(res[0][2], res[1][2]) and res[0][2].shape[0] == 3
    P = rxo.make_P(2.048e6, [1000e3, 1040e3, 1075e3], ['USB', 'AM', 'CW'], foffset=100e3, af_bw=[2e3, 5e3, 500.], nfilt=101)
    x = _capture(P, 3)
    rxo.create_receivers(P)
    C = P.IN_CHUNK_SIZE
    for irx in range(3):
        ref = np.concatenate([P.rx[irx].demod_data(x[c * C:(c + 1) * C]) for c in range(3)]).astype(np.float32)
        assert np.array_equal(res[0][2][irx], ref)
