"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): a capture time-sharded over 2 ranks with NCCL
(pysdr_b200.dist.ShardedCapture: K1 halo in place, warm-up chunk, ONE all-gather of AGC block peaks) equals the
single-GPU single-stream result."""
import os
import socket

import numpy as np
import pytest
import torch

from tests.util import assert_parity

pytestmark = pytest.mark.gpu

FCS = [-500, 700, 1400, 3100]
MODES = ['AM', 'NFM', 'USB', 'CW']
CPRS = [4, 1, 9]                          # chunks per rank; 1: rank 1's warm-up starts at the very first sample;
                                          # 9: >= 8 blocks per shard -> the O(1) AGC summary exchange instead of the peak gather


def _P():
    from pysdr_b200.params import RUN_TIME_PARAMS
    return RUN_TIME_PARAMS(['-fs', '8', '-fc'] + [str(f) for f in FCS] + ['-mode'] + MODES +
                           ['-foffset', '100', '-af_bw', '5', '10', '2', '0.5'])


def _worker(rank, world, port, outdir, CPR, carry="nccl", steps=2):
    import torch.distributed as dist
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.dist import ShardedCapture
    from pysdr_b200.receiver import receiver_offsets
    from pysdr_b200.synth import synth_iq
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        P = _P()
        offs = receiver_offsets(P)
        C = P.IN_CHUNK_SIZE
        bank = ReceiverBank(P, offs, max_in=(CPR + 1) * C, device=dev)
        sh = ShardedCapture(bank, P, rank, world, CPR, carry=carry)
        assert (sh.peer is not None) == (carry == "peer" and CPR >= 8)
        pl = sh.plan
        xbuf = synth_iq(pl['lead'] + pl['n'], P.SRATE, offs, MODES, seed=77, device=dev, n0=pl['first_sample'], block=1 << 16)
        for _ in range(steps):                                    # repeatable; > 16 steps reuse the exchange ring's slots
            am, iq, dc = sh.step(xbuf, want_dc=True)
        torch.cuda.synchronize()
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), **{"am%d" % r: am[r].cpu().numpy() for r in range(4)},
                 **{"iq%d" % r: iq[r].cpu().numpy() for r in range(4)})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("CPR,carry,steps", [(c, "nccl", 2) for c in CPRS] + [(9, "peer", 2), (9, "peer", 40)])
def test_two_gpu_time_shard_equals_single_gpu(tmp_path, CPR, carry, steps):
    """carry = "peer": the AGC summaries travel through NVLink peer memory written by agc_summary_push_kernel and are
    flag-waited inside the fused back kernel (no NCCL call on the data path); 40 steps wrap the 16-deep slot ring twice,
    so the acknowledgement path (slot reuse) is exercised too."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pysdr_b200.bank import ReceiverBank
    from pysdr_b200.receiver import receiver_offsets
    from pysdr_b200.synth import synth_iq
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path), CPR, carry, steps), nprocs=world, join=True)
    P = _P()
    offs = receiver_offsets(P)
    C = P.IN_CHUNK_SIZE
    n = world * CPR * C
    x = synth_iq(n, P.SRATE, offs, MODES, seed=77, device="cuda:0", block=1 << 16)
    bank = ReceiverBank(P, offs, max_in=n, device="cuda:0")
    am, iq, _ = bank.process(x)
    ref_am = [a.cpu().numpy() for a in am]
    ref_iq = [a.cpu().numpy() for a in iq]
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for r in range(4):
        got_am = np.concatenate([p["am%d" % r] for p in parts])
        got_iq = np.concatenate([p["iq%d" % r] for p in parts])
        assert got_am.shape == ref_am[r].shape
        np.testing.assert_array_equal(got_iq, ref_iq[r])                         # K1 is shard-invariant bit for bit
        assert_parity(got_am, ref_am[r], "sharded vs single rx%d" % r, rel_tol=2e-5, snr_min=90)


def _chan_worker(rank, world, port, outdir, raster, cpr=3):
    import torch.distributed as dist
    from pysdr_b200.channelizer import ChannelBank, ShardedChannelBank, raster_offsets
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.synth import synth_iq
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        P = RUN_TIME_PARAMS(['-fs', '10', '-fc', '7000', '-mode', 'USB', '-af_bw', '2'])
        offs, modes, afs = _chan_cfg()
        C = P.IN_CHUNK_SIZE
        cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=(cpr + 2) * C, device=dev,
                         raster=(offs[0], 9600.0) if raster else None, group=8)
        sh = ShardedChannelBank(cb, rank, world, cpr)
        pl = sh.plan
        xbuf = synth_iq(pl['lead'] + pl['n'], P.SRATE, offs[:4], modes[:4], seed=78, device=dev, n0=pl['first_sample'], block=1 << 16)
        am, iq = sh.step(xbuf)
        torch.cuda.synchronize()
        np.savez(os.path.join(outdir, "chan%d.npz" % rank), **{"am%d" % r: am[r].cpu().numpy() for r in range(len(offs))})
    finally:
        dist.destroy_process_group()


def _chan_cfg():
    from pysdr_b200.channelizer import raster_offsets
    n_ch = 20
    offs = raster_offsets(n_ch, 9600.0, 150e3)
    modes = [['AM', 'NFM', 'USB', 'CW'][k % 4] for k in range(n_ch)]
    afs = [[5e3, 10e3, 2e3, 500.][k % 4] for k in range(n_ch)]
    return offs, modes, afs


@pytest.mark.parametrize("raster,cpr", [(False, 3), (True, 3), (True, 9), (False, 9)])
def test_two_gpu_many_channel_time_shard(tmp_path, raster, cpr):
    """Config 5's shape at test size: 20 channels (3 groups) on a 10 MS/s stream, time-sharded over 2 ranks with one
    all-gather of every channel's AGC peaks, equals the single-GPU single-stream result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pysdr_b200.channelizer import ChannelBank
    from pysdr_b200.params import RUN_TIME_PARAMS
    from pysdr_b200.synth import synth_iq
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_chan_worker, args=(2, port, str(tmp_path), raster, cpr), nprocs=2, join=True)
    P = RUN_TIME_PARAMS(['-fs', '10', '-fc', '7000', '-mode', 'USB', '-af_bw', '2'])
    offs, modes, afs = _chan_cfg()
    C = P.IN_CHUNK_SIZE
    x = synth_iq(2 * cpr * C, P.SRATE, offs[:4], modes[:4], seed=78, device="cuda:0", block=1 << 16)
    cb = ChannelBank(P, offs, modes, af_bw=afs, bfo=700.0, max_in=2 * cpr * C, device="cuda:0")
    for b in cb.banks:
        b.set_k1_mma(0)                 # the shards' banks (groups of 8) run the FP32 K1 / wola.cu: compare like with like at 2e-5
    am, _ = cb.process(x)
    parts = [np.load(os.path.join(str(tmp_path), "chan%d.npz" % r)) for r in range(2)]
    for r in range(len(offs)):
        got = np.concatenate([p["am%d" % r] for p in parts])
        assert_parity(got, am[r].cpu().numpy(), "channel %d" % r, rel_tol=2e-5, snr_min=90)
