"""Many-channel receiver bank (BASELINE config 5: 1024 simultaneous channel receivers on one 10 MS/s stream).

The reference's surface stops at MAX_RX = 6 (params.py:33); this is the north-star extension.  Channels are served in
groups of up to 8 receivers, one ReceiverBank (= one set of K1/K2 launches) per group, all groups reading the SAME
device-resident block of IQ.  Every number still comes from libpysdr_b200.so; this class is only the loop over
groups.  At 3/625 the contraction is compute-bound (13 kFLOP per input sample for 1024 channels, 474 flop/B): the
tap-stationary K1 runs it at its FP32-FMA rate; a tensor-core formulation is the named next step."""
import copy

import numpy as np
import torch

from . import design
from .bank import ReceiverBank


class ChannelBank:
    GROUP = 8

    def __init__(self, P, offsets_hz, modes, af_bw=0.0, bfo=0.0, max_in=None, device=None):
        n = len(offsets_hz)
        modes = list(modes) if isinstance(modes, (list, tuple)) else [modes] * n
        af_bw = list(af_bw) if isinstance(af_bw, (list, tuple, np.ndarray)) else [af_bw] * n
        bfo = list(bfo) if isinstance(bfo, (list, tuple, np.ndarray)) else [bfo] * n
        if not (len(modes) == len(af_bw) == len(bfo) == n):
            raise ValueError("one mode / AF bandwidth / BFO per channel")
        self.P, self.n_ch = P, n
        self.banks, self.slices = [], []
        for g0 in range(0, n, self.GROUP):
            g1 = min(n, g0 + self.GROUP)
            Pg = copy.copy(P)
            Pg.MODE, Pg.AF_BW, Pg.BFO = modes[g0:g1], af_bw[g0:g1], bfo[g0:g1]
            Pg.AF_FILTER_NUM = None
            Pg.NUM_RX = g1 - g0
            self.banks.append(ReceiverBank(Pg, list(offsets_hz[g0:g1]), max_in=max_in, device=device))
            self.slices.append((g0, g1))
        self.device = self.banks[0].device
        self.n_out = 0

    def process(self, x, want_dc=False):
        """x: device complex64 block (whole IN_CHUNK_SIZE chunks).  Returns (am, iq): lists of n_ch device views, valid
        until the next call."""
        am, iq = [], []
        for b in self.banks:
            a, q, _ = b.process(x, want_dc=want_dc)
            am.extend(a)
            iq.extend(q)
        self.n_out = self.banks[0].n_out
        return am, iq

    def launch_count(self):
        return sum(b.lib.pysdr_bank_launch_count(b.h) for b in self.banks)


def raster_offsets(n_ch, spacing_hz, centre_hz=0.0):
    """n_ch offsets on a uniform raster centred on centre_hz (config 5: 1024 channels, 9.6 kHz apart)."""
    return [centre_hz + spacing_hz * (k - (n_ch - 1) / 2.0) for k in range(n_ch)]


class ShardedChannelBank:
    """Config 5's multi-GPU shape: the capture is split in time across ranks, every rank runs all channel groups on its
    shard, and the AGC carry of ALL channels travels in ONE all-gather of [n_ch, n_blocks] block peaks."""

    def __init__(self, cb, rank, world, chunks_per_rank):
        from .dist import ShardedCapture
        self.cb, self.rank, self.world = cb, rank, world
        self.shards = [ShardedCapture(b, b.P, rank, world, chunks_per_rank) for b in cb.banks]
        self.plan = self.shards[0].plan

    def step(self, xbuf):
        """xbuf: device samples [first_sample, start+n) of this rank's shard.  Returns (am, iq) lists over channels."""
        from .dist import exchange_agc_peaks
        own = torch.cat([sh.front(xbuf) for sh in self.shards])          # [n_ch, n_blocks]
        prev = exchange_agc_peaks(own, self.rank, self.world)            # one collective for every channel
        am, iq = [], []
        for sh, (g0, g1) in zip(self.shards, self.cb.slices):
            a, q, _ = sh.back(None if prev is None else prev[g0:g1].contiguous())
            am.extend(a)
            iq.extend(q)
        return am, iq
