"""Receive executive over the B200 path — the L2 loop of reference receiver.py restated so that all
receivers of a chunk share ONE read of the IQ samples on the device.

Reference semantics kept:
  * receiver offsets            receiver.py:826-835   frq = FOFFSET + FC[irx] - FC[0]  (or FC[irx]-FC[SOURCE])
  * replay chunking             receiver.py:538-559   idx = praw + arange(IN_CHUNK); continue while
                                                      praw+IN_CHUNK < len(raw) (strict); on EOF the stale
                                                      chunk is demodulated once more (:715-725)
  * per-chunk DC removal        receiver.py:250-252   for AM / USB, on the PSD/file copy only
  * audio gain                  receiver.py:197-200   af_gain = 10**AF_GAIN - 1, muted -> 0
  * duration                    receiver.py:764
"""
import ctypes

import numpy as np
import torch

from . import design
from ._lib import check
from .bank import ReceiverBank


def receiver_offsets(P):
    frq = []
    for irx in range(P.NUM_RX):
        if P.SOURCE[irx] >= 0:
            frq.append(P.FC[irx] - P.FC[P.SOURCE[irx]])
        else:
            frq.append(P.FOFFSET + P.FC[irx] - P.FC[0])
    return frq


def af_gain(P, irx=0):
    if P.MUTED[irx] or P.AUTO_MUTED:
        return 0.
    return pow(10., P.AF_GAIN) - 1


def audio_out(P, am):
    """What reference receiver.py:153-225 pushes to the audio players for one chunk (compute part only; the players and
    their ring buffers are out of scope).  am: list of per-receiver audio of this chunk.  Returns one payload per player:
    AUDIO_SCHEME 2 (:158-188) routes two mono receivers to one player as ``am1*g1 + 1j*am2*g2`` with the partner
    irx + (NUM_RX+1)//2; otherwise (:190-225) every receiver has its own player and gets ``am*gain`` (0 when muted or
    auto-muted).  The gain is the slider law 10**AF_GAIN - 1 (:173,200)."""
    n_rx = int(P.NUM_RX)
    g = pow(10., P.AF_GAIN) - 1
    if getattr(P, 'AUDIO_SCHEME', 1) == 2:
        n2 = int((n_rx + 1) / 2)
        out = []
        for irx in range(n2):
            a1 = np.asarray(am[irx]).real
            g1 = 0. if P.MUTED[irx] else g
            if irx + n2 < n_rx:
                a2 = np.asarray(am[irx + n2]).real
                g2 = 0. if P.MUTED[irx + n2] else g
            else:
                a2, g2 = 0, 0.
            out.append(a1 * g1 + 1j * a2 * g2)
        return out
    return [np.asarray(am[irx]) * af_gain(P, irx) for irx in range(n_rx)]


class SDR_EXECUTIVE:
    """Replay-mode executive (reference receiver.py:408-782, MP_SCHEME 1 data plane)."""

    def __init__(self, P, max_chunks_per_call=1):
        self.P = P
        P.SDR_EXEC = self
        P.RX_DONE = False
        P.nchunks = 0
        self.bank = ReceiverBank(P, receiver_offsets(P), max_in=int(P.IN_CHUNK_SIZE) * int(max_chunks_per_call))
        self.x = np.zeros(P.IN_CHUNK_SIZE, np.complex64)            # receiver.py:445

    def Run(self, raw, sink=None):
        """raw: host complex64 capture.  sink(irx, am*af_gain, am_dc, iq) is called per receiver per chunk
        (audio_out + PSD/file routing).  Returns the number of loop iterations."""
        P = self.P
        dt = float(P.IN_CHUNK_SIZE) / P.SRATE
        t = 0.
        praw = 0
        iters = 0
        P.RX_DONE = False
        while not P.RX_DONE:
            t += dt
            P.nchunks += 1
            iters += 1
            if praw + P.IN_CHUNK_SIZE < len(raw):                   # receiver.py:544 (strict)
                self.x = raw[praw:praw + P.IN_CHUNK_SIZE]
                praw += P.IN_CHUNK_SIZE
            else:
                P.RX_DONE = True                                    # stale self.x is processed again
            am, iq, dc = self.bank.process_host(self.x)
            if sink is not None:
                for irx in range(P.NUM_RX):
                    sink(irx, am[irx] * af_gain(P, irx), dc[irx], iq[irx])
            P.RX_DONE = P.RX_DONE or t >= P.DURATION               # receiver.py:764
        return iters


def demod_capture(P, x_dev, bank=None, want_dc=False):
    """Batch form: a device-resident capture of k whole chunks through every receiver in one pass
    (identical numbers to k successive demod_data calls).  Returns (bank, am, iq, am_dc)."""
    if bank is None:
        bank = ReceiverBank(P, receiver_offsets(P), max_in=int(x_dev.numel()))
    am, iq, dc = bank.process(x_dev, want_dc=want_dc)
    return bank, am, iq, dc


class ReplayStreamer:
    """Host-resident capture -> audio, the L2 executive on GPU streams: the capture stays in
    pinned host memory and is fed in segments of `seg_chunks` IN_CHUNK_SIZE blocks through two device buffers — the
    H2D copy of segment i+1 (copy stream) overlaps the kernels of segment i (compute stream) — and each segment's
    audio is copied back to pinned host memory asynchronously.  Same numbers as chunk-at-a-time `demod_data`
    (per-block semantics are defined on absolute block indices); MP_SCHEME 3's broadcast + barrier collapses to one
    process per GPU."""

    CS16_SCALE = 1.0 / 2048.0                                   # reference receiver.py:614

    def __init__(self, P, seg_chunks=64, device=None, want_iq=False, fmt='cf32'):
        """fmt 'cf32': the host capture is complex64 (replay files, SOAPY_SDR_CF32).  fmt 'cs16': interleaved int16 I/Q
        as SDR hardware delivers it (reference receiver.py:609-617); it crosses PCIe at 4 bytes per sample and is
        scaled by 1/2048 to complex64 on the device."""
        if fmt not in ('cf32', 'cs16'):
            raise ValueError("fmt must be 'cf32' or 'cs16'")
        self.fmt = fmt
        self.P = P
        self.C = int(P.IN_CHUNK_SIZE)
        self.seg_chunks = int(seg_chunks)
        self.bank = ReceiverBank(P, receiver_offsets(P), max_in=self.seg_chunks * self.C, device=device)
        dev = self.bank.device
        self.dbuf = [torch.empty(self.seg_chunks * self.C, dtype=torch.complex64, device=dev) for _ in range(2)]
        self.dbuf16 = [torch.empty(2 * self.seg_chunks * self.C, dtype=torch.int16, device=dev) for _ in range(2)] \
            if fmt == 'cs16' else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.freed = [torch.cuda.Event() for _ in range(2)]
        self.want_iq = want_iq
        self.h_am = self.h_iq = None
        self.n_out_seg = []

    def pin(self, raw):
        """Pinned copy of a host capture (numpy complex64 or CPU tensor)."""
        if isinstance(raw, np.ndarray):
            raw = torch.from_numpy(np.ascontiguousarray(raw, np.int16 if self.fmt == 'cs16' else np.complex64))
        return raw if raw.is_pinned() else raw.pin_memory()

    def run(self, hx, start_sample=0):
        """hx: pinned CPU complex64 tensor holding whole chunks.  Returns (h_am, n_out_per_segment): pinned float32
        [n_segments, n_rx, n_out_max] (complex receivers: interleaved re/im in 2*n_out floats)."""
        C, sc, bank = self.C, self.seg_chunks, self.bank
        cs16 = self.fmt == 'cs16'
        if cs16 and hx.dtype != torch.int16 or not cs16 and hx.dtype != torch.complex64:
            raise ValueError("capture dtype %s does not match fmt %r" % (hx.dtype, self.fmt))
        n_chunks = (hx.numel() // 2 if cs16 else hx.numel()) // C
        segs = [(s, min(n_chunks, s + sc)) for s in range(0, n_chunks, sc)]
        n_out_max = 2 * ((int(self.P.UP) * sc * C) // int(self.P.DOWN) + 2)
        if self.h_am is None or self.h_am.shape[0] < len(segs):
            self.h_am = torch.empty((len(segs), bank.n_rx, n_out_max), dtype=torch.float32, pin_memory=True)
            if self.want_iq:
                self.h_iq = torch.empty((len(segs), bank.n_rx, n_out_max // 2), dtype=torch.complex64, pin_memory=True)
        cs = torch.cuda.current_stream(bank.device)
        cp = self.copy_stream
        cp.wait_stream(cs)
        bank.seek(start_sample)
        self.n_out_seg = []
        any_cplx = any(bank._mode_of(r) in ('IQ', 'RTTY') for r in range(bank.n_rx))
        for i, (a, b) in enumerate(segs):
            k = i & 1
            with torch.cuda.stream(cp):
                if i >= 2:
                    cp.wait_event(self.freed[k])
                if cs16:
                    self.dbuf16[k][:2 * (b - a) * C].copy_(hx[2 * a * C:2 * b * C], non_blocking=True)
                else:
                    self.dbuf[k][:(b - a) * C].copy_(hx[a * C:b * C], non_blocking=True)
                self.ready[k].record(cp)
            cs.wait_event(self.ready[k])
            if cs16:
                check(bank.lib.pysdr_cs16_to_cf32(ctypes.c_void_p(self.dbuf16[k].data_ptr()),
                                                  ctypes.c_void_p(self.dbuf[k].data_ptr()), (b - a) * C, self.CS16_SCALE,
                                                  ctypes.c_void_p(cs.cuda_stream)))
            bank.process(self.dbuf[k][:(b - a) * C], want_dc=False)
            self.freed[k].record(cs)
            no = bank.n_out
            self.n_out_seg.append(no)
            w = 2 * no if any_cplx else no                      # complex receivers (IQ/RTTY) use 2 floats per sample
            for r in range(bank.n_rx):                          # row by row: contiguous copies stay asynchronous (a strided
                self.h_am[i, r, :w].copy_(bank._am[r, :w], non_blocking=True)   # 2-D device-to-host copy_ blocks the host)
                if self.want_iq:
                    self.h_iq[i, r, :no].copy_(bank.iq_row(r, no), non_blocking=True)
        cs.synchronize()
        return self.h_am, self.n_out_seg

    def audio(self, irx):
        """Concatenated host audio of receiver irx from the last run (float32, complex64 for IQ/RTTY)."""
        cplx = self.bank._mode_of(irx) in ('IQ', 'RTTY')
        parts = []
        for i, no in enumerate(self.n_out_seg):
            row = self.h_am[i, irx]
            parts.append(row[:2 * no].numpy().view(np.complex64).copy() if cplx else row[:no].numpy().copy())
        return np.concatenate(parts) if parts else np.zeros(0, np.float32)
