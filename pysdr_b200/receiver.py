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
import numpy as np
import torch

from . import design
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
