"""Capture files either side of the receive path: the replay source and the raw_iq /
baseband_iq / demod sinks.

Interface kept from the reference's call sites — ``sdr_fileio(fname, 'r'|'w', P, nchan, tag)`` with ``.srate``, ``.fc``,
``read_data()``, ``save_data(x, VERBOSITY=0)`` (reference receiver.py:295-297,526,761,810-813; pySDR.py:118-123;
mp.py:98-100; sigs/iq.py:59-60) and the file naming visible in the tree (``demod_20190321_225218.dat``,
``baseband_iq_20190413_221346.dat``, sigs/nfm.m:40-44).

PARITY UNPINNED: the on-disk layout is defined by the un-vendored ``fileio`` module of github.com/aa2il/libs (and its
Matlab twin ``read_sdr_data.m``), neither of which is in the reference tree.  The only in-tree facts are ``hdr(1) = fs``
and ``hdr(4) = nchan`` plus a string tag returned next to the header (sigs/nfm.m:50-55, sigs/sdr2wav.m:37-43).  The
layout below honours those and is otherwise OURS:

    float32  hdr[16]   hdr[0] = fs [Hz], hdr[1] = fc [kHz], hdr[2] = 0 (reserved), hdr[3] = nchan,
                       hdr[4] = layout version (1), hdr[5] = fc remainder [Hz] (fc = hdr[1]*1e3 + hdr[5], exact)
    uint8    tag[64]   ASCII, zero padded (e.g. 'RAW_IQ', 'BASEBAND_IQ', the demod mode)
    float32  data[]    nchan = 2: interleaved I,Q (= complex64);  nchan = 1: real samples
"""
import os
import time

import numpy as np

HDR_LEN = 16
TAG_LEN = 64
VERSION = 1.0


class sdr_fileio:
    def __init__(self, fname, rw, P=None, nchan=2, tag=''):
        self.rw = rw
        self.P = P
        self.nchan = int(nchan)
        self.tag = tag
        self.fp = None
        self.nsamps = 0
        if rw == 'r':
            self.fname = fname
            with open(fname, 'rb') as f:
                hdr = np.fromfile(f, np.float32, HDR_LEN)
                raw_tag = f.read(TAG_LEN)
            if len(hdr) != HDR_LEN or len(raw_tag) != TAG_LEN or hdr[4] != VERSION or int(hdr[3]) not in (1, 2):
                raise ValueError("%s is not a pysdr_b200 capture file (layout version 1)" % fname)
            self.hdr = hdr
            self.srate = float(hdr[0])
            self.fc = float(hdr[1]) * 1e3 + float(hdr[5])
            self.nchan = int(hdr[3])
            self.tag = raw_tag.rstrip(b'\0').decode('ascii', 'replace')
        elif rw == 'w':
            # writers are created up front for every stream and only touch the disk on the first save_data
            # (reference pySDR.py:118-123 opens all three unconditionally)
            self.base = fname
            self.fname = None
            self.srate = float(getattr(P, 'SRATE', 0.0)) if tag == 'RAW_IQ' else float(getattr(P, 'FS_OUT', 0.0))
            fc = getattr(P, 'FC', [0.0])
            self.fc = float(fc[0] if isinstance(fc, (list, tuple, np.ndarray)) else fc)
        else:
            raise ValueError("rw must be 'r' or 'w'")

    # ---- reading -------------------------------------------------------------------------------------------
    def read_data(self, pinned=False):
        """All samples: complex64 (nchan 2) or float32 (nchan 1).  pinned=True returns a pinned torch tensor, ready for
        ReplayStreamer.run (the H2D copies then run at PCIe rate without a staging copy)."""
        off = 4 * HDR_LEN + TAG_LEN
        n = (os.path.getsize(self.fname) - off) // 4
        if self.nchan == 2:
            n -= n & 1
        dt = np.complex64 if self.nchan == 2 else np.float32
        if not pinned:
            return np.fromfile(self.fname, np.float32, n, offset=off).view(dt)
        import torch
        t = torch.empty(n, dtype=torch.float32, pin_memory=True)
        with open(self.fname, 'rb') as f:
            f.seek(off)
            f.readinto(memoryview(t.numpy()).cast('B'))
        return torch.view_as_complex(t.view(-1, 2)) if self.nchan == 2 else t

    # ---- writing -------------------------------------------------------------------------------------------
    def _open(self):
        stamp = time.strftime('%Y%m%d_%H%M%S', time.gmtime())
        d = getattr(self.P, 'SAVE_DIR', '.') if self.P is not None else '.'
        self.fname = os.path.join(d, '%s_%s.dat' % (self.base, stamp)) if not self.base.endswith('.dat') else self.base
        self.fp = open(self.fname, 'wb')
        hdr = np.zeros(HDR_LEN, np.float32)
        fc_khz = np.float32(np.floor(self.fc / 1e3))
        hdr[0], hdr[1], hdr[3], hdr[4] = self.srate, fc_khz, self.nchan, VERSION
        hdr[5] = self.fc - float(fc_khz) * 1e3
        self.hdr = hdr
        hdr.tofile(self.fp)
        self.fp.write(self.tag.encode('ascii', 'replace')[:TAG_LEN].ljust(TAG_LEN, b'\0'))

    def save_data(self, x, VERBOSITY=0):
        if self.rw != 'w':
            raise IOError("file was opened for reading")
        if self.fp is None:
            self._open()
        x = np.asarray(x)
        if self.nchan == 2:
            y = np.ascontiguousarray(x, np.complex64).view(np.float32)
        else:
            y = np.ascontiguousarray(x.real if np.iscomplexobj(x) else x, np.float32)
        y.tofile(self.fp)
        self.nsamps += len(x)
        if VERBOSITY > 0:
            print('sdr_fileio: wrote', len(x), 'samples to', self.fname)

    def close(self):
        if self.fp is not None:
            self.fp.close()
            self.fp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SDR_FILEIO = sdr_fileio            # reference sigs/iq.py:59 spells it in capitals


def open_replay(P, fname, literal=False):
    """Replay set-up of reference receiver.py:808-822: rates and chunk size follow the file's own sample rate.

    Baseband captures are already at the audio rate, so FS_OUT = SRATE for them.  The reference tests this with
    ``if P.REPLAY.find('baseband_iq'):`` (receiver.py:815) — str.find's truthiness, which is true for every name that does
    NOT start with 'baseband_iq' (-1) and false only when the name starts with it (0).  literal=True reproduces exactly
    that (what SDR_EXECUTIVE.create_SDR does, pinned by tests/golden/ref_callers.npz); the default implements the evident
    intent ('baseband_iq' in the file name)."""
    from . import design
    P.REPLAY = fname
    P.sdr = sdr_fileio(fname, 'r', P)
    P.SRATE = P.sdr.srate
    P.REPLAY_FC = P.sdr.fc
    P.FC[0] = P.sdr.fc
    if (fname.find('baseband_iq') if literal else 'baseband_iq' in os.path.basename(fname)):   # receiver.py:815-816
        P.FS_OUT = P.SRATE
    P.UP, P.DOWN = design.up_dn(P.SRATE, P.FS_OUT)
    P.FS_OUT = int(P.SRATE * P.UP / P.DOWN)
    P.IN_CHUNK_SIZE = int(P.OUT_CHUNK_SIZE * P.DOWN / float(P.UP) + 0 * 0.5)
    return P.sdr
