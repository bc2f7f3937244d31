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


class BankReceiver:
    """One row of a ReceiverBank behind the attribute surface the reference's callers touch on ``P.rx[irx]``
    (SURVEY.md 8b): ``.am .iq .sub .lo.change_freq() .lo.fo .dec.h .dec.filter_bank .demod.filter_bank_real/_cmpx
    .demod.am_pll.reset() .agc.reset() .agc.{agc,gain,maxbuf,ref,err} .auto_mute(x)``.  The samples of all rows are
    produced together by the bank (one read of the chunk); ``demod_data`` on a row returns that row's result."""

    AUTO_MUTE_THRESH = 0.25

    def __init__(self, executive, bank, P, irx):
        from . import sig_proc as dsp
        self._exec, self._bank, self._slot, self.P, self.irx = executive, bank, irx, P, irx
        self.name = str(irx + 1)
        self.sub = 0
        self._wfm = None
        self.lo, self.dec, self.agc = dsp._Lo(self), dsp._Dec(self), dsp._Agc(self)
        self.demod = dsp._Demod(self)
        self.am = np.zeros(0, np.float32)
        self.am_dc = np.zeros(0, np.float32)
        self.iq = np.zeros(0, np.complex64)
        self.mute_cnt = 0

    def demod_data(self, x):
        self._exec._process_chunk(x)                    # all rows at once; a no-op when x was already processed
        return self.am

    def auto_mute(self, x):
        """Reference receiver.py:238-245 with the detector of DESIGN.md section 3: mean|x|^2 above the threshold holds the
        mute for MUTE_CHUNKS calls.  The chunk power is measured once on the device for all rows."""
        if self._exec._chunk_power(x) > self.AUTO_MUTE_THRESH:
            self.mute_cnt = int(self.P.MUTE_CHUNKS)
        elif self.mute_cnt > 0:
            self.mute_cnt -= 1
        return self.mute_cnt > 0


def demodulate_data(P, x, irx):
    """Per-receiver post-processing of one chunk, reference receiver.py:231-297: demodulate, auto-mute, DC removal for
    AM / USB on the copy that feeds the PSD and the demod file (the audio keeps its DC), routing to the AF / baseband PSD
    buffers, the RTTY decoder queue and the baseband / demod files."""
    rx = P.rx[irx]
    am = rx.demod_data(x)
    if P.ENABLE_AUTO_MUTE:
        mute = rx.auto_mute(x)
        if mute != bool(P.AUTO_MUTED):
            P.AUTO_MUTED = mute
            if getattr(P, 'gui', None) is not None:
                P.gui.btn9.setColor('red' if mute else 'lime')
    if design.per_rx(P.MODE, irx) in ('AM', 'USB'):
        dc = getattr(rx, 'am_dc', None)
        am = dc if dc is not None and len(dc) == len(am) else am - np.mean(am)
    if P.SHOW_AF_PSD and irx == P.PLOT_RX:
        if P.PANADAPTOR:
            P.rb_af.push(rx.iq)
        elif P.MP_SCHEME == 1:
            P.rb_af.push(am)
        else:
            P.af_psd_Q.put(am)
    if P.SHOW_BASEBAND_PSD and irx == P.PLOT_RX:
        if P.MP_SCHEME == 1:
            P.rb_baseband.push(rx.iq)
        else:
            P.bb_psd_Q.put(rx.iq)
    if getattr(P, 'ENABLE_RTTY', False) and P.gui is not None and P.gui.rtty.active:
        P.gui.rtty.q_in.put(('IQ', rx.iq))
    if P.SAVE_BASEBAND and irx == 0:
        P.baseband_iq_io.save_data(rx.iq)
    if P.SAVE_DEMOD and irx == 0:
        P.demod_io.save_data(am)
    return am


def push_audio(P):
    """The player side of reference receiver.py:153-225 (the payloads come from audio_out above): one push per player,
    playback started once DELAY samples are queued."""
    if not P.audio_playback:
        return
    payload = audio_out(P, [rx.am for rx in P.rx])
    for player, a in zip(P.players, payload):
        player.rb.push(a)
        if player and not player.active:
            player.start_playback(P.DELAY, False)


class SDR_EXECUTIVE:
    """Replay-mode executive (reference receiver.py:408-782, MP_SCHEME 1 data plane) over ONE ReceiverBank: the chunk is
    copied to the device once and every receiver is served from that read; ``P.rx[irx]`` are BankReceiver rows.
    WFM / WFM2 (demodulate first, then resample, gui.py:1703) run through per-receiver ``sig_proc.Receiver`` objects.

    ``Run(raw)`` keeps the loop's observable behaviour: strict ``<`` chunking with the stale last chunk demodulated once
    more at EOF (:544-559, :715-725), the optional replay pre-mixer ``P.lo`` (:552-555), mode changes taking effect
    between chunks with 'FM' read as 'NFM' and the AGC / PLL of receiver 0 restarted (:634-650), audio routing, the
    RF-PSD and raw-IQ taps (:742-758), the DURATION stop (:764) and ``SHUT_DOWN`` after a replay (:776-777)."""

    def __init__(self, P, GUI=False, max_chunks_per_call=1):
        from . import sig_proc as dsp
        self.P = P
        P.SDR_EXEC = self
        P.RX_DONE = False
        P.nchunks = 0
        self.raw = None
        self.praw = 0
        self._seen = None
        self._pw = None
        self.create_SDR()
        self.create_Receivers(max_chunks_per_call)
        self.create_Audio_Players()
        self.x = np.zeros(P.IN_CHUNK_SIZE, np.complex64)            # receiver.py:445
        if not hasattr(P, 'lo'):
            P.lo = dsp.signal_generator(0, P.IN_CHUNK_SIZE, P.SRATE, True)      # receiver.py:822

    # -- construction ---------------------------------------------------------------------------------------------
    def create_SDR(self):
        """Replay source, reference receiver.py:808-822.  The reference decides `FS_OUT = SRATE` with
        ``if P.REPLAY.find('baseband_iq'):`` — the truthiness of str.find, i.e. for every name that does NOT start with
        'baseband_iq'; kept literally (fileio.open_replay(literal=False) implements the evident intent instead)."""
        P = self.P
        if getattr(P, 'REPLAY_MODE', False) and getattr(P, 'REPLAY', None) and getattr(P, 'sdr', None) is None:
            from .fileio import open_replay
            open_replay(P, P.REPLAY, literal=True)

    def create_Receivers(self, max_chunks_per_call=1):
        P = self.P
        self.wfm = any(design.per_rx(P.MODE, i) in ('WFM', 'WFM2') for i in range(P.NUM_RX))
        if self.wfm:
            from . import sig_proc as dsp
            self.bank = None
            P.rx = [dsp.Receiver(P, f, i, str(i + 1)) for i, f in enumerate(receiver_offsets(P))]
            return
        self.bank = ReceiverBank(P, receiver_offsets(P), max_in=int(P.IN_CHUNK_SIZE) * int(max_chunks_per_call))
        P.rx = [BankReceiver(self, self.bank, P, i) for i in range(P.NUM_RX)]

    def create_Audio_Players(self):
        """Players are the caller's (audio sink is out of scope): P.players is used as found; with P.player_factory set,
        one player per NUM_PLAYERS is built on a ring buffer of RB_SIZE like reference receiver.py:838-850."""
        P = self.P
        make = getattr(P, 'player_factory', None)
        if make is not None:
            from . import sig_proc as dsp
            P.players = [make(P, P.FS_OUT + getattr(P, 'FS_OUT_CORR', 0), dsp.ring_buffer2('Audio' + str(i + 1), P.RB_SIZE), i)
                         for i in range(P.NUM_PLAYERS)]

    # -- per chunk ------------------------------------------------------------------------------------------------
    def _process_chunk(self, x):
        key = (id(x), self.P.nchunks)
        if self._seen == key:
            return
        self._seen = key
        self._pw = None
        am, iq, dc = self.bank.process_host(x)
        for i, rx in enumerate(self.P.rx):
            rx.am, rx.iq, rx.am_dc = am[i], iq[i], dc[i]

    def _chunk_power(self, x):
        if self._pw is None:
            b = self.bank
            n = len(x)
            d_in = ctypes.c_void_p()
            check(b.lib.pysdr_bank_host_chunk_ptr(b.h, ctypes.byref(d_in)))
            check(b.lib.pysdr_mean_power(d_in, n, ctypes.c_void_p(b.scratch1().data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
            self._pw = float(b.scratch1().item())
        return self._pw

    def read_chunk(self):
        P = self.P
        C = int(P.IN_CHUNK_SIZE)
        if self.praw + C < len(self.raw):                           # receiver.py:544 (strict)
            x1 = self.raw[self.praw:self.praw + C]
            self.praw += C
            self.x = P.lo.quad_mixer(x1) if P.lo.fo != 0 else x1    # receiver.py:552-555
        else:
            P.RX_DONE = True                                        # the stale self.x is demodulated once more

    def mode_freq_change(self):
        P = self.P
        if getattr(P, 'MODE_CHANGE', False):
            if P.NEW_MODE == 'FM':                                  # receiver.py:640-641
                P.NEW_MODE = 'NFM'
            if P.MODE != P.NEW_MODE:
                P.MODE = P.NEW_MODE
                if getattr(P, 'gui', None) is not None and P.MP_SCHEME == 1:
                    P.gui.ModeSelect(-1)
                if any(design.per_rx(P.MODE, i) in ('WFM', 'WFM2') for i in range(P.NUM_RX)) != self.wfm:
                    self.create_Receivers()                         # the order of detection and rate reduction changes
                P.rx[0].agc.reset()                                 # receiver.py:648-649
                P.rx[0].demod.am_pll.reset()
            P.MODE_CHANGE = False

    def Run(self, raw=None, sink=None):
        """raw: host complex64 capture (default: ``P.sdr.read_data()``, receiver.py:531).  sink(irx, am*af_gain, am_dc, iq),
        if given, is called per receiver per chunk.  Returns the number of loop iterations."""
        P = self.P
        self.raw = raw if raw is not None else P.sdr.read_data()
        self.praw = 0
        dt = float(P.IN_CHUNK_SIZE) / P.SRATE
        t = 0.
        iters = 0
        P.RX_DONE = False

        def stopped():
            return bool(getattr(P, 'Stopper', None)) and P.Stopper.isSet()
        while not P.RX_DONE:
            t += dt
            P.nchunks += 1
            iters += 1
            if stopped():
                P.RX_DONE = True
                break
            self.read_chunk()
            self.mode_freq_change()
            dcs = [demodulate_data(P, self.x, irx) for irx in range(P.NUM_RX)]
            push_audio(P)
            if sink is not None:
                for irx in range(P.NUM_RX):
                    sink(irx, P.rx[irx].am * af_gain(P, irx), dcs[irx], P.rx[irx].iq)
            if P.SHOW_RF_PSD:                                       # receiver.py:742-752
                if P.MP_SCHEME == 1:
                    P.rb_rf.push(self.x)
                else:
                    P.rf_psd_Q.put(self.x)
            if P.SAVE_IQ:                                           # receiver.py:755-758
                P.raw_iq_io.save_data(self.x, VERBOSITY=0)
            P.RX_DONE = P.RX_DONE or t >= P.DURATION or stopped()   # receiver.py:764
        for io in (getattr(P, 'raw_iq_io', None), getattr(P, 'baseband_iq_io', None), getattr(P, 'demod_io', None)):
            if io:
                io.close()                                          # quit_rx, receiver.py:495-500
        if getattr(P, 'REPLAY_MODE', False):
            P.SHUT_DOWN = True                                      # receiver.py:776-777
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
