"""Synthetic IQ of the benchmark shapes: complex Gaussian noise at -40 dBFS plus one test
carrier per receiver — AM (1 kHz tone, m=0.5), NFM (1 kHz tone, +-3 kHz deviation), USB (two-tone
700/1900 Hz), CW (keyed carrier, 20 wpm dots), amplitude 0.1 each.  Generated with torch so the same code
fills a host array for parity tests or a device tensor for the bench (no datasets, no network)."""
import math

import torch


def synth_iq(n, srate, offsets_hz, modes, seed=1234, device="cpu", n0=0, noise_db=-40.0, amp=0.1, block=1 << 22):
    """complex64[n] starting at absolute sample index n0 (carriers are phase-continuous functions of the
    absolute index, so time shards of one capture can be generated independently)."""
    dev = torch.device(device)
    out = torch.empty(n, dtype=torch.complex64, device=dev)
    g = torch.Generator(device=dev)
    sigma = 10.0 ** (noise_db / 20.0) / math.sqrt(2.0)
    n0 = int(n0)
    for kb in range(n0 // block, (n0 + n - 1) // block + 1):          # absolute blocks: shards of one capture agree
        lo, hi = max(n0, kb * block), min(n0 + n, (kb + 1) * block)
        b0, b1 = lo - n0, hi - n0
        g.manual_seed(seed + 7919 * kb)
        t = torch.arange(lo, hi, device=dev, dtype=torch.float64) / float(srate)
        z = torch.randn(block, 2, generator=g, device=dev, dtype=torch.float32)[lo - kb * block:hi - kb * block] * sigma
        acc = torch.view_as_complex(z.contiguous()).to(torch.complex128)
        for f0, mode in zip(offsets_hz, modes):
            w = 2.0 * math.pi * t
            if mode in ('AM', 'AM-Synch'):
                base = 1.0 + 0.5 * torch.sin(w * 1000.0)
                ph = torch.zeros_like(t)
            elif mode == 'NFM':
                base = torch.ones_like(t)
                ph = (3000.0 / 1000.0) * torch.sin(w * 1000.0)          # beta = dev/fm
            elif mode in ('USB', 'SSB', 'LSB'):
                sgn = -1.0 if mode == 'LSB' else 1.0
                acc = acc + 0.5 * amp * (torch.exp(1j * (w * (f0 + sgn * 700.0))) + torch.exp(1j * (w * (f0 + sgn * 1900.0))))
                continue
            elif mode == 'CW':
                dot = 1.2 / 20.0                                         # 20 wpm
                base = ((t / dot).floor() % 2 == 0).to(torch.float64)
                ph = torch.zeros_like(t)
            else:                                                        # IQ / RTTY: plain carrier 1 kHz off centre
                base = torch.ones_like(t)
                ph = w * 1000.0
            acc = acc + amp * base * torch.exp(1j * (w * f0 + ph))
        out[b0:b1] = acc.to(torch.complex64)
    return out
