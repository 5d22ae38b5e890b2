"""Kaiser-windowed sinc low-pass design (reference: util/dsp.py:5-16).  Host-side, runs once at
module construction; fp32 like the reference so the taps are bit-identical."""
import math

import torch


def kaiser_filter(L, fc, beta=2.5):
    if L % 2 == 0:
        raise Exception("Even length filter not implemented")
    half = (L - 1) // 2
    n = torch.arange(-half, half + 1).float()
    taps = torch.sin(math.pi * fc * n) / (math.pi * n + 1e-8)
    taps[half] = fc
    taps = taps * torch.kaiser_window(L, False, beta)
    return taps / torch.sum(taps)
