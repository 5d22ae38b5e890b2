"""`util.dsp` of the reference exposes one function, the Kaiser-windowed sinc low-pass design used for the
discriminators' half-band decimator and the generator's sub-scale filters (reference util/dsp.py:5-16).  It is host
code that runs once at module construction; kept in fp32 with the reference's operation order so that the taps -- which
are baked into the convolution weights of `tdvc_conv1d_fwd` -- are bit-identical (tests/test_host_cpu.py)."""
import math

import torch

__all__ = ["kaiser_filter"]


def _windowed_sinc(length: int, cutoff: float, beta: float) -> torch.Tensor:
    centre = (length - 1) // 2
    n = torch.arange(-centre, centre + 1).float()
    taps = torch.sin(math.pi * cutoff * n) / (math.pi * n + 1e-8)
    taps[centre] = cutoff                                     # the n = 0 sample of the sinc
    taps = taps * torch.kaiser_window(length, False, beta)
    return taps / torch.sum(taps)                             # unit DC gain


def kaiser_filter(L, fc, beta=2.5):
    """Odd-length linear-phase low-pass with cutoff `fc` (1.0 = Nyquist) and Kaiser parameter `beta`."""
    if L % 2 == 0:
        raise Exception("Even length filter not implemented")
    return _windowed_sinc(int(L), float(fc), beta)
