"""YIN pitch estimation (reference util/yin.py; de Cheveigne & Kawahara 2002, https://asa.scitation.org/doi/10.1121/1.1458024).

`estimate` keeps the reference's signature and behaviour.  CUDA fp32 signals with the hard search (`soft=False`, the form
train.py uses) run on one tdvc kernel (csrc/yin.cu: difference function by direct fp64 sums, cumulative-mean normalisation,
threshold / first-local-minimum search, one CTA per frame); everything else -- CPU tensors, lists, `soft=True` -- runs the
torch formulation below (FFT autocorrelation, as the reference computes it).
"""
import typing as T

import numpy as np
import torch
import torch.nn.functional as F


def _geometry(sample_rate, pitch_min, pitch_max, frame_stride):
    tau_min = int(sample_rate / pitch_max)
    tau_max = int(sample_rate / pitch_min)
    return tau_min, tau_max, 2 * tau_max, int(frame_stride * sample_rate)      # a window holds two periods of pitch_min


def estimate(signal: T.Union[T.List, np.ndarray, torch.Tensor], sample_rate: float, pitch_min: float = 20,
             pitch_max: float = 20000, frame_stride: float = 0.01, threshold: float = 0.1, soft: bool = False) -> torch.Tensor:
    """Pitch per frame of `frame_stride` seconds ([..., frames]); 0 where no period passes the threshold."""
    signal = torch.as_tensor(signal)
    tau_min, tau_max, W, hop = _geometry(sample_rate, pitch_min, pitch_max, frame_stride)
    if signal.is_cuda and signal.dtype == torch.float32 and not soft and signal.numel() > 0:
        return _estimate_device(signal, sample_rate, tau_min, tau_max, W, hop, threshold)
    cmdf = _cmdf(_frames(signal, W, hop), tau_max)[..., tau_min:]
    tau = _soft_period(cmdf, threshold) if soft else _first_dip(cmdf, tau_max, threshold)
    zero = torch.tensor(0, device=signal.device).type(signal.dtype)
    return torch.where(tau > 0, sample_rate / (tau + tau_min + 1).type(signal.dtype), zero)


def _estimate_device(signal, sample_rate, tau_min, tau_max, W, hop, threshold):
    from tdvc import _lib
    lead = signal.shape[:-1]
    x = signal.reshape(-1, signal.shape[-1]).contiguous()
    B, Tn = x.shape
    n_frames = (max(Tn, W) - 1) // hop + 1
    out = torch.empty(B, n_frames, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().tdvc_yin_estimate(x.data_ptr(), out.data_ptr(), B, Tn, n_frames, W, hop, tau_min, tau_max,
                                             float(threshold), float(sample_rate),
                                             torch.cuda.current_stream(x.device).cuda_stream), "yin_estimate")
    return out.reshape(lead + (n_frames,))


def _frames(signal, W, hop):
    """overlapping windows centred on multiples of the hop: at least one window, W/2 zeros before, W/2 - 1 after"""
    short = W - signal.shape[-1]
    if short > 0:
        signal = F.pad(signal, [0, short])
    return F.pad(signal, [W // 2, W // 2 - 1]).unfold(-1, W, hop)


def _cmdf(frames, tau_max):
    """cumulative-mean-normalised difference function of every window, lags 1 .. tau_max - 1"""
    W = frames.shape[-1]
    n_fft = 2 ** (-int(-np.log(W) // np.log(2)) + 1)             # >= 2 W: the circular correlation is the linear one
    spec = torch.fft.rfft(frames, n_fft, dim=-1)
    acorr = torch.fft.irfft(spec * spec.conj())[..., :tau_max]
    energy = F.pad((frames * frames).cumsum(-1), [1, 0])         # energy[i] = sum of the first i squares
    both = energy[..., -1:] + (energy.flip(-1)[..., :tau_max] - energy[..., :tau_max])
    d = (both - 2 * acorr)[..., 1:]
    lag = torch.arange(1, tau_max, device=d.device)
    return d * lag / torch.maximum(d.cumsum(-1), torch.tensor([1e-5], device=d.device))


def _first_dip(cmdf, tau_max, threshold):
    """index of the first local minimum at or after the first value below the threshold (0: none -- also when that first value
    is at index 0, as in the reference)"""
    below = (cmdf < threshold).int().argmax(-1, keepdim=True)
    below = torch.where(below > 0, below, tau_max)
    after = torch.arange(cmdf.shape[-1], device=below.device) >= below
    rising = F.pad(cmdf.diff() >= 0.0, [0, 1], value=1)
    return (after & rising).int().argmax(-1)


def _soft_period(cmdf, threshold):
    """softmax(-100 c)-weighted mean lag, zero for windows without any value below the threshold"""
    voiced = (cmdf < threshold).any(dim=-1).int()
    w = F.softmax(-cmdf * 100, dim=-1)
    return (w * torch.arange(cmdf.shape[-1], device=w.device)).sum(-1) * voiced
