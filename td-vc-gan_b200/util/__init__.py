"""Drop-in `util` package (reference: util/__init__.py).  Factories and small signal helpers used by
train.py / generate_with_target.py; the arithmetic-heavy pieces live in tdvc.  Sub-modules outside the
hot-path scope (util.yin, util.crepe, util.audio, util.hparams) resolve to the reference's own files when
a reference checkout is present (TDVC_REFERENCE or /root/reference); nothing is copied."""
import math
import os as _os

import numpy as np
import torch
import torch.nn.functional as F

_ref = _os.environ.get("TDVC_REFERENCE", "/root/reference")
_ref_util = _os.path.join(_ref, "util")
if _os.path.isdir(_ref_util) and _ref_util not in __path__:
    __path__.append(_ref_util)

from model.conditional_instance_norm import ConditionalInstanceNorm  # noqa: E402
from tdvc.layers import Identity, InstanceNorm1d  # noqa: E402


def get_norm_layer(norm):
    """reference util/__init__.py:8-14 -- returns the layer *class*."""
    if norm is None:
        return Identity
    if norm == 'instance_norm':
        return InstanceNorm1d
    if norm == 'conditional_instance_norm':
        return ConditionalInstanceNorm


def _no_weight_norm(module):
    return module


def _weight_norm(module):
    """Stands where the reference returns nn.utils.weight_norm: tdvc layers take a flag at construction,
    so this marker is what their constructors look for; applied to a foreign torch module it falls back
    to torch's own reparametrisation."""
    if hasattr(module, "weight_norm"):
        raise RuntimeError("tdvc layers take weight_norm=True at construction")
    return torch.nn.utils.weight_norm(module)


_weight_norm.tdvc_weight_norm = True


def get_weight_norm(norm):
    """reference util/__init__.py:16-20."""
    if norm is None:
        return _no_weight_norm
    if norm == 'weight_norm':
        return _weight_norm


def f0_to_excitation(f0, step_size, sampling_rate=16000, linear=True):
    """F0 track [B,1,frames] -> sine + noise excitation [B,1,(frames-1)*step_size]
    (reference util/__init__.py:22-50; host-side torch, runs once per batch)."""
    f0 = f0[:, :, :-1]
    sin_gain, noise_std = 0.1, 0.003
    noise_gain = sin_gain / (3 * noise_std)
    omega = 2 * torch.pi * f0 / sampling_rate
    up = F.interpolate(omega, scale_factor=step_size, mode='nearest')
    if linear:
        up_lin = F.interpolate(omega, scale_factor=step_size, mode='linear')
        voiced_both = F.interpolate(torch.log(omega), scale_factor=step_size, mode='linear') != -torch.inf
        up[voiced_both] = up_lin[voiced_both]
    phase = torch.cumsum(up, -1)
    start = torch.rand(1, device=f0.device) * 2 * torch.pi
    exc = sin_gain * torch.sin(phase + start) + torch.randn(phase.shape, device=f0.device) * noise_std
    unvoiced = up == 0
    exc[unvoiced] = torch.randn(exc[unvoiced].shape, device=f0.device) * noise_std * noise_gain
    return exc


def eq_rms(signal, target_rms):
    rms = np.sqrt((signal ** 2).mean())
    return signal * (10 ** (target_rms / 20) / rms)


def eq_rms_signals(signal_eq, signal_tgt):
    rms_eq = np.sqrt((signal_eq ** 2).mean())
    rms_tgt = np.sqrt((signal_tgt ** 2).mean())
    return signal_eq * rms_tgt / (rms_eq + 1e-8)


def load_possible(model, state_dict):
    """Permissive checkpoint load (reference util/__init__.py:64-89): copies matching tensors, and the
    overlapping slice of size-mismatched ones, into the model's own state_dict tensors."""
    own = model.state_dict()
    messages = {'matched': [], 'mismatched_size': [], 'unmatched_keys': [], 'missing_keys': []}
    for name, src in state_dict.items():
        if name not in own:
            messages['unmatched_keys'].append(name)
            continue
        dst = own[name]
        if src.shape == dst.shape:
            own[name] = src
            messages['matched'].append(name)
        else:
            window = tuple(slice(0, min(a, b)) for a, b in zip(dst.shape, src.shape))
            dst[window] = src[window]
            messages['mismatched_size'].append(name)
    messages['missing_keys'] = [name for name in own if name not in state_dict]
    return messages


def roll_batches(input, shifts, dim):
    """Per-sample circular shift along `dim` (reference util/__init__.py:91-102)."""
    n = input.shape[dim]
    shape = [1] * input.ndim
    shape[dim] = n
    idx = torch.arange(n, device=input.device).view(shape).expand(input.shape)
    bshape = [1] * input.ndim
    bshape[0] = -1
    idx = (idx - shifts.view(bshape)) % n
    return torch.gather(input, dim, idx)


def kaiser_filter(L, w):
    """L+1 tap Kaiser(beta 2.5) windowed sinc, cutoff w (reference util/__init__.py:104-113), [1,1,L+1]."""
    half = L // 2
    n = torch.arange(-half, half + 1).float() if L % 2 == 0 else torch.arange(-L // 2, L // 2 + 1).float()
    taps = torch.sin(math.pi * w * n) / (math.pi * n + 1e-8)
    taps[n.shape[0] // 2] = w
    taps = taps * torch.kaiser_window(L + 1, False, 2.5)
    taps = taps / torch.sum(taps)
    return taps.view(1, 1, -1)
