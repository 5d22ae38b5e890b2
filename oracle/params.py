"""Deterministic parameter / input synthesis shared by the golden generator and the tests.

TEST INFRASTRUCTURE (see oracle/tdvc_oracle.py header).  Parameters are a pure function of
(checkpoint key, shape, seed) so that golden vectors produced from the real reference in the
build container can be reproduced on the GPU box, where /root/reference does not exist, without
shipping 130 MB of weights.  torch's CPU Philox/MT generators are bit-stable for a fixed torch
build; the GPU box runs the same image.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Sequence

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return g


def make_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Kaiming-uniform-like values per tensor (bound 1/sqrt(fan_in)); weight_g = ||v|| * U(0.8,1.2) so the
    weight-norm scale is exercised away from its g=||v|| initial point; biases U(+-bound/2)."""
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        if name.endswith("weight_g"):
            continue
        shape = tuple(shape)
        g = _gen(name, seed)
        if name.endswith("bias"):
            wname = name[:-4] + ("weight_v" if name[:-4] + "weight_v" in shapes else "weight")
            wshape = tuple(shapes[wname])
            # fan_in of the owning weight; ConvTranspose weights are [Cin, Cout, K] but the bound only
            # needs to be "reasonable", so the same rule is used for every layout.
            fan = max(1, int(math.prod(wshape[1:])))
            b = 0.5 / math.sqrt(fan)
        else:
            fan = max(1, int(math.prod(shape[1:])))
            b = 1.0 / math.sqrt(fan)
        sd[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
    for name, shape in shapes.items():
        if not name.endswith("weight_g"):
            continue
        v = sd[name[:-1] + "v"].double()
        n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(tuple(shape))
        u = torch.rand(tuple(shape), generator=_gen(name, seed), dtype=torch.float64) * 0.4 + 0.8
        sd[name] = (n * u).to(dtype)
    # keep the caller's key order
    return {k: sd[k] for k in shapes}


def make_batch(B: int, T: int, nspk: int, seed: int = 1234, n_neg: int = 100, frames_div: int = 320,
               dtype=torch.float64, permute: bool = True) -> dict:
    """Synthetic 16 kHz batch of SURVEY.md 8(d): x = 0.05 sin(2 pi f0 t) + 0.01 eps, f0 ~ U(100,300) Hz.

    The excitation c_var the Decoder needs is synthesised directly (sine at the known f0 plus noise,
    amplitude as util/__init__.py:25-27) instead of running a pitch tracker: the hot path only sees the
    resulting [B,1,T] tensor and both sides of every comparison get the same one."""
    g = torch.Generator(); g.manual_seed(seed)
    t = torch.arange(T, dtype=torch.float64) / 16000.0
    f0 = torch.rand(B, 1, 1, generator=g, dtype=torch.float64) * 200 + 100
    x = 0.05 * torch.sin(2 * math.pi * f0 * t) + 0.01 * torch.randn(B, 1, T, generator=g, dtype=torch.float64)
    xc = x + 0.005 * torch.randn(B, 1, T, generator=g, dtype=torch.float64)
    lab_s = torch.randint(0, nspk, (B,), generator=g)
    perm = torch.randperm(B, generator=g) if permute else torch.arange(B)
    lab_t = lab_s[perm]
    f0_t = f0[perm]

    def excite(f):
        ph = torch.rand(1, generator=g, dtype=torch.float64) * 2 * math.pi
        return 0.1 * torch.sin(2 * math.pi * f * t + ph) + 0.003 * torch.randn(B, 1, T, generator=g, dtype=torch.float64)

    Tf = T // frames_div
    from . import tdvc_oracle as O
    neg = [O.contrastive_raw_draws(B, Tf, n_neg, g) for _ in range(2)] if Tf > 1 else []
    return {
        "signal_real": x.to(dtype), "signal_corrupted": xc.to(dtype),
        "label_src": lab_s, "label_tgt": lab_t,
        "c_f0_conv": excite(f0_t).to(dtype), "c_f0_src": excite(f0).to(dtype),
        "neg_idx": neg,
    }
