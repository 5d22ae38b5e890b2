"""Shared helpers for the parity tests."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_shapes(g):
    return {str(k): ast.literal_eval(str(s)) for k, s in zip(g["keys"], g["shapes"])}


def relerr(a, b):
    """max |a-b| / max |b|  (the 'max-abs-normalised' error of SURVEY.md 8c)."""
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    den = b.abs().max().item()
    # 1e-12 absolute floor: tensors that are analytically zero (e.g. a bias feeding an instance norm)
    # hold only rounding noise on both sides
    return max((a - b).abs().max().item() - 1e-12, 0.0) / max(den, 1e-30)


def stats(t):
    t = t.detach().double().reshape(-1).cpu()
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()])
