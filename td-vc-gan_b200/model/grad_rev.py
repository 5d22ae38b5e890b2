"""Gradient-reversal layer in front of the latent classifier -- drop-in names for the reference's `model/grad_rev.py`.
The op itself lives with the other autograd ops (`tdvc.ops.grad_reverse`: identity forward, gradient negated by the
`add3_scale` kernel)."""
import torch.nn as nn

from tdvc import ops

GradRevFunction = ops._GradReverse        # reference name of the autograd Function (model/grad_rev.py:3)


class GradRevLayer(nn.Module):
    """`lamb` is stored as in the reference, whose backward does not use it (model/grad_rev.py:8-10 always returns
    `-grad`); the gradient scale applied here is therefore the constant -1."""

    def __init__(self, lamb=1):
        super().__init__()
        self.lamb = lamb

    def forward(self, x):
        return ops.grad_reverse(x)
