"""Gradient reversal layer (reference: model/grad_rev.py:3-18): identity forward, negated gradient."""
import torch

from tdvc import ops


class GradRevFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad):
        return ops.add_scale(grad, alpha=-1.0)      # lamb = 1 in the reference


class GradRevLayer(torch.nn.Module):
    def __init__(self, lamb=1):
        super().__init__()
        self.lamb = lamb

    def forward(self, x):
        return GradRevFunction.apply(x)
