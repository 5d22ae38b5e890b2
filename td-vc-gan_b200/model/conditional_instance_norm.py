"""Speaker-conditioned instance normalisation on the tdvc kernels.

Drop-in for the reference's `model/conditional_instance_norm.py:4-19`: `(1 + gamma) * InstanceNorm1d(x) + beta` with
(gamma, beta) predicted from the conditioning code.  The statistics and the affine (+ the LeakyReLU of the layer that
follows, when the caller asks for it) are two kernel launches instead of five elementwise passes.
"""
import torch.nn as nn

from tdvc import ops
from tdvc.layers import Conv1d, InstanceNorm1d, Linear


class ConditionalInstanceNorm(nn.Module):
    """Both predictors are always allocated, like the reference, so the checkpoint keys match:
    `embedding.{weight,bias}` (a Linear for a static [B, n_cond] code) and `embedding_conv.{weight,bias}`
    (a k=5 'same' Conv1d for a time-varying [B, n_cond + 1, T] code)."""

    def __init__(self, n_channel, n_cond, n_conf_var=0):
        super().__init__()
        self.norm = InstanceNorm1d(n_channel)
        self.embedding = Linear(n_cond, 2 * n_channel)
        self.embedding_conv = Conv1d(n_cond + 1, 2 * n_channel, kernel_size=5, padding="same")

    def _gamma_beta(self, c):
        """[B, 2C, 1] from a static code, [B, 2C, T] from a time-varying one."""
        return self.embedding(c).unsqueeze(2) if c.dim() == 2 else self.embedding_conv(c)

    def forward(self, x, c, out_slope: float = 1.0):
        return ops.cond_instance_norm(x, self._gamma_beta(c), self.norm.eps, out_slope)
