"""ConditionalInstanceNorm on the tdvc kernels (reference: model/conditional_instance_norm.py:4-19)."""
import torch.nn as nn

from tdvc import ops
from tdvc.layers import Conv1d, InstanceNorm1d, Linear


class ConditionalInstanceNorm(nn.Module):
    """(1 + gamma) * InstanceNorm1d(x) + beta, (gamma, beta) from a Linear on a [B, n_cond] code or from
    a k=5 'same' Conv1d on a time-varying [B, n_cond+1, T] code.  Both embedding layers are always
    allocated, as in the reference, so the state_dict keys match: embedding.{weight,bias},
    embedding_conv.{weight,bias}."""

    def __init__(self, n_channel, n_cond, n_conf_var=0):
        super().__init__()
        self.norm = InstanceNorm1d(n_channel)
        self.embedding = Linear(n_cond, n_channel * 2)
        self.embedding_conv = Conv1d(n_cond + 1, n_channel * 2, kernel_size=5, padding="same")

    def forward(self, x, c, out_slope: float = 1.0):
        if len(c.shape) == 2:
            h = self.embedding(c).unsqueeze(2)
        else:
            h = self.embedding_conv(c)
        # instance-norm statistics + affine (+ optional LeakyReLU of the following layer) in two kernels
        return ops.cond_instance_norm(x, h, self.norm.eps, out_slope)
