"""Parameter-holding layers with the reference's checkpoint layout.

`Conv1d` / `ConvTranspose1d` here stand where the reference has `weight_norm(nn.Conv1d(...))`
or a plain `nn.Conv1d`: same constructor meaning, same parameter names (`bias`, `weight_g`,
`weight_v` for old-style weight norm -- registered in that order -- or `weight`, `bias`), same
default initialisation (and RNG consumption, so `torch.manual_seed(s)` gives the reference's
initial weights).  Their forward runs the tdvc CUDA ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _same_padding(kernel_size: int, dilation: int) -> int:
    total = dilation * (kernel_size - 1)
    if total % 2:
        raise ValueError("padding='same' needs an odd effective kernel (as every call site in the reference has)")
    return total // 2


class _ConvBase(nn.Module):
    transposed = False

    def _init_params(self, ref: nn.Module, weight_norm: bool):
        """Adopt the tensors of a freshly initialised torch layer (exactly the reference's init)."""
        w = ref.weight.detach()
        b = ref.bias.detach() if ref.bias is not None else None
        self.weight_norm = bool(weight_norm)
        if weight_norm:
            # torch.nn.utils.weight_norm: bias stays first, then weight_g (= ||v|| over dims != 0), weight_v
            self.bias = nn.Parameter(b) if b is not None else None
            g = w.reshape(w.shape[0], -1).norm(dim=1).reshape(w.shape[0], *([1] * (w.dim() - 1)))
            self.weight_g = nn.Parameter(g)
            self.weight_v = nn.Parameter(w)
        else:
            self.weight = nn.Parameter(w)
            self.bias = nn.Parameter(b) if b is not None else None

    def effective_weight(self) -> torch.Tensor:
        if self.weight_norm:
            return ops.weight_norm(self.weight_v, self.weight_g)
        return self.weight


class Conv1d(_ConvBase):
    """nn.Conv1d (optionally under old-style weight norm) on the tdvc kernels.

    forward(x, in_slope=1, out_act=None, residual=None) fuses the LeakyReLU that precedes the conv in
    the reference's Sequentials, the bias, a residual add and an output activation into one kernel."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode="zeros", weight_norm=False):
        super().__init__()
        if padding == "same":
            padding = _same_padding(kernel_size, dilation)
        if padding_mode not in ("zeros", "reflect"):
            raise ValueError(f"padding_mode {padding_mode!r} not supported")
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.padding_mode = padding_mode
        ref = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=0, dilation=dilation,
                        groups=groups, bias=bias)
        self._init_params(ref, weight_norm)

    def forward(self, x, in_slope: float = 1.0, out_act=None, out_slope: float = 0.2, residual=None):
        return ops.conv1d(x, self.effective_weight(), self.bias, stride=self.stride, padding=self.padding,
                          dilation=self.dilation, groups=self.groups, reflect=self.padding_mode == "reflect",
                          in_slope=in_slope, out_act=out_act, out_slope=out_slope, residual=residual)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, groups={self.groups}, "
                f"padding_mode={self.padding_mode}, weight_norm={self.weight_norm}")


class ConvTranspose1d(_ConvBase):
    transposed = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bias=True,
                 weight_norm=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.output_padding = stride, padding, output_padding
        ref = nn.ConvTranspose1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                 output_padding=output_padding, bias=bias)
        self._init_params(ref, weight_norm)

    def forward(self, x, in_slope: float = 1.0):
        if in_slope != 1.0:
            x = ops.leaky_relu(x, in_slope)
        return ops.conv_transpose1d(x, self.effective_weight(), self.bias, stride=self.stride,
                                    padding=self.padding, output_padding=self.output_padding)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, output_padding={self.output_padding}, weight_norm={self.weight_norm}")


class Linear(nn.Module):
    """nn.Linear on the conv kernel (a length-1 convolution)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        ref = nn.Linear(in_features, out_features)
        self.weight = nn.Parameter(ref.weight.detach())
        self.bias = nn.Parameter(ref.bias.detach())

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class LeakyReLU(nn.Module):
    """Slot-compatible stand-in for nn.LeakyReLU inside the reference's Sequential/ModuleList layouts
    (keeps child indices, hence state_dict keys).  Parents fuse it into the next conv; called on its own
    it runs the standalone kernel."""

    def __init__(self, negative_slope=0.01, inplace=False):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return ops.leaky_relu(x, self.negative_slope)


class Tanh(nn.Module):
    """Slot stand-in for nn.Tanh; parents fuse it into the preceding conv's epilogue."""

    def forward(self, x):  # pragma: no cover - only reached if someone calls the slot directly
        raise RuntimeError("tdvc.layers.Tanh is fused into the preceding convolution")


class Identity(nn.Module):
    """nn.Identity(*args) -- what util.get_norm_layer(None) returns in the reference."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


class InstanceNorm1d(nn.Module):
    """nn.InstanceNorm1d(num_features, eps) with affine=False, no running stats (the only form the
    reference constructs: util/__init__.py:11-12).  NB the Decoder calls norm_layer(ch, conditional_dim),
    which lands in `eps` exactly as it does with torch's class (model/generator.py:308,342)."""

    def __init__(self, num_features, eps=1e-5, *args, **kwargs):
        super().__init__()
        self.num_features, self.eps = num_features, float(eps)

    def forward(self, x, out_slope: float = 1.0):
        return ops.instance_norm(x, self.eps, out_slope)


def maybe_weight_norm(weight_norm) -> bool:
    """util.get_weight_norm(None | 'weight_norm') as a flag for the layer constructors."""
    if weight_norm is None or weight_norm is False:
        return False
    if weight_norm is True or weight_norm == "weight_norm":
        return True
    raise ValueError(f"unknown weight_norm {weight_norm!r}")
