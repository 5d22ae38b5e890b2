"""TEST INFRASTRUCTURE: torch (CPU) stand-ins for the entry points of tdvc.ops, installed by monkeypatching.

The product has no CPU path (tdvc.ops raises on non-CUDA tensors).  The host-side logic above the kernels -- which
passes tdvc.train_step.TrainStep stacks along the batch, which rows feed which loss, the order of the optimiser
updates, the CUDA-graph segment plan -- can still be checked on a machine without a GPU by swapping every op for
its plain torch definition.  Only tests may import this module."""
import contextlib

import torch
import torch.nn.functional as F


def _act(y, act, slope):
    if act in (None, "none"):
        return y
    if act == "lrelu":
        return F.leaky_relu(y, slope)
    if act == "tanh":
        return torch.tanh(y)
    raise ValueError(act)


def conv1d(x, weight, bias=None, *, stride=1, padding=0, dilation=1, groups=1, reflect=False, in_slope=1.0, out_act=None,
           out_slope=0.2, residual=None):
    if in_slope != 1.0:
        x = F.leaky_relu(x, in_slope)
    if reflect and padding > 0:
        x = F.pad(x, (padding, padding), mode="reflect")
        padding = 0
    y = F.conv1d(x, weight, bias, stride=stride, padding=padding, dilation=dilation, groups=groups)
    if residual is not None:
        y = y + residual
    return _act(y, out_act, out_slope)


def linear(x, weight, bias=None):
    return F.linear(x, weight, bias)


def conv_transpose1d(x, weight, bias=None, *, stride=1, padding=0, output_padding=0):
    return F.conv_transpose1d(x, weight, bias, stride=stride, padding=padding, output_padding=output_padding)


def weight_norm(v, g):
    return g * v / v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)


def film(h, gb):
    gamma, beta = gb.chunk(2, dim=1)
    return h * (1 + gamma) + beta


def add_scale(a, b=None, c=None, alpha=1.0):
    y = a
    if b is not None:
        y = y + b
    if c is not None:
        y = y + c
    return y * alpha


def cond_concat(c, e):
    return torch.cat([c.unsqueeze(2).expand(-1, -1, e.shape[2]), e], dim=1)


def cond_concat_front(x, c):
    return torch.cat([x, c.unsqueeze(2).expand(-1, -1, x.shape[2])], dim=1)


def instance_norm(x, eps=1e-5, out_slope=1.0):
    y = F.instance_norm(x, eps=eps)
    return F.leaky_relu(y, out_slope) if out_slope != 1.0 else y


def cond_instance_norm(x, gb, eps=1e-5, out_slope=1.0):
    gamma, beta = gb.chunk(2, dim=1)
    y = (1 + gamma) * F.instance_norm(x, eps=eps) + beta
    return F.leaky_relu(y, out_slope) if out_slope != 1.0 else y


class _GradReverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, dy):
        return -dy


def select_channel(x, label):
    return x.gather(1, label.view(-1, 1, 1).expand(-1, 1, x.shape[2]))


def conv1d_select(x, weight, label):
    return select_channel(F.conv1d(x, weight, None, padding=(weight.shape[2] - 1) // 2), label)


def mse_to_const_sum(tensors, target):
    return sum(((t - target) ** 2).mean() for t in tensors)


def l1_mean_sum(sig, ref):
    return sum((a - b.detach()).abs().mean() for a, b in zip(sig, ref))


def l1_mean_sum_rows(sig, row0, nrows, ref):
    return sum((a[row0:row0 + nrows] - b.detach()).abs().mean() for a, b in zip(sig, ref))


def contrastive_loss(X, Y, raw_X, raw_Y):
    from oracle import tdvc_oracle as O          # the pinned restatement of util/losses.py:70-116
    return O.contrastive_loss(X, Y, raw_X, raw_Y)


def _false(*a, **k):
    return False


@contextlib.contextmanager
def installed():
    """Swap tdvc.ops' entry points for the torch definitions above for the duration of the block."""
    from tdvc import ops
    table = dict(conv1d=conv1d, linear=linear, conv_transpose1d=conv_transpose1d, weight_norm=weight_norm,
                 leaky_relu=lambda x, slope=0.2: F.leaky_relu(x, slope), film=film, add_scale=add_scale,
                 grad_reverse=_GradReverse.apply, l2_normalize=lambda x: F.normalize(x, dim=1), cond_concat=cond_concat,
                 cond_concat_front=cond_concat_front, instance_norm=instance_norm, cond_instance_norm=cond_instance_norm,
                 time_mean=lambda x: x.mean(dim=2), avg_pool_4_2_1=lambda x: F.avg_pool1d(x, 4, 2, 1, count_include_pad=False),
                 select_channel=select_channel, conv1d_select=conv1d_select, mse_to_const_sum=mse_to_const_sum, l1_mean_sum=l1_mean_sum,
                 l1_mean_sum_rows=l1_mean_sum_rows, contrastive_loss=contrastive_loss, mrf_cond_path_eligible=_false,
                 film_posconv_eligible=_false)
    saved = {k: getattr(ops, k) for k in table}
    for k, v in table.items():
        setattr(ops, k, v)
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
