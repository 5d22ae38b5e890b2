"""Multi-scale grouped-Conv1d discriminators on the tdvc kernels.

Drop-in for the reference's `model/discriminator.py` (same class names, constructor arguments, sub-module names and
therefore the same `state_dict` keys and the same initial weights for a given seed); the computation is ours: every
conv + LeakyReLU(0.2) pair is ONE kernel launch (bias and activation in the conv's epilogue), the label gather and
the pooling / band-split filters are tdvc kernels.
"""
import torch.nn as nn

from tdvc import ops
from tdvc.layers import Conv1d, LeakyReLU
from util.dsp import kaiser_filter

_SLOPE = 0.2          # LeakyReLU slope of every hidden layer (reference model/discriminator.py:12)
_MAX_CHANNELS = 1024  # channel cap of the strided stack (reference model/discriminator.py:13)


def _hidden_layer_plan(num_layers, base, mult, down):
    """Conv1d keyword sets of the hidden stack, in construction order (reference model/discriminator.py:17-35):
    a k15 reflect-padded stem, `num_layers` grouped k(10*down+1) stride-`down` layers that multiply the width by `mult`
    up to 1024 with `width/mult` groups, and a dense k5 layer."""
    plan = [dict(in_channels=1, out_channels=base, kernel_size=15, padding=7, padding_mode='reflect')]
    width = base
    for _ in range(num_layers):
        wider = min(width * mult, _MAX_CHANNELS)
        plan.append(dict(in_channels=width, out_channels=wider, kernel_size=10 * down + 1, stride=down, padding=5 * down,
                         groups=width // mult))
        width = wider
    plan.append(dict(in_channels=width, out_channels=width, kernel_size=5, padding=2))
    return plan, width


class Discriminator(nn.Module):
    """One scale: `discriminator.{i}.0` are the weight-normed hidden convs (`.1` the activation slot, kept so that the
    key layout matches), `output` maps to one logit track per speaker, of which the target speaker's is returned.
    The feature maps handed back are post-activation, as with the reference's in-place LeakyReLU."""

    def __init__(self, num_classes, num_layers, num_channels_base, num_channel_mult=4, downsampling_factor=4,
                 conditional_dim=32, conditional='both'):
        super().__init__()
        plan, width = _hidden_layer_plan(num_layers, num_channels_base, num_channel_mult, downsampling_factor)
        self.slope = _SLOPE
        self.discriminator = nn.ModuleList(
            nn.Sequential(Conv1d(weight_norm=True, **kw), LeakyReLU(_SLOPE, inplace=True)) for kw in plan)
        self.output = Conv1d(width, num_classes, kernel_size=3, stride=1, padding=1, bias=False, weight_norm=True)

    def forward(self, x, label_tgt):
        features = []
        for conv, act in self.discriminator:
            x = conv(x, out_act="lrelu", out_slope=act.negative_slope)      # conv + bias + LeakyReLU: one launch
            features.append(x)
        # output conv + x.gather(1, label) (discriminator.py:36,49-51): only the target speaker's logit track is computed
        out = self.output
        return ops.conv1d_select(x, out.effective_weight(), label_tgt.view(-1)), features


class _DiscriminatorBank(nn.Module):
    """`discriminators.{j}`: `num_disc` independent scales built in order (shared by the two multi-scale wrappers)."""

    def __init__(self, num_disc, num_classes, num_layers, num_channels_base, num_channel_mult=4,
                 downsampling_factor=4, conditional_dim=32, conditional='both'):
        super().__init__()
        self.discriminators = nn.ModuleList(
            Discriminator(num_classes, num_layers, num_channels_base, num_channel_mult, downsampling_factor,
                          conditional_dim, conditional) for _ in range(num_disc))

    def pooling(self, x):
        return ops.avg_pool_4_2_1(x)        # AvgPool1d(4, 2, 1, count_include_pad=False), discriminator.py:63

    @staticmethod
    def _split(results):
        outs, feats = zip(*results)
        return list(outs), list(feats)


class MultiscaleDiscriminator(_DiscriminatorBank):
    """Average-pooled pyramid (reference model/discriminator.py:55-75)."""

    def forward(self, x, label_tgt):
        results = []
        for disc in self.discriminators:
            results.append(disc(x, label_tgt))
            x = self.pooling(x)
        return self._split(results)


class CollaborativeMultibandDiscriminator(_DiscriminatorBank):
    """Scales see x, LP|2(x), LP|4(x) -- a 129-tap Kaiser(beta 10) half-band decimator, kept as a non-persistent buffer
    so that it never enters a checkpoint -- and, in reverse order, the generator's sub-scale heads
    (reference model/discriminator.py:77-118)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.L = 129
        self.register_buffer('down_filter', kaiser_filter(self.L, 0.5, 10).view(1, 1, -1), persistent=False)

    def _down(self, x):
        return ops.conv1d(x, self.down_filter, None, stride=2, padding=(self.L - 1) // 2)

    def forward(self, x, label_tgt, subscales=()):
        results = []
        last = len(self.discriminators) - 1
        for i, disc in enumerate(self.discriminators):
            results.append(disc(x, label_tgt))
            if i < last:                    # the reference also filters after the last scale and drops the result
                x = self._down(x)
        results += [disc(x_sub, label_tgt) for x_sub, disc in zip(subscales, reversed(self.discriminators))]
        return self._split(results)

    def get_subsamples(self, x):
        bands = []
        for _ in range(len(self.discriminators) - 1):
            x = self._down(x)
            bands.append(x)
        return bands[::-1]
