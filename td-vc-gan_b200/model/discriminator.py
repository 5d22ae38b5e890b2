"""Multi-scale grouped-Conv1d discriminators on the tdvc kernels (reference: model/discriminator.py)."""
import torch
import torch.nn as nn

from tdvc import ops
from tdvc.layers import Conv1d, LeakyReLU
from util.dsp import kaiser_filter


class Discriminator(nn.Module):
    """reference model/discriminator.py:7-53.  7 weight-normed convs; every conv + LeakyReLU(0.2) pair runs as
    one kernel (bias and activation in the epilogue); the returned feature maps are the post-activation
    tensors, as with the reference's inplace LeakyReLU."""

    def __init__(self, num_classes, num_layers, num_channels_base, num_channel_mult=4, downsampling_factor=4,
                 conditional_dim=32, conditional='both'):
        super().__init__()
        leaky_relu_slope = 0.2
        num_channel_max = 1024
        self.slope = leaky_relu_slope
        self.discriminator = nn.ModuleList()
        self.discriminator += [nn.Sequential(Conv1d(1, num_channels_base, kernel_size=15, padding=7,
                                                    padding_mode='reflect', weight_norm=True),
                                             LeakyReLU(leaky_relu_slope, inplace=True))]
        nf = num_channels_base
        for i in range(num_layers):
            nf_prev = nf
            nf = min(nf * num_channel_mult, num_channel_max)
            self.discriminator += [nn.Sequential(Conv1d(nf_prev, nf, kernel_size=downsampling_factor * 10 + 1,
                                                        stride=downsampling_factor, padding=downsampling_factor * 5,
                                                        groups=nf_prev // num_channel_mult, weight_norm=True),
                                                 LeakyReLU(leaky_relu_slope, inplace=True))]
        self.discriminator += [nn.Sequential(Conv1d(nf, nf, kernel_size=5, padding=2, weight_norm=True),
                                             LeakyReLU(leaky_relu_slope, inplace=True))]
        self.output = Conv1d(nf, num_classes, kernel_size=3, stride=1, padding=1, bias=False, weight_norm=True)

    def forward(self, x, label_tgt):
        features = []
        for layer in self.discriminator:
            x = layer[0](x, out_act="lrelu", out_slope=layer[1].negative_slope)
            features.append(x)
        x = self.output(x)
        out = ops.select_channel(x, label_tgt.view(-1))
        return out, features


class MultiscaleDiscriminator(nn.Module):
    """reference model/discriminator.py:55-75 (AvgPool1d(4, 2, 1, count_include_pad=False) between scales)."""

    def __init__(self, num_disc, num_classes, num_layers, num_channels_base, num_channel_mult=4,
                 downsampling_factor=4, conditional_dim=32, conditional='both'):
        super().__init__()
        self.discriminators = nn.ModuleList()
        for i in range(num_disc):
            self.discriminators += [Discriminator(num_classes, num_layers, num_channels_base, num_channel_mult,
                                                  downsampling_factor, conditional_dim, conditional)]

    def pooling(self, x):
        return ops.avg_pool_4_2_1(x)

    def forward(self, x, label_tgt):
        ret = []
        for disc in self.discriminators:
            ret.append(disc(x, label_tgt))
            x = self.pooling(x)
        out, features = zip(*ret)
        return list(out), list(features)


class CollaborativeMultibandDiscriminator(nn.Module):
    """reference model/discriminator.py:77-118: discriminators on x, LP|2(x), LP|4(x) plus the generator's
    sub-scale heads; 129-tap Kaiser (beta 10) half-band decimator kept as a non-persistent buffer."""

    def __init__(self, num_disc, num_classes, num_layers, num_channels_base, num_channel_mult=4,
                 downsampling_factor=4, conditional_dim=32, conditional='both'):
        super().__init__()
        self.discriminators = nn.ModuleList()
        for i in range(num_disc):
            self.discriminators += [Discriminator(num_classes, num_layers, num_channels_base, num_channel_mult,
                                                  downsampling_factor, conditional_dim, conditional)]
        L = 129
        f = kaiser_filter(L, 0.5, 10)
        f = f.view(1, 1, -1)
        self.L = L
        self.register_buffer('down_filter', f, persistent=False)

    def pooling(self, x):
        return ops.avg_pool_4_2_1(x)

    def _down(self, x):
        return ops.conv1d(x, self.down_filter, None, stride=2, padding=(self.L - 1) // 2)

    def forward(self, x, label_tgt, subscales=[]):
        ret = []
        n = len(self.discriminators)
        for i, disc in enumerate(self.discriminators):
            ret.append(disc(x, label_tgt))
            if i + 1 < n:          # the reference also filters after the last scale and drops the result
                x = self._down(x)
        for x_sub, disc in zip(subscales, reversed(self.discriminators)):
            ret.append(disc(x_sub, label_tgt))
        out, features = zip(*ret)
        return list(out), list(features)

    def get_subsamples(self, x):
        ret = []
        for i in range(len(self.discriminators) - 1):
            x = self._down(x)
            ret.append(x)
        return list(reversed(ret))
