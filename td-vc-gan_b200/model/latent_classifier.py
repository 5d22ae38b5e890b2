"""Speaker classifier on the content embedding, behind a gradient-reversal layer (reference:
model/latent_classifier.py:8-38), on the tdvc kernels: every conv + LeakyReLU pair is one launch."""
import torch.nn as nn

from tdvc import ops
from tdvc.layers import Conv1d, LeakyReLU

from .grad_rev import GradRevLayer


class LatentClassifier(nn.Module):
    def __init__(self, num_classes, num_channels_input, num_layers=3, num_channel_mult=2, downsampling_factor=2):
        super().__init__()
        leaky_relu_slope = 0.2
        self.classifier = nn.ModuleList()
        self.classifier += [GradRevLayer()]
        nf = num_channels_input
        for i in range(num_layers):
            nf_prev = nf
            nf = nf * num_channel_mult
            self.classifier += [Conv1d(nf_prev, nf, kernel_size=downsampling_factor * 10 + 1, stride=downsampling_factor,
                                       padding=downsampling_factor * 5, weight_norm=True),
                                LeakyReLU(leaky_relu_slope, inplace=True)]
        self.classifier += [Conv1d(nf, nf, kernel_size=5, padding=2, weight_norm=True),
                            LeakyReLU(leaky_relu_slope, inplace=True)]
        self.classifier += [Conv1d(nf, num_classes, kernel_size=3, padding=1, bias=False, weight_norm=True)]

    def forward(self, x):
        mods = list(self.classifier)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, Conv1d):
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if isinstance(nxt, LeakyReLU):
                    x = m(x, out_act="lrelu", out_slope=nxt.negative_slope)
                    i += 1
                else:
                    x = m(x)
            else:
                x = m(x)
            i += 1
        return ops.time_mean(x)          # F.avg_pool1d(x, x.size(2)).squeeze(2)
