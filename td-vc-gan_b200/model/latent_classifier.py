"""Speaker classifier on the content embedding, trained adversarially through a gradient-reversal layer.

Drop-in for the reference's `model/latent_classifier.py` (same constructor, `classifier.{i}` module indices and
therefore checkpoint keys); on the tdvc kernels every conv + LeakyReLU pair is one launch and the temporal average is
the row-mean kernel.
"""
import torch.nn as nn

from tdvc import ops
from tdvc.layers import Conv1d, LeakyReLU

from .grad_rev import GradRevLayer

_SLOPE = 0.2


class LatentClassifier(nn.Module):
    """`classifier` = [GradRev, (strided wn-conv, LeakyReLU) x num_layers, (wn-conv k5, LeakyReLU), wn-conv k3 -> classes]
    (reference model/latent_classifier.py:8-32); the output is the time average of the class logits (:36)."""

    def __init__(self, num_classes, num_channels_input, num_layers=3, num_channel_mult=2, downsampling_factor=2):
        super().__init__()
        widths = [num_channels_input * num_channel_mult ** i for i in range(num_layers + 1)]
        down = downsampling_factor
        mods = [GradRevLayer()]
        for cin, cout in zip(widths[:-1], widths[1:]):
            mods += [Conv1d(cin, cout, kernel_size=10 * down + 1, stride=down, padding=5 * down, weight_norm=True),
                     LeakyReLU(_SLOPE, inplace=True)]
        top = widths[-1]
        mods += [Conv1d(top, top, kernel_size=5, padding=2, weight_norm=True), LeakyReLU(_SLOPE, inplace=True)]
        mods += [Conv1d(top, num_classes, kernel_size=3, padding=1, bias=False, weight_norm=True)]
        self.classifier = nn.ModuleList(mods)

    def forward(self, x):
        x = self.classifier[0](x)                                   # gradient reversal
        hidden = list(self.classifier)[1:-1]
        for conv, act in zip(hidden[0::2], hidden[1::2]):
            x = conv(x, out_act="lrelu", out_slope=act.negative_slope)
        return ops.time_mean(self.classifier[-1](x))                # F.avg_pool1d(x, x.size(2)).squeeze(2)
