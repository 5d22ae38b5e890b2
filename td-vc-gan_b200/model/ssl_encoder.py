"""SSL content encoder (reference model/ssl_encoder.py): a frozen WavLM front end followed by a WaveNet-style stack that
maps the 1024-dim SSL features to the content embedding.

The WN stack (`WN`, `Encoder`, reference lines 16-116) runs on the tdvc kernels: weight-normed convs through
`tdvc.layers.Conv1d` (checkpoint keys `bias / weight_g / weight_v`, the reference's init and RNG consumption), the gate
`tanh(.) * sigmoid(.)` (`fused_add_tanh_sigmoid_multiply`, lines 7-14) as one kernel forward and one backward.  At the
content rate (T/320 = 28 frames per 0.56 s segment) the 16 k5 convs are short sequences: in bf16 mode they take the
batch-flattened tensor-core path (tdvc.ops._Conv1dTC).

WavLM itself (`wavlm/`, 316 M frozen parameters, forward only) is the reference's vendored third-party model and outside the
path this package rebuilds: `SSLEncoder` imports it from a reference checkout on `sys.path` and fails
with the import error when there is none.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from tdvc import ops
from tdvc.layers import Conv1d


def fused_add_tanh_sigmoid_multiply(input_a, input_b, n_channels=None):
    """reference model/ssl_encoder.py:7-14 (n_channels is implied by the shapes)."""
    return ops.gated_tanh_sigmoid(input_a, input_b)


class WN(nn.Module):
    """reference model/ssl_encoder.py:16-88."""

    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0):
        super().__init__()
        assert kernel_size % 2 == 1
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size,          # the reference stores a 1-tuple here (trailing comma, line 21)
        self.dilation_rate = dilation_rate
        self.n_layers = n_layers
        self.gin_channels = gin_channels
        self.p_dropout = p_dropout
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        self.drop = nn.Dropout(p_dropout)
        if gin_channels != 0:
            self.cond_layer = Conv1d(gin_channels, 2 * hidden_channels * n_layers, 1, weight_norm=True)
        for i in range(n_layers):
            dilation = dilation_rate ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(Conv1d(hidden_channels, 2 * hidden_channels, kernel_size, dilation=dilation, padding=padding,
                                         weight_norm=True))
            res_skip_channels = 2 * hidden_channels if i < n_layers - 1 else hidden_channels      # the last layer has no residual
            self.res_skip_layers.append(Conv1d(hidden_channels, res_skip_channels, 1, weight_norm=True))

    def forward(self, x, x_mask, g=None, **kwargs):
        H = self.hidden_channels
        output = None
        if g is not None:
            g = self.cond_layer(g)
        for i in range(self.n_layers):
            x_in = self.in_layers[i](x)
            g_l = g[:, i * 2 * H:(i + 1) * 2 * H, :] if g is not None else None
            acts = self.drop(ops.gated_tanh_sigmoid(x_in, g_l))
            res_skip_acts = self.res_skip_layers[i](acts)
            if i < self.n_layers - 1:
                x = x + res_skip_acts[:, :H, :]
                if torch.is_tensor(x_mask):
                    x = x * x_mask
                skip = res_skip_acts[:, H:, :]
            else:
                skip = res_skip_acts
            output = skip if output is None else output + skip
        return output * x_mask if torch.is_tensor(x_mask) else output

    def remove_weight_norm(self):
        raise NotImplementedError("weight norm is folded per call scope (tdvc.ops.inference_cache), not removed")


class Encoder(nn.Module):
    """reference model/ssl_encoder.py:91-116."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.dilation_rate = dilation_rate
        self.n_layers = n_layers
        self.gin_channels = gin_channels
        self.pre = Conv1d(in_channels, hidden_channels, 1)
        self.enc = WN(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.proj = Conv1d(hidden_channels, out_channels * 2, 1)

    def forward(self, x, x_lengths=None, g=None):
        x_mask = 1                              # the reference's sequence mask is commented out (line 108)
        x = self.pre(x)
        x = self.enc(x, x_mask, g=g)
        stats = self.proj(x)
        m, logs = torch.split(stats, self.out_channels, dim=1)
        z = m + torch.randn_like(m) * torch.exp(logs)      # drawn as in the reference (keeps the RNG stream aligned)
        return z, m, logs, x_mask


class SSLEncoder(nn.Module):
    """reference model/ssl_encoder.py:118-148: frozen WavLM-Large features -> WN stack -> mean of the posterior."""

    def __init__(self, encoder_model='wavlm', num_layers=16, emb_dim=128, kernel_size=5, dilation_rate=1,
                 weight_norm=lambda x: x):
        super().__init__()
        self.encoder_model = encoder_model
        if encoder_model == 'wavlm':
            from wavlm import WavLM, WavLMConfig      # the reference's vendored model (needs a reference checkout on sys.path)
            print("Loading WavLM for content...")
            checkpoint = torch.load('wavlm/WavLM-Large.pt')
            cfg = WavLMConfig(checkpoint['cfg'])
            self.cmodel = WavLM(cfg).cuda()
            self.cmodel.load_state_dict(checkpoint['model'])
            self.cmodel.eval()
            print("Loaded WavLM.")
            ssl_dim = 1024
        else:
            raise NotImplementedError("Unknown encoder model")
        self.encoder = Encoder(ssl_dim, emb_dim, emb_dim, kernel_size, dilation_rate, num_layers)

    def forward(self, x):
        with torch.no_grad():
            x = F.pad(x, (160, 0))
            c = self.cmodel.extract_features(x.squeeze(1))[0]
            c = c.transpose(1, 2)
        z, m, logs, _ = self.encoder(c.contiguous())
        return m
