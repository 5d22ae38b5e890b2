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
    """reference model/ssl_encoder.py:16-88: n_layers x [dilated conv H -> 2H, gate, 1x1 conv H -> 2H (residual | skip)]."""

    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0):
        super().__init__()
        if kernel_size % 2 != 1:
            raise AssertionError("WN: odd kernel sizes only")
        H = hidden_channels
        self.hidden_channels, self.dilation_rate, self.n_layers = H, dilation_rate, n_layers
        self.kernel_size = (kernel_size,)        # a 1-tuple in the reference too (stray comma, line 21)
        self.gin_channels, self.p_dropout = gin_channels, p_dropout
        dil = [dilation_rate ** i for i in range(n_layers)]
        # registration order = the reference's: in_layers, res_skip_layers, drop, then (optionally) cond_layer
        self.in_layers = nn.ModuleList(Conv1d(H, 2 * H, kernel_size, dilation=d, padding=int((kernel_size * d - d) / 2),
                                              weight_norm=True) for d in dil)
        # the last layer feeds the skip sum only: no residual half
        self.res_skip_layers = nn.ModuleList(Conv1d(H, 2 * H if i < n_layers - 1 else H, 1, weight_norm=True)
                                             for i in range(n_layers))
        self.drop = nn.Dropout(p_dropout)
        if gin_channels != 0:
            self.cond_layer = Conv1d(gin_channels, 2 * H * n_layers, 1, weight_norm=True)

    def forward(self, x, x_mask, g=None, **kwargs):
        H, last = self.hidden_channels, self.n_layers - 1
        masked = torch.is_tensor(x_mask)
        cond = self.cond_layer(g) if g is not None else None
        skip_sum = None
        for i, (conv_in, conv_rs) in enumerate(zip(self.in_layers, self.res_skip_layers)):
            g_i = cond[:, 2 * H * i:2 * H * (i + 1), :] if cond is not None else None
            rs = conv_rs(self.drop(ops.gated_tanh_sigmoid(conv_in(x), g_i)))
            if i < last:
                x = x + rs[:, :H, :]
                if masked:
                    x = x * x_mask
                rs = rs[:, H:, :]
            skip_sum = rs if skip_sum is None else skip_sum + rs
        return skip_sum * x_mask if masked else skip_sum

    def remove_weight_norm(self):
        raise NotImplementedError("weight norm is folded per call scope (tdvc.ops.inference_cache), not removed")


class Encoder(nn.Module):
    """reference model/ssl_encoder.py:91-116: 1x1 conv in, WN stack, 1x1 conv to (mean | log-scale), sampled z."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0):
        super().__init__()
        for name, val in (("in_channels", in_channels), ("out_channels", out_channels), ("hidden_channels", hidden_channels),
                          ("kernel_size", kernel_size), ("dilation_rate", dilation_rate), ("n_layers", n_layers),
                          ("gin_channels", gin_channels)):
            setattr(self, name, val)
        self.pre = Conv1d(in_channels, hidden_channels, 1)
        self.enc = WN(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.proj = Conv1d(hidden_channels, out_channels * 2, 1)

    def forward(self, x, x_lengths=None, g=None):
        x_mask = 1                               # the reference's sequence mask is commented out (line 108)
        stats = self.proj(self.enc(self.pre(x), x_mask, g=g))
        m, logs = torch.split(stats, self.out_channels, dim=1)
        z = m + torch.randn_like(m) * torch.exp(logs)      # drawn as in the reference (keeps the RNG stream aligned)
        return z, m, logs, x_mask


def _load_wavlm():
    """the frozen WavLM-Large front end, as model/ssl_encoder.py:126-134 builds it (reference's vendored `wavlm` package and its
    checkpoint path relative to the working directory)"""
    from wavlm import WavLM, WavLMConfig
    print("Loading WavLM for content...")
    ckpt = torch.load('wavlm/WavLM-Large.pt')
    net = WavLM(WavLMConfig(ckpt['cfg'])).cuda()
    net.load_state_dict(ckpt['model'])
    net.eval()
    print("Loaded WavLM.")
    return net, 1024


class SSLEncoder(nn.Module):
    """reference model/ssl_encoder.py:118-148: frozen WavLM-Large features -> WN stack -> mean of the posterior."""

    def __init__(self, encoder_model='wavlm', num_layers=16, emb_dim=128, kernel_size=5, dilation_rate=1,
                 weight_norm=lambda x: x):
        super().__init__()
        self.encoder_model = encoder_model
        if encoder_model != 'wavlm':
            raise NotImplementedError("Unknown encoder model")
        self.cmodel, ssl_dim = _load_wavlm()
        self.encoder = Encoder(ssl_dim, emb_dim, emb_dim, kernel_size, dilation_rate, num_layers)

    def forward(self, x):
        with torch.no_grad():                    # 160 samples of left context, features at the 20 ms rate, channels first
            feats = self.cmodel.extract_features(F.pad(x, (160, 0)).squeeze(1))[0].transpose(1, 2).contiguous()
        return self.encoder(feats)[1]
