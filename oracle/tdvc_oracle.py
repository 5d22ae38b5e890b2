"""CPU oracle for the td-vc-gan hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state_dict in, tensors out) restatement of the reference's Generator /
ConditionalInstanceNorm / Discriminator / GAN-loss arithmetic in plain CPU PyTorch
(fp32 or fp64).  It exists only so that tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg can check (or time beside) the CUDA path.  Nothing
under td-vc-gan_b200/ may import it.

Pinned: yes.  tests/golden/*.npz were produced by importing the *real* reference modules
from /root/reference in fp64 (oracle/gen_golden.py, committed) and tests/test_oracle_golden.py
holds this file to those vectors (1e-9 in fp64).  The reference itself has no golden
vectors or unit tests.

Every function cites the reference file:line (relative to /root/reference) it restates.
The restatement is deliberately structured differently from the reference (no nn.Module,
parameters looked up by checkpoint key) so that it also pins the state_dict key layout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
LRELU = 0.2  # leaky_relu_slope everywhere: model/generator.py:70,209,292 ; model/discriminator.py:12


# ----------------------------------------------------------------------------- primitives

def wn_weight(sd: SD, prefix: str) -> torch.Tensor:
    """Effective conv weight.  Old-style torch.nn.utils.weight_norm (util/__init__.py:16-20,
    model/discriminator.py:11): w = g * v / ||v||, norm over every dim but 0.  Falls back to a
    plain `weight` (e.g. ExciteDownsampleBlock.shortcut, model/generator.py:157)."""
    if prefix + ".weight_v" in sd:
        v = sd[prefix + ".weight_v"]
        g = sd[prefix + ".weight_g"]
        n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v * (g / n)
    return sd[prefix + ".weight"]


def conv(sd: SD, prefix: str, x: torch.Tensor, *, stride: int = 1, padding: int = 0, dilation: int = 1,
         groups: int = 1, reflect: bool = False) -> torch.Tensor:
    """nn.Conv1d with optional padding_mode='reflect' (ATen pads with reflection_pad1d, then a
    zero-pad-free convolution)."""
    w = wn_weight(sd, prefix)
    b = sd.get(prefix + ".bias")
    if reflect and padding > 0:
        x = F.pad(x, (padding, padding), mode="reflect")
        padding = 0
    return F.conv1d(x, w, b, stride=stride, padding=padding, dilation=dilation, groups=groups)


def conv_transpose(sd: SD, prefix: str, x: torch.Tensor, *, stride: int, padding: int,
                   output_padding: int) -> torch.Tensor:
    """nn.ConvTranspose1d under weight_norm: the norm runs over dim 0 = in-channels
    (model/generator.py:311-315)."""
    w = wn_weight(sd, prefix)
    return F.conv_transpose1d(x, w, sd.get(prefix + ".bias"), stride=stride, padding=padding,
                              output_padding=output_padding)


def lrelu(x: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(x, LRELU)


def kaiser_lowpass_gen(L: int, w: float, dtype=torch.float32) -> torch.Tensor:
    """util/__init__.py:104-113  (even L -> L+1 taps, beta 2.5).  The reference builds the taps in
    fp32 and the buffer follows the module dtype afterwards; we build in fp32 then cast, as
    `module.double()` does."""
    n = torch.arange(-L // 2, L // 2 + 1).float()
    f = torch.sin(math.pi * w * n) / (math.pi * n + 1e-8)
    f[n.shape[0] // 2] = w
    f = f * torch.kaiser_window(L + 1, False, 2.5)
    f = f / torch.sum(f)
    return f.view(1, 1, -1).to(dtype)


def kaiser_lowpass_dsp(L: int, fc: float, beta: float = 2.5, dtype=torch.float32) -> torch.Tensor:
    """util/dsp.py:5-16 (odd L only)."""
    if L % 2 == 0:
        raise Exception("Even length filter not implemented")
    L -= 1
    n = torch.arange(-L // 2, L // 2 + 1).float()
    f = torch.sin(math.pi * fc * n) / (math.pi * n + 1e-8)
    f[n.shape[0] // 2] = fc
    f = f * torch.kaiser_window(L + 1, False, beta)
    f = f / torch.sum(f)
    return f.to(dtype)


# ----------------------------------------------------------------------------- CIN

def instance_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm1d(affine=False, track_running_stats=False): per (b, c) over T, biased var."""
    m = x.mean(dim=2, keepdim=True)
    v = x.var(dim=2, unbiased=False, keepdim=True)
    return (x - m) / torch.sqrt(v + eps)


def cond_instance_norm(sd: SD, prefix: str, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """model/conditional_instance_norm.py:11-19."""
    if c.dim() == 2:
        h = F.linear(c, sd[prefix + ".embedding.weight"], sd[prefix + ".embedding.bias"]).unsqueeze(2)
    else:
        h = F.conv1d(c, sd[prefix + ".embedding_conv.weight"], sd[prefix + ".embedding_conv.bias"], padding=2)
    gamma, beta = torch.chunk(h, 2, dim=1)
    return (1 + gamma) * instance_norm(x) + beta


def cin_resnet_block(sd: SD, prefix: str, x: torch.Tensor, c: torch.Tensor, *, dilation: int = 1,
                     kernel_size: int = 3) -> torch.Tensor:
    """model/generator.py:113-139 (plain convs, no weight-norm)."""
    h = cond_instance_norm(sd, prefix + ".block.0", x, c)
    h = conv(sd, prefix + ".block.2", lrelu(h), dilation=dilation,
             padding=(kernel_size * dilation - dilation) // 2, reflect=True)
    h = cond_instance_norm(sd, prefix + ".block.3", h, c)
    h = conv(sd, prefix + ".block.5", lrelu(h))
    return h + conv(sd, prefix + ".shortcut", x)


# ----------------------------------------------------------------------------- legacy residual blocks

def decoder_resnet_block(sd: SD, prefix: str, x: torch.Tensor, *, dilation: int = 1) -> torch.Tensor:
    """DecoderResnetBlock.forward, model/generator.py:11-26 (k=3, reflect pad = dilation, wn 1x1 shortcut)."""
    h = conv(sd, prefix + ".block.1", lrelu(x), dilation=dilation, padding=dilation, reflect=True)
    h = conv(sd, prefix + ".block.3", lrelu(h))
    return h + conv(sd, prefix + ".shortcut", x)


def transform_resnet_block(sd: SD, prefix: str, x: torch.Tensor, *, dilation: int = 1) -> torch.Tensor:
    """TranformResnetBlock.forward, model/generator.py:29-46 (relu-conv-norm order, InstanceNorm1d)."""
    h = instance_norm(conv(sd, prefix + ".block.1", lrelu(x), dilation=dilation, padding=dilation, reflect=True))
    h = instance_norm(conv(sd, prefix + ".block.4", lrelu(h)))
    return h + conv(sd, prefix + ".shortcut", x)


def resnet_block(sd: SD, prefix: str, x: torch.Tensor, *, dilation: int = 1) -> torch.Tensor:
    """ResnetBlock.forward, model/generator.py:48-67 (norm-relu-conv, identity shortcut)."""
    h = conv(sd, prefix + ".block.2", lrelu(instance_norm(x)), dilation=dilation, padding=dilation, reflect=True)
    h = conv(sd, prefix + ".block.5", lrelu(instance_norm(h)))
    return h + x


# ----------------------------------------------------------------------------- generator blocks

def film_block(sd: SD, prefix: str, x: torch.Tensor, c: Optional[torch.Tensor], *, kernel_size: int,
               dilation: int) -> torch.Tensor:
    """FiLMResnetBlock.forward, model/generator.py:96-111 (3-D `c` branch; the 2-D branch is dead)."""
    h = conv(sd, prefix + ".conv.1", lrelu(x), dilation=dilation,
             padding=(kernel_size * dilation - dilation) // 2, reflect=True)
    if c is not None:
        g = conv(sd, prefix + ".cond_var.0", c, padding=1)
        g = conv(sd, prefix + ".cond_var.2", lrelu(g), padding=1)
        gamma, beta = g.chunk(2, dim=1)
        h = h * (1 + gamma)
        h = h + beta
    return conv(sd, prefix + ".posconv.1", lrelu(h)) + x


MRF_KERNELS = (3, 7, 11)   # model/generator.py:211,293
MRF_DILATIONS = (1, 3, 5)  # model/generator.py:210,292


def mrf_block(sd: SD, prefix: str, x: torch.Tensor, c: Optional[torch.Tensor]) -> torch.Tensor:
    """MRFBlock.forward, model/generator.py:186-194."""
    y = 0
    for i, k in enumerate(MRF_KERNELS):
        xs = x
        for j, d in enumerate(MRF_DILATIONS):
            xs = film_block(sd, f"{prefix}.blocks.{i}.{j}", xs, c, kernel_size=k, dilation=d)
        y = y + xs
    return y / len(MRF_KERNELS)


def excite_downsample(sd: SD, prefix: str, x: torch.Tensor, r: int) -> torch.Tensor:
    """ExciteDownsampleBlock.forward, model/generator.py:162-173."""
    sh = conv(sd, prefix + ".shortcut", x)
    filt = kaiser_lowpass_gen(16 * r, 1 / r, x.dtype).expand(sh.shape[1], 1, -1)
    sh = F.conv1d(sh, filt, stride=r, padding=8 * r, groups=sh.shape[1])
    h = conv(sd, prefix + ".block.0", x, stride=r, padding=r // 2)
    h = conv(sd, prefix + ".block.2", lrelu(h), padding=2)
    h = conv(sd, prefix + ".block.4", lrelu(h), padding=2)
    return h + sh


def encoder(sd: SD, prefix: str, x: torch.Tensor, ratios: Sequence[int]) -> torch.Tensor:
    """Encoder.forward (non-CIN path), model/generator.py:255-272; `ratios` is already reversed
    (model/generator.py:456)."""
    p = prefix + ".encoder"
    x = conv(sd, f"{p}.0", x, padding=3, reflect=True)
    for i, r in enumerate(ratios):
        x = conv(sd, f"{p}.{3 + 4 * i}", lrelu(x), stride=r, padding=r // 2 + r % 2)
        x = mrf_block(sd, f"{p}.{4 + 4 * i}", x, None)
    n = 1 + 4 * len(ratios)
    x = conv(sd, f"{p}.{n + 1}", lrelu(x), padding=3)
    if f"{p}.{n + 3}.weight_v" in sd or f"{p}.{n + 3}.weight" in sd:
        x = conv(sd, f"{p}.{n + 3}", lrelu(x), padding=3)
    return F.normalize(x, dim=1)


def decoder(sd: SD, prefix: str, x: torch.Tensor, c: torch.Tensor, c_var: torch.Tensor,
            ratios: Sequence[int]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Decoder.forward (cin=True, c_var given), model/generator.py:375-406 and the excitation
    pyramid 364-372.  Returns (y, [sub-scale tanh heads])."""
    p = prefix + ".decoder"
    nst = len(ratios)
    # pyramid: full rate first, then each ExciteDownsampleBlock from the last ratio to the first
    scales = [conv(sd, f"{prefix}.excite_downsample.{nst}", c_var, padding=3, reflect=True)]
    for i in range(nst - 1, -1, -1):
        scales.append(excite_downsample(sd, f"{prefix}.excite_downsample.{i}", scales[-1], ratios[i]))
    has_embed = (f"{p}.3.weight_v" in sd) or (f"{p}.3.weight" in sd and f"{p}.1.weight" in sd)
    idx = 0
    if has_embed:  # content_dim given: two stem convs (model/generator.py:298-305)
        x = conv(sd, f"{p}.1", lrelu(x), padding=3)
        x = conv(sd, f"{p}.3", lrelu(x), padding=3)
        idx = 4
    else:
        x = conv(sd, f"{p}.1", lrelu(x), padding=3)
        idx = 2
    c_const = c.unsqueeze(2).repeat(1, 1, x.size(2))
    subs: List[torch.Tensor] = []
    for i, r in enumerate(ratios):
        x = conv_transpose(sd, f"{p}.{idx + 2}", lrelu(x), stride=r, padding=r // 2 + r % 2,
                           output_padding=r % 2)
        head = f"{prefix}.subsample_out_layers.{i}.1"
        if head + ".weight_v" in sd or head + ".weight" in sd:
            subs.append(torch.tanh(conv(sd, head, lrelu(x), padding=3, reflect=True)))
        c_const = c_const.repeat(1, 1, r)       # tiles, does not interleave (generator.py:397)
        cc = torch.cat([c_const, scales[nst - 1 - i]], dim=1)
        x = mrf_block(sd, f"{p}.{idx + 3}", x, cc)
        idx += 4
    x = torch.tanh(conv(sd, f"{p}.{idx + 2}", lrelu(x), padding=3, reflect=True))
    return x, subs


def generator(sd: SD, x: torch.Tensor, c_tgt: torch.Tensor, c_var: torch.Tensor,
              ratios: Sequence[int]) -> Tuple[torch.Tensor, List[torch.Tensor], torch.Tensor]:
    """Generator.forward, model/generator.py:490-508, with 0 bottleneck layers.
    Returns (y, sub-scale outputs, content embedding)."""
    emb = F.linear(c_tgt, sd["embedding.weight"], sd["embedding.bias"])
    content = encoder(sd, "encoder", x, list(ratios)[::-1])
    y, subs = decoder(sd, "decoder", content, emb, c_var, ratios)
    return y, subs, content


# ----------------------------------------------------------------------------- discriminators

def discriminator(sd: SD, prefix: str, x: torch.Tensor, label: torch.Tensor, *, num_layers: int = 4,
                  mult: int = 4, down: int = 4) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Discriminator.forward, model/discriminator.py:40-53; layer geometry 17-38."""
    feats = []
    p = prefix + ".discriminator"
    x = lrelu(conv(sd, f"{p}.0.0", x, padding=7, reflect=True))
    feats.append(x)
    for i in range(num_layers):
        cin = x.shape[1]
        x = lrelu(conv(sd, f"{p}.{i + 1}.0", x, stride=down, padding=down * 5, groups=cin // mult))
        feats.append(x)
    x = lrelu(conv(sd, f"{p}.{num_layers + 1}.0", x, padding=2))
    feats.append(x)
    x = conv(sd, prefix + ".output", x, padding=1)
    idx = label.view(-1, 1, 1).expand(-1, 1, x.shape[2])
    return x.gather(1, idx), feats


def cmb_discriminator(sd: SD, x: torch.Tensor, label: torch.Tensor, subscales: Sequence[torch.Tensor] = (),
                      *, num_disc: int = 3, **kw) -> Tuple[List[torch.Tensor], List[List[torch.Tensor]]]:
    """CollaborativeMultibandDiscriminator.forward, model/discriminator.py:94-108."""
    filt = kaiser_lowpass_dsp(129, 0.5, 10, x.dtype).view(1, 1, -1)
    outs, feats = [], []
    for k in range(num_disc):
        o, f = discriminator(sd, f"discriminators.{k}", x, label, **kw)
        outs.append(o); feats.append(f)
        x = F.conv1d(x, filt, stride=2, padding=64)
    for xs, k in zip(subscales, reversed(range(num_disc))):
        o, f = discriminator(sd, f"discriminators.{k}", xs, label, **kw)
        outs.append(o); feats.append(f)
    return outs, feats


def cmb_subsamples(x: torch.Tensor, num_disc: int = 3) -> List[torch.Tensor]:
    """CollaborativeMultibandDiscriminator.get_subsamples, model/discriminator.py:110-118."""
    filt = kaiser_lowpass_dsp(129, 0.5, 10, x.dtype).view(1, 1, -1)
    ret = []
    for _ in range(num_disc - 1):
        x = F.conv1d(x, filt, stride=2, padding=64)
        ret.append(x)
    return ret[::-1]


def multiscale_discriminator(sd: SD, x: torch.Tensor, label: torch.Tensor, *, num_disc: int = 3, **kw):
    """MultiscaleDiscriminator.forward, model/discriminator.py:67-75 (AvgPool1d(4,2,1,
    count_include_pad=False) between scales)."""
    outs, feats = [], []
    for k in range(num_disc):
        o, f = discriminator(sd, f"discriminators.{k}", x, label, **kw)
        outs.append(o); feats.append(f)
        x = F.avg_pool1d(x, 4, 2, 1, count_include_pad=False)
    return outs, feats


def latent_classifier(sd: SD, x: torch.Tensor, *, num_layers: int = 3, down: int = 2) -> torch.Tensor:
    """LatentClassifier.forward, model/latent_classifier.py:34-38: gradient reversal (identity forward,
    model/grad_rev.py:5-10), 3 x (wn Conv k=21 s=2 p=10 + LeakyReLU), wn Conv k5, LeakyReLU, wn Conv k3 (no bias),
    global average pool over time."""
    x = GradRev.apply(x)
    idx = 1
    for _ in range(num_layers):
        x = lrelu(conv(sd, f"classifier.{idx}", x, stride=down, padding=down * 5))
        idx += 2
    x = lrelu(conv(sd, f"classifier.{idx}", x, padding=2))
    x = conv(sd, f"classifier.{idx + 2}", x, padding=1)
    return x.mean(dim=2)


def ssl_wn(sd: SD, prefix: str, x: torch.Tensor, g=None, *, hidden: int, kernel_size: int, dilation_rate: int,
           n_layers: int) -> torch.Tensor:
    """WN.forward, model/ssl_encoder.py:53-81 (x_mask = 1, dropout 0): per layer a dilated wn conv to 2H channels, the
    tanh * sigmoid gate (fused_add_tanh_sigmoid_multiply, :7-14) with the optional conditioning slice added first, a 1x1 wn
    conv whose first H outputs are the residual and the rest the skip sum (the last layer has skip outputs only)."""
    out = torch.zeros_like(x)
    if g is not None:
        g = conv(sd, prefix + "cond_layer", g)
    for i in range(n_layers):
        d = dilation_rate ** i
        a = conv(sd, f"{prefix}in_layers.{i}", x, padding=int((kernel_size * d - d) / 2), dilation=d)
        if g is not None:
            a = a + g[:, i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = conv(sd, f"{prefix}res_skip_layers.{i}", acts)
        if i < n_layers - 1:
            x = x + rs[:, :hidden]
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out


def ssl_wn_encoder(sd: SD, x: torch.Tensor, *, out_channels: int, hidden: int, kernel_size: int, dilation_rate: int,
                   n_layers: int):
    """Encoder.forward of model/ssl_encoder.py:105-116 without the sampled z: (m, logs)."""
    h = conv(sd, "pre", x)
    h = ssl_wn(sd, "enc.", h, hidden=hidden, kernel_size=kernel_size, dilation_rate=dilation_rate, n_layers=n_layers)
    stats = conv(sd, "proj", h)
    return stats[:, :out_channels], stats[:, out_channels:]


def yin_estimate(signal: torch.Tensor, sample_rate: float, pitch_min: float = 20, pitch_max: float = 20000,
                 frame_stride: float = 0.01, threshold: float = 0.1) -> torch.Tensor:
    """util/yin.py:24-84 + :87-133 (`estimate`, hard search) with the difference function in its defining form (equation 6 of
    the YIN paper, d(tau) = sum_j (x_j - x_{j+tau})^2 over the part of the window where both samples exist) in fp64 instead of
    the reference's fp32 FFT; the cumulative-mean normalisation (equation 8) is rounded to fp32 before the threshold / slope
    comparisons because that is where the reference decides."""
    x = signal.double()
    tau_min, tau_max = int(sample_rate / pitch_max), int(sample_rate / pitch_min)
    W, hop = 2 * tau_max, int(frame_stride * sample_rate)
    if x.shape[-1] < W:
        x = F.pad(x, [0, W - x.shape[-1]])
    fr = F.pad(x, [W // 2, W // 2 - 1]).unfold(-1, W, hop)                      # [..., frames, W]
    d = torch.stack([((fr[..., :W - tau] - fr[..., tau:]) ** 2).sum(-1) for tau in range(1, tau_max)], dim=-1)
    lag = torch.arange(1, tau_max, dtype=torch.float64)
    c = (d * lag / d.cumsum(-1).clamp_min(1e-5)).float()[..., tau_min:]
    thr = torch.tensor(threshold, dtype=torch.float32)
    n = c.shape[-1]
    below = (c < thr).int().argmax(-1)
    rising = F.pad(c.diff() >= 0.0, [0, 1], value=True)
    idx = torch.arange(n)
    ok = (idx >= below.unsqueeze(-1)) & rising & (below.unsqueeze(-1) > 0)
    tau = ok.int().argmax(-1)
    # `sample_rate / tensor` is Tensor.__rtruediv__ = reciprocal(tensor) * sample_rate: two fp32 roundings, reproduced here
    return torch.where(tau > 0, (tau + tau_min + 1).float().reciprocal() * sample_rate, torch.zeros((), dtype=torch.float32))


class GradRev(torch.autograd.Function):
    """model/grad_rev.py:3-10: identity forward, negated gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -g


# ----------------------------------------------------------------------------- losses

def lsgan_d_loss(outs_real: Sequence[torch.Tensor], outs_fake: Sequence[torch.Tensor]):
    """train.py:271-281."""
    real = sum(((o - 1) ** 2).mean() for o in outs_real)
    fake = sum((o ** 2).mean() for o in outs_fake)
    return real, fake


def lsgan_g_loss(outs_fake: Sequence[torch.Tensor]) -> torch.Tensor:
    """train.py:327-331."""
    return sum(((o - 1) ** 2).mean() for o in outs_fake)


def feat_loss(feats_sig, feats_ref) -> torch.Tensor:
    """util/losses.py:55-68 with norm_p=1."""
    tot = 0
    for fs, fr in zip(feats_sig, feats_ref):
        for a, b in zip(fs, fr):
            tot = tot + (a - b.detach()).abs().mean()
    return tot


def _hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs: int, n_mels: int, sr: int, dtype=torch.float32) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(n_freqs, 0, sr/2, n_mels, sr, norm='slaney',
    mel_scale='htk') restated (the defaults util/losses.py:29-31 ends up with)."""
    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_min, m_max = _hz_to_mel_htk(0.0), _hz_to_mel_htk(sr / 2)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    fb = fb * enorm.unsqueeze(0)
    return fb.to(dtype)  # [n_freqs, n_mels]


def log_mel(signal: torch.Tensor, n_fft: int = 2048, n_mels: int = 80, sr: int = 16000) -> torch.Tensor:
    """MelSpectrogram(sr, n_fft, hop=n_fft//4, n_mels, norm='slaney') -> log(clamp(., 1e-5)),
    util/losses.py:38-42.  torchaudio defaults: hann window (periodic), power 2, center=True,
    pad_mode='reflect', onesided."""
    shp = signal.shape
    x = signal.reshape(-1, shp[-1])
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32).to(signal.dtype)
    spec = torch.stft(x, n_fft, hop_length=n_fft // 4, win_length=n_fft, window=win, center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    power = spec.real ** 2 + spec.imag ** 2                     # [N, F, frames]
    fb = mel_filterbank(n_fft // 2 + 1, n_mels, sr, signal.dtype)
    mel = torch.matmul(power.transpose(1, 2), fb).transpose(1, 2)
    mel = mel.reshape(shp[:-1] + mel.shape[-2:])
    return torch.log(torch.clamp(mel, min=1e-5))


def mel_loss(signal: torch.Tensor, ref: torch.Tensor, fft_sizes=(2048, 1024, 512)) -> torch.Tensor:
    """util/losses.py:33-53: the `return` sits inside the loop, so only fft_sizes[0] is used."""
    return (log_mel(signal, fft_sizes[0]) - log_mel(ref, fft_sizes[0]).detach()).abs().mean()


def contrastive_raw_draws(B: int, T: int, n_neg: int, gen: Optional[torch.Generator] = None) -> torch.Tensor:
    """The torch.randint(0, T-1, (B, T, n_neg)) draw of util/losses.py:77-79, taken outside the loss so
    both sides of a comparison can share it."""
    return torch.randint(0, T - 1, (B, T, n_neg), generator=gen)


def contrastive_loss(X: torch.Tensor, Y: torch.Tensor, raw_X: torch.Tensor, raw_Y: torch.Tensor) -> torch.Tensor:
    """util/losses.py:70-116 with the raw negative draws supplied by the caller (`temp` is ignored by
    the reference's inner call, losses.py:107-108, so logits are divided by 1)."""
    def sim(A, Bm, raw):
        Bsz, C, T = A.shape
        idx = raw.clone()
        self_idx = torch.arange(T).unsqueeze(-1).expand(-1, idx.shape[-1])
        idx[idx >= self_idx] += 1                        # skip self: losses.py:80-81
        with torch.no_grad():
            negs = A.unsqueeze(2).expand(-1, -1, T, -1).gather(3, idx.unsqueeze(1).expand(-1, C, -1, -1))
        targets = torch.cat([Bm.unsqueeze(-1), negs], dim=-1)
        return F.cosine_similarity(A.unsqueeze(-1), targets, dim=1)
    logits = torch.cat((sim(X, Y, raw_X), sim(Y, X, raw_Y)), dim=0)       # [2B, T, 1+N]
    tgt = torch.zeros(logits.shape[:-1], dtype=torch.long)
    return F.cross_entropy(logits.transpose(1, 2), tgt)


# ----------------------------------------------------------------------------- the train step

def one_hot(labels: torch.Tensor, n: int, dtype) -> torch.Tensor:
    """train.py:39-44."""
    return F.one_hot(labels, n).to(dtype)
