"""Waveform Generator on the tdvc kernels (reference: model/generator.py).

Same constructors, forward signatures, sub-module names and state_dict keys as the reference, so
`train.py` / `generate_with_target.py` and existing checkpoints work unchanged.  The module tree keeps
the reference's Sequential / ModuleList slot layout (LeakyReLU / Tanh / Identity slots are kept as
parameter-free placeholders so child indices, and therefore checkpoint keys, are identical); the
forward passes do not run slot by slot but fuse every LeakyReLU into the following convolution's input
read, every bias / residual / tanh into its epilogue, and build the 136-channel conditioning tensor with
one kernel per scale.
"""
import torch
import torch.nn as nn

import util
from model.conditional_instance_norm import ConditionalInstanceNorm
from tdvc import ops
from tdvc.layers import (Conv1d, ConvTranspose1d, Identity, InstanceNorm1d, LeakyReLU, Linear, Tanh,
                         maybe_weight_norm)


def _wn(weight_norm):
    """The reference passes either a callable (lambda x: x / nn.utils.weight_norm) or, through
    util.get_weight_norm, our flag object.  Map all of them to a bool."""
    if callable(weight_norm):
        return bool(getattr(weight_norm, "tdvc_weight_norm", False)) or weight_norm is nn.utils.weight_norm
    return maybe_weight_norm(weight_norm)


def _run_chain(mods, x, c=None):
    """Interpreter for the reference's [norm, LeakyReLU, conv, ...] slot lists: a LeakyReLU slot is folded into
    the next conv (or instance norm output), Tanh into the previous conv."""
    mods = list(mods)
    slope = 1.0
    i = 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if isinstance(m, LeakyReLU):
            slope = m.negative_slope
        elif isinstance(m, Identity) or m is None:
            pass
        elif isinstance(m, Conv1d):
            act = "tanh" if isinstance(nxt, Tanh) else None
            x = m(x, in_slope=slope, out_act=act)
            slope = 1.0
            if act:
                i += 1
        elif isinstance(m, ConvTranspose1d):
            x = m(x, in_slope=slope)
            slope = 1.0
        elif isinstance(m, (InstanceNorm1d,)):
            if slope != 1.0:
                x = ops.leaky_relu(x, slope)
                slope = 1.0
            if isinstance(nxt, LeakyReLU):
                x = m(x, out_slope=nxt.negative_slope)
                i += 1
            else:
                x = m(x)
        elif isinstance(m, ConditionalInstanceNorm):
            if isinstance(c, ops.CondParts):
                c = c.tensor()
            if slope != 1.0:
                x = ops.leaky_relu(x, slope)
                slope = 1.0
            if isinstance(nxt, LeakyReLU):
                x = m(x, c, out_slope=nxt.negative_slope)
                i += 1
            else:
                x = m(x, c)
        elif isinstance(m, (CINResnetBlock, FiLMResnetBlock, MRFBlock)):
            if slope != 1.0:
                x = ops.leaky_relu(x, slope)
                slope = 1.0
            if isinstance(c, ops.CondParts) and not isinstance(m, MRFBlock):
                c = c.tensor()
            x = m(x, c)
        else:
            if slope != 1.0:
                x = ops.leaky_relu(x, slope)
                slope = 1.0
            x = m(x)
        i += 1
    if slope != 1.0:
        x = ops.leaky_relu(x, slope)
    return x


class DecoderResnetBlock(nn.Module):
    """reference model/generator.py:11-26 (legacy, weight-normed)."""

    def __init__(self, n_channel, dilation=1, kernel_size=3, leaky_relu_slope=0.2):
        super().__init__()
        self.block = nn.Sequential(
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=kernel_size, dilation=dilation, padding=dilation,
                   padding_mode='reflect', weight_norm=True),
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=1, weight_norm=True))
        self.shortcut = Conv1d(n_channel, n_channel, kernel_size=1, weight_norm=True)

    def forward(self, x):
        s = self.block[0].negative_slope
        h = self.block[1](x, in_slope=s)
        return self.block[3](h, in_slope=self.block[2].negative_slope, residual=self.shortcut(x))


class TranformResnetBlock(nn.Module):
    """reference model/generator.py:29-46 (legacy; relu-conv-norm order)."""

    def __init__(self, n_channel, dilation=1, kernel_size=3, leaky_relu_slope=0.2, norm_layer=InstanceNorm1d):
        super().__init__()
        self.block = nn.Sequential(
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=kernel_size, dilation=dilation, padding=dilation,
                   padding_mode='reflect'),
            norm_layer(n_channel),
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=1),
            norm_layer(n_channel))
        self.shortcut = Conv1d(n_channel, n_channel, kernel_size=1)

    def forward(self, x):
        return ops.add_scale(_run_chain(self.block, x), self.shortcut(x))


class ResnetBlock(nn.Module):
    """reference model/generator.py:48-67 (legacy; identity shortcut)."""

    def __init__(self, n_channel, dilation=1, kernel_size=3, leaky_relu_slope=0.2, norm_layer=InstanceNorm1d,
                 weight_norm=lambda x: x):
        super().__init__()
        wn = _wn(weight_norm)
        self.block = nn.Sequential(
            norm_layer(n_channel),
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=kernel_size, dilation=dilation, padding=dilation,
                   padding_mode='reflect', weight_norm=wn),
            norm_layer(n_channel),
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=1, weight_norm=wn))
        self.shortcut = Identity()

    def forward(self, x):
        return ops.add_scale(_run_chain(self.block, x), x)


class FiLMResnetBlock(nn.Module):
    """reference model/generator.py:69-111.
    h = conv_{k,d,reflect}(lrelu(x)); (gamma, beta) = cond_var(c); out = conv1x1(lrelu(h*(1+gamma)+beta)) + x."""

    def __init__(self, n_channel, n_cond_const, n_cond_var=0, dilation=1, kernel_size=3, leaky_relu_slope=0.2,
                 weight_norm=lambda x: x):
        super().__init__()
        wn = _wn(weight_norm)
        self.use_scale = True
        self.conv = nn.Sequential(
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=kernel_size, dilation=dilation,
                   padding=(kernel_size * dilation - dilation) // 2, padding_mode='reflect', weight_norm=wn))
        self.posconv = nn.Sequential(
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=1, weight_norm=wn))
        if n_cond_const or n_cond_var:
            nc = n_cond_const + n_cond_var
            self.cond_var = nn.Sequential(
                Conv1d(nc, nc, kernel_size=3, padding='same', weight_norm=wn),
                LeakyReLU(leaky_relu_slope),
                Conv1d(nc, n_channel * 2, kernel_size=3, padding='same', weight_norm=wn))
        self.shortcut = Identity()

    def forward(self, x, c=None, gb=None):
        """`gb`: gamma|beta already computed for this block by the stage-level fused conditioning path
        (MRFBlock.forward); otherwise cond_var runs here."""
        h = self.conv[1](x, in_slope=self.conv[0].negative_slope)
        pos = self.posconv[1]
        if (c is None or c.ndim == 3) and x.is_cuda and ops.film_posconv_eligible(pos.in_channels) and pos.kernel_size == 1:
            # bf16 mode: FiLM + LeakyReLU + pack in one pass, mask of the posconv dgrad in its epilogue
            if gb is None and c is not None:
                g = self.cond_var[0](c)
                gb = self.cond_var[2](g, in_slope=self.cond_var[1].negative_slope)
            return ops.film_posconv(h, gb, pos.effective_weight(), pos.bias, x, self.posconv[0].negative_slope)
        if gb is not None:
            h = ops.film(h, gb)
        elif c is not None:
            if c.ndim == 2:
                # the reference dereferences an attribute that is never defined here (generator.py:100)
                raise AttributeError("'FiLMResnetBlock' object has no attribute 'cond'")
            g = self.cond_var[0](c)
            g = self.cond_var[2](g, in_slope=self.cond_var[1].negative_slope)
            h = ops.film(h, g)
        return self.posconv[1](h, in_slope=self.posconv[0].negative_slope, residual=x)


class CINResnetBlock(nn.Module):
    """reference model/generator.py:113-139: CIN -> LReLU -> conv(k,d,reflect) -> CIN -> LReLU -> conv1x1, plus a
    1x1 shortcut.  CIN's affine and the following LeakyReLU run in one kernel."""

    def __init__(self, n_channel, n_cond, dilation=1, kernel_size=3, leaky_relu_slope=0.2):
        super().__init__()
        self.block = nn.ModuleList([
            ConditionalInstanceNorm(n_channel, n_cond),
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=kernel_size, dilation=dilation,
                   padding=(kernel_size * dilation - dilation) // 2, padding_mode='reflect'),
            ConditionalInstanceNorm(n_channel, n_cond),
            LeakyReLU(leaky_relu_slope),
            Conv1d(n_channel, n_channel, kernel_size=1)])
        self.shortcut = Conv1d(n_channel, n_channel, kernel_size=1)

    def _residual(self, x, c):
        return _run_chain(self.block, x, c)

    def forward(self, x, c):
        b = self.block
        h = b[0](x, c, out_slope=b[1].negative_slope)
        h = b[2](h)
        h = b[3](h, c, out_slope=b[4].negative_slope)
        return b[5](h, residual=self.shortcut(x))


class ExciteDownsampleBlock(nn.Module):
    """reference model/generator.py:141-173: strided conv stack plus a (1x1 conv -> Kaiser low-pass decimator)
    shortcut.  The decimator is a depthwise FIR with 16r+1 taps held in a non-persistent buffer."""

    def __init__(self, in_channel, out_channel, scale_factor, n_layers=2, kernel_size=5, leaky_relu_slope=0.2,
                 weight_norm=lambda x: x):
        super().__init__()
        wn = _wn(weight_norm)
        self.scale_factor = scale_factor
        self.block = nn.ModuleList()
        self.block += [Conv1d(in_channel, out_channel, kernel_size=2 * scale_factor, stride=scale_factor,
                              padding=scale_factor // 2, weight_norm=wn)]
        for _ in range(n_layers):
            self.block += [LeakyReLU(leaky_relu_slope),
                           Conv1d(out_channel, out_channel, kernel_size=kernel_size, stride=1, padding='same',
                                  weight_norm=wn)]
        self.shortcut = Conv1d(in_channel, out_channel, kernel_size=1)
        f = util.kaiser_filter(16 * scale_factor, 1 / scale_factor)
        f = f.expand(out_channel, 1, -1)
        self.register_buffer('shortcut_filter', f, persistent=False)

    def forward(self, x):
        x_sh = self.shortcut(x)
        x_sh = ops.conv1d(x_sh, self.shortcut_filter, None, stride=self.scale_factor,
                          padding=8 * self.scale_factor, groups=x_sh.shape[1])
        mods = list(self.block)
        h = mods[0](x)
        slope = 1.0
        last = len(mods) - 1
        for i, m in enumerate(mods[1:], start=1):
            if isinstance(m, LeakyReLU):
                slope = m.negative_slope
            else:
                h = m(h, in_slope=slope, residual=x_sh if i == last else None)
                slope = 1.0
        if not isinstance(mods[last], Conv1d) or last == 0:
            h = ops.add_scale(h, x_sh)
        return h


class MRFBlock(nn.Module):
    """Multi-receptive-field fusion (reference model/generator.py:175-194): 3 kernel sizes x 3 dilations of
    FiLMResnetBlocks, branch outputs averaged."""

    def __init__(self, n_channel, n_cond_const=0, n_cond_var=0, dilations=[1, 3, 5], kernel_sizes=[3, 7, 11],
                 leaky_relu_slope=0.2, weight_norm=lambda x: x):
        super().__init__()
        self.blocks = nn.ModuleList([nn.ModuleList() for i in range(len(kernel_sizes))])
        self.has_cond = n_cond_const > 0 or n_cond_var > 0
        self.kernel_sizes, self.dilations, self.n_channel = list(kernel_sizes), list(dilations), n_channel
        for i, kernel_size in enumerate(kernel_sizes):
            for dilation in dilations:
                self.blocks[i].append(FiLMResnetBlock(n_channel, n_cond_const, n_cond_var, dilation, kernel_size,
                                                      leaky_relu_slope, weight_norm))

    def _branch(self, i, x, c, gbs):
        """One kernel-size branch: its FiLM blocks in sequence."""
        nd = len(self.blocks[i])
        xs = x
        for j, mod in enumerate(self.blocks[i]):
            if gbs is not None:
                xs = mod(xs, None, gb=gbs[i * nd + j])
            else:
                xs = mod(xs, c)
        return xs

    def _forward_branches_concurrent(self, x, c, gbs):
        """The kernel-size branches are independent chains of small kernels: run them on forked CUDA streams (autograd
        replays each node's backward on its forward stream, and a CUDA-graph capture records the fork/join), so the
        GPU overlaps three latency-bound chains instead of idling between their launches."""
        cur = torch.cuda.current_stream()
        side = ops.branch_streams(x.device, len(self.blocks) - 1)
        fork = torch.cuda.Event()
        fork.record(cur)
        outs = [None] * len(self.blocks)
        joins = []
        for i in range(1, len(self.blocks)):
            s = side[i - 1]
            s.wait_event(fork)
            with torch.cuda.stream(s):
                outs[i] = self._branch(i, x, c, gbs)
                ev = torch.cuda.Event()
                ev.record(s)
                joins.append(ev)
            # tensors made on `cur` are read on `s` and vice versa: tell the caching allocator
            for t in (x, c) + tuple(gbs or ()):
                if t is not None:
                    t.record_stream(s)
            outs[i].record_stream(cur)
        outs[0] = self._branch(0, x, c, gbs)
        for ev in joins:
            cur.wait_event(ev)
        return outs

    def _fused_cond(self, c):
        """All blocks' gamma|beta from one grouped tensor-core pass over c (bf16 mode), or None."""
        if c is None or c.ndim != 3 or not self.has_cond:
            return None
        mods = [m for block in self.blocks for m in block]
        cv0 = mods[0].cond_var[0]
        if not ops.mrf_cond_path_eligible(c.shape[1], mods[0].cond_var[2].out_channels, c.shape[2]):
            return None
        if cv0.kernel_size != 3 or cv0.in_channels != c.shape[1]:
            return None
        wb = [(m.cond_var[0].effective_weight(), m.cond_var[0].bias, m.cond_var[2].effective_weight(), m.cond_var[2].bias)
              for m in mods]
        return ops.mrf_cond_path(c, wb, slope=mods[0].cond_var[1].negative_slope)

    def _fused_stage(self, x, c):
        """bf16 mode: the whole stage as one autograd node whose intermediates stay bf16 channels-last (tdvc.ops._MRFStage),
        or None when the stage is not of the form that path covers."""
        if not x.is_cuda or (c is not None and (c.ndim != 3 or not self.has_cond)) or (c is None and self.has_cond):
            return None
        if not ops.mrf_stage_eligible(self.n_channel, x.shape[2], self.kernel_sizes, self.dilations, c is not None,
                                      c.shape[1] if c is not None else 0):
            return None
        slopes = {m.conv[0].negative_slope for row in self.blocks for m in row} | \
                 {m.posconv[0].negative_slope for row in self.blocks for m in row}
        if len(slopes) != 1 or any(m.posconv[1].kernel_size != 1 for row in self.blocks for m in row):
            return None
        rows = []
        for row in self.blocks:
            r = []
            for m in row:
                t = (m.conv[1].effective_weight(), m.conv[1].bias, m.posconv[1].effective_weight(), m.posconv[1].bias)
                if c is not None:
                    t += (m.cond_var[0].effective_weight(), m.cond_var[0].bias, m.cond_var[2].effective_weight(),
                          m.cond_var[2].bias)
                r.append(t)
            rows.append(r)
        cond_slope = self.blocks[0][0].cond_var[1].negative_slope if c is not None else 0.2
        return ops.mrf_stage(x, c, rows, self.kernel_sizes, self.dilations, slope=slopes.pop(), cond_slope=cond_slope)

    def forward(self, x, c=None):
        y = self._fused_stage(x, c)
        if y is not None:
            return y
        if isinstance(c, ops.CondParts):
            c = c.tensor()
        gbs = self._fused_cond(c)
        if ops.branch_streams_enabled() and x.is_cuda and len(self.blocks) > 1:
            outs = self._forward_branches_concurrent(x, c, gbs)
        else:
            outs = [self._branch(i, x, c, gbs) for i in range(len(self.blocks))]
        n = len(outs)
        if n <= 3:
            return ops.add_scale(*outs, alpha=1.0 / n)
        y = ops.add_scale(*outs[:3])
        for o in outs[3:-1]:
            y = ops.add_scale(y, o)
        return ops.add_scale(y, outs[-1], alpha=1.0 / n)


class Encoder(nn.Module):
    """Conv content encoder (reference model/generator.py:197-272)."""

    def __init__(self, downsample_ratios, channel_sizes, n_res_blocks, conditional_dim=0, embedding_dim=None,
                 norm_layer=InstanceNorm1d, weight_norm=lambda x: x):
        super().__init__()
        wn = _wn(weight_norm)
        model = nn.ModuleList()
        self.spk_conditioning = conditional_dim > 0
        self.cin = norm_layer is ConditionalInstanceNorm
        if self.cin and not self.spk_conditioning:
            print('WARNING: Using conditional instance normalization but conditional dimension is 0')
        leaky_relu_slope = 0.2
        resblock_dilations = [1, 3, 5]
        resblock_kernel_sizes = [3, 7, 11]
        model += [Conv1d(1, channel_sizes[0], kernel_size=7, padding=3, padding_mode='reflect', weight_norm=wn)]
        channel_sizes[0] += conditional_dim if not self.cin else 0
        for i, r in enumerate(downsample_ratios):
            model += [norm_layer(channel_sizes[i]) if not self.cin else norm_layer(channel_sizes[i], conditional_dim),
                      LeakyReLU(leaky_relu_slope),
                      Conv1d(channel_sizes[i], channel_sizes[i + 1], kernel_size=2 * r, stride=r,
                             padding=r // 2 + r % 2, weight_norm=wn)]
            model += [MRFBlock(channel_sizes[i + 1], n_cond_const=0, n_cond_var=0, dilations=resblock_dilations,
                               kernel_sizes=resblock_kernel_sizes, leaky_relu_slope=leaky_relu_slope,
                               weight_norm=weight_norm)]
        model += [LeakyReLU(leaky_relu_slope),
                  Conv1d(channel_sizes[-1], channel_sizes[-1], kernel_size=7, stride=1, padding=3, weight_norm=wn)]
        if embedding_dim:
            model += [LeakyReLU(leaky_relu_slope),
                      Conv1d(channel_sizes[-1], embedding_dim, kernel_size=7, stride=1, padding=3, bias=False,
                             weight_norm=wn)]
        self.encoder = model

    def forward(self, x, c=None):
        if not self.cin:
            x = self.encoder[0](x)
            if self.spk_conditioning:
                x = ops.cond_concat_front(x, c)
            x = _run_chain(list(self.encoder)[1:], x)
        else:
            x = _run_chain(self.encoder, x, c)
        return ops.l2_normalize(x)


class Decoder(nn.Module):
    """Speaker / excitation conditioned upsampling decoder (reference model/generator.py:276-406)."""

    def __init__(self, upsample_ratios, channel_sizes, n_res_blocks, conditional_dim=0, embedding_dim=None,
                 norm_layer=InstanceNorm1d, weight_norm=lambda x: x):
        super().__init__()
        wn = _wn(weight_norm)
        model = nn.ModuleList()
        self.spk_conditioning = conditional_dim > 0
        self.cin = True
        if self.cin and not self.spk_conditioning:
            print('WARNING: Using conditional instance normalization but conditional dimension is 0')
        channel_sizes[0] += conditional_dim if not self.cin else 0
        leaky_relu_slope = 0.2
        self.upsample_ratios = upsample_ratios
        self.upsample_idxs = []
        excite_channels = [8, 8, 8, 8, 8]
        resblock_dilations = [1, 3, 5]
        resblock_kernel_sizes = [3, 7, 11]
        self.subsample_out_layers = nn.ModuleList()
        subsample_out = [False, True, True, False]
        if embedding_dim:
            model += [LeakyReLU(leaky_relu_slope),
                      Conv1d(embedding_dim, channel_sizes[0], kernel_size=7, stride=1, padding=3, bias=False,
                             weight_norm=wn)]
        model += [LeakyReLU(leaky_relu_slope),
                  Conv1d(channel_sizes[0], channel_sizes[0], kernel_size=7, stride=1, padding=3, weight_norm=wn)]
        for i, r in enumerate(upsample_ratios):
            model += [norm_layer(channel_sizes[i]) if not self.cin else norm_layer(channel_sizes[i], conditional_dim),
                      LeakyReLU(leaky_relu_slope),
                      ConvTranspose1d(channel_sizes[i], channel_sizes[i + 1], kernel_size=2 * r, stride=r,
                                      padding=r // 2 + r % 2, output_padding=r % 2, weight_norm=wn)]
            self.upsample_idxs.append(len(model))
            model += [MRFBlock(channel_sizes[i + 1], conditional_dim, excite_channels[i + 1],
                               dilations=resblock_dilations, kernel_sizes=resblock_kernel_sizes,
                               leaky_relu_slope=leaky_relu_slope, weight_norm=weight_norm)]
            if subsample_out[i]:
                out_block = nn.Sequential(LeakyReLU(leaky_relu_slope),
                                          Conv1d(channel_sizes[i + 1], 1, kernel_size=7, padding=3,
                                                 padding_mode='reflect', weight_norm=wn),
                                          Tanh())
                self.subsample_out_layers.append(out_block)
            else:
                self.subsample_out_layers.append(None)
        model += [norm_layer(channel_sizes[-1]) if not self.cin else norm_layer(channel_sizes[-1], conditional_dim),
                  LeakyReLU(leaky_relu_slope),
                  Conv1d(channel_sizes[-1], 1, kernel_size=7, padding=3, padding_mode='reflect', weight_norm=wn),
                  Tanh()]
        self.upsample_idxs.append(len(model))
        self.decoder = model
        self.excite_downsample = nn.ModuleList()
        for r, ch_in, ch_out in zip(self.upsample_ratios, excite_channels[:-1], excite_channels[1:]):
            self.excite_downsample += [ExciteDownsampleBlock(ch_in, ch_out, r, weight_norm=weight_norm)]
        self.excite_downsample += [Conv1d(1, excite_channels[0], kernel_size=7, padding=3, padding_mode='reflect',
                                          weight_norm=wn)]

    def get_scaled_conditioning(self, c):
        scaled_c = []
        for mod in reversed(self.excite_downsample):
            c = mod(c)
            scaled_c.append(c)
        return scaled_c

    def forward(self, x, c=None, c_var=None, out_subsample=False):
        subsample_out = []
        if not self.cin:  # unreachable in the reference too (cin is hard-coded True, generator.py:284)
            if self.spk_conditioning:
                x = ops.cond_concat_front(x, c)
            x = _run_chain(self.decoder, x)
        else:
            if c_var is None:
                # the reference fails on an undefined local here (generator.py:384-391)
                raise UnboundLocalError("cannot access local variable 'curr_scale' where it is not associated with a value")
            curr_scale = 0
            c_var_scales = self.get_scaled_conditioning(c_var)
            # c is time-constant: cat([c.repeat(time), excitation_scale]) is built by one kernel per scale
            cc = ops.cond_parts(c, c_var_scales[-1])
            mods = list(self.decoder)
            seg_start = 0
            bounds = [i for i in self.upsample_idxs if i < len(mods)]
            for b in bounds:
                x = _run_chain(mods[seg_start:b], x, cc)
                head = self.subsample_out_layers[curr_scale]
                if head is not None:
                    subsample_out.append(_run_chain(head, x))
                curr_scale += 1
                cc = ops.cond_parts(c, c_var_scales[-1 - curr_scale])
                seg_start = b
            x = _run_chain(mods[seg_start:], x, cc)
        if out_subsample:
            return x, subsample_out
        return x


class Generator(nn.Module):
    """reference model/generator.py:409-508."""

    def __init__(self, decoder_ratios, decoder_channels, num_bottleneck_layers, num_classes, conditional_dim,
                 content_dim=None, num_res_blocks=3, num_enc_layers=0, encoder_model=None,
                 norm_layer=None, weight_norm=None, bot_cond='target', enc_cond=None, dec_cond=None,
                 output_content_emb=False):
        super().__init__()
        self.output_content_emb = output_content_emb
        if type(norm_layer) is not tuple:
            nl = util.get_norm_layer(norm_layer)
            enc_norm_layer = dec_norm_layer = bot_norm_layer = nl
        else:
            bot_norm_layer = util.get_norm_layer(norm_layer[0])
            enc_norm_layer = util.get_norm_layer(norm_layer[1])
            dec_norm_layer = util.get_norm_layer(norm_layer[2])
        if type(weight_norm) is not tuple:
            nl = util.get_weight_norm(weight_norm)
            enc_weight_norm = dec_weight_norm = bot_weight_norm = nl
        else:
            bot_weight_norm = util.get_weight_norm(weight_norm[0])
            enc_weight_norm = util.get_weight_norm(weight_norm[1])
            dec_weight_norm = util.get_weight_norm(weight_norm[2])
        bot_cond_dim = conditional_dim if bot_cond == 'target' else 2 * conditional_dim
        enc_cond_dim = 0 if enc_cond is None else conditional_dim
        dec_cond_dim = 0 if dec_cond is None else conditional_dim
        self.both_cond = bot_cond == 'both'
        self.cin = True   # hard-coded in the reference (generator.py:449-450)
        self.decoder = Decoder(decoder_ratios, decoder_channels[:], num_res_blocks, dec_cond_dim, content_dim,
                               dec_norm_layer, dec_weight_norm)
        if encoder_model in ['wavlm']:
            from model.ssl_encoder import SSLEncoder   # WN stack on the tdvc kernels; WavLM itself comes from a reference checkout
            self.encoder = SSLEncoder(encoder_model, num_enc_layers, content_dim, weight_norm=_as_torch_wn(enc_weight_norm))
        else:
            self.encoder = Encoder(decoder_ratios[::-1], decoder_channels[::-1], num_res_blocks, enc_cond_dim,
                                   content_dim, enc_norm_layer, enc_weight_norm)
        bottleneck = nn.ModuleList()
        bot_dim = decoder_channels[0]
        for i in range(num_bottleneck_layers):
            bottleneck += [FiLMResnetBlock(bot_dim, bot_cond_dim, dilation=1, weight_norm=bot_weight_norm)]
        self.bottleneck = bottleneck
        self.embedding = Linear(num_classes, conditional_dim)

    def _bottleneck(self, x, c):
        for mod in self.bottleneck:
            x = mod(x, c)
        return x

    def forward(self, x, c_tgt, c_src=None, c_var=None, out_subsample=False):
        c_tgt = self.embedding(c_tgt)
        c_src = self.embedding(c_src) if c_src is not None else None
        x = self.encoder(x)
        if self.output_content_emb:
            self.content_embedding = x
        if self.both_cond:
            c = ops.cat_channels_2d(c_src, c_tgt)
            x = self._bottleneck(x, c)
        else:
            x = self._bottleneck(x, c_tgt)
        return self.decoder(x, c_tgt, c_var, out_subsample=out_subsample)


def _as_torch_wn(flag):
    """SSLEncoder (reference file) expects a callable that wraps torch modules."""
    if callable(flag) and getattr(flag, "tdvc_weight_norm", False):
        return nn.utils.weight_norm
    if callable(flag):
        return flag
    return nn.utils.weight_norm if maybe_weight_norm(flag) else (lambda m: m)
