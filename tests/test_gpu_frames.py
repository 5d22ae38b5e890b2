"""The discriminators' grouped k41 s4 convolutions (model/discriminator.py:26-30) on the tensor-core path: channel-major
frame view, bundled block-diagonal grouped convolution (forward, data gradient), grouped weight-gradient kernel
(conv_tc_wgrad2_k) -- against fp64 PyTorch.  bf16 operands / fp32 accumulation: asserted at 1e-2 (observed ~3e-3)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import relerr

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64) * scale


def dev(t):
    return t.detach().float().cuda().requires_grad_(t.requires_grad)


@pytest.fixture(autouse=True)
def bf16_mode():
    from tdvc import ops
    ops.set_precision("bf16")
    yield
    ops.set_precision("fp32")


@pytest.mark.parametrize("B,C,T,s,pad,Tq", [(2, 16, 2240, 4, 20, 570), (3, 8, 70, 4, 20, 28), (2, 40, 333, 2, 10, 180),
                                            (1, 1024, 70, 4, 20, 28)])
def test_frame_pack_and_unpack(B, C, T, s, pad, Tq):
    from tdvc import _lib, ops
    lib = _lib.load()
    x = rnd(B, C, T, seed=1).float().cuda()
    xf = torch.full((B, Tq, C * s), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.tdvc_frame_pack_bf16(ops._p(x), ops._p(xf), B, C, T, s, pad, Tq, ops._st()), "pack")
    xpad = F.pad(x, (pad, s * Tq))[:, :, :s * Tq]                       # sample s*q + p - pad at index s*q + p
    ref = xpad.view(B, C, Tq, s).permute(0, 2, 1, 3).reshape(B, Tq, C * s).to(torch.bfloat16)
    assert torch.equal(xf.view(torch.int16), ref.view(torch.int16))
    dxf = rnd(B, C * s, Tq, seed=2).float().cuda()
    dx = torch.full((B, C, T), float("nan"), device="cuda")
    _lib.check(lib.tdvc_frame_unpack(ops._p(dxf), ops._p(dx), B, C, T, s, pad, Tq, ops._st()), "unpack")
    full = dxf.view(B, C, s, Tq).permute(0, 1, 3, 2).reshape(B, C, Tq * s)    # index s*q + p = u + pad
    want = torch.zeros(B, C, T, device="cuda")
    n = min(T, Tq * s - pad)
    want[:, :, :n] = full[:, :, pad:pad + n]
    assert torch.equal(dx, want)


GROUPED = [
    # B, Cin, Cout, groups, T, K, stride, pad, act
    (2, 16, 64, 4, 2240, 41, 4, 20, "lrelu"),        # discriminator.1.0 (4 -> 16 per group), bundle = 1 group
    (2, 64, 256, 16, 560, 41, 4, 20, "lrelu"),       # discriminator.2.0
    (1, 256, 1024, 64, 140, 41, 4, 20, "lrelu"),     # discriminator.3.0
    (2, 1024, 1024, 256, 70, 41, 4, 20, "lrelu"),    # discriminator.4.0: 4 -> 4 per group (bundles of 4), ragged 70 -> 18
    (3, 64, 64, 16, 35, 41, 4, 20, None),            # T_out 9, no activation
    (2, 32, 64, 4, 301, 9, 2, 4, "lrelu"),           # another stride / kernel: 8 -> 16 per group, k9 s2
]


@pytest.mark.parametrize("case", GROUPED, ids=[str(i) for i in range(len(GROUPED))])
def test_grouped_strided_conv_tc(case):
    from tdvc import ops
    B, Cin, Cout, groups, T, K, s, p, act = case
    x = rnd(B, Cin, T, seed=1).requires_grad_(True)
    w = rnd(Cout, Cin // groups, K, seed=2, scale=(Cin // groups * K) ** -0.5).requires_grad_(True)
    b = rnd(Cout, seed=3, scale=0.1).requires_grad_(True)
    assert ops._grouped_frame_plan(Cin, Cout, K, s, groups, 1, False) is not None
    xd, wd, bd = dev(x), dev(w), dev(b)
    n0 = ops._lib.load().tdvc_flop_count(4)
    y = ops.conv1d(xd, wd, bd, stride=s, padding=p, groups=groups, out_act=act, out_slope=0.2)
    torch.cuda.synchronize()
    assert ops._lib.load().tdvc_flop_count(4) == n0          # nothing went to the fp32 CUDA-core kernels
    y0 = F.conv1d(x, w, b, stride=s, padding=p, groups=groups)
    assert relerr(y, F.leaky_relu(y0, 0.2) if act else y0) < 1e-2
    # gradients: the LeakyReLU branch is taken from the device (0.3 % of the pre-activations sit below the bf16 rounding
    # error; a flipped branch there is not a kernel error but would show as percent-level noise in dx)
    yref = y0 * torch.where(y.detach().double().cpu() > 0, 1.0, 0.2) if act else y0
    proj = rnd(*yref.shape, seed=4)
    (yref * proj).sum().backward()
    (y * proj.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert relerr(xd.grad, x.grad) < 1e-2
    assert relerr(wd.grad, w.grad) < 1e-2
    assert relerr(bd.grad, b.grad) < 1e-2


WG2 = [
    # B, T, Cin, Cout, [K per group], dil, haloed, taps on M (-1 = default: on for Cin 16 / 32 / 64)
    (2, 300, 16, 16, [11], 5, 1, -1),         # 8 taps per MMA (SWIZZLE_32B rows), 2 tap blocks
    (2, 300, 16, 16, [11], 5, 1, 0),          # same conv, channels only on M
    (2, 300, 16, 16, [11], 5, 0, -1),         # one copy of the x tile per tap
    (2, 520, 16, 16, [3, 7, 11], 3, 1, -1),   # the three branches of an MRF depth, one launch
    (2, 520, 16, 16, [3, 7, 11], 1, 1, -1),   # dilation 1: M atoms one row apart
    (2, 520, 32, 32, [3, 7, 11], 1, 1, -1),   # 4 taps per MMA (SWIZZLE_64B rows)
    (2, 520, 32, 32, [3, 7, 11], 5, 1, 0),
    (3, 200, 64, 64, [3, 7, 11], 5, 1, -1),   # 2 taps per MMA (SWIZZLE_128B rows)
    (3, 200, 64, 64, [3, 7, 11], 5, 1, 0),    # 11 taps x 64 columns do not fit TMEM: two tap groups
    (3, 200, 64, 64, [3, 7, 11], 5, 0, -1),
    (2, 2100, 16, 16, [11], 1, 1, -1),        # many time units per CTA (pipeline wrap-around)
    (1, 130, 128, 128, [3, 7, 11], 3, 1, -1),
    (2, 130, 128, 128, [3, 7, 11], 5, 1, -1),  # four tap groups of 3 taps, dilation 5
    (1, 28, 256, 256, [3, 7, 11], 5, 1, -1),   # encoder's deepest stage: one time unit, two ci tiles, four tap groups
    (2, 28, 256, 256, [7], 1, 1, -1),         # two ci tiles, T < one time unit
    (2, 333, 136, 32, [3], 1, 1, -1),         # Cin not a multiple of 64 (cond_var.2 shape)
    (2, 333, 24, 16, [1], 1, 1, -1),          # k = 1
    (2, 333, 64, 16, [11], 1, 1, -1),         # the frame view of discriminator.4.0: 64 frame channels -> 16 outputs
]


@pytest.mark.parametrize("case", WG2, ids=[str(i) for i in range(len(WG2))])
def test_wgrad2_groups(case):
    """conv_tc_wgrad2_k on branch-concatenated operands: x[B, T + 2H, G*Cin] (reflect-free, zero halo H = max pad) and
    dy[B, T, G*Cout]; group g is a 'same' conv with kg[g] taps at dilation dil."""
    from tdvc import ops
    B, T, Cin, Cout, ks, dil, haloed, tapsm = case
    G = len(ks)
    H = max(dil * (k - 1) // 2 for k in ks)
    x = rnd(B, G * Cin, T, seed=1)
    dy = rnd(B, G * Cout, T, seed=2)
    xd = x.float().cuda()
    xp = torch.zeros(B, T + 2 * H, G * Cin, device="cuda", dtype=torch.bfloat16)
    xp[:, H:H + T] = xd.transpose(1, 2).to(torch.bfloat16)
    dyp = dy.float().cuda().transpose(1, 2).contiguous().to(torch.bfloat16)
    dws = [torch.full((Cout, Cin, k), float("nan"), device="cuda") for k in ks]
    dbs = [torch.full((Cout,), float("nan"), device="cuda") for _ in ks]
    ops.wgrad2(dyp=dyp, xp=xp, B=B, Cdp=G * Cout, Tout=T, Cp=G * Cin, Tp=T + 2 * H, Cout=Cout, Cin=Cin, K=max(ks), dilation=dil,
               ngroups=G, per_group=True, x_ch_stride=Cin, dy_ch_stride=Cout, kg=ks, t_off=[H - dil * (k - 1) // 2 for k in ks],
               dw=dws, db=dbs, haloed=haloed, tapsm=tapsm)
    torch.cuda.synchronize()
    xb = xp.double().cpu().transpose(1, 2)          # bf16-rounded operands, padded
    dyb = dyp.double().cpu().transpose(1, 2)
    for g, k in enumerate(ks):
        xg = xb[:, g * Cin:(g + 1) * Cin]
        dg = dyb[:, g * Cout:(g + 1) * Cout]
        off = H - dil * (k - 1) // 2
        ref = torch.stack([torch.einsum("bot,bit->oi", dg, xg[:, :, off + tap * dil: off + tap * dil + T]) for tap in range(k)], 2)
        assert relerr(dws[g], ref) < 2e-3, g            # same rounded operands on both sides: only the fp32 sum order differs
        assert relerr(dbs[g], dg.sum(dim=(0, 2))) < 2e-3, g
    # the persistent workspace is left zeroed
    ws = ops._WGRAD_WS[dyp.device]
    assert float(ws.abs().max()) == 0.0


def test_latent_classifier_bf16_vs_golden():
    """LatentClassifier (k21 s2 dense convs as frame convolutions with one appended zero tap) in bf16 mode."""
    import numpy as np
    from helpers import golden
    from model.latent_classifier import LatentClassifier
    from oracle.cases import rand_like
    from oracle.params import make_state_dict
    g = golden("latcls")
    m = LatentClassifier(6, 16)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(make_state_dict(shapes, seed=9, dtype=torch.float32), strict=True)
    m.cuda()
    x = rand_like(torch.empty(3, 16, 28), 71).float().cuda().requires_grad_(True)
    out = m(x)
    assert relerr(out, g["out"]) < 2e-2
    loss = F.cross_entropy(out, torch.tensor([1, 4, 0]).cuda())
    assert abs(loss.item() - float(g["loss"])) < 2e-2
    loss.backward()
    # 28 -> 14 -> 7 -> 4 time steps: a handful of LeakyReLU branch flips weigh percents (the kernels of these layers are
    # held at 1e-2 without an activation in test_gpu_tc.py::test_strided_conv_as_frames): relative L2 of dL/dx within 1e-1
    ref = torch.as_tensor(np.asarray(g["dx"])).double()
    assert float((x.grad.double().cpu() - ref).norm() / ref.norm()) < 1e-1


WG2U = [
    # B, T, groups, Cin, x pitch, Cout, dy pitch, K, dil, swap
    (2, 333, 9, 136, 144, 32, 32, 3, 1, -1),      # the 9 cond_var.2 blocks of a stage: operand roles swapped (M = co)
    (2, 333, 9, 136, 144, 32, 32, 3, 1, 0),       # the same through the channels-on-M form (two M tiles, 8 real lanes in the second)
    (1, 700, 1, 137, 144, 1296, 1296, 3, 1, -1),  # the stacked cond_var.0 gradient (+ the constant-one channel): 11 M tiles of co
    (2, 70, 3, 200, 208, 48, 64, 5, 2, -1),       # another width / dilation, T barely over one time unit
]


@pytest.mark.parametrize("case", WG2U, ids=[str(i) for i in range(len(WG2U))])
def test_wgrad2_uniform_groups_and_swapped_operands(case):
    """conv_tc_wgrad2(s)_k on uniform groups with zero 'same' padding through TMA out-of-bounds fill (t_off = -pad), incl. the
    operand-swapped form used when the input width is just over one 128-lane tile."""
    from tdvc import ops
    B, T, G, Cin, xpitch, Cout, dpitch, K, dil, swap = case
    pad = dil * (K - 1) // 2
    x = rnd(B, G * xpitch, T, seed=1)
    dy = rnd(B, G * dpitch, T, seed=2)
    xp = x.float().cuda().transpose(1, 2).contiguous().to(torch.bfloat16)
    dyp = dy.float().cuda().transpose(1, 2).contiguous().to(torch.bfloat16)
    dw = torch.full((G, Cout, Cin, K), float("nan"), device="cuda")
    db = torch.full((G, Cout), float("nan"), device="cuda")
    ops.wgrad2(dyp=dyp, xp=xp, B=B, Cdp=G * dpitch, Tout=T, Cp=G * xpitch, Tp=T, Cout=Cout, Cin=Cin, K=K, dilation=dil, ngroups=G,
               x_ch_stride=xpitch, dy_ch_stride=dpitch, t_off=[-pad], dw=[dw], db=[db], dw_grp_stride=Cout * Cin * K,
               db_grp_stride=Cout, swap=swap)
    torch.cuda.synchronize()
    xb = F.pad(xp.double().cpu().transpose(1, 2), (pad, pad))
    dyb = dyp.double().cpu().transpose(1, 2)
    for g in range(G):
        xg = xb[:, g * xpitch:g * xpitch + Cin]
        dg = dyb[:, g * dpitch:g * dpitch + Cout]
        ref = torch.stack([torch.einsum("bot,bit->oi", dg, xg[:, :, tap * dil: tap * dil + T]) for tap in range(K)], 2)
        assert relerr(dw[g], ref) < 2e-3, g
        assert relerr(db[g], dg.sum(dim=(0, 2))) < 2e-3, g
    assert float(ops._WGRAD_WS[dyp.device].abs().max()) == 0.0
