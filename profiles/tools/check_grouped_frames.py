"""Design check for DESIGN.md §8 item 2 (CPU, plain PyTorch fp64, no kernels): the discriminators' grouped
Conv1d(k=41, stride=4, groups=Cin/4) equals a stride-1 GROUPED convolution over frames of 4 samples whose groups are
bundles of 8 original groups with block-diagonal weights -- the shape the grouped tcgen05 entry point takes.
  1. pad the kernel to 44 taps (11 frames);
  2. frame view in CHANNEL-major order: xs[b, c*s + p, q] = x_pad[b, c, s*q + p]  (a group's 4x4 frame channels contiguous);
  3. per bundle of `gb` groups: weight tile [gb*cout_g, gb*cin_g*s, 11], block diagonal.
usage: python profiles/tools/check_grouped_frames.py"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)


def frames_c_major(x, s, pad, Tq):
    B, C, T = x.shape
    xp = F.pad(x, (pad, s * Tq + s))[:, :, :s * Tq].reshape(B, C, Tq, s)        # [b, c, q, p]
    return xp.permute(0, 1, 3, 2).reshape(B, C * s, Tq)                           # channel index c*s + p


def bundled_weights(w, groups, s, gb):
    """w[Cout, cin_g, K] (K already a multiple of s) -> [Cout, gb*cin_g*s, K/s] for a conv with groups/gb groups."""
    Cout, cin_g, K = w.shape
    m, cout_g = K // s, Cout // groups
    wf = w.view(Cout, cin_g, m, s).permute(0, 1, 3, 2).reshape(Cout, cin_g * s, m)    # [(co), (ci*s + p), j]
    out = torch.zeros(Cout, gb * cin_g * s, m, dtype=w.dtype)
    for g in range(groups):
        slot = g % gb                                                                   # position inside the bundle
        out[g * cout_g:(g + 1) * cout_g, slot * cin_g * s:(slot + 1) * cin_g * s] = wf[g * cout_g:(g + 1) * cout_g]
    return out


for (B, Cin, T, Cout, groups) in [(2, 16, 600, 64, 4), (2, 64, 150, 256, 16), (1, 256, 70, 1024, 64), (1, 1024, 35, 1024, 256)]:
    K, s, pad = 41, 4, 20
    x = torch.randn(B, Cin, T, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin // groups, K, dtype=torch.float64, requires_grad=True)
    ref = F.conv1d(x, w, stride=s, padding=pad, groups=groups)
    proj = torch.randn_like(ref)
    gx, gw = torch.autograd.grad((ref * proj).sum(), (x, w))

    Kp = -(-K // s) * s                                      # 44
    wpad = F.pad(w, (0, Kp - K))
    gb = min(8, groups)
    Tq = (T + 2 * pad + (Kp - K)) // s                       # the zero taps may look past the right edge: pad the view
    xs = frames_c_major(x, s, pad, Tq)
    wb = bundled_weights(wpad, groups, s, gb)
    y = F.conv1d(xs, wb, groups=groups // gb)
    y = y[:, :, :ref.shape[2]]
    gx2, gw2 = torch.autograd.grad((y * proj[:, :, :y.shape[2]]).sum(), (x, w))
    print(f"Cin={Cin:4d} Cout={Cout:4d} groups={groups:3d}: out {tuple(y.shape)} == {tuple(ref.shape)}  "
          f"|dy|={float((y - ref).detach().abs().max()):.1e} |dx|={float((gx2 - gx).abs().max()):.1e} |dw|={float((gw2 - gw).abs().max()):.1e}  "
          f"tensor-core group: {gb * (Cin // groups) * s} in x {gb * (Cout // groups)} out channels, 11 taps, {groups // gb} groups")
