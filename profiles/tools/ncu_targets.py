"""Each kernel family of the training step launched in isolation at its benchmark shape, twice (the second launch of
every kernel is the one profiled: warm instruction cache, cold data like inside the step) -- the program `ncu --set full`
is pointed at (profiles/tools/gpu_run_ncu.sh).  Shapes: conv_enc-stage1, B = 16 per pass, two passes stacked (B = 32),
T = 8960, 136 conditioning channels.

    python profiles/tools/ncu_targets.py [mini|mrf|disc|hbm|all]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(REPO, "td-vc-gan_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from tdvc import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ops.set_precision("bf16")
B, T = 32, 8960


def rnd(*shape, scale=1.0, grad=False):
    return (torch.randn(*shape, device=dev) * scale).requires_grad_(grad)


def twice(fn):
    """first call outside the profiler range (ncu --profile-from-start off), second one inside"""
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def mrf_blocks(C, Cc, ks, ds):
    def blk(k):
        t = [rnd(C, C, k, scale=(C * k) ** -0.5, grad=True), rnd(C, scale=0.1, grad=True),
             rnd(C, C, 1, scale=C ** -0.5, grad=True), rnd(C, scale=0.1, grad=True)]
        if Cc:
            t += [rnd(Cc, Cc, 3, scale=(3 * Cc) ** -0.5, grad=True), rnd(Cc, scale=0.1, grad=True),
                  rnd(2 * C, Cc, 3, scale=0.5 * (3 * Cc) ** -0.5, grad=True), rnd(2 * C, scale=0.1, grad=True)]
        return t
    return [[blk(k) for _ in ds] for k in ks]


def run_stage(C, Cc, Tq, ks, ds):
    x = rnd(B, C, Tq, grad=True)
    c = rnd(B, Cc, Tq, grad=True) if Cc else None
    blocks, proj = mrf_blocks(C, Cc, ks, ds), rnd(B, C, Tq)

    def stage():
        y = ops.mrf_stage(x, c, blocks, ks, ds)
        (y * proj).sum().backward()
    twice(stage)


if what in ("mini", "all"):
    # ONE FiLM block (k = 7, d = 3) of the full-rate decoder stage (C = 16, 136-channel conditioning) and of an encoder stage
    # (C = 64, T / 4): every kernel of the bf16-resident stage at its benchmark shape, ~30 launches instead of ~300
    run_stage(16, 136, T, (7,), (3,))
    run_stage(64, 0, T // 4, (7,), (3,))
    run_stage(256, 0, 28, (11,), (5,))

if what in ("mrf",):
    # the full 9-block stages: the stacked cond_var.0 launch (9 x 136 output rows) and the 3-branch chain kernels
    run_stage(16, 136, T, (3, 7, 11), (1, 3, 5))
    run_stage(64, 0, T // 4, (3, 7, 11), (1, 3, 5))

if what in ("disc", "mini", "all"):
    # discriminator.1.0 (16 -> 64, k41 s4, 4 groups) on the full-rate branch and the dense 1024 -> 1024 k5 layer
    x = rnd(B, 16, T, grad=True)
    w, b = rnd(64, 4, 41, scale=164 ** -0.5, grad=True), rnd(64, scale=0.1, grad=True)

    def grouped():
        y = ops.conv1d(x, w, b, stride=4, padding=20, groups=4, out_act="lrelu")
        y.sum().backward()
    twice(grouped)
    x3 = rnd(B, 1024, 35, grad=True)
    w3, b3 = rnd(1024, 1024, 5, scale=5120 ** -0.5, grad=True), rnd(1024, scale=0.1, grad=True)

    def dense():
        y = ops.conv1d(x3, w3, b3, padding=2, out_act="lrelu")
        y.sum().backward()
    twice(dense)

if what in ("hbm", "mini", "all"):
    # bandwidth-bound kernels: feature-matching L1 over one full-rate map, LSGAN term, CIN, AdamW, batched weight norm
    a, r = rnd(B, 16, T, grad=True), rnd(B // 2, 16, T)

    def l1():
        ops.l1_mean_sum_rows([a], B // 2, B // 2, [r]).backward()
    twice(l1)
    o = rnd(B, 1, 35, grad=True)
    twice(lambda: ops.mse_to_const_sum([o], 1.0).backward())
    xc, gbc = rnd(16, 64, 2240, grad=True), rnd(16, 128, 1, grad=True)
    twice(lambda: ops.cond_instance_norm(xc, gbc, out_slope=0.2).sum().backward())
    import bench
    from tdvc.optim import FusedAdamW
    G, _ = bench.build_models(dev)
    opt = FusedAdamW(G.parameters(), 1e-4, (0.8, 0.99))
    for p_ in G.parameters():
        p_.grad = torch.randn_like(p_) * 1e-3
    twice(opt.step)
    host = bench.synth_batch(2, T, 100, 1)
    xg, cv = host["signal_real"].to(dev), host["c_f0_conv"].to(dev)
    ct = torch.zeros(2, 100, device=dev)
    ct[:, 3] = 1

    def gfwd():
        with torch.no_grad(), ops.step_cache("G"):
            G(xg, ct, c_var=cv)
    for _ in range(2):       # the third scope runs the batched weight-norm / pack launches of the recorded plan
        gfwd()
        torch.cuda.synchronize()
    twice(gfwd)
print("ncu targets done:", what)
