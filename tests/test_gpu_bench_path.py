"""Parity of the configuration bench.py measures: bf16 tensor-core mode + the whole step replayed from a CUDA graph
+ the batched weight-norm / operand-pack launches of the step scopes + the persistent wgrad workspace -- as a
combination, not piece by piece (the per-kernel and per-module tests live in test_gpu_tc.py / test_gpu_models.py).

Tolerances (BASELINE.json north_star: 2e-2 in bf16):
  * loss scalars and the generated waveform: 2e-2 of the fp64 reference (golden vectors from the real reference);
  * gradients: LeakyReLU / |.| kinks flip branch where a pre-activation is below the bf16 rounding error, so individual
    entries cannot be held; asserted on the L2 norm of EVERY tensor with a stated worst-case bound next to the percentile
    bounds (values measured on B200 are written to gpurun_out/parity_bf16.json by the tests)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import golden, relerr, stats
from oracle.cases import CASES, HP_LATCLS, HP_STAGE1, HP_STAGE2_1, HP_STAGE2_2, HP_WAVE_CLIP, rand_like
from oracle.params import make_batch, make_state_dict
from test_host_cpu import build_D, build_G

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(key, value):
    """Measured parity numbers travel back from the GPU box in gpurun_out/ (scratch; summaries are copied to profiles/)."""
    d = os.path.join(REPO, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "parity_bf16.json")
    try:
        with open(p) as f:
            data = json.load(f)
    except Exception:
        data = {}
    data[key] = value
    with open(p, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


@pytest.fixture(autouse=True)
def bf16_mode():
    from tdvc import ops
    ops.set_precision("bf16")
    yield
    ops.set_precision("fp32")


def load_det(mod, seed):
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    mod.load_state_dict(make_state_dict(shapes, seed=seed, dtype=torch.float32), strict=True)
    return mod.cuda()


def cu(t):
    return t.float().cuda() if t.is_floating_point() else t.cuda()


def _models(cfg, hp, seeds=None):
    sg, sd = seeds or (cfg["seed"], cfg["seed"] + 100)
    G, D = load_det(build_G(cfg), sg), load_det(build_D(cfg), sd)
    C = None
    if hp["lambda_latcls"] != 0:
        from model.latent_classifier import LatentClassifier
        C = load_det(LatentClassifier(cfg["nspk"], cfg["content_dim"]), cfg["seed"] + 200)
    return G, D, C


def _batch(cfg, hp, seed=None):
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1 if seed is None else seed,
                   frames_div=int(np.prod(cfg["ratios"])), permute=not hp["no_conv"])
    return b, {k: (cu(v) if torch.is_tensor(v) else v) for k, v in b.items()}


def _norm_errs(grads, g, prefix):
    errs = {}
    for k, gr in grads.items():
        ref = g[f"{prefix}_grad/{k}"]
        if ref[2] < 1e-12:
            continue
        errs[f"{prefix}/{k}"] = abs(stats(gr)[2] - ref[2]) / ref[2]
    return errs


# --------------------------------------------------------------------------- batched launches == per-weight kernels
def test_batched_weight_norm_and_pack_are_bit_identical():
    """tdvc_weight_norm_fwd_multi / tdvc_pack_weight_bf16_multi (one launch each for every weight of a step scope)
    against the per-weight kernels tdvc_weight_norm_fwd / tdvc_pack_weight_bf16: same bits."""
    from tdvc import _lib, ops
    lib = _lib.load()
    shapes = [(16, 1, 7), (32, 16, 4), (136, 136, 3), (272, 136, 3), (5, 3, 1), (1024, 4, 41), (64, 64, 11), (256, 128, 20)]
    gen = torch.Generator().manual_seed(7)
    live = [(torch.nn.Parameter(torch.randn(*s, generator=gen).cuda()),
             torch.nn.Parameter((torch.rand(s[0], 1, 1, generator=gen) + 0.5).cuda())) for s in shapes]
    sc = ops._StepCache()
    pl = sc._plan("t")
    pl["plan_p"] = [(1, 32, 64, False), (1, 16, 64, True), (2, 144, 144, False), (2, 144, 144, True), (3, 272, 144, False),
                    (6, 64, 64, False), (6, 64, 64, True), (7, 256, 128, False)]
    t = sc._build_tables(pl, live, torch.device("cuda"))
    st = ops._st()
    flat_w = torch.full((t["w_elems"],), float("nan"), device="cuda")
    flat_inv = torch.empty(t["total_rows"], device="cuda")
    _lib.check(lib.tdvc_weight_norm_fwd_multi(ops._p(t["table"]), ops._p(t["rows_dev"]), t["n"], t["total_rows"],
                                              ops._p(flat_w), ops._p(flat_inv), st), "multi")
    singles = []
    for j, (v, g) in enumerate(live):
        rows, cols = v.shape[0], v.numel() // v.shape[0]
        w = torch.empty_like(v)
        inv = torch.empty(rows, device="cuda")
        _lib.check(lib.tdvc_weight_norm_fwd(ops._p(v), ops._p(g), ops._p(w), ops._p(inv), rows, cols, st), "single")
        singles.append(w)
        got = flat_w[t["w_off"][j]:t["w_off"][j] + v.numel()].view_as(v)
        assert torch.equal(got, w), j
        assert torch.equal(flat_inv[t["row_start"][j]:t["row_start"][j + 1]], inv), j
        ref = (g.double() * v.double() / v.double().reshape(rows, -1).norm(dim=1).reshape(g.shape))
        assert relerr(w, ref) < 1e-6
    flat_wp = torch.empty(t["p_elems"], device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.tdvc_pack_weight_bf16_multi(ops._p(t["jobs"]), len(pl["plan_p"]), t["blocks_per_job"], ops._p(flat_w),
                                               ops._p(flat_wp), st), "pack multi")
    for i, (idx, rows_p, cols_p, flip) in enumerate(pl["plan_p"]):
        Cout, Cin, K = shapes[idx]
        one = torch.empty(K, rows_p, cols_p, device="cuda", dtype=torch.bfloat16)
        coutp, cinp = (cols_p, rows_p) if flip else (rows_p, cols_p)
        _lib.check(lib.tdvc_pack_weight_bf16(ops._p(singles[idx]), ops._p(one), Cout, Cin, K, coutp, cinp, int(flip), 0, 0, 0, 0,
                                             st), "pack single")
        got = flat_wp[t["p_off"][i]:t["p_off"][i] + K * rows_p * cols_p].view(K, rows_p, cols_p)
        assert torch.equal(got.view(torch.int16), one.view(torch.int16)), i


# --------------------------------------------------------------------------- the graphed bf16 step
def _grads_after_one_iteration(cfg, hp, graphed):
    """Gradients (and losses) of one iteration at fixed weights: lr = 0 and no weight decay, so the eager warm-up
    iterations GraphedTrainStep runs before capturing leave the weights where they were."""
    from tdvc.optim import FusedAdamW
    from tdvc.train_step import GraphedTrainStep, TrainStep
    G, D, C = _models(cfg, hp, seeds=(11, 12))
    _, bd = _batch(cfg, hp, seed=9)
    bd.pop("neg_idx", None)
    mk = lambda m: FusedAdamW(m.parameters(), 0.0, (0.8, 0.99), weight_decay=0.0)
    oG, oD = mk(G), mk(D)
    oC = mk(C) if C is not None else None
    ts = TrainStep(G, D, hp, oG, oD, cfg["nspk"], C=C, optimizer_C=oC)
    grads = {}
    if graphed:
        gs = GraphedTrainStep(ts, bd, warmup=2)
        out = gs.step()
        torch.cuda.synchronize()
        for opt, mod, pre in ((oD, D, "D"), (oG, G, "G")) + (((oC, C, "C"),) if C is not None else ()):
            names = {id(p): k for k, p in mod.named_parameters()}
            for bank in opt._banks:
                for p, v in zip(bank["params"], bank["views"]):
                    grads[pre + "." + names[id(p)]] = v.detach().clone()
    else:
        out = ts.d_step(bd)
        for k, p in D.named_parameters():
            grads["D." + k] = p.grad.detach().clone()
        if C is not None:
            for k, p in C.named_parameters():
                grads["C." + k] = p.grad.detach().clone()
        out.update(ts.g_step(bd))
        torch.cuda.synchronize()
        for k, p in G.named_parameters():
            grads["G." + k] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    return grads, {k: float(v) for k, v in out.items() if torch.is_tensor(v) and v.numel() == 1}


@pytest.mark.parametrize("name,hp", [("s21", HP_STAGE2_1), ("s1", HP_STAGE1), ("s22", HP_STAGE2_2), ("latcls", HP_LATCLS)])
def test_graphed_bf16_step_equals_eager_bf16_step(name, hp):
    """GraphedTrainStep in bf16 mode (step scopes with the batched weight-norm / pack launches, persistent wgrad
    workspace, grad banks) produces the eager bf16 iteration's losses and gradients: every gradient tensor within 2e-2
    (max-abs-normalised; fp32 atomics in the split-K weight gradients reorder sums run to run, nothing else may differ --
    measured 1e-4 .. 5e-4 for the worst tensor of most runs, up to 6.6e-3 on 8-element weight_g gradients, which are
    near-total cancellations of the weight gradient against the weight direction).
    Gradients, not updated weights, are compared: Adam's m / sqrt(v) turns a sign flip of a noise-level gradient entry into
    a full lr-sized difference (measured: 4e-2 after three updates at lr 1e-3), which says nothing about the kernels.
    The contrastive term draws fresh negatives per call, so it is switched off here (pinned in test_gpu_models.py)."""
    cfg = CASES["step_tiny"]
    hp = dict(hp, lambda_cont_emb=0)
    g_e, out_e = _grads_after_one_iteration(cfg, hp, graphed=False)
    g_g, out_g = _grads_after_one_iteration(cfg, hp, graphed=True)
    for k in ("d_loss", "g_loss"):
        assert abs(out_e[k] - out_g[k]) <= 1e-5 * abs(out_e[k]), (k, out_e[k], out_g[k])
    assert set(g_e) == set(g_g)
    worst = max(relerr(g_g[k], g_e[k]) for k in g_e if g_e[k].abs().max() > 0)
    _record(f"graphed_vs_eager_{name}_worst_grad_relerr", worst)
    for k in g_e:
        if g_e[k].abs().max() > 0:
            assert relerr(g_g[k], g_e[k]) < 2e-2, k


def test_graphed_bf16_step_full_size_vs_golden():
    """The benchmarked path at the full model size (conv_enc-stage1, B=2) against the reference's fp64 golden step:
    loss scalars and waveform 2e-2; gradient norms of every G and D tensor -- read back from the optimisers' flat
    gradient banks after ONE graph replay with lr = 0 -- 90 % within 2e-2, 99 % within 5e-2, ALL within 1e-1 (measured on
    B200: median 2.5e-3, p90 1.1e-2, p99 2.2e-2, worst 3.6e-2 = the 8-element bias of decoder.excite_downsample.4)."""
    from tdvc.optim import FusedAdamW
    from tdvc.train_step import GraphedTrainStep, TrainStep
    g = golden("step_full_s1")
    cfg, hp = CASES["step_full"], HP_STAGE1
    G, D, _ = _models(cfg, hp)
    b, bd = _batch(cfg, hp)
    # lr 0 / wd 0: warm-up iterations and the replay leave the weights at the golden's; the banks keep the last gradients
    oG = FusedAdamW(G.parameters(), 0.0, (0.8, 0.99), weight_decay=0.0)
    oD = FusedAdamW(D.parameters(), 0.0, (0.8, 0.99), weight_decay=0.0)
    ts = TrainStep(G, D, hp, oG, oD, cfg["nspk"])
    bd.pop("neg_idx")
    draws = [d.cuda() for d in b["neg_idx"]]
    real_randint = torch.randint
    it = {"i": 0}

    def fixed_draws(low, high, size, **kw):       # the golden's contrastive negatives, also under graph capture
        if len(size) == 3:
            it["i"] += 1
            return draws[(it["i"] - 1) % 2]
        return real_randint(low, high, size, **kw)
    torch.randint = fixed_draws
    try:
        gs = GraphedTrainStep(ts, bd, warmup=2)
    finally:
        torch.randint = real_randint
    out = gs.step()
    torch.cuda.synchronize()
    rec = {}
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "g_idt", "g_cont", "g_loss"):
        ref = float(np.asarray(g[k]).reshape(-1)[0])
        rec[k] = abs(float(out[k]) - ref) / max(1.0, abs(ref))
        assert rec[k] <= 2e-2, (k, float(out[k]), ref)
    rec["fake"] = relerr(out["fake"], g["fake"])
    assert rec["fake"] < 2e-2
    errs = {}
    for opt, mod, pre in ((oD, D, "D"), (oG, G, "G")):
        names = {id(p): k for k, p in mod.named_parameters()}
        grads = {names[id(p)]: v for bank in opt._banks for p, v in zip(bank["params"], bank["views"])}
        errs.update(_norm_errs(grads, g, pre))
    e = np.sort(np.array(list(errs.values())))
    rec.update(grad_norm_p50=float(e[len(e) // 2]), grad_norm_p90=float(e[int(0.9 * len(e))]),
               grad_norm_p99=float(e[int(0.99 * len(e))]), grad_norm_max=float(e[-1]),
               grad_norm_argmax=max(errs, key=errs.get), n_tensors=len(e))
    _record("graphed_full_s1", rec)
    assert e[int(0.9 * len(e))] < 2e-2, rec
    assert e[int(0.99 * len(e))] < 5e-2, rec
    assert e[-1] < 1e-1, rec


@pytest.mark.parametrize("name,hp", [("step_tiny_s1", HP_STAGE1), ("step_tiny_s21", HP_STAGE2_1),
                                     ("step_tiny_s22", HP_STAGE2_2), ("step_tiny_latcls", HP_LATCLS),
                                     ("step_tiny_wave", HP_WAVE_CLIP)])
def test_bf16_steps_of_every_stage_config_vs_golden(name, hp):
    """stage1 / stage2_1 / stage2_2 (rec pass) / latent-classifier / wave+clip iterations in bf16 mode against the
    reference's fp64 goldens (tiny model: its 16- and 32-channel layers run on tcgen05, the 8-channel ones on the fp32
    kernels): losses and waveform 2e-2 (measured <= 2.2e-4 / 1.7e-3).  The host logic of these configurations is pinned
    exactly, in fp64, by tests/test_step_logic_cpu.py and the kernels by test_gpu_tc.py; what this adds is that the bf16
    kernels run in every configuration and give gradients of the right size.  The tiny model's gradient tensors have 8-32
    elements that are heavily cancelling sums over (batch, time), so bf16 rounding noise shows far more in their norms
    than at full size (measured: median 1.2e-2..3e-2, p90 5e-2..1.5e-1, worst 0.27..0.89 on 8-element bias gradients;
    the full-size bound -- every tensor within 1e-1 -- is asserted in test_graphed_bf16_step_full_size_vs_golden):
    median within 5e-2, 90 % within 2.5e-1, every tensor within 1.5."""
    from tdvc.train_step import TrainStep
    g = golden(name)
    cfg = CASES["step_tiny"]
    G, D, C = _models(cfg, hp)
    b, bd = _batch(cfg, hp)
    ts = TrainStep(G, D, hp, None, None, cfg["nspk"], C=C)
    out = ts.d_step(bd)
    grads = {"D": {k: p.grad.clone() for k, p in D.named_parameters()}}
    if C is not None:
        grads["C"] = {k: p.grad.clone() for k, p in C.named_parameters()}
        C.zero_grad()
    D.zero_grad(); G.zero_grad()
    out.update(ts.g_step(bd, raw_draws=b["neg_idx"]))
    grads["G"] = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in G.named_parameters()}
    rec = {}
    keys = ["d_loss_real", "d_loss_fake", "g_adv", "g_idt", "g_cont", "g_rec", "g_loss"]
    if C is not None:
        keys += ["c_loss", "g_latcls"]
    for k in keys:
        ref = float(np.asarray(g[k]).reshape(-1)[0])
        rec[k] = abs(float(out[k]) - ref) / max(1.0, abs(ref))
        assert rec[k] <= 2e-2, (k, float(out[k]), ref)
    assert relerr(out["fake"], g["fake"]) < 2e-2
    errs = {}
    for pre, gr in grads.items():
        if hp.get("grad_max_norm_" + pre) is not None:
            # eager step without a grad bank: clip_grad_norm_ ran on .grad in place, as in the golden
            pass
        errs.update(_norm_errs(gr, g, pre))
    e = np.sort(np.array(list(errs.values())))
    rec.update(grad_norm_p50=float(e[len(e) // 2]), grad_norm_p90=float(e[int(0.9 * len(e))]), grad_norm_max=float(e[-1]),
               grad_norm_argmax=max(errs, key=errs.get))
    _record(name, rec)
    assert e[len(e) // 2] < 5e-2, rec
    assert e[int(0.9 * len(e))] < 2.5e-1, rec
    assert e[-1] < 1.5, rec


def test_discriminator_feature_maps_bf16_vs_golden():
    """Every feature map of the full-size discriminator in bf16 mode (they carry the feature-matching loss, the largest
    term of g_loss): L2 norm of each of the 30 maps within 2e-2 of the reference (measured 2.8e-4), stored maps elementwise
    within 2e-2, and the data gradients that flow back to the generator -- through six LeakyReLUs whose branch flips where a
    pre-activation is below the bf16 rounding error -- within 1e-1 at the worst element (measured 4.9e-2 .. 6.5e-2) and
    5e-2 in relative L2."""
    g = golden("d_full")
    cfg = CASES["d_full"]
    D = load_det(build_D(cfg, "cmb"), cfg["seed"])
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1, frames_div=320)
    x = cu(b["signal_real"]).requires_grad_(True)
    subs = [(cu(rand_like(torch.empty(cfg["B"], 1, cfg["T"] // 4), 31)) * 0.1).requires_grad_(True),
            (cu(rand_like(torch.empty(cfg["B"], 1, cfg["T"] // 2), 32)) * 0.1).requires_grad_(True)]
    outs, feats = D(x, b["label_src"].cuda(), subs)
    worst = 0.0
    loss = sum(((o - 1) ** 2).mean() for o in outs)
    for i, fl in enumerate(feats):
        for j, f in enumerate(fl):
            ref = g[f"fstat/{i}.{j}"]
            e = abs(stats(f)[2] - ref[2]) / ref[2]
            worst = max(worst, e)
            assert e < 2e-2, (i, j, e)
            if f"feat/{i}.{j}" in g.files:
                assert relerr(f, g[f"feat/{i}.{j}"]) < 2e-2, (i, j)
            loss = loss + (f * cu(rand_like(f, 100 + 10 * i + j))).mean()
    for i, o in enumerate(outs):
        assert relerr(o, g[f"outs/{i}"]) < 2e-2
    loss.backward()
    def l2(a, b):
        b = torch.as_tensor(np.asarray(b)).double()
        return float((a.detach().double().cpu() - b).norm() / b.norm())
    rec = {"feature_norm_worst": worst, "dx": relerr(x.grad, g["dx"]),
           "dsubs": [relerr(s.grad, g[f"dsubs/{i}"]) for i, s in enumerate(subs)],
           "dx_l2": l2(x.grad, g["dx"]), "dsubs_l2": [l2(s.grad, g[f"dsubs/{i}"]) for i, s in enumerate(subs)]}
    _record("d_full_features", rec)
    assert rec["dx"] < 1e-1 and max(rec["dsubs"]) < 1e-1, rec
    assert rec["dx_l2"] < 5e-2 and max(rec["dsubs_l2"]) < 5e-2, rec


def test_inference_folded_weights_bf16():
    """generate_with_target.py:169: forward only, weight norm + bf16 operand packing done once (ops.inference_cache),
    several calls and a CUDA-graph replay give the eager result; the result is within 2e-2 of the fp64 reference."""
    from tdvc import ops
    g = golden("g_full")
    cfg = CASES["g_full"]
    G = load_det(build_G(cfg), cfg["seed"]).eval()
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1, frames_div=320)
    c_tgt = F.one_hot(b["label_tgt"], cfg["nspk"]).float().cuda()
    x, cv = cu(b["signal_real"]), cu(b["c_f0_conv"])
    with torch.no_grad():
        y_plain = G(x, c_tgt, c_var=cv)
        with ops.inference_cache():
            y1 = G(x, c_tgt, c_var=cv)
            n0 = ops._lib.load().tdvc_launch_count()
            y2 = G(x, c_tgt, c_var=cv)
            n_fold = ops._lib.load().tdvc_launch_count() - n0
        n0 = ops._lib.load().tdvc_launch_count()
        G(x, c_tgt, c_var=cv)
        n_plain = ops._lib.load().tdvc_launch_count() - n0
    assert torch.equal(y1, y_plain) and torch.equal(y2, y_plain)
    assert n_fold < n_plain, (n_fold, n_plain)          # no weight-norm / weight-pack launches once folded
    assert relerr(y_plain, g["y"]) < 2e-2
    _record("inference_launches", {"folded": int(n_fold), "per_call_weight_norm": int(n_plain)})
