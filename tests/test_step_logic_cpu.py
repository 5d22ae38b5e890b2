"""Host logic of tdvc.train_step.TrainStep on a machine without a GPU: every tdvc op is swapped for its plain torch
definition (tests/cpu_shim.py), the modules and the step run in fp64 on the CPU, and the result is held against the
golden vectors the real reference produced.  What this pins is the part the restructured step adds on top of the
reference's loop body: ONE generator forward shared by the D and the G step, the encoder pass shared by the fake /
identity / corrupted inputs, decoder and discriminator passes stacked along the batch, the row ranges each loss reads,
and the wave-L1 / gradient-clipping / jitter options."""
import numpy as np
import pytest
import torch

import cpu_shim
from helpers import golden, relerr, stats
from oracle.cases import CASES, HP_LATCLS, HP_STAGE1, HP_STAGE2_1, HP_STAGE2_2, HP_WAVE_CLIP
from oracle.params import make_batch, make_state_dict
from test_host_cpu import build_D, build_G


def _load(mod, seed):
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    mod.load_state_dict(make_state_dict(shapes, seed=seed, dtype=torch.float64), strict=True)
    return mod.double()


def _run(cfg, hp, batched=True):
    from tdvc.train_step import TrainStep
    G = _load(build_G(cfg), cfg["seed"])
    D = _load(build_D(cfg), cfg["seed"] + 100)
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1, frames_div=int(np.prod(cfg["ratios"])),
                   permute=not hp["no_conv"])
    bd = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in b.items()}
    C = None
    if hp["lambda_latcls"] != 0:
        from model.latent_classifier import LatentClassifier
        C = _load(LatentClassifier(cfg["nspk"], cfg["content_dim"]), cfg["seed"] + 200)
    ts = TrainStep(G, D, hp, None, None, cfg["nspk"], C=C)
    ts.batched = ts.batched and batched
    out = ts.d_step(bd)
    out["D_grad"] = {k: p.grad.clone() for k, p in D.named_parameters()}
    if C is not None:
        out["C_grad"] = {k: p.grad.clone() for k, p in C.named_parameters()}
        C.zero_grad()
    D.zero_grad(); G.zero_grad()
    out.update(ts.g_step(bd, raw_draws=b["neg_idx"]))
    out["G_grad"] = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in G.named_parameters()}
    return out


@pytest.mark.parametrize("name,hp", [("step_tiny_s1", HP_STAGE1), ("step_tiny_s21", HP_STAGE2_1),
                                     ("step_tiny_s22", HP_STAGE2_2), ("step_tiny_latcls", HP_LATCLS),
                                     ("step_tiny_wave", HP_WAVE_CLIP)])
@pytest.mark.parametrize("batched", [True, False])
def test_step_logic_vs_reference_golden(name, hp, batched):
    g = golden(name)
    with cpu_shim.installed():
        out = _run(CASES["step_tiny"], hp, batched)
    if hp["lambda_latcls"] != 0:
        for k in ("c_loss", "g_latcls"):
            assert abs(float(out[k]) - float(g[k])) <= 1e-9 * max(1.0, abs(float(g[k]))), k
    for k in ("d_loss_real", "d_loss_fake", "g_adv", "g_idt", "g_cont", "g_rec", "g_loss"):
        ref = float(np.asarray(g[k]).reshape(-1)[0])
        # 1e-7: the mel filterbank / Hann window are fp32 constants on both sides (torchaudio builds them in fp32), built
        # by two different formula orders; everything else agrees to 1e-12
        assert abs(float(out[k]) - ref) <= 1e-7 * max(1.0, abs(ref)), (k, float(out[k]), ref)
    assert relerr(out["fake"], g["fake"]) < 1e-6          # the golden waveform is stored in fp32
    # fp64 arithmetic on both sides, but the Kaiser filters / mel filterbank / Hann window are fp32 constants built by
    # different formula orders (measured: 5e-7 without the mel term, 3e-6 with it, elementwise): norms to 2e-5
    for which in ("D", "G") + (("C",) if hp["lambda_latcls"] != 0 else ()):
        for k, gr in out[which + "_grad"].items():
            ref = g[f"{which}_grad/{k}"]
            got = stats(gr)
            assert abs(got[2] - ref[2]) <= 2e-5 * max(ref[2], 1e-12) + 1e-12, (which, k, got, ref)
            assert abs(got[0] - ref[0]) <= 2e-5 * max(ref[1], 1e-12) + 1e-12, (which, k, got, ref)


@pytest.mark.parametrize("hp", [HP_STAGE1, HP_STAGE2_2, HP_LATCLS, HP_WAVE_CLIP], ids=["s1", "s22", "latcls", "wave_clip"])
def test_stacked_passes_equal_separate_passes(hp):
    """The strong form of the check above: with the same constants on both sides, the step that shares the encoder
    pass and stacks decoder / discriminator passes along the batch reproduces the step that calls G and D once per
    signal, as the reference does -- every loss and every gradient element to 1e-10."""
    with cpu_shim.installed():
        a = _run(CASES["step_tiny"], hp, batched=True)
        b = _run(CASES["step_tiny"], hp, batched=False)
    for k in ("d_loss", "g_adv", "g_idt", "g_cont", "g_rec", "g_loss"):
        assert abs(float(a[k]) - float(b[k])) <= 1e-12 * max(1.0, abs(float(b[k]))), k
    for which in ("D", "G") + (("C",) if hp["lambda_latcls"] != 0 else ()):
        for k, gr in a[which + "_grad"].items():
            assert relerr(gr, b[which + "_grad"][k]) < 1e-10, (which, k)


def test_unsupported_hp_is_refused():
    from tdvc.train_step import TrainStep
    cfg = CASES["step_tiny"]
    G, D = build_G(cfg), build_D(cfg)
    with pytest.raises(NotImplementedError):
        TrainStep(G, D, dict(HP_STAGE1, lambda_f0=1000), None, None, cfg["nspk"])
    with pytest.raises(ValueError):
        TrainStep(G, D, dict(HP_STAGE1, lambda_latcls=1), None, None, cfg["nspk"])


def test_jitter_rolls_the_reference_signal():
    """jitter_amp > 0 (train.py:334-337, util/audio.py:27-30): the feature / mel reference is the per-sample rolled
    signal; with the draw forced to zero shift the losses equal the jitter-free step."""
    cfg, hp = CASES["step_tiny"], dict(HP_STAGE1, jitter_amp=3)
    g = golden("step_tiny_s1")
    real_randint = torch.randint

    def zero_shift(low, high, size, **kw):
        if low < 0 and tuple(size) == (cfg["B"],):       # the jitter draw (the only one with a negative lower bound)
            return torch.zeros(size, dtype=torch.long)
        return real_randint(low, high, size, **kw)
    torch.randint = zero_shift
    try:
        with cpu_shim.installed():
            out = _run(cfg, hp)
    finally:
        torch.randint = real_randint
    assert abs(float(out["g_loss"]) - float(np.asarray(g["g_loss"]).reshape(-1)[0])) <= 1e-7 * abs(float(out["g_loss"]))
