"""Oracle G+D train step with gradients (TEST INFRASTRUCTURE; see tdvc_oracle.py header).

Restates the body of train.py:259-491 on top of the functional oracle, differentiating with
CPU autograd.  Used by tests/ (parity), and by bench.py's cpu_baseline / `--impl reference`
legs, which time it on the host cores.
"""
from __future__ import annotations

import numpy as np
import torch

from . import tdvc_oracle as O
from .params import make_batch, make_state_dict


def generator_shapes(cfg) -> dict:
    """state_dict key -> shape of Generator(...) for a conv-encoder config; derived analytically from
    model/generator.py:197-347,409-475 (pinned against the reference's own keys by the golden tests)."""
    ratios, ch = list(cfg["ratios"]), list(cfg["channels"])
    cd, cond, nspk = cfg["content_dim"], cfg["cond_dim"], cfg["nspk"]
    S = {}

    def wn(prefix, co, ci, k, bias=True, transpose=False):
        if bias:
            S[prefix + ".bias"] = (co,)
        S[prefix + ".weight_g"] = ((ci if transpose else co), 1, 1)
        S[prefix + ".weight_v"] = (ci, co, k) if transpose else (co, ci, k)

    def mrf(prefix, c, ncond):
        for i, k in enumerate(O.MRF_KERNELS):
            for j in range(len(O.MRF_DILATIONS)):
                p = f"{prefix}.blocks.{i}.{j}"
                wn(p + ".conv.1", c, c, k)
                wn(p + ".posconv.1", c, c, 1)
                if ncond:
                    wn(p + ".cond_var.0", ncond, ncond, 3)
                    wn(p + ".cond_var.2", 2 * c, ncond, 3)

    # decoder first (registered first, model/generator.py:452)
    p = "decoder.decoder"
    wn(f"{p}.1", ch[0], cd, 7, bias=False)
    wn(f"{p}.3", ch[0], ch[0], 7)
    idx = 4
    for i, r in enumerate(ratios):
        wn(f"{p}.{idx + 2}", ch[i + 1], ch[i], 2 * r, transpose=True)
        mrf(f"{p}.{idx + 3}", ch[i + 1], cond + 8)
        idx += 4
    wn(f"{p}.{idx + 2}", 1, ch[-1], 7)
    for i in (1, 2):
        if i < len(ratios):
            wn(f"decoder.subsample_out_layers.{i}.1", 1, ch[i + 1], 7)
    for i, r in enumerate(ratios):
        q = f"decoder.excite_downsample.{i}"
        wn(q + ".block.0", 8, 8, 2 * r)
        wn(q + ".block.2", 8, 8, 5)
        wn(q + ".block.4", 8, 8, 5)
        S[q + ".shortcut.weight"] = (8, 8, 1)
        S[q + ".shortcut.bias"] = (8,)
    wn(f"decoder.excite_downsample.{len(ratios)}", 8, 1, 7)
    # encoder
    er, ec = ratios[::-1], ch[::-1]
    p = "encoder.encoder"
    wn(f"{p}.0", ec[0], 1, 7)
    for i, r in enumerate(er):
        wn(f"{p}.{3 + 4 * i}", ec[i + 1], ec[i], 2 * r)
        mrf(f"{p}.{4 + 4 * i}", ec[i + 1], 0)
    n = 1 + 4 * len(er)
    wn(f"{p}.{n + 1}", ec[-1], ec[-1], 7)
    wn(f"{p}.{n + 3}", cd, ec[-1], 7, bias=False)
    S["embedding.weight"] = (cond, nspk)
    S["embedding.bias"] = (cond,)
    return S


def discriminator_shapes(cfg) -> dict:
    """Keys/shapes of CollaborativeMultibandDiscriminator(...), model/discriminator.py:7-38,77-92."""
    S = {}
    for d in range(cfg["num_disc"]):
        p = f"discriminators.{d}"
        nf = cfg["d_base"]
        S[f"{p}.discriminator.0.0.bias"] = (nf,)
        S[f"{p}.discriminator.0.0.weight_g"] = (nf, 1, 1)
        S[f"{p}.discriminator.0.0.weight_v"] = (nf, 1, 15)
        for i in range(cfg["d_layers"]):
            prev, nf = nf, min(nf * 4, 1024)
            S[f"{p}.discriminator.{i + 1}.0.bias"] = (nf,)
            S[f"{p}.discriminator.{i + 1}.0.weight_g"] = (nf, 1, 1)
            S[f"{p}.discriminator.{i + 1}.0.weight_v"] = (nf, prev // (prev // 4), 41)
        L = cfg["d_layers"] + 1
        S[f"{p}.discriminator.{L}.0.bias"] = (nf,)
        S[f"{p}.discriminator.{L}.0.weight_g"] = (nf, 1, 1)
        S[f"{p}.discriminator.{L}.0.weight_v"] = (nf, nf, 5)
        S[f"{p}.output.weight_g"] = (cfg["nspk"], 1, 1)
        S[f"{p}.output.weight_v"] = (cfg["nspk"], nf, 3)
    return S


def latent_classifier_shapes(cfg) -> dict:
    """Keys/shapes of LatentClassifier(nspk, content_dim), model/latent_classifier.py:8-32."""
    S = {}
    nf = cfg["content_dim"]
    idx = 1
    for _ in range(3):
        prev, nf = nf, nf * 2
        S[f"classifier.{idx}.bias"] = (nf,)
        S[f"classifier.{idx}.weight_g"] = (nf, 1, 1)
        S[f"classifier.{idx}.weight_v"] = (nf, prev, 21)
        idx += 2
    S[f"classifier.{idx}.bias"] = (nf,)
    S[f"classifier.{idx}.weight_g"] = (nf, 1, 1)
    S[f"classifier.{idx}.weight_v"] = (nf, nf, 5)
    S[f"classifier.{idx + 2}.weight_g"] = (cfg["nspk"], 1, 1)
    S[f"classifier.{idx + 2}.weight_v"] = (cfg["nspk"], nf, 3)
    return S


def make_models(cfg, dtype=torch.float64):
    sdG = make_state_dict(generator_shapes(cfg), seed=cfg["seed"], dtype=dtype)
    sdD = make_state_dict(discriminator_shapes(cfg), seed=cfg["seed"] + 100, dtype=dtype)
    return sdG, sdD


def step_batch(cfg, hp, dtype=torch.float64):
    return make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1,
                      frames_div=int(np.prod(cfg["ratios"])), dtype=dtype, permute=not hp["no_conv"])


def oracle_step(cfg, hp, dtype=torch.float64, sdG=None, sdD=None, batch=None, sdC=None) -> dict:
    """Runs the D-step and G-step losses and both backward passes.  Returns loss scalars,
    'fake' and {'D_grad','G_grad'[,'C_grad']}: name -> grad tensor."""
    if sdG is None:
        sdG, sdD = make_models(cfg, dtype)
    if sdC is None and hp.get("lambda_latcls", 0) != 0:
        sdC = make_state_dict(latent_classifier_shapes(cfg), seed=cfg["seed"] + 200, dtype=dtype)
    if batch is None:
        batch = step_batch(cfg, hp, dtype)
    for v in list(sdG.values()) + list(sdD.values()) + list((sdC or {}).values()):
        v.requires_grad_(True)
        v.grad = None
    kw = dict(num_disc=cfg["num_disc"], num_layers=cfg["d_layers"])
    out = {}
    d = _d_step(sdG, sdD, batch, hp, cfg, kw)
    d["d_loss"].backward()
    if hp.get("grad_max_norm_D") is not None:      # train.py:288-289
        torch.nn.utils.clip_grad_norm_(list(sdD.values()), hp["grad_max_norm_D"])
    out.update({k: v.detach() for k, v in d.items()})
    out["D_grad"] = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in sdD.items()}
    if sdC is not None:      # latent classifier step, train.py:300-309
        c_loss = torch.nn.functional.cross_entropy(O.latent_classifier(sdC, d["emb"].detach()), batch["label_src"])
        c_loss.backward()
        out["c_loss"] = c_loss.detach()
        out["C_grad"] = {k: v.grad.clone() for k, v in sdC.items()}
    for v in list(sdG.values()) + list(sdD.values()) + list((sdC or {}).values()):
        v.grad = None
    g = _g_step(sdG, sdD, batch, hp, cfg, kw, sdC)
    g["g_loss"].backward()
    if hp.get("grad_max_norm_G") is not None:      # train.py:488-489
        torch.nn.utils.clip_grad_norm_(list(sdG.values()), hp["grad_max_norm_G"])
    out.update({k: v.detach() for k, v in g.items()})
    out["G_grad"] = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in sdG.items()}
    return out


def _d_step(sdG, sdD, b, hp, cfg, kw):
    """train.py:259-291."""
    x = b["signal_real"]
    nspk = cfg["nspk"]
    c_tgt = O.one_hot(b["label_tgt"], nspk, x.dtype)
    fake, fake_subs, emb = O.generator(sdG, x, c_tgt, b["c_f0_conv"], cfg["ratios"])
    o_real, _ = O.cmb_discriminator(sdD, x, b["label_src"], O.cmb_subsamples(x, cfg["num_disc"]), **kw)
    # reference passes the heads un-detached (train.py:269); the G grads that deposits are zeroed before
    # G's own backward (train.py:485-486) so detaching changes no result, only skips dead work.
    subs_in = fake_subs if hp.get("faithful_waste") else [s.detach() for s in fake_subs]
    o_fake, _ = O.cmb_discriminator(sdD, fake.detach(), b["label_tgt"], subs_in, **kw)
    d_real, d_fake = O.lsgan_d_loss(o_real, o_fake)
    return {"d_loss_real": d_real, "d_loss_fake": d_fake, "d_loss": d_real + d_fake, "fake": fake, "emb": emb}


def _g_step(sdG, sdD, b, hp, cfg, kw, sdC=None):
    """train.py:320-480 (lambda_f0 = 0)."""
    x = b["signal_real"]
    nspk, ratios = cfg["nspk"], cfg["ratios"]
    lab_s, lab_t = b["label_src"], b["label_tgt"]
    c_src = O.one_hot(lab_s, nspk, x.dtype)
    c_tgt = O.one_hot(lab_t, nspk, x.dtype)
    out = {}
    fake, fake_subs, emb_real = O.generator(sdG, x, c_tgt, b["c_f0_conv"], ratios)
    o_fake, _ = O.cmb_discriminator(sdD, fake, lab_t, fake_subs, **kw)
    g_adv = O.lsgan_g_loss(o_fake)
    f_real = None
    if (hp["lambda_rec"] > 0 or hp["lambda_idt"] > 0) and hp["lambda_feat"] > 0:
        _, f_real = O.cmb_discriminator(sdD, x, lab_s, O.cmb_subsamples(x, cfg["num_disc"]), **kw)
    g_rec = x.new_zeros(())
    if (not hp["no_conv"]) and hp["lambda_rec"] > 0:
        rec, rec_subs, _ = O.generator(sdG, fake.detach(), c_src, b["c_f0_src"], ratios)
        if hp["lambda_feat"] > 0:
            _, f_rec = O.cmb_discriminator(sdD, rec, lab_s, rec_subs, **kw)
            g_rec = g_rec + hp["lambda_feat"] * O.feat_loss(f_rec, f_real)
        if hp["lambda_spec"] > 0:
            g_rec = g_rec + hp["lambda_spec"] * O.mel_loss(rec, x)
        if hp.get("lambda_wave", 0) > 0:      # train.py:358-361
            g_rec = g_rec + hp["lambda_wave"] * torch.mean(torch.abs(x - rec))
    g_idt = x.new_zeros(())
    if hp["lambda_idt"] > 0:
        if not hp["no_conv"]:
            idt, idt_subs, _ = O.generator(sdG, x, c_src, b["c_f0_src"], ratios)
        else:
            idt, idt_subs = fake, fake_subs
        if hp["lambda_feat"] > 0:
            _, f_idt = O.cmb_discriminator(sdD, idt, lab_s, idt_subs, **kw)
            g_idt = g_idt + hp["lambda_feat"] * O.feat_loss(f_idt, f_real)
        if hp["lambda_spec"] > 0:
            g_idt = g_idt + hp["lambda_spec"] * O.mel_loss(idt, x)
        if hp.get("lambda_wave", 0) > 0:      # train.py:382-385: added to g_loss_REC by the reference
            g_rec = g_rec + hp["lambda_wave"] * torch.mean(torch.abs(x - idt))
    g_cont = x.new_zeros(())
    if hp["lambda_cont_emb"] > 0 and hp["lambda_corrupted"]:
        emb_corr = O.encoder(sdG, "encoder", b["signal_corrupted"], list(ratios)[::-1])
        g_cont = g_cont + O.contrastive_loss(emb_real, emb_corr, b["neg_idx"][0], b["neg_idx"][1])
    out.update(g_adv=g_adv, g_rec=g_rec, g_idt=g_idt, g_cont=g_cont)
    out["g_loss"] = (g_adv + hp["lambda_rec"] * g_rec + hp["lambda_idt"] * g_idt
                     + hp["lambda_cont_emb"] * g_cont)
    if sdC is not None and hp.get("lambda_latcls", 0) != 0:      # train.py:420-425, through gradient reversal
        out["g_latcls"] = torch.nn.functional.cross_entropy(O.latent_classifier(sdC, emb_real), lab_s)
        out["g_loss"] = out["g_loss"] + hp["lambda_latcls"] * out["g_latcls"]
    return out
