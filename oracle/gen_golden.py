"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

TEST INFRASTRUCTURE.  Usage (from the repo root):  python oracle/gen_golden.py
Imports /root/reference's own `model` / `util` packages (fp64, CPU), loads the deterministic
parameters of oracle/params.py into them by state_dict, runs forward/backward and stores the
outputs.  /root/reference does not exist on the GPU box, so only the vectors (and this script)
are committed.  Nothing here is imported by the product.
"""
import os
import sys
import warnings

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TDVC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)          # reference's `model`, `util`, `wavlm`
sys.path.insert(1, REPO)         # `oracle.params`
warnings.filterwarnings("ignore")

import numpy as np
import torch
import torch.nn.functional as F

from model.generator import Generator, CINResnetBlock                      # noqa: E402  (reference)
from model.discriminator import (CollaborativeMultibandDiscriminator,       # noqa: E402  (reference)
                                 MultiscaleDiscriminator)
from model.conditional_instance_norm import ConditionalInstanceNorm        # noqa: E402  (reference)
import util.losses as ref_losses                                            # noqa: E402  (reference)
from oracle.params import make_state_dict, make_batch                       # noqa: E402
from oracle.cases import CASES, HP_LATCLS, HP_STAGE1, HP_STAGE2_1, HP_STAGE2_2, HP_WAVE_CLIP, rand_like  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(os.cpu_count())


def shapes_of(mod):
    return {k: tuple(v.shape) for k, v in mod.state_dict().items()}


def load(mod, seed):
    sd = make_state_dict(shapes_of(mod), seed=seed, dtype=torch.float64)
    mod.double()
    mod.load_state_dict(sd, strict=True)
    return sd


def stats(t):
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()], dtype=np.float64)


def grads_of(mod, full_limit):
    """name -> full grad (small tensors) and name -> [sum, abssum, l2] for every parameter."""
    full, st = {}, {}
    for k, p in mod.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        st[k] = stats(g)
        if g.numel() <= full_limit:
            full[k] = g.detach().double().numpy()
    return full, st


def build_G(cfg):
    return Generator(list(cfg["ratios"]), list(cfg["channels"]), 0, cfg["nspk"], cfg["cond_dim"], cfg["content_dim"],
                     3, 0, "conv", norm_layer=(None, None, None),
                     weight_norm=("weight_norm",) * 3, bot_cond="target", enc_cond=None, dec_cond="target",
                     output_content_emb=True)


def build_D(cfg, cls=CollaborativeMultibandDiscriminator):
    return cls(cfg["num_disc"], cfg["nspk"], cfg["d_layers"], cfg["d_base"], 4, 4, 128, "target")


def save(name, **arrs):
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}/{kk}"] = np.asarray(vv)
        elif isinstance(v, (list, tuple)):
            for i, vv in enumerate(v):
                flat[f"{k}/{i}"] = vv.detach().double().numpy() if torch.is_tensor(vv) else np.asarray(vv)
        else:
            flat[k] = v.detach().double().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"{name}: {len(flat)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


def case_generator(name, cfg, full_limit):
    G = build_G(cfg)
    load(G, cfg["seed"])
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1, frames_div=int(np.prod(cfg["ratios"])))
    c_tgt = F.one_hot(b["label_tgt"], cfg["nspk"]).double()
    y, subs = G(b["signal_real"], c_tgt, c_var=b["c_f0_conv"], out_subsample=True)
    emb = G.content_embedding
    loss = (y * rand_like(y, 11)).sum() + sum((s * rand_like(s, 12 + i)).sum() for i, s in enumerate(subs)) \
        + (emb * rand_like(emb, 20)).sum()
    loss.backward()
    full, st = grads_of(G, full_limit)
    save(name, y=y, subs=subs, emb=emb, loss=loss, grad=full, gstat=st,
         keys=np.array(list(G.state_dict().keys())),
         shapes=np.array([str(tuple(v.shape)) for v in G.state_dict().values()]))


def case_discriminator(name, cfg, full_limit, cls=CollaborativeMultibandDiscriminator):
    D = build_D(cfg, cls)
    load(D, cfg["seed"])
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1, frames_div=320)
    x = b["signal_real"].clone().requires_grad_(True)
    if cls is CollaborativeMultibandDiscriminator:
        subs = [(rand_like(torch.empty(cfg["B"], 1, cfg["T"] // 4), 31) * 0.1).requires_grad_(True),
                (rand_like(torch.empty(cfg["B"], 1, cfg["T"] // 2), 32) * 0.1).requires_grad_(True)]
        outs, feats = D(x, b["label_src"], subs)
    else:
        subs = []
        outs, feats = D(x, b["label_src"])
    loss = sum(((o - 1) ** 2).mean() for o in outs)
    for i, fl in enumerate(feats):
        for j, f in enumerate(fl):
            loss = loss + (f * rand_like(f, 100 + 10 * i + j)).mean()
    loss.backward()
    full, st = grads_of(D, full_limit)
    fstat = {f"{i}.{j}": stats(f) for i, fl in enumerate(feats) for j, f in enumerate(fl)}
    feat_small = {f"{i}.{j}": f.detach().numpy() for i, fl in enumerate(feats) for j, f in enumerate(fl)
                  if f.numel() <= full_limit}
    save(name, outs=outs, fstat=fstat, feat=feat_small, loss=loss, grad=full, gstat=st, dx=x.grad,
         dsubs=[s.grad for s in subs],
         keys=np.array(list(D.state_dict().keys())),
         shapes=np.array([str(tuple(v.shape)) for v in D.state_dict().values()]))


def case_cin(name):
    torch.manual_seed(0)
    C, ncond, B, T = 12, 7, 3, 50
    m = ConditionalInstanceNorm(C, ncond)
    load(m, 5)
    x = (rand_like(torch.empty(B, C, T), 41) * 2 + 0.3).requires_grad_(True)
    c2 = rand_like(torch.empty(B, ncond), 42).requires_grad_(True)
    c3 = rand_like(torch.empty(B, ncond + 1, T), 43).requires_grad_(True)
    y2 = m(x, c2)
    (y2 * rand_like(y2, 44)).sum().backward()
    g2 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    dx2, dc2 = x.grad.clone(), c2.grad.clone()
    m.zero_grad(); x.grad = None
    y3 = m(x, c3)
    (y3 * rand_like(y3, 45)).sum().backward()
    g3 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    blk = CINResnetBlock(C, ncond, dilation=3, kernel_size=7)
    load(blk, 6)
    xb = (rand_like(torch.empty(B, C, T), 46)).requires_grad_(True)
    yb = blk(xb, c2.detach())
    (yb * rand_like(yb, 47)).sum().backward()
    gb = {k: p.grad.clone() for k, p in blk.named_parameters() if p.grad is not None}
    save(name, y2=y2, dx2=dx2, dc2=dc2, g2={k: v.numpy() for k, v in g2.items()},
         y3=y3, dx3=x.grad, dc3=c3.grad, g3={k: v.numpy() for k, v in g3.items()},
         yb=yb, dxb=xb.grad, gb={k: v.numpy() for k, v in gb.items()},
         cin_keys=np.array(list(m.state_dict().keys())), blk_keys=np.array(list(blk.state_dict().keys())))


def case_losses(name):
    B, T = 3, 8960
    a = rand_like(torch.empty(B, 1, T), 51) * 0.1
    r = rand_like(torch.empty(B, 1, T), 52) * 0.1
    a.requires_grad_(True)
    import functools
    orig = ref_losses.get_melspec_transform
    ref_losses.get_melspec_transform = lambda *args: orig(*args).double()
    mel = ref_losses.multiscale_spec_loss(a, r, [2048, 1024, 512])
    mel.backward()
    ref_losses.get_melspec_transform = orig
    X = F.normalize(rand_like(torch.empty(B, 16, 28), 53), dim=1).requires_grad_(True)
    Y = F.normalize(rand_like(torch.empty(B, 16, 28), 54), dim=1).requires_grad_(True)
    g = torch.Generator(); g.manual_seed(77)
    raws = [torch.randint(0, 27, (B, 28, 100), generator=g) for _ in range(2)]
    it = iter([r_.clone() for r_ in raws])
    real_randint = torch.randint
    torch.randint = lambda *a_, **k_: next(it)
    try:
        con = ref_losses.contrastive_loss(X, Y, num_negatives=100, temp=0.1)
    finally:
        torch.randint = real_randint
    con.backward()
    fs = [[rand_like(torch.empty(2, 4, 30), 60 + i * 3 + j) for j in range(3)] for i in range(2)]
    fr = [[rand_like(torch.empty(2, 4, 30), 80 + i * 3 + j) for j in range(3)] for i in range(2)]
    fl = ref_losses.multiscale_feat_loss(fs, fr)
    save(name, mel=mel, dmel=a.grad, con=con, dX=X.grad, dY=Y.grad, raw0=raws[0], raw1=raws[1], feat=fl)


def case_legacy_blocks(name):
    """DecoderResnetBlock / TranformResnetBlock / ResnetBlock (model/generator.py:11-67): constructible legacy blocks that
    no shipped config instantiates; pinned so the drop-in constructors stay honest."""
    from model.generator import DecoderResnetBlock, TranformResnetBlock, ResnetBlock
    import torch.nn as nn
    C, B, T = 10, 2, 64
    out = {}
    for i, (tag, mk) in enumerate((("dec", lambda: DecoderResnetBlock(C, dilation=3)),
                                   ("trf", lambda: TranformResnetBlock(C, dilation=1)),
                                   ("res", lambda: ResnetBlock(C, dilation=3, weight_norm=nn.utils.weight_norm)))):
        m = mk()
        load(m, 20 + i)
        x = rand_like(torch.empty(B, C, T), 81).requires_grad_(True)
        y = m(x)
        (y * rand_like(y, 82)).sum().backward()
        out[tag + "_y"] = y
        out[tag + "_dx"] = x.grad
        out[tag + "_keys"] = np.array(list(m.state_dict().keys()))
        out[tag + "_shapes"] = np.array([str(tuple(v.shape)) for v in m.state_dict().values()])
        for k, p_ in m.named_parameters():
            out[f"{tag}_grad/{k}"] = p_.grad.detach().double().numpy()
    save(name, **out)


def case_latent_classifier(name):
    """LatentClassifier + gradient reversal (model/latent_classifier.py:8-38, model/grad_rev.py:3-18)."""
    from model.latent_classifier import LatentClassifier
    nspk, cdim, B, T = 6, 16, 3, 28
    m = LatentClassifier(nspk, cdim)
    load(m, 9)
    x = rand_like(torch.empty(B, cdim, T), 71).requires_grad_(True)
    lab = torch.tensor([1, 4, 0])
    out = m(x)
    loss = F.cross_entropy(out, lab)
    loss.backward()
    full, st = grads_of(m, 20000)
    save(name, out=out, loss=loss, dx=x.grad, grad=full, gstat=st,
         keys=np.array(list(m.state_dict().keys())),
         shapes=np.array([str(tuple(v.shape)) for v in m.state_dict().values()]))


def case_ssl_wn(name):
    """The WaveNet-style stack behind WavLM (model/ssl_encoder.py:16-116: WN + Encoder): m, logs and every gradient for a
    small geometry (64 SSL dims -> 32 hidden / 32 out, 4 layers k5) and a conditioned WN (gin_channels, dilation rate 2)."""
    from model.ssl_encoder import Encoder as SSLWNEncoder, WN
    B, T = 3, 28
    m = SSLWNEncoder(64, 32, 32, 5, 1, 4)
    load(m, 31)
    x = rand_like(torch.empty(B, 64, T), 91).requires_grad_(True)
    torch.manual_seed(5)
    z, mean, logs, _ = m(x)
    loss = (mean * rand_like(mean, 92)).sum() + (logs * rand_like(logs, 93)).sum()
    loss.backward()
    full, st = grads_of(m, 1 << 30)
    w = WN(16, 3, 2, 3, gin_channels=8)
    load(w, 32)
    xw = rand_like(torch.empty(2, 16, 40), 94).requires_grad_(True)
    gw = rand_like(torch.empty(2, 8, 40), 95).requires_grad_(True)
    yw = w(xw, 1, g=gw)
    (yw * rand_like(yw, 96)).sum().backward()
    wfull, wst = grads_of(w, 1 << 30)
    save(name, m=mean, logs=logs, dx=x.grad, grad=full, gstat=st,
         keys=np.array(list(m.state_dict().keys())),
         shapes=np.array([str(tuple(v.shape)) for v in m.state_dict().values()]),
         wn_y=yw, wn_dx=xw.grad, wn_dg=gw.grad, wn_grad=wfull,
         wn_keys=np.array(list(w.state_dict().keys())))


def case_yin(name):
    """util/yin.py `estimate` as train.py:238 would call it (pitch 50..550 Hz, one frame per 64 samples) on four 0.56 s
    signals: sines with noise (one with a silent stretch, one with a glide), a two-tone mixture, and white noise."""
    import util.yin as ref_yin
    g = torch.Generator().manual_seed(77)
    T, sr = 8960, 16000
    t = torch.arange(T, dtype=torch.float32) / sr
    x = torch.zeros(4, T)
    x[0] = 0.05 * torch.sin(2 * np.pi * 140.0 * t) + 0.01 * torch.randn(T, generator=g)
    x[0, 3000:4500] = 0.002 * torch.randn(1500, generator=g)
    x[1] = 0.05 * torch.sin(2 * np.pi * (110.0 * t + 150.0 * t * t)) + 0.005 * torch.randn(T, generator=g)
    x[2] = 0.04 * torch.sin(2 * np.pi * 220.0 * t) + 0.02 * torch.sin(2 * np.pi * 331.0 * t + 0.3) + 0.004 * torch.randn(T, generator=g)
    x[3] = 0.02 * torch.randn(T, generator=g)
    f0 = ref_yin.estimate(x, sr, pitch_min=50, pitch_max=550, frame_stride=64 / sr)
    f0_soft = ref_yin.estimate(x, sr, pitch_min=50, pitch_max=550, frame_stride=64 / sr, soft=True)
    short = ref_yin.estimate(x[:2, :500], sr, pitch_min=50, pitch_max=550, frame_stride=64 / sr)
    save(name, x=x, f0=f0, f0_soft=f0_soft, f0_short=short)


def ref_step(G, D, b, hp, nspk, C=None):
    """train.py:259-491 driven through the reference's own modules (lambda_f0 term = 0)."""
    x = b["signal_real"]
    c_src = F.one_hot(b["label_src"], nspk).double()
    c_tgt = F.one_hot(b["label_tgt"], nspk).double()
    lab_s, lab_t = b["label_src"], b["label_tgt"]
    out = {}
    # D step
    fake, fake_subs = G(x, c_tgt, c_var=b["c_f0_conv"], out_subsample=True)
    o_real, _ = D(x, lab_s, D.get_subsamples(x))
    o_fake, _ = D(fake.detach(), lab_t, fake_subs)
    d_real = sum(F.mse_loss(o, torch.ones_like(o)) for o in o_real)
    d_fake = sum(F.mse_loss(o, torch.zeros_like(o)) for o in o_fake)
    D.zero_grad()
    (d_real + d_fake).backward()
    if hp.get("grad_max_norm_D") is not None:      # train.py:288-289
        torch.nn.utils.clip_grad_norm_(D.parameters(), hp["grad_max_norm_D"])
    out["d_loss_real"], out["d_loss_fake"] = d_real.detach(), d_fake.detach()
    out["D_grad"] = grads_of(D, 0)[1]
    if C is not None:      # latent classifier step, train.py:300-309
        emb = G.content_embedding.clone()
        c_loss = F.cross_entropy(C(emb.detach()), lab_s)
        C.zero_grad()
        c_loss.backward()
        out["c_loss"] = c_loss.detach()
        out["C_grad"] = grads_of(C, 0)[1]
    # G step (D weights NOT updated in between, so the fixture is optimiser-independent)
    fake, fake_subs = G(x, c_tgt, c_var=b["c_f0_conv"], out_subsample=True)
    emb_real = G.content_embedding.clone()
    o_fake, f_fake = D(fake, lab_t, fake_subs)
    g_adv = sum(F.mse_loss(o, torch.ones_like(o)) for o in o_fake)
    _, f_real = D(x, lab_s, D.get_subsamples(x))
    g_rec = torch.zeros(())
    orig = ref_losses.get_melspec_transform
    ref_losses.get_melspec_transform = lambda *args: orig(*args).double()
    if (not hp["no_conv"]) and hp["lambda_rec"] > 0:
        rec, rec_subs = G(fake.detach(), c_src, c_var=b["c_f0_src"], out_subsample=True)
        _, f_rec = D(rec, lab_s, rec_subs)
        g_rec = hp["lambda_feat"] * ref_losses.multiscale_feat_loss(f_rec, f_real) + \
            hp["lambda_spec"] * ref_losses.multiscale_spec_loss(rec, x, [2048, 1024, 512])
        if hp["lambda_wave"] > 0:      # train.py:358-361
            g_rec = g_rec + hp["lambda_wave"] * torch.mean(torch.abs(x - rec))
    if not hp["no_conv"]:
        idt, idt_subs = G(x, c_src, c_var=b["c_f0_src"], out_subsample=True)
    else:
        idt, idt_subs = fake, fake_subs
    _, f_idt = D(idt, lab_s, idt_subs)
    idt_feat = ref_losses.multiscale_feat_loss(f_idt, f_real)
    idt_spec = ref_losses.multiscale_spec_loss(idt, x, [2048, 1024, 512])
    ref_losses.get_melspec_transform = orig
    g_idt = hp["lambda_feat"] * idt_feat + hp["lambda_spec"] * idt_spec
    if hp["lambda_wave"] > 0:          # train.py:382-385: the identity pass' wave term is added to g_loss_REC
        g_rec = g_rec + hp["lambda_wave"] * torch.mean(torch.abs(x - idt))
    emb_corr = G.encoder(b["signal_corrupted"])
    it = iter([r.clone() for r in b["neg_idx"]])
    real_randint = torch.randint
    torch.randint = lambda *a_, **k_: next(it)
    try:
        g_cont = ref_losses.contrastive_loss(emb_real, emb_corr, num_negatives=100, temp=0.1)
    finally:
        torch.randint = real_randint
    g_loss = g_adv + hp["lambda_rec"] * g_rec + hp["lambda_idt"] * g_idt + hp["lambda_cont_emb"] * g_cont
    if C is not None and hp["lambda_latcls"] != 0:      # train.py:420-425 (through the gradient-reversal layer)
        g_lat = F.cross_entropy(C(emb_real), lab_s)
        out["g_latcls"] = g_lat.detach()
        g_loss = g_loss + hp["lambda_latcls"] * g_lat
    D.zero_grad(); G.zero_grad()
    g_loss.backward()
    if hp.get("grad_max_norm_G") is not None:      # train.py:488-489
        torch.nn.utils.clip_grad_norm_(G.parameters(), hp["grad_max_norm_G"])
    out.update(g_adv=g_adv.detach(), g_rec=g_rec.detach(), g_idt=g_idt.detach(), g_idt_feat=idt_feat.detach(),
               g_idt_spec=idt_spec.detach(), g_cont=g_cont.detach(), g_loss=g_loss.detach())
    out["G_grad"] = grads_of(G, 0)[1]
    out["fake"] = fake.detach()
    return out


def case_step(name, cfg, hp):
    G, D = build_G(cfg), build_D(cfg)
    load(G, cfg["seed"]); load(D, cfg["seed"] + 100)
    C = None
    if hp["lambda_latcls"] != 0:
        from model.latent_classifier import LatentClassifier
        C = LatentClassifier(cfg["nspk"], cfg["content_dim"])
        load(C, cfg["seed"] + 200)
    b = make_batch(cfg["B"], cfg["T"], cfg["nspk"], seed=cfg["seed"] + 1, frames_div=int(np.prod(cfg["ratios"])),
                   permute=not hp["no_conv"])
    out = ref_step(G, D, b, hp, cfg["nspk"], C)
    save(name, **out)


def case_host_utils(name):
    """Host-side helpers of the path's callers: util.f0_to_excitation (util/__init__.py:22-50; consumes the
    global torch RNG: start phase, noise, unvoiced noise -- so the seed is part of the fixture) and the Kaiser filters
    (util/dsp.py:5-16, util/__init__.py kaiser_filter wrapper)."""
    import util as ref_util
    from util.dsp import kaiser_filter
    g = torch.Generator().manual_seed(77)
    f0 = 80.0 + 200.0 * torch.rand(2, 1, 15, generator=g)
    f0[0, 0, 3:6] = 0.0                       # an unvoiced stretch
    f0[1, 0, 10:] = 0.0
    out = {}
    for linear in (True, False):
        torch.manual_seed(2024)
        out[f"exc_linear{int(linear)}"] = ref_util.f0_to_excitation(f0.clone(), 64, sampling_rate=16000, linear=linear)
    out["f0"] = f0
    out["kaiser_129"] = kaiser_filter(129, 0.5, 10)
    for r in (2, 8, 10):
        out[f"kaiser_r{r}"] = ref_util.kaiser_filter(16 * r, 1 / r)
    save(name, **{k: v.float().numpy() for k, v in out.items()})


if __name__ == "__main__":
    only = sys.argv[1:]
    def want(n): return not only or n in only
    if want("host"): case_host_utils("host")
    if want("g_tiny"): case_generator("g_tiny", CASES["g_tiny"], full_limit=1 << 30)
    if want("d_tiny"): case_discriminator("d_tiny", CASES["d_tiny"], full_limit=20000)
    if want("msd_tiny"): case_discriminator("msd_tiny", CASES["d_tiny"], full_limit=20000, cls=MultiscaleDiscriminator)
    if want("cin"): case_cin("cin")
    if want("losses"): case_losses("losses")
    if want("latcls"): case_latent_classifier("latcls")
    if want("ssl_wn"): case_ssl_wn("ssl_wn")
    if want("yin"): case_yin("yin")
    if want("legacy"): case_legacy_blocks("legacy")
    if want("g_full"): case_generator("g_full", CASES["g_full"], full_limit=4096)
    if want("d_full"): case_discriminator("d_full", CASES["d_full"], full_limit=4096)
    if want("step_tiny_s1"): case_step("step_tiny_s1", CASES["step_tiny"], HP_STAGE1)
    if want("step_tiny_s21"): case_step("step_tiny_s21", CASES["step_tiny"], HP_STAGE2_1)
    if want("step_tiny_s22"): case_step("step_tiny_s22", CASES["step_tiny"], HP_STAGE2_2)
    if want("step_tiny_latcls"): case_step("step_tiny_latcls", CASES["step_tiny"], HP_LATCLS)
    if want("step_tiny_wave"): case_step("step_tiny_wave", CASES["step_tiny"], HP_WAVE_CLIP)
    if want("step_full_s1"): case_step("step_full_s1", CASES["step_full"], HP_STAGE1)
