"""One G+D training iteration, restating the loop body of the reference's train.py:209-511 on the tdvc
modules.  `train.py` itself runs unchanged against the drop-in `model` / `util` packages;
this class is the same computation packaged for benchmarking, CUDA-graph capture and data-parallel use.

Work that cannot change any result is done once instead of several times (each case cited where it happens):

  * the reference runs `G(signal_real, c_tgt, c_f0_conv)` in the D step (train.py:262) and again, on the same
    weights, in the G step (train.py:322): here the generator runs ONCE per iteration, the D step takes the
    detached outputs, the G step back-propagates through the same graph;
  * `G(x, c_tgt)`, `G(x, c_src)` (identity pass, train.py:370) and `G.encoder(signal_corrupted)` (train.py:405)
    share one encoder pass over `cat(x, signal_corrupted)` (the conv encoder takes no conditioning) and one decoder
    pass over the two conditionings stacked along the batch; every op of the path is per-sample, so each
    sample's outputs and gradients are those of the separate calls;
  * `D(real)` / `D(fake)` of the D step, and `D(fake)` / `D(idt)` / `D(rec)` of the G step, are one batched
    discriminator call each.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

import util.losses as losses
from tdvc import ops


def label2onehot(labels: torch.Tensor, n_classes: int, dtype=torch.float32) -> torch.Tensor:
    """train.py:39-44, built on the labels' device."""
    out = torch.zeros(labels.shape[0], n_classes, device=labels.device, dtype=dtype)
    return out.scatter_(1, labels.view(-1, 1), 1.0)      # scalar-valued scatter: no host tensor, graph-capturable


@contextlib.contextmanager
def frozen(module: torch.nn.Module):
    """Temporarily stop producing weight gradients for `module` (its inputs still get gradients).  During the
    G step the reference computes D's weight gradients and throws them away (optimizer_D.zero_grad() precedes,
    optimizer_D.step() never follows: train.py:485-491)."""
    flags = [p.requires_grad for p in module.parameters()]
    for p in module.parameters():
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p, f in zip(module.parameters(), flags):
            p.requires_grad_(f)


_KNOWN_HP = {"no_conv", "lambda_rec", "lambda_idt", "lambda_feat", "lambda_spec", "lambda_wave", "lambda_latcls",
             "lambda_cont_emb", "lambda_corrupted", "lambda_converted", "lambda_f0", "jitter_amp", "grad_max_norm_G",
             "grad_max_norm_D"}


class TrainStep:
    def __init__(self, G, D, hp: dict, optimizer_G=None, optimizer_D=None, num_spk: Optional[int] = None,
                 grad_hook=None, C=None, optimizer_C=None):
        """hp: the `train:` section of a config/*.yaml as a dict (lambda_*, no_conv, jitter_amp, grad_max_norm_* ...;
        keys that only drive the data loader / schedule -- batch_size, lr, ... -- are ignored).
        grad_hook.reduce_flat(bank) / grad_hook(name, params) is called after each backward and before the optimiser
        step: the data-parallel gradient all-reduce plugs in there."""
        self.G, self.D, self.hp = G, D, hp
        self.opt_G, self.opt_D = optimizer_G, optimizer_D
        # latent classifier (train.py:153-154,192): only when lambda_latcls != 0
        self.C, self.opt_C = C, optimizer_C
        if hp.get("lambda_latcls", 0) != 0 and C is None:
            raise ValueError("lambda_latcls != 0 needs the LatentClassifier C")
        if hp.get("lambda_f0", 0) != 0:
            # train.py:427-470: needs torchcrepe's pitch activations of the converted signal (util/crepe.py), which is
            # neither part of this path nor installable offline -- refuse rather than train with a different loss
            raise NotImplementedError("lambda_f0 != 0 needs util.crepe (torchcrepe); set lambda_f0 to 0")
        if hp.get("lambda_converted", 0):
            # train.py:409-413 computes this term and then adds it to itself, never to the loss: nothing to restate,
            # but its G.encoder(signal_fake.detach()) pass is not run here either
            pass
        self.num_spk = num_spk if num_spk is not None else G.embedding.weight.shape[1]
        self.grad_hook = grad_hook
        # gradient all-reduce overlapped with the backward (tdvc.dp.BucketedReducer per optimiser), see enable_overlap()
        self.reducers = {}
        self._gen = None          # generator passes of the current iteration (shared by the D and the G step)
        from model.generator import Encoder
        enc = getattr(G, "encoder", None)
        # one encoder pass / one stacked decoder pass need: the conv encoder without speaker conditioning and
        # target-only conditioning further on (true for every shipped conv_enc config)
        self.batched = (isinstance(enc, Encoder) and not enc.cin and not enc.spk_conditioning
                        and not getattr(G, "both_cond", False) and len(getattr(G, "bottleneck", [])) == 0)

    # ------------------------------------------------------------------ plumbing
    def _clip(self, module, opt, max_norm):
        """torch.nn.utils.clip_grad_norm_(params, max_norm) (train.py:288-289,488-489); on the optimiser's flat
        gradient bank when there is one (after the data-parallel mean, as DDP would see it)."""
        if max_norm is None:
            return
        banked = opt is not None and getattr(opt, "_banks", None) is not None
        if banked:
            flats = [opt.bank(gi) for gi in range(len(opt._banks))]
            total = torch.sqrt(sum((f * f).sum() for f in flats))
            coef = torch.clamp(float(max_norm) / (total + 1e-6), max=1.0)
            for f in flats:
                f.mul_(coef)
        else:
            torch.nn.utils.clip_grad_norm_(module.parameters(), max_norm)

    def enable_overlap(self, bucket_mb: float = 16.0, group=None):
        """Data parallel: average the gradients of D, G (and C) bucket by bucket WHILE the backward that produces them still
        runs (north_star: "gradient allreduce over NCCL overlapped with the discriminator and generator backward"), instead
        of one all-reduce per network after its backward.  Needs FusedAdamW optimisers (the buckets are slices of their flat
        gradient banks)."""
        from tdvc.dp import BucketedReducer
        for name, opt in (("D", self.opt_D), ("G", self.opt_G), ("C", self.opt_C)):
            if opt is not None:
                self.reducers[name] = BucketedReducer(opt, group=group, bucket_mb=bucket_mb)
        return self

    def _arm(self, name):
        r = self.reducers.get(name)
        if r is not None:
            r.arm()

    def _reduce_and_step(self, name, module, opt, max_norm=None):
        """(data-parallel gradient mean) + (clip) + optimiser update.  With a FusedAdamW grad bank the all-reduce runs
        on the optimiser's flat bucket, otherwise on the parameters' .grad tensors."""
        banked = opt is not None and getattr(opt, "_banks", None) is not None
        if name in self.reducers:
            self.reducers[name].finish()        # buckets were reduced during the backward; wait for the stragglers
        elif banked:
            opt.gather_grads()
            if self.grad_hook is not None:
                for gi in range(len(opt._banks)):
                    self.grad_hook.reduce_flat(opt.bank(gi))
        elif self.grad_hook is not None:
            self.grad_hook(name, list(module.parameters()))
        self._clip(module, opt, max_norm)
        if opt is not None:
            opt.step()

    # ------------------------------------------------------------------ generator passes (train.py:262,322,346,370,405)
    def _generator_passes(self, batch) -> dict:
        """Every generator forward of the iteration, with autograd graph: fake = G(x, c_tgt, c_f0_conv), idt =
        G(x, c_src, c_f0_src) (when lambda_idt > 0 and not no_conv), emb_corr = G.encoder(signal_corrupted) (when the
        contrastive term is on), rec = G(fake.detach(), c_src, c_f0_src) (when lambda_rec > 0 and not no_conv)."""
        G, hp = self.G, self.hp
        x = batch["signal_real"]
        B = x.shape[0]
        c_tgt = label2onehot(batch["label_tgt"], self.num_spk, x.dtype)
        c_src = label2onehot(batch["label_src"], self.num_spk, x.dtype)
        want_idt = hp["lambda_idt"] > 0 and not hp["no_conv"]
        want_corr = hp["lambda_cont_emb"] > 0 and bool(hp["lambda_corrupted"])
        want_rec = (not hp["no_conv"]) and hp["lambda_rec"] > 0
        gen = {}
        if self.batched:
            enc_in = torch.cat([x, batch["signal_corrupted"]], dim=0) if want_corr else x
            emb_all = G.encoder(enc_in)
            emb = emb_all[:B]
            gen["emb_real"] = emb
            if want_corr:
                gen["emb_corr"] = emb_all[B:]
            if want_idt:
                c = G.embedding(torch.cat([c_tgt, c_src], dim=0))
                h = torch.cat([emb, emb], dim=0)
                cv = torch.cat([batch["c_f0_conv"], batch["c_f0_src"]], dim=0)
            else:
                c, h, cv = G.embedding(c_tgt), emb, batch["c_f0_conv"]
            y, subs = G.decoder(h, c, cv, out_subsample=True)
            if getattr(G, "output_content_emb", False):
                G.content_embedding = emb
            gen["y"], gen["subs"] = y, subs                       # [fake ; idt] stacked along the batch
            gen["fake"], gen["fake_subs"] = y[:B], [s[:B] for s in subs]
            if want_idt:
                gen["idt"], gen["idt_subs"] = y[B:], [s[B:] for s in subs]
        else:
            fake, fake_subs = G(x, c_tgt, c_var=batch["c_f0_conv"], out_subsample=True)
            gen["emb_real"] = getattr(G, "content_embedding", None)
            gen["fake"], gen["fake_subs"] = fake, fake_subs
            if want_idt:
                gen["idt"], gen["idt_subs"] = G(x, c_src, c_var=batch["c_f0_src"], out_subsample=True)
            if want_corr:
                gen["emb_corr"] = G.encoder(batch["signal_corrupted"])
        if hp["lambda_idt"] > 0 and hp["no_conv"]:
            gen["idt"], gen["idt_subs"] = gen["fake"], gen["fake_subs"]       # train.py:373-375
        if want_rec:
            gen["rec"], gen["rec_subs"] = G(gen["fake"].detach(), c_src, c_var=batch["c_f0_src"], out_subsample=True)
            if getattr(G, "output_content_emb", False):
                G.content_embedding = gen["emb_real"]
        return gen

    def _gen_for(self, batch) -> dict:
        """The iteration's generator passes (run on first use).  Their step scope stays open until the G backward has
        consumed the graph: G's weights do not change in between, so the normalised / packed weights of the forward
        serve the backward (D's and C's scopes nest inside and are dropped before their optimiser runs)."""
        if self._gen is None or self._gen[0] is not batch:
            self._close_gen()
            scope = ops.step_cache("G")
            scope.__enter__()
            try:
                self._gen = (batch, self._generator_passes(batch), scope)
            except BaseException:
                scope.__exit__(None, None, None)
                raise
        return self._gen[1]

    def _close_gen(self):
        if self._gen is not None:
            self._gen[2].__exit__(None, None, None)
            self._gen = None

    # ------------------------------------------------------------------ D step: train.py:259-296
    def d_step(self, batch) -> dict:
        out = self.d_forward_backward(batch)
        self._reduce_and_step("D", self.D, self.opt_D, self.hp.get("grad_max_norm_D"))
        if self.C is not None and self.hp.get("lambda_latcls", 0) != 0:
            out.update(self.c_step(batch, out["emb_real"]))
        return out

    def c_forward_backward(self, batch, emb_real) -> dict:
        """Latent classifier loss and backward, train.py:300-305 (the embedding is detached here, so only C receives
        gradients -- the reference zeroes what this deposits in G)."""
        with ops.step_cache("C"):
            out_lat = self.C(emb_real.detach())
            c_loss = torch.nn.functional.cross_entropy(out_lat, batch["label_src"])
            if self.opt_C is not None:
                self.opt_C.zero_grad(set_to_none=True)
            self._arm("C")
            c_loss.backward()
        return {"c_loss": c_loss.detach()}

    def c_step(self, batch, emb_real) -> dict:
        out = self.c_forward_backward(batch, emb_real)
        self._reduce_and_step("C", self.C, self.opt_C)
        return out

    def d_forward_backward(self, batch) -> dict:
        gen = self._gen_for(batch)
        with ops.step_cache("D"):
            return self._d_forward_backward(batch, gen)

    def _d_forward_backward(self, batch, gen) -> dict:
        D = self.D
        x = batch["signal_real"]
        B = x.shape[0]
        # The reference detaches the full-rate output (train.py:269); the sub-scale heads are not detached there, but
        # the gradients they deposit in G are zeroed before G's own backward (train.py:485-486): detaching skips them.
        fake = gen["fake"].detach()
        fake_subs = [s.detach() for s in gen["fake_subs"]]
        with torch.no_grad():
            real_subs = D.get_subsamples(x)
        # D(real) and D(fake) as one call on [real ; fake]
        sig = torch.cat([x, fake], dim=0)
        subs = [torch.cat([r, f], dim=0) for r, f in zip(real_subs, fake_subs)]
        lab = torch.cat([batch["label_src"], batch["label_tgt"]], dim=0)
        o, _ = D(sig, lab, subs)
        d_real = ops.mse_to_const_sum([t[:B] for t in o], 1.0)
        d_fake = ops.mse_to_const_sum([t[B:] for t in o], 0.0)
        d_loss = d_real + d_fake
        if self.opt_D is not None:
            self.opt_D.zero_grad(set_to_none=True)
        self._arm("D")
        d_loss.backward()
        return {"d_loss_real": d_real.detach(), "d_loss_fake": d_fake.detach(), "d_loss": d_loss.detach(),
                "fake": fake, "emb_real": gen["emb_real"]}

    # ------------------------------------------------------------------ G step: train.py:320-491
    def g_step(self, batch, raw_draws=None) -> dict:
        out = self.g_forward_backward(batch, raw_draws)
        self._reduce_and_step("G", self.G, self.opt_G, self.hp.get("grad_max_norm_G"))
        return out

    def g_forward_backward(self, batch, raw_draws=None) -> dict:
        gen = self._gen_for(batch)
        try:
            # D (and C) are frozen BEFORE their scope opens: the scope's batched weight norm then builds no autograd node
            # for them, and their convolutions compute no weight gradients in this step
            with contextlib.ExitStack() as stack:
                stack.enter_context(frozen(self.D))
                if self.C is not None:
                    stack.enter_context(frozen(self.C))
                stack.enter_context(ops.step_cache("Gstep"))
                return self._g_forward_backward(batch, gen, raw_draws)
        finally:
            self._close_gen()                 # the graph is consumed by this backward

    def _g_forward_backward(self, batch, gen, raw_draws=None) -> dict:
        G, D, hp = self.G, self.D, self.hp
        x = batch["signal_real"]
        B = x.shape[0]
        lab_s, lab_t = batch["label_src"], batch["label_tgt"]
        out = {}
        want_feat = (hp["lambda_rec"] > 0 or hp["lambda_idt"] > 0) and hp["lambda_feat"] > 0
        x_ref = x
        if (hp["lambda_rec"] > 0 or hp["lambda_idt"] > 0) and hp.get("jitter_amp", 0) > 0:
            import util
            a = int(hp["jitter_amp"])                               # util/audio.py:27-30
            jitter = torch.randint(-a, a + 1, (B,), device=x.device)
            x_ref = util.roll_batches(x, jitter, x.ndim - 1)
        with frozen(D):
            # every generated signal that goes through D in this step, as one batched call:
            # fake (adversarial) [+ idt] [+ rec] (feature matching)
            sigs = [("fake", gen["fake"], gen["fake_subs"], lab_t)]
            if hp["lambda_idt"] > 0 and hp["lambda_feat"] > 0 and not hp["no_conv"]:
                sigs.append(("idt", gen["idt"], gen["idt_subs"], lab_s))
            if "rec" in gen and hp["lambda_feat"] > 0:
                sigs.append(("rec", gen["rec"], gen["rec_subs"], lab_s))
            if self.batched and len(sigs) >= 2 and sigs[1][0] == "idt" and "y" in gen:
                head_sig, head_subs = gen["y"], gen["subs"]          # [fake ; idt] is already one tensor
                rest = sigs[2:]
            else:
                head_sig, head_subs = sigs[0][1], sigs[0][2]
                rest = sigs[1:]
            if rest:
                sig = torch.cat([head_sig] + [s[1] for s in rest], dim=0)
                subs = [torch.cat([hs] + [s[2][i] for s in rest], dim=0) for i, hs in enumerate(head_subs)]
            else:
                sig, subs = head_sig, head_subs
            lab = torch.cat([s[3] for s in sigs], dim=0)
            o, feats = D(sig, lab, subs)
            pos = {s[0]: i * B for i, s in enumerate(sigs)}
            g_adv = ops.mse_to_const_sum([t[:B] for t in o], 1.0)
            f_real = None
            if want_feat:
                with torch.no_grad():   # reference features are .detach()ed inside the loss (losses.py:63)
                    _, f_real = D(x_ref, lab_s, D.get_subsamples(x_ref))
            zero = torch.zeros((), device=x.device, dtype=x.dtype)

            def rec_like_terms(name, signal):
                """lambda_feat * feat + lambda_spec * mel (+ wave L1 returned apart: train.py:358-361,382-385)."""
                total, wave = zero, zero
                if hp["lambda_feat"] > 0:
                    if name in pos:
                        total = total + hp["lambda_feat"] * losses.multiscale_feat_loss_rows(feats, pos[name], B, f_real)
                    else:   # no_conv: idt is fake (train.py:373-375)
                        total = total + hp["lambda_feat"] * losses.multiscale_feat_loss_rows(feats, pos["fake"], B, f_real)
                if hp["lambda_spec"] > 0:
                    total = total + hp["lambda_spec"] * losses.multiscale_spec_loss(signal, x_ref, [2048, 1024, 512])
                if hp.get("lambda_wave", 0) > 0:
                    wave = ops.l1_mean_sum([signal], [x])
                return total, wave

            g_rec = zero
            if "rec" in gen:
                g_rec, wave = rec_like_terms("rec", gen["rec"])
                g_rec = g_rec + hp.get("lambda_wave", 0) * wave
            g_idt = zero
            if hp["lambda_idt"] > 0:
                g_idt, wave = rec_like_terms("idt", gen["idt"])
                # the reference adds the identity pass' wave term to g_loss_rec (train.py:384)
                g_rec = g_rec + hp.get("lambda_wave", 0) * wave
            g_cont = zero
            if "emb_corr" in gen:
                g_cont = g_cont + losses.contrastive_loss(gen["emb_real"], gen["emb_corr"], num_negatives=100, temp=0.1,
                                                          _raw_draws=raw_draws)
            g_loss = g_adv + hp["lambda_rec"] * g_rec + hp["lambda_idt"] * g_idt + hp["lambda_cont_emb"] * g_cont
            if self.C is not None and hp.get("lambda_latcls", 0) != 0:
                # train.py:420-425: speaker classification of the content embedding through the gradient-reversal
                # layer; C's own weights are not updated by this loss (optimizer_C.step() is not called here)
                with frozen(self.C):
                    g_lat = torch.nn.functional.cross_entropy(self.C(gen["emb_real"]), lab_s)
                out["g_latcls"] = g_lat.detach()
                g_loss = g_loss + hp["lambda_latcls"] * g_lat
            if self.opt_G is not None:
                self.opt_G.zero_grad(set_to_none=True)
            self._arm("G")
            g_loss.backward()
        out.update(g_adv=g_adv.detach(), g_rec=g_rec.detach(), g_idt=g_idt.detach(), g_cont=g_cont.detach(),
                   g_loss=g_loss.detach(), fake=gen["fake"].detach())
        return out

    def step(self, batch, raw_draws=None) -> dict:
        out = self.d_step(batch)
        out.update(self.g_step(batch, raw_draws))
        return out


class GraphedTrainStep:
    """The G+D iteration captured into CUDA graphs and replayed: the step issues thousands of kernel launches from
    Python, which costs more host time than the GPU needs to run them.  Inputs are copied into static tensors;
    outputs are static tensors overwritten by every replay.  Needs optimisers with a grad bank (fixed gradient
    addresses) -- FusedAdamW.use_grad_bank().

    Single GPU: one graph for the whole iteration.  Data parallel: the iteration is cut at every gradient all-reduce
    -- [generator passes, D fwd/bwd, gather] -> all-reduce(D) -> [AdamW(D), (C fwd/bwd, gather] -> all-reduce(C) ->
    [AdamW(C),) G-step fwd/bwd, gather] -> all-reduce(G) -> [AdamW(G)] -- so NCCL never runs inside a capture."""

    def __init__(self, ts: TrainStep, example_batch: dict, warmup: int = 3):
        self.ts = ts
        if ts.reducers and getattr(ts.grad_hook, "world", 1) > 1:
            # measured on 2 x B200 (torch 2.11, NCCL 2.28.9): capturing the hook-driven asynchronous all-reduces into the
            # step graph never returns.  The graph path cuts the iteration at the all-reduces instead (below).
            raise RuntimeError("GraphedTrainStep: the hook-driven gradient all-reduce (TrainStep.enable_overlap) is for eager "
                               "steps; under CUDA graphs the all-reduces run between graph segments")
        self.with_c = ts.C is not None and ts.hp.get("lambda_latcls", 0) != 0
        opts = [ts.opt_G, ts.opt_D] + ([ts.opt_C] if self.with_c else [])
        for opt in opts:
            if opt is None or not hasattr(opt, "use_grad_bank"):
                raise RuntimeError("GraphedTrainStep needs FusedAdamW optimisers (grad bank) for G, D and, when "
                                   "lambda_latcls != 0, for C")
            opt.use_grad_bank()
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        # data parallel: the iteration is cut at the gradient all-reduces, which are issued eagerly between the segments
        self.split = ts.grad_hook is not None and getattr(ts.grad_hook, "world", 1) > 1
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                ts.step(self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops._pack_cache.clear()
        hp = ts.hp
        if not self.split:
            self.graphs = [torch.cuda.CUDAGraph()]
            self.reduce_after = [None]
            with torch.cuda.graph(self.graphs[0]):
                self.out = ts.step(self.static)
        else:
            # segments, each followed by the all-reduce of one optimiser's bank
            def seg_d():
                self.out = ts.d_forward_backward(self.static)
                ts.opt_D.gather_grads()

            def seg_c():
                self._finish(ts.D, ts.opt_D, hp.get("grad_max_norm_D"))
                self.out.update(ts.c_forward_backward(self.static, self.out["emb_real"]))
                ts.opt_C.gather_grads()

            def seg_g():
                if self.with_c:
                    self._finish(ts.C, ts.opt_C, None)
                else:
                    self._finish(ts.D, ts.opt_D, hp.get("grad_max_norm_D"))
                self.out.update(ts.g_forward_backward(self.static))
                ts.opt_G.gather_grads()

            def seg_end():
                self._finish(ts.G, ts.opt_G, hp.get("grad_max_norm_G"))

            plan = [(seg_d, ts.opt_D)] + ([(seg_c, ts.opt_C)] if self.with_c else []) + [(seg_g, ts.opt_G), (seg_end, None)]
            self.graphs, self.reduce_after = [], []
            pool = None
            for fn, opt in plan:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    fn()
                pool = g.pool()
                ops._pack_cache.clear()
                self.graphs.append(g)
                self.reduce_after.append(opt)
        ops._pack_cache.clear()

    def _finish(self, module, opt, max_norm):
        """clip + optimiser update on an already gathered (and all-reduced) bank"""
        self.ts._clip(module, opt, max_norm)
        opt._gathered = True
        opt.step()

    def load(self, batch: dict, non_blocking: bool = True):
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=non_blocking)

    def step(self, batch: Optional[dict] = None) -> dict:
        if batch is not None:
            self.load(batch)
        hook = self.ts.grad_hook
        for g, opt in zip(self.graphs, self.reduce_after):
            g.replay()
            if opt is not None:
                for gi in range(len(opt._banks)):
                    hook.reduce_flat(opt.bank(gi))
        return self.out
