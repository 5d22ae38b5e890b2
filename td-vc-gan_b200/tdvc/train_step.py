"""One G+D training iteration, restating the loop body of the reference's train.py:209-511 on the tdvc
modules.  `train.py` itself runs unchanged against the drop-in `model` / `util` packages;
this class is the same computation packaged for benchmarking, CUDA-graph capture and data-parallel use
(it skips work that provably cannot change any result, each case cited below)."""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

import util.losses as losses
from tdvc import ops


def label2onehot(labels: torch.Tensor, n_classes: int) -> torch.Tensor:
    """train.py:39-44, built on the labels' device."""
    out = torch.zeros(labels.shape[0], n_classes, device=labels.device, dtype=torch.float32)
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


class TrainStep:
    def __init__(self, G, D, hp: dict, optimizer_G=None, optimizer_D=None, num_spk: Optional[int] = None,
                 grad_hook=None, C=None, optimizer_C=None):
        """hp: the `train:` section of a config/*.yaml as a dict (lambda_*, no_conv, jitter_amp ...).
        grad_hook(module_name, params) is called after each backward and before the optimiser step: the
        data-parallel gradient all-reduce plugs in there."""
        self.G, self.D, self.hp = G, D, hp
        self.opt_G, self.opt_D = optimizer_G, optimizer_D
        # latent classifier (train.py:153-154,192): only when lambda_latcls != 0
        self.C, self.opt_C = C, optimizer_C
        if hp.get("lambda_latcls", 0) != 0 and C is None:
            raise ValueError("lambda_latcls != 0 needs the LatentClassifier C")
        self.num_spk = num_spk if num_spk is not None else G.embedding.weight.shape[1]
        self.grad_hook = grad_hook

    def _reduce_and_step(self, name, module, opt):
        """(data-parallel gradient mean) + optimiser update.  With a FusedAdamW grad bank the all-reduce runs on
        the optimiser's flat bucket, otherwise on the parameters' .grad tensors."""
        banked = opt is not None and getattr(opt, "_banks", None) is not None
        if banked:
            opt.gather_grads()
            if self.grad_hook is not None:
                for gi in range(len(opt._banks)):
                    self.grad_hook.reduce_flat(opt.bank(gi))
        elif self.grad_hook is not None:
            self.grad_hook(name, list(module.parameters()))
        if opt is not None:
            opt.step()

    # ---- D step: train.py:259-296
    def d_step(self, batch) -> dict:
        out = self.d_forward_backward(batch)
        self._reduce_and_step("D", self.D, self.opt_D)
        if self.C is not None and self.hp.get("lambda_latcls", 0) != 0:
            out.update(self.c_step(batch, out["emb_real"]))
        return out

    def c_step(self, batch, emb_real) -> dict:
        """Latent classifier update, train.py:300-309 (the embedding comes from the D-step generator pass and carries
        no graph here, so only C receives gradients -- the reference zeroes what it deposits in G)."""
        out_lat = self.C(emb_real.detach())
        c_loss = torch.nn.functional.cross_entropy(out_lat, batch["label_src"])
        if self.opt_C is not None:
            self.opt_C.zero_grad(set_to_none=True)
        c_loss.backward()
        self._reduce_and_step("C", self.C, self.opt_C)
        return {"c_loss": c_loss.detach()}

    def d_forward_backward(self, batch) -> dict:
        with ops.step_cache():
            return self._d_forward_backward(batch)

    def g_forward_backward(self, batch, raw_draws=None) -> dict:
        with ops.step_cache():
            return self._g_forward_backward(batch, raw_draws)

    def _d_forward_backward(self, batch) -> dict:
        G, D = self.G, self.D
        x = batch["signal_real"]
        c_tgt = label2onehot(batch["label_tgt"], self.num_spk)
        # The reference builds G's graph here and never back-propagates it through G's weights: the full-rate
        # output is detached (train.py:269); the sub-scale heads are not, but the gradients they deposit in G
        # are zeroed before G's own backward (train.py:485-486).  no_grad skips that dead graph.
        with torch.no_grad():
            fake, fake_subs = G(x, c_tgt, c_var=batch["c_f0_conv"], out_subsample=True)
            emb_real = getattr(G, "content_embedding", None)
            real_subs = D.get_subsamples(x)
        o_real, _ = D(x, batch["label_src"], real_subs)
        o_fake, _ = D(fake, batch["label_tgt"], fake_subs)
        d_real = ops.mse_to_const_sum(o_real, 1.0)
        d_fake = ops.mse_to_const_sum(o_fake, 0.0)
        d_loss = d_real + d_fake
        if self.opt_D is not None:
            self.opt_D.zero_grad(set_to_none=True)
        d_loss.backward()
        return {"d_loss_real": d_real.detach(), "d_loss_fake": d_fake.detach(), "d_loss": d_loss.detach(),
                "fake": fake, "emb_real": emb_real}

    # ---- G step: train.py:320-491 (lambda_f0 needs torchcrepe, lambda_latcls the latent classifier: both 0 here)
    def g_step(self, batch, raw_draws=None) -> dict:
        out = self.g_forward_backward(batch, raw_draws)
        self._reduce_and_step("G", self.G, self.opt_G)
        return out

    def _g_forward_backward(self, batch, raw_draws=None) -> dict:
        G, D, hp = self.G, self.D, self.hp
        x = batch["signal_real"]
        lab_s, lab_t = batch["label_src"], batch["label_tgt"]
        c_src = label2onehot(lab_s, self.num_spk)
        c_tgt = label2onehot(lab_t, self.num_spk)
        out = {}
        with frozen(D):
            fake, fake_subs = G(x, c_tgt, c_var=batch["c_f0_conv"], out_subsample=True)
            emb_real = G.content_embedding
            o_fake, _ = D(fake, lab_t, fake_subs)
            g_adv = ops.mse_to_const_sum(o_fake, 1.0)
            f_real = None
            if (hp["lambda_rec"] > 0 or hp["lambda_idt"] > 0) and hp["lambda_feat"] > 0:
                with torch.no_grad():   # reference features are .detach()ed inside the loss (losses.py:63)
                    _, f_real = D(x, lab_s, D.get_subsamples(x))
            zero = torch.zeros((), device=x.device)
            g_rec = zero
            if (not hp["no_conv"]) and hp["lambda_rec"] > 0:
                rec, rec_subs = G(fake.detach(), c_src, c_var=batch["c_f0_src"], out_subsample=True)
                if hp["lambda_feat"] > 0:
                    _, f_rec = D(rec, lab_s, rec_subs)
                    g_rec = g_rec + hp["lambda_feat"] * losses.multiscale_feat_loss(f_rec, f_real, norm_p=1)
                if hp["lambda_spec"] > 0:
                    g_rec = g_rec + hp["lambda_spec"] * losses.multiscale_spec_loss(rec, x, [2048, 1024, 512])
            g_idt = zero
            if hp["lambda_idt"] > 0:
                if not hp["no_conv"]:
                    idt, idt_subs = G(x, c_src, c_var=batch["c_f0_src"], out_subsample=True)
                else:
                    idt, idt_subs = fake, fake_subs
                if hp["lambda_feat"] > 0:
                    _, f_idt = D(idt, lab_s, idt_subs)
                    g_idt = g_idt + hp["lambda_feat"] * losses.multiscale_feat_loss(f_idt, f_real, norm_p=1)
                if hp["lambda_spec"] > 0:
                    g_idt = g_idt + hp["lambda_spec"] * losses.multiscale_spec_loss(idt, x, [2048, 1024, 512])
            g_cont = zero
            if hp["lambda_cont_emb"] > 0 and hp["lambda_corrupted"]:
                emb_corr = G.encoder(batch["signal_corrupted"])
                g_cont = g_cont + losses.contrastive_loss(emb_real, emb_corr, num_negatives=100, temp=0.1,
                                                          _raw_draws=raw_draws)
            g_loss = g_adv + hp["lambda_rec"] * g_rec + hp["lambda_idt"] * g_idt + hp["lambda_cont_emb"] * g_cont
            if self.C is not None and hp.get("lambda_latcls", 0) != 0:
                # train.py:420-425: speaker classification of the content embedding through the gradient-reversal
                # layer; C's own weights are not updated by this loss (optimizer_C.step() is not called here)
                with frozen(self.C):
                    g_lat = torch.nn.functional.cross_entropy(self.C(emb_real), lab_s)
                out["g_latcls"] = g_lat.detach()
                g_loss = g_loss + hp["lambda_latcls"] * g_lat
            if self.opt_G is not None:
                self.opt_G.zero_grad(set_to_none=True)
            g_loss.backward()
        out.update(g_adv=g_adv.detach(), g_rec=g_rec.detach(), g_idt=g_idt.detach(), g_cont=g_cont.detach(),
                   g_loss=g_loss.detach(), fake=fake.detach())
        return out

    def step(self, batch, raw_draws=None) -> dict:
        out = self.d_step(batch)
        out.update(self.g_step(batch, raw_draws))
        return out


class GraphedTrainStep:
    """The G+D iteration captured into CUDA graphs and replayed: the step issues ~8000 kernel launches from
    Python, which costs more host time than the GPU needs to run them.  Inputs are copied into static tensors;
    outputs are static tensors overwritten by every replay.  Needs optimisers with a grad bank (fixed gradient
    addresses) -- FusedAdamW.use_grad_bank().

    Single GPU: one graph for the whole iteration.  Data parallel: three graphs with the two gradient all-reduces
    issued eagerly between them on the same stream ([D fwd/bwd + gather] -> all-reduce(D bank) -> [AdamW(D) +
    G fwd/bwd + gather] -> all-reduce(G bank) -> [AdamW(G)]), so NCCL never runs inside a capture."""

    def __init__(self, ts: TrainStep, example_batch: dict, warmup: int = 3):
        from tdvc import ops
        self.ts = ts
        for opt in (ts.opt_G, ts.opt_D):
            if opt is not None and hasattr(opt, "use_grad_bank"):
                opt.use_grad_bank()
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        self.split = ts.grad_hook is not None and getattr(ts.grad_hook, "world", 1) > 1
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                ts.step(self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops._pack_cache.clear()
        if not self.split:
            self.graphs = [torch.cuda.CUDAGraph()]
            with torch.cuda.graph(self.graphs[0]):
                self.out = ts.step(self.static)
        else:
            g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self.out = ts.d_forward_backward(self.static)
                ts.opt_D.gather_grads()
            ops._pack_cache.clear()
            with torch.cuda.graph(g2, pool=g1.pool()):
                ts.opt_D._gathered = True
                ts.opt_D.step()
                self.out.update(ts.g_forward_backward(self.static))
                ts.opt_G.gather_grads()
            with torch.cuda.graph(g3, pool=g1.pool()):
                ts.opt_G._gathered = True
                ts.opt_G.step()
            self.graphs = [g1, g2, g3]
        ops._pack_cache.clear()

    def load(self, batch: dict, non_blocking: bool = True):
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=non_blocking)

    def step(self, batch: Optional[dict] = None) -> dict:
        if batch is not None:
            self.load(batch)
        if not self.split:
            self.graphs[0].replay()
        else:
            ts = self.ts
            self.graphs[0].replay()
            for gi in range(len(ts.opt_D._banks)):
                ts.grad_hook.reduce_flat(ts.opt_D.bank(gi))
            self.graphs[1].replay()
            for gi in range(len(ts.opt_G._banks)):
                ts.grad_hook.reduce_flat(ts.opt_G.bank(gi))
            self.graphs[2].replay()
        return self.out
