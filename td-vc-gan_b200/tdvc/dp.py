"""Batch-sharded data parallelism: one process per GPU, full G and D replicas, gradient mean over ranks.

The reference is single-device; every op in the hot path is per-sample (InstanceNorm is
per (b, c), there is no BatchNorm, losses are means), so averaging gradients over ranks that each hold B
samples is numerically the large-batch step.  torch.distributed (NCCL over NVLink / NVSwitch) is the
plumbing; gradients travel as one flat fp32 bucket per network so a step costs two all-reduces
(71 MB for D, 59 MB for G) instead of ~800 small ones."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradAverager:
    """grad_hook for tdvc.train_step.TrainStep: flattens the gradients of one network into a persistent
    bucket, all-reduces it (AVG on NCCL, SUM then scale elsewhere) and scatters the result back."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._buckets = {}

    def reduce_flat(self, flat: torch.Tensor):
        """Average one persistent flat gradient bucket in place (FusedAdamW.bank())."""
        if self.world == 1:
            return
        if flat.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)

    def __call__(self, name: str, params: List[torch.nn.Parameter]):
        if self.world == 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        n = sum(g.numel() for g in grads)
        flat = self._buckets.get(name)
        if flat is None or flat.numel() != n or flat.device != grads[0].device:
            flat = torch.empty(n, device=grads[0].device, dtype=grads[0].dtype)
            self._buckets[name] = flat
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(views, grads)
        if flat.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        torch._foreach_copy_(grads, views)


class BucketedReducer:
    """Gradient mean overlapped with the backward pass that produces the gradients: the optimiser's flat
    gradient bank is cut into buckets of ~`bucket_mb` MB in REVERSE parameter order (backward reaches the last layers
    first); a post-accumulate-grad hook per parameter counts arrivals, and the moment a bucket is complete its gradients are
    copied into the bank and ONE asynchronous all-reduce of that slice is issued -- on the process group's communication
    stream, so it runs while the rest of the backward still computes.  `finish()` flushes the buckets a backward did not
    complete (parameters without gradient count as zeros) and waits for every outstanding collective.

    For eager steps.  Capturing these hook-driven collectives into the CUDA graph of the whole step was tried and never
    returned on 2 x B200 (torch 2.11 / NCCL 2.28.9); GraphedTrainStep therefore keeps the all-reduces between graph
    segments (tdvc/train_step.py)."""

    def __init__(self, opt, group: Optional[dist.ProcessGroup] = None, bucket_mb: float = 16.0):
        if getattr(opt, "_banks", None) is None:
            opt.use_grad_bank()
        self.opt, self.group = opt, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = []          # dict(flat slice, params, views)
        self.where = {}            # id(param) -> bucket index
        cap = int(bucket_mb * (1 << 20) / 4)
        for bank in opt._banks:
            params, views, flat = bank["params"], bank["views"], bank["flat"]
            offs, o = [], 0
            for p in params:
                offs.append(o)
                o += p.numel()
            hi = len(params)
            while hi > 0:
                lo = hi - 1
                while lo > 0 and offs[hi - 1] + params[hi - 1].numel() - offs[lo - 1] <= cap:
                    lo -= 1
                end = offs[hi - 1] + params[hi - 1].numel()
                self.buckets.append(dict(flat=flat[offs[lo]:end], params=params[lo:hi], views=views[lo:hi]))
                for p in params[lo:hi]:
                    self.where[id(p)] = len(self.buckets) - 1
                hi = lo
        self.pending = [0] * len(self.buckets)
        self.done = [True] * len(self.buckets)
        self.works = []
        self.order = []            # bucket indices in the order they were reduced (tests / diagnostics)
        self.armed = False
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for b in self.buckets for p in b["params"]]
        import weakref
        from tdvc import ops
        self._listener = weakref.WeakMethod(self._on_grad)     # gradients the batched weight-norm backward writes itself
        ops.grad_deposit_listeners.append(self._listener)

    def arm(self):
        """call before the backward pass whose gradients are to be averaged"""
        self.pending = [len(b["params"]) for b in self.buckets]
        self.done = [False] * len(self.buckets)
        self.works, self.order, self.armed = [], [], True

    def _on_grad(self, p):
        if not self.armed or id(p) not in self.where:
            return
        bi = self.where[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0 and not self.done[bi]:
            self._reduce(bi)

    @torch.no_grad()
    def _reduce(self, bi):
        b = self.buckets[bi]
        src = [p.grad for p in b["params"] if p.grad is not None]
        dst = [v for p, v in zip(b["params"], b["views"]) if p.grad is not None]
        missing = [v for p, v in zip(b["params"], b["views"]) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        if src:
            torch._foreach_copy_(dst, src)
        self.done[bi] = True
        self.order.append(bi)
        if self.world > 1:
            flat = b["flat"]
            if flat.is_cuda and dist.get_backend(self.group) == "nccl":
                self.works.append((dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
            else:
                self.works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat))

    def finish(self):
        """flush what the backward left incomplete, wait for the collectives; the bank then holds the averaged gradients"""
        for bi in range(len(self.buckets)):
            if not self.done[bi]:
                self._reduce(bi)
        for w, flat in self.works:
            w.wait()
            if flat is not None:
                flat.div_(self.world)
        self.works, self.armed = [], False
        self.opt._gathered = True

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        from tdvc import ops
        if self._listener in ops.grad_deposit_listeners:
            ops.grad_deposit_listeners.remove(self._listener)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Make every rank start from rank `src`'s weights (replicas are built with the same seed, this is a guard)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Split a global batch along dim 0 into `world` equal shards and return shard `rank`."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            if v.shape[0] % world:
                raise ValueError(f"batch dimension {v.shape[0]} of {k} is not divisible by world size {world}")
            n = v.shape[0] // world
            out[k] = v[rank * n:(rank + 1) * n]
        elif isinstance(v, (list, tuple)):
            out[k] = [shard_batch({"x": t}, rank, world)["x"] if torch.is_tensor(t) else t for t in v]
        else:
            out[k] = v
    return out
