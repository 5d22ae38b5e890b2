"""Fused multi-tensor AdamW: one kernel launch for a whole parameter list (reference uses
torch.optim.AdamW(params, lr, betas) at train.py:188-189; defaults eps=1e-8, weight_decay=1e-2).

Two ways to feed it gradients:
  * default: reads each p.grad in place (pointer tables are rebuilt when a gradient tensor moves);
  * `use_grad_bank()`: gradients are first gathered into one persistent flat fp32 buffer per parameter group
    (a single multi-tensor copy).  The bank has a fixed address, so the step is CUDA-graph capturable, and it is
    the bucket the data-parallel all-reduce operates on (`bank(i)`), between `gather_grads()` and `step()`.
The step counter is kept on the device so graph replays advance the bias correction."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_BLOCK = 4096          # elements per CTA of tdvc_adamw_blocks (ADAMW_BLOCK in csrc/misc.cu)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}
        self._banks = None
        self._gathered = False

    # ------------------------------------------------------------------ grad bank
    def use_grad_bank(self):
        if self._banks is not None:
            return self
        self._banks = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            n = sum(p.numel() for p in ps)
            flat = torch.zeros(n, device=ps[0].device, dtype=torch.float32)
            views, off = [], 0
            for p in ps:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self._banks.append(dict(params=ps, flat=flat, views=views))
        return self

    def bank(self, gi=0) -> torch.Tensor:
        return self._banks[gi]["flat"]

    @torch.no_grad()
    def gather_grads(self):
        """Copy every p.grad into the bank (zeros where a parameter received no gradient)."""
        for b in self._banks:
            src, dst = [], []
            missing = [v for p, v in zip(b["params"], b["views"]) if p.grad is None]
            for p, v in zip(b["params"], b["views"]):
                if p.grad is not None:
                    src.append(p.grad)
                    dst.append(v)
            if missing:
                torch._foreach_zero_(missing)
            if src:
                torch._foreach_copy_(dst, src)
        self._gathered = True

    # ------------------------------------------------------------------ tables
    def _ensure_state(self, params):
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)

    def _table(self, gi, params, grads):
        # keyed on every address the kernel will dereference (parameters, gradients -- 0 where a parameter has none, which
        # the kernel skips like torch.optim.AdamW does -- and both moments)
        self._ensure_state(params)
        key = tuple((p.data_ptr(), g.data_ptr() if g is not None else 0, self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p, g in zip(params, grads))
        cached = self._tables.get(gi)
        if cached is not None and cached["key"] == key:
            return cached
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("FusedAdamW: pointer tables must exist before CUDA-graph capture "
                               "(call use_grad_bank() and run one eager step first)")
        dev = params[0].device
        mk = lambda vals: torch.tensor(vals, dtype=torch.int64).to(dev)
        tab = dict(key=key,
                   p=mk([p.data_ptr() for p in params]), g=mk([g.data_ptr() if g is not None else 0 for g in grads]),
                   m=mk([self.state[p]["exp_avg"].data_ptr() for p in params]),
                   v=mk([self.state[p]["exp_avg_sq"].data_ptr() for p in params]),
                   n=mk([p.numel() for p in params]), max_n=max(p.numel() for p in params), count=len(params),
                   step=torch.zeros(1, device=dev, dtype=torch.float32))
        # equal blocks of _BLOCK elements over all tensors: (tensor index, first element) per CTA of tdvc_adamw_blocks
        blocks = [(ti, s) for ti, p in enumerate(params) for s in range(0, p.numel(), _BLOCK)]
        tab["blk"] = torch.tensor(blocks, dtype=torch.int32).reshape(-1, 2).to(dev)
        tab["n_blk"] = len(blocks)
        if cached is not None:
            tab["step"] = cached["step"]
        elif gi in getattr(self, "_restored_steps", {}):
            tab["step"].fill_(self._restored_steps.pop(gi))
        self._tables[gi] = tab
        return tab

    # ------------------------------------------------------------------ checkpointing
    def state_dict(self):
        """torch's layout plus the device-side step counters (one per parameter group), so that bias correction
        continues where it stopped."""
        sd = super().state_dict()
        sd["tdvc_steps"] = {gi: float(t["step"].item()) for gi, t in self._tables.items()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        steps = state_dict.pop("tdvc_steps", {})
        super().load_state_dict(state_dict)
        self._tables = {}                       # the moments were replaced: cached pointer tables are stale
        self._restored_steps = {int(k): float(v) for k, v in steps.items()}

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        if self._banks is not None and not self._gathered:
            self.gather_grads()
        self._gathered = False
        for gi, group in enumerate(self.param_groups):
            if self._banks is not None:
                bank = self._banks[gi]
                params = bank["params"]
                # a parameter that received no gradient is skipped (null gradient pointer), as torch.optim.AdamW skips it
                grads = [v if p.grad is not None else None for p, v in zip(params, bank["views"])]
            else:
                params = [p for p in group["params"] if p.grad is not None]
                grads = [p.grad for p in params]
            if not params:
                continue
            for p, g in zip(params, grads):
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and (g is None or g.is_contiguous())):
                    raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients")
            tab = self._table(gi, params, grads)
            b1, b2 = group["betas"]
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            vp = lambda t: C.c_void_p(t.data_ptr())
            _lib.check(lib.tdvc_adamw_blocks(vp(tab["p"]), vp(tab["g"]), vp(tab["m"]), vp(tab["v"]), vp(tab["n"]), vp(tab["blk"]),
                                             tab["n_blk"], group["lr"], b1, b2, group["eps"], group["weight_decay"], 0,
                                             float(grad_scale), vp(tab["step"]), st), "adamw_blocks")
        return loss
