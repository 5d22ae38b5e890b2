"""Fused multi-tensor AdamW: one kernel launch for a whole parameter list (reference uses
torch.optim.AdamW(params, lr, betas) at train.py:188-189; defaults eps=1e-8, weight_decay=1e-2)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}

    def _table(self, gi, params):
        """Device pointer tables for one param group; rebuilt only when a pointer changed."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        cached = self._tables.get(gi)
        if cached is not None and cached["key"] == key:
            return cached
        dev = params[0].device
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        mk = lambda vals: torch.tensor(vals, dtype=torch.int64).to(dev, non_blocking=False)
        tab = dict(key=key,
                   p=mk([p.data_ptr() for p in params]), g=mk([p.grad.data_ptr() for p in params]),
                   m=mk([self.state[p]["exp_avg"].data_ptr() for p in params]),
                   v=mk([self.state[p]["exp_avg_sq"].data_ptr() for p in params]),
                   n=mk([p.numel() for p in params]), max_n=max(p.numel() for p in params), count=len(params))
        self._tables[gi] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients")
            group["step"] = group.get("step", 0) + 1
            tab = self._table(gi, params)
            b1, b2 = group["betas"]
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            vp = lambda t: C.c_void_p(t.data_ptr())
            _lib.check(lib.tdvc_adamw_multi(vp(tab["p"]), vp(tab["g"]), vp(tab["m"]), vp(tab["v"]), vp(tab["n"]),
                                            tab["count"], tab["max_n"], group["lr"], b1, b2, group["eps"],
                                            group["weight_decay"], group["step"], float(grad_scale), st), "adamw_multi")
        return loss
