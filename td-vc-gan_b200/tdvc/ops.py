"""torch.autograd.Functions over the C ABI (include/tdvc_b200.h).

Each op mirrors one ATen call site of the reference (cited in the header) and keeps its
argument meaning.  PyTorch is used for device memory, streams and autograd bookkeeping only:
all arithmetic happens in libtdvc_b200.so.  Inputs must be CUDA tensors; fp32, NCW.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, PAD_REFLECT, PAD_ZEROS, ConvGeom

_ACT = {None: ACT_NONE, "none": ACT_NONE, "lrelu": ACT_LRELU, "tanh": ACT_TANH}

# precision mode of the dense stride-1 convolutions: "fp32" (CUDA-core exact path, rel 1e-5 vs the
# reference) or "bf16" (tcgen05 tensor-core path, bf16 operands / fp32 accumulate, rel 2e-2)
_PRECISION = "fp32"


def set_precision(mode: str):
    global _PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError(mode)
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


_BRANCH_STREAMS = os.environ.get("TDVC_BRANCH_STREAMS", "0") == "1"
_branch_stream_pool = {}


def branch_streams_enabled() -> bool:
    return _BRANCH_STREAMS


def set_branch_streams(on: bool):
    """Run the independent kernel-size branches of every MRF block on forked CUDA streams."""
    global _BRANCH_STREAMS
    _BRANCH_STREAMS = bool(on)


def branch_streams(device, n):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _branch_stream_pool.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(*ts):
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("tdvc ops need CUDA tensors (there is no CPU fallback); got a "
                               f"{t.device} tensor")
        if t.dtype != torch.float32:
            raise RuntimeError(f"tdvc ops are fp32 at the boundary; got {t.dtype}")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            # kernels are launched on the current device's current stream
            raise RuntimeError(f"tdvc ops run on the current CUDA device (cuda:{cur}); got a {t.device} tensor -- wrap the call "
                               "in torch.cuda.device(...)")


def _c(t: Optional[torch.Tensor]):
    """contiguous(); with branch streams on, also tells the caching allocator that the tensor is read on the current
    stream: tensors (activations forward, gradients backward) cross the forked MRF branch streams, and a block
    freed by its allocating stream must not be recycled while another stream still reads it."""
    if t is None:
        return None
    if _BRANCH_STREAMS and t.is_cuda:
        t.record_stream(torch.cuda.current_stream())
    return t.contiguous()


def _geom(B, Cin, Tin, Cout, Tout, K, stride, pad, dilation, groups, pad_mode, in_slope, out_act, out_slope):
    return ConvGeom(B, Cin, Tin, Cout, Tout, K, stride, pad, dilation, groups, pad_mode, float(in_slope),
                    out_act, float(out_slope))


# ----------------------------------------------------------------------------- weight norm

class _WeightNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, g):
        _req(v, g)
        pre = _step_cache.lookup_w(v, g)
        if pre is not None and pre[1] is not None:
            w, inv = pre
        else:
            v, g = _c(v), _c(g)
            rows = v.shape[0]
            cols = v.numel() // rows
            w = torch.empty_like(v)
            inv = torch.empty(rows, device=v.device, dtype=torch.float32)
            _lib.check(_lib.load().tdvc_weight_norm_fwd(_p(v), _p(g), _p(w), _p(inv), rows, cols, _st()), "weight_norm_fwd")
            _step_cache.note_w(v, g, w)
        ctx.save_for_backward(v, g, inv)
        return w

    @staticmethod
    def backward(ctx, dw):
        v, g, inv = ctx.saved_tensors
        dw = _c(dw)
        rows = v.shape[0]
        cols = v.numel() // rows
        dv = torch.empty_like(v)
        dg = torch.empty_like(g)
        _lib.check(_lib.load().tdvc_weight_norm_bwd(_p(dw), _p(v), _p(g), _p(inv), _p(dv), _p(dg), rows, cols, _st()),
                   "weight_norm_bwd")
        return dv, dg


def _wn_multi_forward(t):
    lib = _lib.load()
    flat_w = torch.empty(t["w_elems"], device=t["dev"], dtype=torch.float32)
    flat_inv = torch.empty(t["total_rows"], device=t["dev"], dtype=torch.float32)
    _lib.check(lib.tdvc_weight_norm_fwd_multi(_p(t["table"]), _p(t["rows_dev"]), t["n"], t["total_rows"], _p(flat_w),
                                              _p(flat_inv), _st()), "weight_norm_fwd_multi")
    return flat_w, flat_inv


# called with every parameter whose .grad _WeightNormMulti.backward has just written (the engine's post-accumulate-grad hooks
# do not see those deposits); weakref.WeakMethod entries, tdvc.dp.BucketedReducer listens here
grad_deposit_listeners = []


class _WeightNormMulti(torch.autograd.Function):
    """w_j = g_j * v_j / ||v_j|| for every weight of a step scope: one launch forward (all w_j are views of one flat
    buffer), one launch backward (tdvc_weight_norm_bwd_multi) once every dL/dw_j has arrived.

    The parameters enter as plain tensors next to a per-scope `anchor` that carries the autograd edge, and the backward
    deposits dL/dv_j, dL/dg_j into the parameters' .grad itself (set, or added in place when a gradient is already there --
    what AccumulateGrad does).  Routing ~500 parameter gradients of one node through the engine's AccumulateGrad nodes
    made the engine synchronise the legacy default stream with the capturing stream under CUDA-graph capture (fp32 mode,
    measured: cudaErrorStreamCaptureImplicit).  The engine's post-accumulate-grad hooks do not fire for these parameters;
    `grad_deposit_listeners` are called instead."""

    @staticmethod
    def forward(ctx, t, live, anchor):
        n = t["n"]
        vs, gs = [v for v, _ in live], [g for _, g in live]
        flat_w, _ = _wn_multi_forward(t)
        ws = [flat_w[t["w_off"][j]:t["w_off"][j] + vs[j].numel()].view_as(vs[j]) for j in range(n)]
        ctx.t = t
        ctx.params = (vs, gs)
        return tuple(ws)

    @staticmethod
    def backward(ctx, *dws):
        t = ctx.t
        n = t["n"]
        vs, gs = ctx.params
        lib = _lib.load()
        dev = t["dev"]
        flat_dw = torch.empty(t["w_elems"], device=dev, dtype=torch.float32)
        flat_dv = torch.empty(t["w_elems"], device=dev, dtype=torch.float32)
        flat_dg = torch.empty(t["total_rows"], device=dev, dtype=torch.float32)
        src, dst, missing = [], [], []
        for j in range(n):
            view = flat_dw[t["w_off"][j]:t["w_off"][j] + vs[j].numel()].view_as(vs[j])
            if dws[j] is None:
                missing.append(view)
            else:
                src.append(dws[j])
                dst.append(view)
        if missing:
            torch._foreach_zero_(missing)
        if src:
            torch._foreach_copy_(dst, src)
        _lib.check(lib.tdvc_weight_norm_bwd_multi(_p(t["btable"]), _p(t["rows_dev"]), n, t["total_rows"], _p(flat_dw), _p(flat_dv),
                                                  _p(flat_dg), _st()), "weight_norm_bwd_multi")
        with torch.no_grad():
            for j in range(n):
                if dws[j] is None:
                    continue
                v, g = vs[j], gs[j]
                if v.requires_grad:
                    dv = flat_dv[t["w_off"][j]:t["w_off"][j] + v.numel()].view_as(v)
                    if v.grad is None:
                        v.grad = dv
                    else:
                        v.grad.add_(dv)
                if g.requires_grad:
                    dg = flat_dg[t["row_start"][j]:t["row_start"][j] + g.numel()].view_as(g)
                    if g.grad is None:
                        g.grad = dg
                    else:
                        g.grad.add_(dg)
            for ref in list(grad_deposit_listeners):
                fn = ref()
                if fn is None:                 # its reducer is gone
                    grad_deposit_listeners.remove(ref)
                    continue
                for j in range(n):
                    if dws[j] is not None:
                        for p in (vs[j], gs[j]):
                            if p.requires_grad:
                                fn(p)
        return None, None, None


class _Scope:
    """State of one open step scope: what was normalised / packed inside it."""

    def __init__(self, tag):
        self.tag = tag
        self.wn = {}                # weight_norm() results of this scope
        self.wp = {}                # packed operands / _step_cached() results of this scope
        self.pre_w = {}             # (v ptr, g ptr) -> (w, inv, v version, g version) from the batched launch
        self.pre_p = {}             # (w ptr, rows_p, cols_p, flip) -> packed operand from the batched launch
        self.cur_w = {}             # w ptr -> weight index in this tag's plan


class _StepCache:
    """Inside `with ops.step_cache(tag):` a weight that has not changed is normalised (and packed to bf16) once and
    reused by every pass of the scope.  Reuse is through autograd (the cached tensor carries its grad_fn), so gradients
    from all passes accumulate into one weight-norm backward.  Entries are keyed on the parameter versions and dropped at
    scope exit.

    Scopes nest (a stack): lookups search from the innermost scope outwards, new entries go to the innermost one.  The
    training iteration keeps a generator scope open from the generator passes to the end of the G backward and opens a
    scope for the D step and one for the G step's discriminator passes inside it -- D's weights change between the two.

    Batching: the (v, g) pairs and the bf16 packs a scope asked for are remembered per tag (the PLAN); from the next
    scope of that tag on, entering it normalises every planned weight with ONE multi-tensor launch and packs every
    planned operand with ONE more, instead of hundreds of small launches per step.  `_WeightNorm.forward` / `_pack_w`
    then pick their result up from the flat buffers.  TDVC_NO_WN_BATCH=1 disables it."""

    def __init__(self):
        self.stack = []
        self.plans = {}             # tag -> dict(plan_w, plan_w_idx, plan_p, plan_p_idx, tables)
        self._next_tag = "default"

    @property
    def depth(self):
        return len(self.stack)

    def __call__(self, tag="default"):
        self._next_tag = tag
        return self

    def _plan(self, tag):
        pl = self.plans.get(tag)
        if pl is None:
            # plan_w [(weakref v, weakref g)], plan_w_idx (v ptr, g ptr) -> index, plan_p [(weight index, rows_p, cols_p,
            # flip)], plan_p_idx
            pl = self.plans[tag] = dict(plan_w=[], plan_w_idx={}, plan_p=[], plan_p_idx={}, tables=None)
        return pl

    def __enter__(self):
        tag, self._next_tag = self._next_tag, "default"
        if os.environ.get("TDVC_NO_STEP_CACHE") != "1":
            sc = _Scope(tag)
            self.stack.append(sc)
            self._prelaunch(sc)
        return self

    def __exit__(self, *exc):
        if os.environ.get("TDVC_NO_STEP_CACHE") != "1" and self.stack:
            self.stack.pop()
        return False

    # ---- scope-stack lookups
    def get_wn(self, key):
        for sc in reversed(self.stack):
            hit = sc.wn.get(key)
            if hit is not None:
                return hit
        return None

    def put_wn(self, key, val):
        self.stack[-1].wn[key] = val

    def get_wp(self, key):
        for sc in reversed(self.stack):
            hit = sc.wp.get(key)
            if hit is not None:
                return hit
        return None

    def put_wp(self, key, val):
        self.stack[-1].wp[key] = val

    # ---- plan bookkeeping
    def note_w(self, v, g, w):
        if not self.stack or not (v.is_leaf and g.is_leaf):
            return
        sc = self.stack[-1]
        pl = self._plan(sc.tag)
        key = (v.data_ptr(), g.data_ptr())
        idx = pl["plan_w_idx"].get(key)
        if idx is None:
            idx = len(pl["plan_w"])
            pl["plan_w"].append((weakref.ref(v), weakref.ref(g)))
            pl["plan_w_idx"][key] = idx
            pl["tables"] = None
        sc.cur_w[w.data_ptr()] = idx

    def note_p(self, w, rows_p, cols_p, flip):
        # the pack joins the plan of the scope that normalised w (the generator's weights are packed for their data
        # gradients while the G-step scope is the innermost one)
        for sc in reversed(self.stack):
            idx = sc.cur_w.get(w.data_ptr())
            if idx is None:
                continue
            pl = self._plan(sc.tag)
            key = (idx, rows_p, cols_p, bool(flip))
            if key not in pl["plan_p_idx"]:
                pl["plan_p_idx"][key] = len(pl["plan_p"])
                pl["plan_p"].append(key)
                pl["tables"] = None
            return

    def lookup_w(self, v, g):
        for sc in reversed(self.stack):
            hit = sc.pre_w.get((v.data_ptr(), g.data_ptr()))
            if hit is not None:
                if hit[2] != v._version or hit[3] != g._version:
                    return None
                return hit[0], hit[1]
        return None

    def lookup_p(self, w, rows_p, cols_p, flip):
        key = (w.data_ptr(), rows_p, cols_p, bool(flip))
        for sc in reversed(self.stack):
            hit = sc.pre_p.get(key)
            if hit is not None:
                return hit
        return None

    # ---- the batched launches
    @staticmethod
    def _build_tables(pl, live, dev):
        al = lambda n, a: (n + a - 1) // a * a
        rows, tab, w_off, off, r = [0], [], [], 0, 0
        for v, g in live:
            nr = v.shape[0]
            cols = v.numel() // nr
            tab += [v.data_ptr(), g.data_ptr(), off, cols]
            w_off.append(off)
            off += al(v.numel(), 64)
            r += nr
            rows.append(r)
        jobs, p_off, poff = [], [], 0
        for (idx, rows_p, cols_p, flip) in pl["plan_p"]:
            v = live[idx][0]
            Cout, Cin, K = v.shape
            jobs += [w_off[idx], poff, Cout, Cin, K, rows_p, cols_p, int(flip)]
            p_off.append(poff)
            poff += al(K * rows_p * cols_p, 64)
        max_job = max([jobs[8 * i + 4] * jobs[8 * i + 5] * jobs[8 * i + 6] for i in range(len(jobs) // 8)] or [1])
        # backward table (tdvc_weight_norm_bwd_multi): dw_j and dv_j sit at w_j's offset of their own flat buffers, dg_j at
        # its first global row
        btab = []
        for j, (v, g) in enumerate(live):
            btab += [v.data_ptr(), g.data_ptr(), w_off[j], w_off[j], rows[j], v.numel() // v.shape[0]]
        t = dict(sig=tuple(tab[0::4]) + tuple(tab[1::4]), dev=dev, n=len(live), total_rows=r, w_elems=off, p_elems=poff,
                 btable=torch.tensor(btab, dtype=torch.int64).to(dev),
                 blocks_per_job=max(1, min(256, -(-max_job // (256 * 16)))),      # <= 16 outputs per thread in the largest job
                 w_off=w_off, row_start=rows, p_off=p_off,
                 table=torch.tensor(tab, dtype=torch.int64).to(dev), rows_dev=torch.tensor(rows, dtype=torch.int32).to(dev),
                 jobs=torch.tensor(jobs, dtype=torch.int64).to(dev) if jobs else None)
        return t

    def _prelaunch(self, sc):
        pl = self._plan(sc.tag)
        if not pl["plan_w"] or os.environ.get("TDVC_NO_WN_BATCH") == "1" or branch_streams_enabled() or not torch.cuda.is_available():
            return
        live = [(rv(), rg()) for rv, rg in pl["plan_w"]]
        if any(v is None or g is None for v, g in live):          # a module went away: start the plan over
            self.plans.pop(sc.tag, None)
            return
        dev = live[0][0].device
        if any(v.device != dev or not v.is_contiguous() or not g.is_contiguous() for v, g in live):
            return
        sig = tuple(v.data_ptr() for v, _ in live) + tuple(g.data_ptr() for _, g in live)
        if pl["tables"] is None or pl["tables"]["sig"] != sig or pl["tables"]["dev"] != dev:
            if torch.cuda.is_current_stream_capturing():
                return                                           # host->device table upload is not capturable
            pl["tables"] = self._build_tables(pl, live, dev)
        t = pl["tables"]
        lib = _lib.load()
        need_grad = torch.is_grad_enabled() and any(v.requires_grad or g.requires_grad for v, g in live)
        if need_grad:
            # ONE autograd node for every weight of the scope: its backward gathers the weight gradients and runs a single
            # multi-tensor launch instead of one wn_bwd_k per weight
            anchor = torch.zeros(1, device=dev, dtype=torch.float32, requires_grad=True)
            ws = list(_WeightNormMulti.apply(t, list(live), anchor))
            flat_w = ws[0]               # the first weight sits at offset 0 of the flat buffer: its address is the buffer's
        else:
            flat_w, _ = _wn_multi_forward(t)
            ws = [flat_w[t["w_off"][j]:t["w_off"][j] + v.numel()].view_as(v) for j, (v, g) in enumerate(live)]
        for j, (v, g) in enumerate(live):
            sc.pre_w[(v.data_ptr(), g.data_ptr())] = (ws[j], None, v._version, g._version)
            sc.cur_w[ws[j].data_ptr()] = j
        if t["jobs"] is not None:
            flat_wp = torch.empty(t["p_elems"], device=dev, dtype=torch.bfloat16)
            _lib.check(lib.tdvc_pack_weight_bf16_multi(_p(t["jobs"]), len(pl["plan_p"]), t["blocks_per_job"], _p(flat_w),
                                                       _p(flat_wp), _st()),
                       "pack_weight_bf16_multi")
            for i, (idx, rows_p, cols_p, flip) in enumerate(pl["plan_p"]):
                K = live[idx][0].shape[2]
                wp = flat_wp[t["p_off"][i]:t["p_off"][i] + K * rows_p * cols_p].view(K, rows_p, cols_p)
                sc.pre_p[(ws[idx].data_ptr(), rows_p, cols_p, flip)] = wp


_step_cache = _StepCache()


def step_cache(tag: str = "default") -> _StepCache:
    return _step_cache(tag)


def inference_cache() -> _StepCache:
    """`with torch.no_grad(), ops.inference_cache():` -- weight norm folded and the bf16 operands packed ONCE for every
    forward inside the block (generate_with_target.py:105-120 loads a checkpoint and then only runs G.forward): the first
    call normalises and packs, later calls -- and a CUDA graph captured inside the block -- launch neither.  The weights
    must not change while the block is open."""
    return _step_cache("inference")


def weight_norm(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """w = g * v / ||v|| (norm over all dims but 0) -- torch.nn.utils.weight_norm's per-forward recompute."""
    if _step_cache.depth > 0:
        pre = _step_cache.lookup_w(v, g)
        if pre is not None and pre[1] is None:        # normalised by the scope's batched launch (carries the scope's grad_fn)
            return pre[0]
        key = (v.data_ptr(), v._version, g.data_ptr(), g._version, torch.is_grad_enabled() and (v.requires_grad or g.requires_grad))
        hit = _step_cache.get_wn(key)
        if hit is None:
            hit = (_WeightNorm.apply(v, g), torch.cuda.current_stream().cuda_stream)
            _step_cache.put_wn(key, hit)
        elif os.environ.get("TDVC_DEBUG_STREAMS") == "1" and hit[1] != torch.cuda.current_stream().cuda_stream:
            print(f"[tdvc] weight {tuple(v.shape)} normalised on stream {hit[1]:#x}, reused on {torch.cuda.current_stream().cuda_stream:#x}",
                  flush=True)
        return hit[0]
    return _WeightNorm.apply(v, g)


# ----------------------------------------------------------------------------- conv1d

class _Conv1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, residual, stride, pad, dilation, groups, pad_mode, in_slope, out_act, out_slope):
        _req(x, w, bias, residual)
        x, w, bias, residual = _c(x), _c(w), _c(bias), _c(residual)
        B, Cin, Tin = x.shape
        Cout, cin_g, K = w.shape
        if cin_g * groups != Cin:
            raise RuntimeError(f"conv1d: weight {tuple(w.shape)} does not match input {tuple(x.shape)} groups={groups}")
        Tout = (Tin + 2 * pad - dilation * (K - 1) - 1) // stride + 1
        if Tout <= 0:
            raise RuntimeError(f"conv1d: input length {Tin} too short for kernel {K} (dilation {dilation})")
        if pad_mode == PAD_REFLECT and pad >= Tin:
            raise RuntimeError(f"conv1d: reflect padding {pad} must be smaller than the input length {Tin}")
        g = _geom(B, Cin, Tin, Cout, Tout, K, stride, pad, dilation, groups, pad_mode, in_slope, out_act, out_slope)
        y = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
        if residual is not None and residual.shape != y.shape:
            raise RuntimeError("conv1d: residual shape mismatch")
        _lib.check(_lib.load().tdvc_conv1d_fwd(C.byref(g), _p(x), _p(w), _p(bias), _p(residual), _p(y), _st()), "conv1d_fwd")
        ctx.geom = g
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, w, y if out_act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        g = ctx.geom
        lib = _lib.load()
        dy = _c(dy)
        if g.out_act != ACT_NONE:
            dz = torch.empty_like(dy)
            _lib.check(lib.tdvc_act_bwd_from_output(_p(dy), _p(y), _p(dz), dy.numel(), g.out_act, g.out_slope, _st()),
                       "act_bwd")
            dy = dz
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            n_ws = lib.tdvc_conv1d_bwd_data_ws(C.byref(g))
            ws = torch.empty(n_ws, device=x.device, dtype=torch.float32) if n_ws > 0 else None
            _lib.check(lib.tdvc_conv1d_bwd_data(C.byref(g), _p(dy), _p(w), _p(x), _p(dx), _p(ws), _st()), "conv1d_bwd_data")
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] or need_b:
            dw = torch.empty_like(w)
            db = torch.empty(g.Cout, device=x.device, dtype=torch.float32) if need_b else None
            _lib.check(lib.tdvc_conv1d_bwd_weight(C.byref(g), _p(dy), _p(x), _p(dw), _p(db), _st()), "conv1d_bwd_weight")
            if not ctx.needs_input_grad[1]:
                dw = None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None, None, None, None, None, None, None


def conv1d(x, weight, bias=None, *, stride=1, padding=0, dilation=1, groups=1, reflect=False,
           in_slope=1.0, out_act=None, out_slope=0.2, residual=None):
    """act(conv1d(leaky_relu(pad(x), in_slope), weight) + bias + residual); nn.Conv1d semantics
    (padding_mode 'zeros' or 'reflect')."""
    stride, padding, dilation, groups = int(stride), int(padding), int(dilation), int(groups)
    is_reflect = bool(reflect and padding > 0)
    if (_frame_conv_eligible(x.shape[1], weight.shape[0], weight.shape[2], stride, groups, dilation, is_reflect)
            and residual is None and tc_eligible(stride * x.shape[1], weight.shape[0], 1, 1)
            and x.shape[2] + 2 * padding >= weight.shape[2]):
        return _strided_conv_as_frames(x, weight, bias, stride, padding, in_slope, out_act, out_slope)
    if residual is None and in_slope == 1.0 and x.is_cuda:
        plan = _grouped_frame_plan(x.shape[1], weight.shape[0], weight.shape[2], stride, groups, dilation, is_reflect)
        if plan is not None and x.shape[2] + 2 * padding >= weight.shape[2]:
            return _GroupedFrameConvTC.apply(x, weight, bias, stride, padding, groups, _ACT[out_act], float(out_slope), plan)
    if _PRECISION == "bf16" and tc_eligible(x.shape[1], weight.shape[0], stride, groups):
        return _Conv1dTC.apply(x, weight, bias, residual, padding, dilation,
                               PAD_REFLECT if is_reflect else PAD_ZEROS, float(in_slope), _ACT[out_act],
                               float(out_slope))
    return _Conv1d.apply(x, weight, bias, residual, stride, padding, dilation, groups,
                         PAD_REFLECT if is_reflect else PAD_ZEROS, float(in_slope), _ACT[out_act],
                         float(out_slope))


def linear(x, weight, bias=None):
    """F.linear on [B, Cin] as a length-1 convolution (Generator.embedding, CIN.embedding)."""
    return conv1d(x.unsqueeze(2), weight.unsqueeze(2), bias).squeeze(2)


# ----------------------------------------------------------------------------- conv transpose

class _ConvTranspose1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, output_padding):
        _req(x, w, bias)
        x, w, bias = _c(x), _c(w), _c(bias)
        B, Cin, Tin = x.shape
        Cin_w, Cout, K = w.shape
        if Cin_w != Cin:
            raise RuntimeError(f"conv_transpose1d: weight {tuple(w.shape)} does not match input {tuple(x.shape)}")
        Tout = (Tin - 1) * stride - 2 * pad + (K - 1) + output_padding + 1
        g = _geom(B, Cin, Tin, Cout, Tout, K, stride, pad, 1, 1, PAD_ZEROS, 1.0, ACT_NONE, 1.0)
        y = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_conv_transpose1d_fwd(C.byref(g), _p(x), _p(w), _p(bias), _p(y), _st()),
                   "conv_transpose1d_fwd")
        ctx.geom = g
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        g = ctx.geom
        lib = _lib.load()
        dy = _c(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.check(lib.tdvc_conv_transpose1d_bwd_data(C.byref(g), _p(dy), _p(w), _p(dx), _st()), "convT_bwd_data")
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] or need_b:
            dw = torch.empty_like(w)
            db = torch.empty(g.Cout, device=x.device, dtype=torch.float32) if need_b else None
            _lib.check(lib.tdvc_conv_transpose1d_bwd_weight(C.byref(g), _p(dy), _p(x), _p(dw), _p(db), _st()),
                       "convT_bwd_weight")
            if not ctx.needs_input_grad[1]:
                dw = None
        return dx, dw, db, None, None, None


def conv_transpose1d(x, weight, bias=None, *, stride=1, padding=0, output_padding=0):
    if (_frame_conv_eligible(weight.shape[0], weight.shape[1], weight.shape[2], int(stride), 1, 1, False)
            and weight.shape[2] % int(stride) == 0 and int(output_padding) == 0 and tc_eligible(weight.shape[0], int(stride) * weight.shape[1], 1, 1)):
        return _conv_transpose_as_frames(x, weight, bias, int(stride), int(padding))
    return _ConvTranspose1d.apply(x, weight, bias, int(stride), int(padding), int(output_padding))


# ----------------------------------------------------------------------------- elementwise

class _LeakyReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope):
        _req(x)
        x = _c(x)
        y = torch.empty_like(x)
        _lib.check(_lib.load().tdvc_leaky_relu_fwd(_p(x), _p(y), x.numel(), slope, _st()), "leaky_relu")
        ctx.slope = slope
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _lib.check(_lib.load().tdvc_act_bwd_from_output(_p(dy), _p(y), _p(dx), dy.numel(), ACT_LRELU, ctx.slope, _st()),
                   "leaky_relu_bwd")
        return dx, None


def leaky_relu(x, slope=0.2):
    return _LeakyReLU.apply(x, float(slope))


class _Gate(torch.autograd.Function):
    """tanh(a[:, :H] + g[:, :H]) * sigmoid(a[:, H:] + g[:, H:])  (model/ssl_encoder.py:7-14)."""

    @staticmethod
    def forward(ctx, a, g):
        _req(a, g)
        a, g = _c(a), _c(g)
        B, C2, T = a.shape
        if C2 % 2 or (g is not None and g.shape != a.shape):
            raise RuntimeError(f"gate: inputs {tuple(a.shape)}, {None if g is None else tuple(g.shape)} must be [B, 2H, T]")
        y = torch.empty(B, C2 // 2, T, device=a.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_gate_fwd(_p(a), _p(g), _p(y), B, C2 // 2, T, _st()), "gate_fwd")
        ctx.save_for_backward(a, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, g = ctx.saved_tensors
        dy = _c(dy)
        B, C2, T = a.shape
        da = torch.empty_like(a)
        _lib.check(_lib.load().tdvc_gate_bwd(_p(dy), _p(a), _p(g), _p(da), B, C2 // 2, T, _st()), "gate_bwd")
        return da, (da.clone() if (g is not None and ctx.needs_input_grad[1]) else None)


def gated_tanh_sigmoid(a, g=None):
    return _Gate.apply(a, g)


class _Film(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, gb):
        _req(h, gb)
        h, gb = _c(h), _c(gb)
        B, Cc, T = h.shape
        if gb.shape != (B, 2 * Cc, T):
            raise RuntimeError(f"film: gamma/beta tensor {tuple(gb.shape)} does not match {tuple(h.shape)}")
        y = torch.empty_like(h)
        _lib.check(_lib.load().tdvc_film_fwd(_p(h), _p(gb), _p(y), B, Cc, T, _st()), "film_fwd")
        ctx.save_for_backward(h, gb)
        return y

    @staticmethod
    def backward(ctx, dy):
        h, gb = ctx.saved_tensors
        dy = _c(dy)
        B, Cc, T = h.shape
        dh = torch.empty_like(h)
        dgb = torch.empty_like(gb)
        _lib.check(_lib.load().tdvc_film_bwd(_p(dy), _p(h), _p(gb), _p(dh), _p(dgb), B, Cc, T, _st()), "film_bwd")
        return dh, dgb


def film(h, gb):
    """h*(1+gamma)+beta with (gamma, beta) = gb.chunk(2, dim=1)  (model/generator.py:104-107)."""
    return _Film.apply(h, gb)


class _Add3Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, a, b, c):
        _req(a, b, c)
        a, b, c = _c(a), _c(b), _c(c)
        y = torch.empty_like(a)
        _lib.check(_lib.load().tdvc_add3_scale(_p(a), _p(b), _p(c), _p(y), a.numel(), alpha, _st()), "add3_scale")
        ctx.alpha = alpha
        ctx.n = 1 + (b is not None) + (c is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        lib = _lib.load()

        def scaled():
            d = torch.empty_like(dy)
            _lib.check(lib.tdvc_add3_scale(_p(dy), None, None, _p(d), dy.numel(), ctx.alpha, _st()), "add3_scale_bwd")
            return d
        if _BRANCH_STREAMS:
            # the summands' producers may live on different streams: autograd accumulates into a gradient in place once
            # its CPU-side refcount drops to one, which would race with kernels another stream has only queued if the
            # three branches were handed the same tensor -- give each its own
            return (None,) + tuple(scaled() if i < ctx.n else None for i in range(3))
        d = scaled()
        return (None, d) + tuple(d if i < ctx.n - 1 else None for i in range(2))


def add_scale(a, b=None, c=None, alpha=1.0):
    """alpha * (a + b + c)"""
    return _Add3Scale.apply(float(alpha), a, b, c)


class _GradReverse(torch.autograd.Function):
    """Identity whose gradient is negated (the adversarial latent classifier's input, model/grad_rev.py:3-10: the
    reference's backward hard-codes lamb = 1 whatever the layer was built with)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, dy):
        return add_scale(dy, alpha=-1.0)


def grad_reverse(x):
    return _GradReverse.apply(x)


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _req(x)
        x = _c(x)
        B, Cc, T = x.shape
        y = torch.empty_like(x)
        inv = torch.empty(B, T, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_l2norm_fwd(_p(x), _p(y), _p(inv), B, Cc, T, _st()), "l2norm_fwd")
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = _c(dy)
        B, Cc, T = y.shape
        dx = torch.empty_like(y)
        _lib.check(_lib.load().tdvc_l2norm_bwd(_p(dy), _p(y), _p(inv), _p(dx), B, Cc, T, _st()), "l2norm_bwd")
        return dx


def l2_normalize(x):
    """F.normalize(x, dim=1) for [B, C, T]."""
    return _L2Norm.apply(x)


class _CondConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c, e, c_first):
        _req(c, e)
        c, e = _c(c), _c(e)
        B, Cc = c.shape
        Be, Ce, T = e.shape
        if B != Be:
            raise RuntimeError("cond_concat: batch mismatch")
        out = torch.empty(B, Cc + Ce, T, device=e.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_cond_concat_fwd(_p(c), _p(e), _p(out), B, Cc, Ce, T, c_first, _st()), "cond_concat_fwd")
        ctx.dims = (B, Cc, Ce, T, c_first)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, Cc, Ce, T, c_first = ctx.dims
        dout = _c(dout)
        dc = torch.empty(B, Cc, device=dout.device, dtype=torch.float32)
        de = torch.empty(B, Ce, T, device=dout.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.load().tdvc_cond_concat_bwd(_p(dout), _p(dc), _p(de), B, Cc, Ce, T, c_first, _st()), "cond_concat_bwd")
        return dc, de, None


def cond_concat(c, e):
    """torch.cat([c.unsqueeze(2).repeat(1, 1, T), e], dim=1)  (model/generator.py:387-399)."""
    return _CondConcat.apply(c, e, 1)


def cond_concat_front(x, c):
    """torch.cat([x, c.unsqueeze(2).repeat(1, 1, T)], dim=1)  (model/generator.py:260-261,380-381)."""
    return _CondConcat.apply(c, x, 0)


def cat_channels_2d(a, b):
    """torch.cat([a, b], dim=1) of two [B, C] speaker codes (model/generator.py:498; glue, not arithmetic)."""
    return torch.cat([a, b], dim=1)


# ----------------------------------------------------------------------------- instance norm / CIN

class _CinApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gb, eps, out_slope):
        _req(x, gb)
        x, gb = _c(x), _c(gb)
        B, Cc, T = x.shape
        Tg = 1
        if gb is not None:
            if gb.dim() != 3 or gb.shape[0] != B or gb.shape[1] != 2 * Cc or gb.shape[2] not in (1, T):
                raise RuntimeError(f"cin: gamma/beta tensor {tuple(gb.shape)} does not match {tuple(x.shape)}")
            Tg = gb.shape[2]
        lib = _lib.load()
        mean = torch.empty(B * Cc, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _lib.check(lib.tdvc_instnorm_stats(_p(x), _p(mean), _p(rstd), B * Cc, T, eps, _st()), "instnorm_stats")
        y = torch.empty_like(x)
        _lib.check(lib.tdvc_cin_apply_fwd(_p(x), _p(mean), _p(rstd), _p(gb), Tg, _p(y), B, Cc, T, out_slope, _st()),
                   "cin_apply_fwd")
        ctx.out_slope, ctx.Tg = out_slope, Tg
        ctx.save_for_backward(x, mean, rstd, gb, y if out_slope != 1.0 else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gb, y = ctx.saved_tensors
        dy = _c(dy)
        B, Cc, T = x.shape
        dx = torch.empty_like(x)
        dgb = torch.empty_like(gb) if gb is not None else None
        _lib.check(_lib.load().tdvc_cin_apply_bwd(_p(dy), _p(x), _p(mean), _p(rstd), _p(gb), ctx.Tg, _p(y), _p(dx),
                                                  _p(dgb), B, Cc, T, ctx.out_slope, _st()), "cin_apply_bwd")
        return dx, dgb, None, None


def instance_norm(x, eps=1e-5, out_slope=1.0):
    """nn.InstanceNorm1d(affine=False) (+ optional fused LeakyReLU)."""
    return _CinApply.apply(x, None, float(eps), float(out_slope))


def cond_instance_norm(x, gb, eps=1e-5, out_slope=1.0):
    """(1 + gamma) * instance_norm(x) + beta with gb = [B, 2C, 1 or T]  (model/conditional_instance_norm.py:18-19)."""
    return _CinApply.apply(x, gb, float(eps), float(out_slope))


class _TimeMean(torch.autograd.Function):
    """x[B,C,T] -> mean over T: F.avg_pool1d(x, x.size(2)).squeeze(2) (model/latent_classifier.py:36).  Forward is the
    row-mean half of the instance-norm statistics kernel, backward a broadcast."""

    @staticmethod
    def forward(ctx, x):
        _req(x)
        x = _c(x)
        B, Cc, T = x.shape
        mean = torch.empty(B * Cc, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        _lib.check(_lib.load().tdvc_instnorm_stats(_p(x), _p(mean), _p(rstd), B * Cc, T, 1.0, _st()), "time_mean")
        ctx.dims = (B, Cc, T)
        return mean.view(B, Cc)

    @staticmethod
    def backward(ctx, dy):
        B, Cc, T = ctx.dims
        dy = _c(dy)
        lib = _lib.load()
        d = torch.empty_like(dy)
        _lib.check(lib.tdvc_add3_scale(_p(dy), None, None, _p(d), dy.numel(), 1.0 / T, _st()), "time_mean_bwd scale")
        dx = torch.empty(B, Cc, T, device=dy.device, dtype=torch.float32)
        _lib.check(lib.tdvc_cond_concat_fwd(_p(d), None, _p(dx), B, Cc, 0, T, 1, _st()), "time_mean_bwd broadcast")
        return dx


def time_mean(x):
    return _TimeMean.apply(x)


# ----------------------------------------------------------------------------- pooling / gather

class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _req(x)
        x = _c(x)
        B, Cc, Tin = x.shape
        Tout = (Tin + 2 - 4) // 2 + 1
        y = torch.empty(B, Cc, Tout, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_avgpool4s2_fwd(_p(x), _p(y), B * Cc, Tin, Tout, _st()), "avgpool_fwd")
        ctx.dims = (B, Cc, Tin, Tout)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, Cc, Tin, Tout = ctx.dims
        dy = _c(dy)
        dx = torch.empty(B, Cc, Tin, device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_avgpool4s2_bwd(_p(dy), _p(dx), B * Cc, Tin, Tout, _st()), "avgpool_bwd")
        return dx


def avg_pool_4_2_1(x):
    """nn.AvgPool1d(kernel_size=4, stride=2, padding=1, count_include_pad=False)."""
    return _AvgPool.apply(x)


class _SelectChannel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label):
        _req(x)
        x = _c(x)
        if not label.is_cuda or label.dtype != torch.int64:
            raise RuntimeError("select_channel: label must be a CUDA int64 tensor")
        label = label.contiguous()
        B, Cc, T = x.shape
        y = torch.empty(B, 1, T, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_select_channel_fwd(_p(x), _p(label), _p(y), B, Cc, T, _st()), "select_fwd")
        ctx.dims = (B, Cc, T)
        ctx.save_for_backward(label)
        return y

    @staticmethod
    def backward(ctx, dy):
        (label,) = ctx.saved_tensors
        B, Cc, T = ctx.dims
        dy = _c(dy)
        dx = torch.empty(B, Cc, T, device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_select_channel_bwd(_p(dy), _p(label), _p(dx), B, Cc, T, _st()), "select_bwd")
        return dx, None


def select_channel(x, label):
    """x.gather(1, label.view(-1,1,1).expand(-1,1,T))  (model/discriminator.py:49-51)."""
    return _SelectChannel.apply(x, label)


class _Conv1dSelect(torch.autograd.Function):
    """conv1d(x, w, padding=(K-1)/2).gather(1, label) computing only the gathered row (tdvc_conv1d_select_*)."""

    @staticmethod
    def forward(ctx, x, w, label):
        _req(x, w)
        x, w = _c(x), _c(w)
        if not label.is_cuda or label.dtype != torch.int64:
            raise RuntimeError("conv1d_select: label must be a CUDA int64 tensor")
        label = label.contiguous()
        B, Cc, T = x.shape
        NC, Cw, K = w.shape
        if Cw != Cc or K % 2 != 1 or K > 7 or label.numel() != B:
            raise RuntimeError(f"conv1d_select: weight {tuple(w.shape)} / label {tuple(label.shape)} do not match input "
                               f"{tuple(x.shape)} (odd K <= 7)")
        y = torch.empty(B, 1, T, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_conv1d_select_fwd(_p(x), _p(w), _p(label), _p(y), B, Cc, T, NC, K, (K - 1) // 2, _st()),
                   "conv1d_select_fwd")
        ctx.save_for_backward(x, w, label)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, label = ctx.saved_tensors
        B, Cc, T = x.shape
        NC, _, K = w.shape
        dy = _c(dy)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None       # accumulated into: samples may share a label
        if dx is not None or dw is not None:
            _lib.check(_lib.load().tdvc_conv1d_select_bwd(_p(dy), _p(x), _p(w), _p(label), _p(dx), _p(dw), B, Cc, T, NC, K,
                                                          (K - 1) // 2, _st()), "conv1d_select_bwd")
        return dx, dw, None


_FUSED_SELECT = os.environ.get("TDVC_FUSED_SELECT", "1") != "0"   # development switch: 0 = full output conv, then the gather


def conv1d_select(x, weight, label):
    """The discriminator's output layer and label gather (model/discriminator.py:36,49-51) as one op:
    conv1d(x, weight, padding=(K-1)//2) restricted to output row label[b] of sample b -> [B, 1, T].  Only the selected
    rows are computed (1 / num_classes of the layer's multiply-adds), in fp32 in every precision mode."""
    if not _FUSED_SELECT:
        return select_channel(conv1d(x, weight, None, padding=(weight.shape[2] - 1) // 2), label)
    return _Conv1dSelect.apply(x, weight, label)


# ----------------------------------------------------------------------------- losses

class _SqErrConstMean(torch.autograd.Function):
    """sum_i weight_i * mean((a_i - target)^2) over a list of tensors: the LSGAN terms of
    train.py:273-281,327-331 in one reduction buffer."""

    @staticmethod
    def forward(ctx, target, *tensors):
        _req(*tensors)
        tensors = [_c(t) for t in tensors]
        out = torch.zeros(1, device=tensors[0].device, dtype=torch.float32)
        lib = _lib.load()
        for t in tensors:
            _lib.check(lib.tdvc_sq_err_const_sum(_p(t), target, 1.0 / t.numel(), _p(out), t.numel(), _st()), "sq_err_sum")
        ctx.target = target
        ctx.save_for_backward(*tensors)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        tensors = ctx.saved_tensors
        g = g.reshape(1).contiguous().float()
        lib = _lib.load()
        grads = []
        for i, t in enumerate(tensors):
            if not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            d = torch.empty_like(t)
            _lib.check(lib.tdvc_sq_err_const_bwd(_p(t), ctx.target, 1.0 / t.numel(), _p(g), _p(d), t.numel(), _st()),
                       "sq_err_bwd")
            grads.append(d)
        return (None, *grads)


def mse_to_const_sum(tensors: Sequence[torch.Tensor], target: float) -> torch.Tensor:
    """sum over tensors of F.mse_loss(t, full_like(t, target))."""
    return _SqErrConstMean.apply(float(target), *tensors)


class _L1MeanSum(torch.autograd.Function):
    """sum_i mean(|a_i - b_i|), b detached: util/losses.py:55-68 (all 30 feature maps, one scalar, one launch)."""

    @staticmethod
    def forward(ctx, n, *tensors):
        a, b = tensors[:n], tensors[n:]
        _req(*a, *b)
        a = [_c(t) for t in a]
        b = [_c(t) for t in b]
        out = torch.zeros(1, device=a[0].device, dtype=torch.float32)
        for u, v in zip(a, b):
            if u.shape != v.shape:
                raise RuntimeError(f"l1 loss: shape mismatch {tuple(u.shape)} vs {tuple(v.shape)}")
        _l1_multi([(u, v, None) for u, v in zip(a, b)], out=out)
        ctx.n = n
        ctx.save_for_backward(*a, *b)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        n = ctx.n
        a, b = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        g = g.reshape(1).contiguous().float()
        grads, pairs = [], []
        for i, (u, v) in enumerate(zip(a, b)):
            if not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            d = torch.empty_like(u)
            pairs.append((u, v, d))
            grads.append(d)
        if pairs:
            _l1_multi(pairs, gscale=g)
        return (None, *grads, *([None] * n))


class _Contrastive(torch.autograd.Function):
    """util/losses.py:70-116 as one kernel per direction (value and unit gradients computed together)."""

    @staticmethod
    def forward(ctx, X, Y, raw_X, raw_Y):
        _req(X, Y)
        X, Y = _c(X), _c(Y)
        B, Cc, T = X.shape
        N = raw_X.shape[-1]
        if Y.shape != X.shape or tuple(raw_X.shape) != (B, T, N) or tuple(raw_Y.shape) != (B, T, N):
            raise RuntimeError("contrastive_loss: shape mismatch")
        raw_X, raw_Y = raw_X.contiguous(), raw_Y.contiguous()
        lib = _lib.load()
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        buf = torch.zeros(1 + (2 * X.numel() if need else 0), device=X.device, dtype=torch.float32)
        loss = buf[:1]
        dX = buf[1:1 + X.numel()].view_as(X) if need else None
        dY = buf[1 + X.numel():].view_as(Y) if need else None
        scale = 1.0 / (2 * B * T)
        _lib.check(lib.tdvc_contrastive_dir(_p(X), _p(Y), _p(raw_X), _p(loss), _p(dX), _p(dY), B, Cc, T, N, scale, _st()),
                   "contrastive X->Y")
        _lib.check(lib.tdvc_contrastive_dir(_p(Y), _p(X), _p(raw_Y), _p(loss), _p(dY), _p(dX), B, Cc, T, N, scale, _st()),
                   "contrastive Y->X")
        if need:
            ctx.save_for_backward(dX, dY)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        dX, dY = ctx.saved_tensors
        g = float(1.0) if g is None else g
        lib = _lib.load()
        outs = []
        for d, need in ((dX, ctx.needs_input_grad[0]), (dY, ctx.needs_input_grad[1])):
            if not need:
                outs.append(None)
                continue
            o = d * g        # scalar upstream gradient (a device tensor): one small elementwise op
            outs.append(o)
        return outs[0], outs[1], None, None


def contrastive_loss(X, Y, raw_X, raw_Y):
    """InfoNCE over frames with in-utterance negatives given by the raw randint draws [B,T,N] (int64)."""
    return _Contrastive.apply(X, Y, raw_X, raw_Y)


def l1_mean_sum(sig: Sequence[torch.Tensor], ref: Sequence[torch.Tensor]) -> torch.Tensor:
    sig, ref = list(sig), [r.detach() for r in ref]
    return _L1MeanSum.apply(len(sig), *sig, *ref)


def _l1_multi(pairs, out=None, gscale=None):
    """pairs: [(a, b, da or None)] contiguous fp32 tensors of equal numel per pair; forward when `out` is given (adds
    sum_j mean|a_j - b_j| into it), backward when `gscale` is given (da_j = gscale / numel_j * sign(a_j - b_j))."""
    lib = _lib.load()
    for i in range(0, len(pairs), _lib.L1_MAX_JOBS):
        chunk = pairs[i:i + _lib.L1_MAX_JOBS]
        jobs = (_lib.L1Job * len(chunk))()
        for j, (a, b, da) in enumerate(chunk):
            jobs[j].a, jobs[j].b = a.data_ptr(), b.data_ptr()
            jobs[j].da = da.data_ptr() if da is not None else None
            jobs[j].n, jobs[j].scale = b.numel(), 1.0 / b.numel()
        if out is not None:
            _lib.check(lib.tdvc_abs_diff_sum_multi(jobs, len(chunk), _p(out), _st()), "abs_diff_sum_multi")
        else:
            _lib.check(lib.tdvc_abs_diff_bwd_multi(jobs, len(chunk), _p(gscale), _st()), "abs_diff_bwd_multi")


class _L1MeanSumRows(torch.autograd.Function):
    """sum_i mean(|a_i[row0:row0+n] - b_i|) where a_i holds several signals stacked along the batch (the batched
    discriminator call of the G step) and b_i is the detached reference of `n` rows: ONE reduction launch for all maps and
    ONE gradient launch.  The backward writes the gradient of the whole stacked tensor (zeros outside the rows) instead of
    leaving a slice for autograd to pad."""

    @staticmethod
    def forward(ctx, n, row0, nrows, *tensors):
        a, b = tensors[:n], tensors[n:]
        _req(*a, *b)
        a = [_c(t) for t in a]
        b = [_c(t) for t in b]
        out = torch.zeros(1, device=a[0].device, dtype=torch.float32)
        pairs = []
        for u, v in zip(a, b):
            if row0 < 0 or row0 + nrows > u.shape[0] or tuple(v.shape) != (nrows,) + tuple(u.shape[1:]):
                raise RuntimeError(f"l1 loss: rows [{row0}, {row0 + nrows}) of {tuple(u.shape)} vs {tuple(v.shape)}")
            pairs.append((u[row0:row0 + nrows], v, None))
        _l1_multi(pairs, out=out)
        ctx.cfg = (n, row0, nrows)
        ctx.save_for_backward(*a, *b)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        n, row0, nrows = ctx.cfg
        a, b = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        g = g.reshape(1).contiguous().float()
        grads, pairs = [], []
        for i, (u, v) in enumerate(zip(a, b)):
            if not ctx.needs_input_grad[3 + i]:
                grads.append(None)
                continue
            d = torch.empty_like(u)
            if row0 > 0:
                d[:row0].zero_()
            if row0 + nrows < u.shape[0]:
                d[row0 + nrows:].zero_()
            pairs.append((u[row0:row0 + nrows], v, d[row0:row0 + nrows]))
            grads.append(d)
        if pairs:
            _l1_multi(pairs, gscale=g)
        return (None, None, None, *grads, *([None] * n))


def l1_mean_sum_rows(sig: Sequence[torch.Tensor], row0: int, nrows: int, ref: Sequence[torch.Tensor]) -> torch.Tensor:
    sig, ref = list(sig), [r.detach() for r in ref]
    return _L1MeanSumRows.apply(len(sig), int(row0), int(nrows), *sig, *ref)


# ----------------------------------------------------------------------------- log-mel loss pieces (mel.cu)

class _StftFrames(torch.autograd.Function):
    """x[B, T] -> windowed frames F[(3*)n_fft, B*NF] with reflect padding n_fft/2 (torch.stft(center=True) framing)."""

    @staticmethod
    def forward(ctx, x, win, n_fft, hop, split):
        _req(x, win)
        x, win = _c(x), _c(win)
        B, T = x.shape
        pad = n_fft // 2
        NF = (T + 2 * pad - n_fft) // hop + 1
        if pad >= T:
            raise RuntimeError(f"stft: reflect padding {pad} must be smaller than the signal length {T}")
        F_ = torch.empty((3 if split else 1) * n_fft, B * NF, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_stft_frames_fwd(_p(x), _p(win), _p(F_), B, T, n_fft, hop, pad, NF, int(split), _st()), "stft_frames")
        ctx.cfg = (B, T, n_fft, hop, pad, NF, int(split))
        ctx.save_for_backward(win)
        return F_

    @staticmethod
    def backward(ctx, dF):
        (win,) = ctx.saved_tensors
        B, T, n_fft, hop, pad, NF, split = ctx.cfg
        dF = _c(dF)
        dx = torch.empty(B, T, device=dF.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_stft_frames_bwd(_p(dF), _p(win), _p(dx), B, T, n_fft, hop, pad, NF, split, _st()), "stft_frames_bwd")
        return dx, None, None, None, None


def stft_frames(x, win, n_fft, hop, split=False):
    return _StftFrames.apply(x, win, int(n_fft), int(hop), bool(split))


class _Power(torch.autograd.Function):
    """S[rows, cols] with real parts in rows [0, nfreq) and imaginary parts in rows [im_off, im_off + nfreq) -> |X|^2
    [nfreq, cols], or its bf16 high / low split [hi | lo | hi] as three row blocks (split)."""

    @staticmethod
    def forward(ctx, S, nfreq, im_off, split):
        _req(S)
        S = _c(S)
        rows, cols = S.shape
        P = torch.empty((3 if split else 1) * nfreq, cols, device=S.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_power_fwd(_p(S), _p(P), nfreq, im_off, cols, int(split), _st()), "power")
        ctx.cfg = (nfreq, im_off, int(split))
        ctx.save_for_backward(S)
        return P

    @staticmethod
    def backward(ctx, dP):
        (S,) = ctx.saved_tensors
        nfreq, im_off, split = ctx.cfg
        dP = _c(dP)
        dS = torch.empty_like(S)
        _lib.check(_lib.load().tdvc_power_bwd(_p(S), _p(dP), _p(dS), nfreq, im_off, S.shape[0], S.shape[1], split, _st()), "power_bwd")
        return dS, None, None, None


def power_spectrum(S, nfreq, im_off, split=False):
    return _Power.apply(S, int(nfreq), int(im_off), bool(split))


class _LogClamp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, floor_):
        _req(x)
        x = _c(x)
        y = torch.empty_like(x)
        _lib.check(_lib.load().tdvc_log_clamp_fwd(_p(x), _p(y), x.numel(), floor_, _st()), "log_clamp")
        ctx.floor_ = floor_
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        _lib.check(_lib.load().tdvc_log_clamp_bwd(_p(x), _p(dy), _p(dx), x.numel(), ctx.floor_, _st()), "log_clamp_bwd")
        return dx, None


def log_clamp(x, floor_=1e-5):
    """torch.log(torch.clamp(x, min=floor_))"""
    return _LogClamp.apply(x, float(floor_))


# ----------------------------------------------------------------------------- bf16 tensor-core conv path

def _ceil(a, m):
    return (a + m - 1) // m * m


def _cp(c):
    """channel padding of a packed (channels-last bf16) operand: the TMA box is 64 channels wide."""
    return 64 if c <= 64 else _ceil(c, 8)


def tc_eligible(Cin, Cout, stride, groups) -> bool:
    """Dense stride-1 convs with enough channels to fill an MMA tile go to tcgen05 in bf16 mode."""
    if stride != 1 or groups != 1 or Cin < 16 or Cout < 16:
        return False
    coutp = _ceil(Cout, 16)
    return coutp <= 256 or coutp % 128 == 0


class _PackCache:
    """The FiLM conditioning tensor feeds 9 cond_var.0 convs per stage: pack it once.  Entries hold a reference
    to the source tensor, so its storage cannot be recycled while the entry is alive."""

    def __init__(self, size=6):
        self.size, self.items = size, []

    def get(self, key, src):
        for k, s, v in self.items:
            if k == key and s.data_ptr() == src.data_ptr() and s._version == src._version:
                return v
        return None

    def put(self, key, src, val):
        self.items.insert(0, (key, src, val))
        del self.items[self.size:]

    def clear(self):
        self.items = []


_pack_cache = _PackCache()


def _pack_act(x, Cp, halo, pad_mode, slope, cache=True, chan_sum=None):
    B, Cc, T = x.shape
    key = (tuple(x.shape), Cp, halo, pad_mode, slope)
    if cache:
        hit = _pack_cache.get(key, x)
        if hit is not None:
            return hit
    xp = torch.empty(B, T + 2 * halo, Cp, device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().tdvc_pack_cl_bf16(_p(x), _p(xp), B, Cc, T, Cp, halo, pad_mode, slope, _p(chan_sum), 0, 0, -1,
                                             None, _st()), "pack_cl_bf16")
    if cache:
        _pack_cache.put(key, x, xp)
    return xp


# bf16 packs of weights that never change (the DFT basis and the mel filterbank of the spectral loss, 13 M elements):
# packed once per process instead of once per step.  The entry keeps the source alive, so its address stays unique.
_CONST_PACKS = {}


def mark_constant(w):
    """Declare a conv weight immutable for the life of the process: its tensor-core operand forms are cached."""
    w._tdvc_const = True
    return w


def _pack_w(w, rows_p, cols_p, transpose_flip):
    Cout, Cin, K = w.shape
    key = None
    if getattr(w, "_tdvc_const", False):
        ckey = (w.data_ptr(), tuple(w.shape), rows_p, cols_p, bool(transpose_flip))
        hit = _CONST_PACKS.get(ckey)
        if hit is None:
            wp = torch.empty(K, rows_p, cols_p, device=w.device, dtype=torch.bfloat16)
            coutp, cinp = (cols_p, rows_p) if transpose_flip else (rows_p, cols_p)
            _lib.check(_lib.load().tdvc_pack_weight_bf16(_p(w), _p(wp), Cout, Cin, K, coutp, cinp, int(transpose_flip), 0, 0, 0, 0,
                                                         _st()), "pack_weight_bf16")
            hit = _CONST_PACKS[ckey] = (w, wp)
        return hit[1]
    if _step_cache.depth > 0:
        pre = _step_cache.lookup_p(w, rows_p, cols_p, transpose_flip)
        if pre is not None:                        # packed by the scope's batched launch
            return pre
        key = (w.data_ptr(), w._version, tuple(w.shape), rows_p, cols_p, bool(transpose_flip))
        hit = _step_cache.get_wp(key)
        if hit is not None and hit[0]() is w:
            return hit[1]
        _step_cache.note_p(w, rows_p, cols_p, transpose_flip)
    wp = torch.empty(K, rows_p, cols_p, device=w.device, dtype=torch.bfloat16)
    coutp, cinp = (cols_p, rows_p) if transpose_flip else (rows_p, cols_p)
    _lib.check(_lib.load().tdvc_pack_weight_bf16(_p(w), _p(wp), Cout, Cin, K, coutp, cinp, int(transpose_flip), 0, 0, 0, 0,
                                                 _st()), "pack_weight_bf16")
    if key is not None:
        _step_cache.put_wp(key, (weakref.ref(w), wp))
    return wp


# Persistent, always-zero split-K workspace of the tcgen05 wgrad: the kernel accumulates into it and its finalize pass
# reads the result and writes the zeros back, so no memset node precedes each of the ~470 wgrad calls of a step.
# One buffer per device (grow-only; superseded buffers are kept alive because captured graphs may still point at them).
# With branch streams calls may overlap, so each call then gets its own scratch and clears it itself.
_WGRAD_WS = {}
_WGRAD_WS_RETIRED = []


def _wgrad_ws(n_floats: int, dev):
    """-> (workspace tensor, ws_is_zero flag for tdvc_conv1d_tc_wgrad)"""
    if branch_streams_enabled():
        return torch.empty(n_floats, device=dev, dtype=torch.float32), 0
    buf = _WGRAD_WS.get(dev)
    if buf is None or buf.numel() < n_floats:
        if buf is not None:
            _WGRAD_WS_RETIRED.append(buf)
        buf = torch.zeros(max(n_floats, 1 << 20), device=dev, dtype=torch.float32)
        _WGRAD_WS[dev] = buf
    return buf, 1


class _Conv1dTC(torch.autograd.Function):
    """Same contract as _Conv1d for stride 1 / groups 1, computed with bf16 operands and fp32 accumulation on the
    tcgen05 implicit-GEMM kernels (forward, data gradient and weight gradient)."""

    @staticmethod
    def forward(ctx, x, w, bias, residual, pad, dilation, pad_mode, in_slope, out_act, out_slope):
        _req(x, w, bias, residual)
        x, w, bias, residual = _c(x), _c(w), _c(bias), _c(residual)
        B, Cin, Tin = x.shape
        Cout, cin_w, K = w.shape
        if cin_w != Cin:
            raise RuntimeError(f"conv1d: weight {tuple(w.shape)} does not match input {tuple(x.shape)}")
        Tout = Tin + 2 * pad - dilation * (K - 1)
        if Tout <= 0:
            raise RuntimeError(f"conv1d: input length {Tin} too short for kernel {K} (dilation {dilation})")
        if pad_mode == PAD_REFLECT and pad >= Tin:
            raise RuntimeError(f"conv1d: reflect padding {pad} must be smaller than the input length {Tin}")
        lib = _lib.load()
        Cp, Coutp = _cp(Cin), _ceil(Cout, 16)
        # short 'same' convs (the discriminator's 1024-channel tail at T = 9 .. 35, the generator's T/320 convs): the batch is
        # run as ONE sequence of zero-separated samples, so a 128-row tile holds several samples instead of one
        flat = (_FLAT_SHORT and pad_mode == PAD_ZEROS and residual is None and B >= 2 and Tout == Tin and pad > 0
                and 2 * (Tin + 2 * pad) <= 128)
        halo = pad if (pad_mode == PAD_REFLECT or flat) else 0   # zero padding is otherwise TMA out-of-bounds fill
        xp = _pack_act(x, Cp, halo, pad_mode, in_slope)
        wp = _pack_w(w, Coutp, Cp, False)
        y = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
        if residual is not None and residual.shape != y.shape:
            raise RuntimeError("conv1d: residual shape mismatch")
        if flat:
            Tp = Tin + 2 * pad
            _tc_conv(xp=xp, wp=wp, bias=bias, y=y, B=1, Tp=B * Tp, Tout=B * Tp, K=K, dilation=dilation, t_off=-pad, Cp_total=Cp,
                     groups=1, a_ch_off=0, a_ch_stride=0, Cinp_g=Cp, Cout_g=Cout, Coutp_g=Coutp, bias_stride=0, out_act=out_act,
                     out_slope=out_slope, out_packed=0, flat_tp=Tp, flat_halo=pad, flat_T=Tout)
        else:
            _lib.check(lib.tdvc_conv1d_tc_fwd(_p(xp), _p(wp), _p(bias), None, _p(residual), _p(y), B, Cp, Tin + 2 * halo, Cout,
                                              Coutp, Tout, K, dilation, halo - pad, out_act, out_slope, _st()), "conv1d_tc_fwd")
        ctx.flat = flat
        ctx.cfg = (pad, dilation, pad_mode, in_slope, out_act, out_slope)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.save_for_backward(x, w, y if out_act != ACT_NONE else None, xp)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y, xp = ctx.saved_tensors
        pad, dilation, pad_mode, in_slope, out_act, out_slope = ctx.cfg
        lib = _lib.load()
        dy = _c(dy)
        B, Cin, Tin = x.shape
        Cout, _, K = w.shape
        Tout = dy.shape[2]
        dx = dw = db = None
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        Cdp = _cp(Cout)
        dyp = None
        db_from_wgrad = need_b and ctx.needs_input_grad[1]      # the wgrad GEMM yields it through a tap of ones
        packs = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        want_dres = ctx.has_res and ctx.needs_input_grad[3]
        # LeakyReLU layers (every layer of D): the activation's backward is applied inside the pack of dL/dy below
        mask_in_pack = out_act == ACT_LRELU and packs and not want_dres and not (need_b and not packs)
        if out_act != ACT_NONE and not mask_in_pack:
            dz = torch.empty_like(dy)
            _lib.check(lib.tdvc_act_bwd_from_output(_p(dy), _p(y), _p(dz), dy.numel(), out_act, out_slope, _st()), "act_bwd")
            dy = dz
        if packs:
            # one pass over dL/dy: bf16 channels-last copy shared by dgrad and wgrad (+ the bias gradient as column
            # sums when no wgrad follows)
            if need_b and not db_from_wgrad:
                db = torch.empty(Cout, device=x.device, dtype=torch.float32)
                need_b = False
            dyh = pad if ctx.flat else 0                        # flat: dL/dy laid out like the packed input (same zero rows)
            if mask_in_pack:
                dyp = torch.empty(B, Tout + 2 * dyh, Cdp, device=x.device, dtype=torch.bfloat16)
                _lib.check(lib.tdvc_pack_cl_bf16_masked(_p(dy), _p(y), out_slope, _p(dyp), B, Cout, Tout, Cdp, dyh, _p(db), _st()),
                           "pack_cl_bf16_masked")
            else:
                dyp = _pack_act(dy, Cdp, dyh, PAD_ZEROS, 1.0, cache=False, chan_sum=db)
        if ctx.needs_input_grad[0]:
            # dgrad = the same implicit GEMM on dy with channel-swapped, tap-flipped weights
            ph = pad if pad_mode == PAD_REFLECT else 0            # reflect halo kept in the staging buffer
            Lout = Tin + 2 * ph
            zpad = (K - 1) * dilation - (pad - ph)                # zero padding of dy in the equivalent forward conv
            Cinp16 = _ceil(Cin, 16)
            wtp = _pack_w(w, Cinp16, Cdp, True)
            need_stage = ph > 0 or in_slope != 1.0
            stage = torch.empty(B, Cin, Lout, device=x.device, dtype=torch.float32)
            if ctx.flat:
                # 'same' conv: the data gradient is the same flattened conv on dL/dy (zpad == pad, Lout == Tin == Tout)
                Tp = Tout + 2 * pad
                _tc_conv(xp=dyp, wp=wtp, y=stage, B=1, Tp=B * Tp, Tout=B * Tp, K=K, dilation=dilation, t_off=-zpad, Cp_total=Cdp,
                         groups=1, a_ch_off=0, a_ch_stride=0, Cinp_g=Cdp, Cout_g=Cin, Coutp_g=Cinp16, bias_stride=0,
                         out_act=ACT_NONE, out_slope=1.0, out_packed=0, flat_tp=Tp, flat_halo=pad, flat_T=Lout)
            else:
                _lib.check(lib.tdvc_conv1d_tc_fwd(_p(dyp), _p(wtp), None, None, None, _p(stage), B, Cdp, Tout, Cin, Cinp16, Lout,
                                                  K, dilation, -zpad, ACT_NONE, 1.0, _st()), "conv1d_tc_dgrad")
            if need_stage:
                dx = torch.empty_like(x)
                _lib.check(lib.tdvc_pad_act_bwd(_p(stage), _p(x), _p(dx), B * Cin, Tin, ph, int(pad_mode == PAD_REFLECT),
                                                in_slope, _st()), "pad_act_bwd")
            else:
                dx = stage
        if ctx.needs_input_grad[1]:
            # wgrad on tcgen05 from the two packed operands (time is the GEMM K dimension)
            halo = pad if pad_mode == PAD_REFLECT else 0
            dw = torch.empty_like(w)
            if db_from_wgrad:
                db = torch.empty(Cout, device=x.device, dtype=torch.float32)
                need_b = False
            if ctx.flat:
                # both operands are [B, T + 2*pad, C] with the same zero rows: one sequence of B * (T + 2*pad) steps
                Tp = Tout + 2 * pad
                wgrad2(dyp=dyp, xp=xp, B=1, Cdp=Cdp, Tout=B * Tp, Cp=xp.shape[2], Tp=B * Tp, Cout=Cout, Cin=Cin, K=K,
                       dilation=dilation, t_off=[-pad], dw=[dw], db=[db if db_from_wgrad else None])
            elif _USE_WGRAD2:
                wgrad2(dyp=dyp, xp=xp, B=B, Cdp=Cdp, Tout=Tout, Cp=xp.shape[2], Tp=xp.shape[1], Cout=Cout, Cin=Cin, K=K,
                       dilation=dilation, t_off=[halo - pad], dw=[dw], db=[db if db_from_wgrad else None])
            else:
                ws, wz = _wgrad_ws(lib.tdvc_conv1d_tc_wgrad_ws(Cout, Cin, K), x.device)
                _lib.check(lib.tdvc_conv1d_tc_wgrad(_p(dyp), _p(xp), _p(dw), _p(ws), B, Cdp, Tout, xp.shape[2], xp.shape[1],
                                                    Cout, Cin, K, dilation, halo - pad, 0, 0, _p(db) if db_from_wgrad else None,
                                                    wz, _st()), "conv1d_tc_wgrad")
        if need_b:
            db = torch.empty(Cout, device=x.device, dtype=torch.float32)
            _lib.check(lib.tdvc_bias_grad(_p(dy), _p(db), B, Cout, Tout, _st()), "bias_grad")
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None, None, None, None, None


# ----------------------------------------------------------------------------- fused FiLM conditioning path

def _tc_conv(**kw):
    c = _lib.TcConv()
    for k, v in kw.items():
        if k == "kg":
            for i, kk in enumerate(v):
                c.kg[i] = int(kk)
            continue
        setattr(c, k, v.data_ptr() if torch.is_tensor(v) else v)
    _lib.check(_lib.load().tdvc_conv1d_tc_fwd_ex(C.byref(c), _st()), "conv1d_tc_fwd_ex")


_STACKED_COND = os.environ.get("TDVC_TC_STACKED", "1") != "0"     # development switch: 0 = time-as-M kernel for cond_var.0


def set_stacked_cond(on: bool) -> None:
    global _STACKED_COND
    _STACKED_COND = bool(on)


def _step_cached(tag, tensors, make):
    """make() once per step scope for this exact set of source tensors (identity + version); outside a scope, every call.
    The entry keeps the sources alive so that their addresses cannot be recycled under the key."""
    if _step_cache.depth == 0:
        return make()
    key = (tag,) + tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors)
    hit = _step_cache.get_wp(key)
    if hit is None:
        hit = (list(tensors), make())
        _step_cache.put_wp(key, hit)
    return hit[1]


# ----------------------------------------------------------------------------- strided convs as frame convolutions

class _SpaceToDepth(torch.autograd.Function):
    """xs[b, p*C + c, q] = x[b, c, s*q + p - pad] (zero outside the signal), q < Tq."""

    @staticmethod
    def forward(ctx, x, s, pad, Tq):
        _req(x)
        x = _c(x)
        B, Cc, T = x.shape
        out = torch.empty(B, s * Cc, Tq, device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_space_to_depth(_p(x), _p(out), B, Cc, T, s, pad, Tq, _st()), "space_to_depth")
        ctx.dims = (B, Cc, T, s, pad, Tq)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, Cc, T, s, pad, Tq = ctx.dims
        dout = _c(dout)
        dx = torch.empty(B, Cc, T, device=dout.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_depth_to_space(_p(dout), _p(dx), B, Cc, Tq, s, pad, T, _st()), "space_to_depth_bwd")
        return dx, None, None, None


class _DepthToSpace(torch.autograd.Function):
    """y[b, c, u] = ys[b, p*C + c, q] with s*q + p = u + pad, u < Tout."""

    @staticmethod
    def forward(ctx, ys, s, pad, Tout):
        _req(ys)
        ys = _c(ys)
        B, sC, Tq = ys.shape
        Cc = sC // s
        y = torch.empty(B, Cc, Tout, device=ys.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_depth_to_space(_p(ys), _p(y), B, Cc, Tq, s, pad, Tout, _st()), "depth_to_space")
        ctx.dims = (B, Cc, Tq, s, pad, Tout)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, Cc, Tq, s, pad, Tout = ctx.dims
        dy = _c(dy)
        dys = torch.empty(B, s * Cc, Tq, device=dy.device, dtype=torch.float32)
        _lib.check(_lib.load().tdvc_space_to_depth(_p(dy), _p(dys), B, Cc, Tout, s, pad, Tq, _st()), "depth_to_space_bwd")
        return dys, None, None, None


def _frame_conv_eligible(Cin, Cout, K, stride, groups, dilation, reflect) -> bool:
    """dense Conv1d(k, stride = s) with k > s: a stride-1 conv with ceil(k / s) taps over frames of s samples (a kernel that
    is not a whole number of strides -- the latent classifier's k = 21, s = 2 -- gets zero taps appended)"""
    return (_PRECISION == "bf16" and _FRAME_CONV and 1 < stride <= 16 and groups == 1 and dilation == 1 and not reflect
            and K > stride)


_FRAME_CONV = os.environ.get("TDVC_FRAME_CONV", "1") != "0"     # development switch: 0 = fp32 CUDA-core strided kernels


def _strided_conv_as_frames(x, weight, bias, stride, padding, in_slope, out_act, out_slope):
    """Conv1d(k = m*s, stride = s, zero padding) = stride-1 conv with m taps over frames of s samples:
    y[t] = sum_j sum_(p,ci) xs[(p,ci), t + j] * w[co, ci, s*j + p].  The frame view costs one pass over x; the
    convolution (forward, data and weight gradients) then runs on the tcgen05 path instead of the fp32 kernels."""
    Cout, Cin, K = weight.shape
    s, m = stride, -(-K // stride)
    Tout = (x.shape[2] + 2 * padding - K) // s + 1
    Tq = Tout + m - 1
    xs = _SpaceToDepth.apply(x, s, padding, Tq)

    def make():
        wk = weight if m * s == K else torch.nn.functional.pad(weight, (0, m * s - K))
        return wk.view(Cout, Cin, m, s).permute(0, 3, 1, 2).reshape(Cout, s * Cin, m)

    w2 = _step_cached(("frames_w", s, torch.is_grad_enabled()), [weight], make)
    return _Conv1dTC.apply(xs, w2, bias, None, 0, 1, PAD_ZEROS, float(in_slope), _ACT[out_act], float(out_slope))


def _conv_transpose_as_frames(x, weight, bias, stride, padding):
    """ConvTranspose1d(k = m*s, stride = s): with u + pad = s*q + p, y[co, u] = sum_j sum_ci x[ci, q - j] * w[ci, co, s*j + p],
    i.e. a stride-1 conv (m taps, padding m-1, taps reversed) producing s*Cout channels (p, co) per frame, then the
    inverse frame view."""
    Cin, Cout, K = weight.shape
    s, m = stride, K // stride
    Tout = (x.shape[2] - 1) * s - 2 * padding + K

    def make():
        w2 = weight.view(Cin, Cout, m, s).flip(2).permute(3, 1, 0, 2).reshape(s * Cout, Cin, m)
        b2 = bias.repeat(s) if bias is not None else None
        return w2, b2

    w2, b2 = _step_cached(("frames_wt", s, torch.is_grad_enabled()), [weight, bias], make)
    yq = _Conv1dTC.apply(x, w2, b2, None, m - 1, 1, PAD_ZEROS, 1.0, ACT_NONE, 1.0)
    return _DepthToSpace.apply(yq, s, padding, Tout)



_USE_WGRAD2 = os.environ.get("TDVC_WGRAD2", "1") != "0"      # development switch: 0 = first-generation tcgen05 wgrad kernel
_GROUPED_FRAMES = os.environ.get("TDVC_GROUPED_FRAMES", "1") != "0"    # development switch: 0 = fp32 CUDA-core grouped kernels
_FLAT_SHORT = os.environ.get("TDVC_FLAT_SHORT", "1") != "0"    # development switch: 0 = one tile per sample for short sequences
_UNFRAME_EPILOGUE = os.environ.get("TDVC_UNFRAME_EPILOGUE", "1") != "0"   # development switch: 0 = separate frame_unpack pass


def _grouped_frame_plan(Cin, Cout, K, stride, groups, dilation, reflect):
    """Conv1d(k, stride = s, groups = G) -- the discriminators' k41 s4 layers, model/discriminator.py:26-30 -- as a stride-1
    GROUPED tensor-core convolution over frames of s samples in channel-major order: a conv group's cin_g * s frame channels
    are contiguous, so `sub` conv groups form one tensor-core group ("bundle") with a block-diagonal weight.  Returns
    (sub, m taps, frame channels per bundle, outputs per bundle, bundles) or None when the layer is not of that form."""
    if _PRECISION != "bf16" or not _GROUPED_FRAMES or groups <= 1 or dilation != 1 or reflect or stride not in (2, 4, 8):
        return None
    if Cin % groups or Cout % groups or K <= stride:
        return None
    cin_g, cout_g = Cin // groups, Cout // groups
    fpg = cin_g * stride
    for sub in (1, 2, 4, 8):
        if groups % sub == 0 and (sub * fpg) % 16 == 0 and (sub * cout_g) % 16 == 0 and sub * fpg <= 128 and sub * cout_g <= 256:
            return (sub, -(-K // stride), sub * fpg, sub * cout_g, groups // sub)
    return None


def _frame_weights(w, stride, plan):
    """(forward, data-gradient) operands of the bundled frame convolution from the grouped conv weight w[Cout, cin_g, K], one
    launch: wp[j][bundle*Cout_b + co][ci] with ci = (local conv group, c, p) <- w[co, c, s*j + p] on the diagonal blocks,
    zero elsewhere and for the appended taps; wtp[j'][bundle*Cin_b + ci][co] = wp[m-1-j'][...][ci] (taps reversed)."""
    sub, m, cin_b, cout_b, nb = plan
    Cout, cin_g, K = w.shape
    w = _c(w)
    wp = torch.empty(m, Cout, cin_b, device=w.device, dtype=torch.bfloat16)
    wtp = torch.empty(m, nb * cin_b, cout_b, device=w.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().tdvc_frame_weights_pack(_p(w), _p(wp), _p(wtp), Cout, cin_g, K, stride, m, sub, cin_b, cout_b, _st()),
               "frame_weights_pack")
    return wp, wtp


class _GroupedFrameConvTC(torch.autograd.Function):
    """act(conv1d(x, w, groups) + bias) for Conv1d(k, stride = s, zero padding, groups) on tcgen05: one pass builds the bf16
    frame view of x, the convolution is a grouped stride-1 launch of the tensor-core kernels writing the NCW fp32 feature
    map (+ bias + LeakyReLU), the data gradient the same launch on dL/dy followed by the inverse frame view, and the weight
    gradients of all groups one grouped launch of the weight-gradient kernel."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, groups, out_act, out_slope, plan):
        _req(x, w, bias)
        x, w, bias = _c(x), _c(w), _c(bias)
        sub, m, cin_b, cout_b, nb = plan
        B, Cin, T = x.shape
        Cout, cin_g, K = w.shape
        if cin_g * groups != Cin:
            raise RuntimeError(f"conv1d: weight {tuple(w.shape)} does not match input {tuple(x.shape)} groups={groups}")
        Tout = (T + 2 * pad - K) // stride + 1
        Tq = Tout + m - 1
        lib = _lib.load()
        xf = torch.empty(B, Tq, Cin * stride, device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.tdvc_frame_pack_bf16(_p(x), _p(xf), B, Cin, T, stride, pad, Tq, _st()), "frame_pack")
        wp, wtp = _step_cached(("gframes", stride, plan), [w], lambda: _frame_weights(w, stride, plan))
        y = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
        _tc_conv(xp=xf, wp=wp, bias=bias, y=y, B=B, Tp=Tq, Tout=Tout, K=m, dilation=1, t_off=0, Cp_total=Cin * stride,
                 groups=nb, a_ch_off=0, a_ch_stride=cin_b, Cinp_g=cin_b, Cout_g=cout_b, Coutp_g=cout_b,
                 bias_stride=cout_b, out_act=out_act, out_slope=out_slope, out_packed=0,
                 y_grp_stride=cout_b * Tout, y_b_stride=Cout * Tout)
        ctx.cfg = (stride, pad, groups, out_act, out_slope, plan, T)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(xf, w, wtp, y if out_act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xf, w, wtp, y = ctx.saved_tensors
        stride, pad, groups, out_act, out_slope, plan, T = ctx.cfg
        sub, m, cin_b, cout_b, nb = plan
        lib = _lib.load()
        dy = _c(dy)
        B, Cout, Tout = dy.shape
        Tq = xf.shape[1]
        Cin = xf.shape[2] // stride
        cin_g, K = w.shape[1], w.shape[2]
        dx = dw = db = None
        need_w = ctx.needs_input_grad[1]
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[0] or need_w or need_b:
            dyp = torch.empty(B, Tout, Cout, device=dy.device, dtype=torch.bfloat16)
            if out_act == ACT_LRELU:      # the activation's backward applied inside the pack
                _lib.check(lib.tdvc_pack_cl_bf16_masked(_p(dy), _p(y), out_slope, _p(dyp), B, Cout, Tout, Cout, 0, None, _st()),
                           "pack_cl_bf16_masked")
            else:
                if out_act != ACT_NONE:
                    dz = torch.empty_like(dy)
                    _lib.check(lib.tdvc_act_bwd_from_output(_p(dy), _p(y), _p(dz), dy.numel(), out_act, out_slope, _st()), "act_bwd")
                    dy = dz
                _lib.check(lib.tdvc_pack_cl_bf16(_p(dy), _p(dyp), B, Cout, Tout, Cout, 0, PAD_ZEROS, 1.0, None, 0, 0, -1, None,
                                                 _st()), "pack dy")
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, Cin, T, device=dy.device, dtype=torch.float32)
            kw = dict(xp=dyp, wp=wtp, B=B, Tp=Tout, Tout=Tq, K=m, dilation=1, t_off=-(m - 1), Cp_total=Cout, groups=nb,
                      a_ch_off=0, a_ch_stride=cout_b, Cinp_g=cout_b, Cout_g=cin_b, Coutp_g=cin_b, bias_stride=0,
                      out_act=ACT_NONE, out_slope=1.0, out_packed=0)
            if _UNFRAME_EPILOGUE and 16 % stride == 0 and cin_b % 16 == 0 and stride * Tq - pad >= T:
                # the inverse frame view in the conv's epilogue: every sample of dx is one (frame, phase) of the Tq rows
                _tc_conv(y=dx, unframe_s=stride, unframe_pad=pad, unframe_T=T, unframe_C=Cin, **kw)
            else:
                dxf = torch.empty(B, Cin * stride, Tq, device=dy.device, dtype=torch.float32)
                _tc_conv(y=dxf, y_grp_stride=cin_b * Tq, y_b_stride=Cin * stride * Tq, **kw)
                _lib.check(lib.tdvc_frame_unpack(_p(dxf), _p(dx), B, Cin, T, stride, pad, Tq, _st()), "frame_unpack")
        if need_w or need_b:
            dw = torch.empty_like(w)
            db = torch.empty(Cout, device=dy.device, dtype=torch.float32) if need_b else None
            wgrad2(dyp=dyp, xp=xf, B=B, Cdp=Cout, Tout=Tout, Cp=Cin * stride, Tp=Tq, Cout=cout_b, Cin=cin_b, K=m, dilation=1,
                   ngroups=nb, x_ch_stride=cin_b, dy_ch_stride=cout_b, t_off=[0], dw=[dw], db=[db], frame_s=stride, kreal=K,
                   cin_conv_g=cin_g, sub=sub)
            if not need_w:
                dw = None
        return dx, dw, db, None, None, None, None, None, None


def wgrad2(*, dyp, xp, B, Cdp, Tout, Cp, Tp, Cout, Cin, K, dilation, ngroups=1, per_group=False, x_ch_off=0, x_ch_stride=0,
           dy_ch_off=0, dy_ch_stride=0, kg=None, t_off=(0,), dw=(None,), db=(None,), dw_grp_stride=0, db_grp_stride=0,
           frame_s=0, kreal=0, cin_conv_g=0, sub=1, haloed=-1, tapsm=-1, swap=-1):
    """tdvc_conv1d_tc_wgrad2 (include/tdvc_b200.h) on the persistent always-zero workspace."""
    lib = _lib.load()
    c = _lib.TcWgrad2()
    c.dyp, c.xp = dyp.data_ptr(), xp.data_ptr()
    for i in range(4):
        c.dw[i] = dw[i].data_ptr() if i < len(dw) and dw[i] is not None else None
        c.db[i] = db[i].data_ptr() if i < len(db) and db[i] is not None else None
        c.kg[i] = int(kg[i]) if kg is not None and i < len(kg) else K
        c.t_off[i] = int(t_off[i]) if i < len(t_off) else int(t_off[0])
    c.dw_grp_stride, c.db_grp_stride = int(dw_grp_stride), int(db_grp_stride)
    c.B, c.Cdp, c.Tout, c.Cp, c.Tp, c.Cout, c.Cin, c.K, c.dilation = B, Cdp, Tout, Cp, Tp, Cout, Cin, K, dilation
    c.ngroups, c.per_group = ngroups, int(bool(per_group))
    c.x_ch_off, c.x_ch_stride, c.dy_ch_off, c.dy_ch_stride = x_ch_off, x_ch_stride, dy_ch_off, dy_ch_stride
    c.want_bias = int(any(b is not None for b in db))
    c.haloed, c.tapsm, c.swap = haloed, tapsm, swap
    c.frame_s, c.kreal, c.cin_conv_g, c.sub = frame_s, kreal, cin_conv_g, sub
    ws, wz = _wgrad_ws(int(lib.tdvc_conv1d_tc_wgrad2_ws(C.byref(c))), dyp.device)
    c.ws, c.ws_is_zero = ws.data_ptr(), wz
    _lib.check(lib.tdvc_conv1d_tc_wgrad2(C.byref(c), _st()), "conv1d_tc_wgrad2")


def _pack_jobs(jobs):
    """tdvc_pack_jobs: jobs = list of dicts (src tensor or None, dst tensor, dst element offset, kind, geometry)."""
    lib = _lib.load()
    for lo in range(0, len(jobs), _lib.PACK_MAX_JOBS):
        part = jobs[lo:lo + _lib.PACK_MAX_JOBS]
        arr = (_lib.PackJob * len(part))()
        for a, j in zip(arr, part):
            a.src = j["src"].data_ptr() if j.get("src") is not None else None
            a.dst = j["dst"].data_ptr() + j.get("dst_off", 0) * j["dst"].element_size()
            for k in ("kind", "Cout", "Cin", "K", "Rp", "Qp", "flip", "R_total", "r_off", "Q_total", "q_off"):
                setattr(a, k, int(j.get(k, 0)))
        _lib.check(lib.tdvc_pack_jobs(arr, len(part), _st()), "pack_jobs")


def _cond_path_forward(c, slope, w0s, b0s, w2s, b2s, parts=None):
    """gamma|beta of all n FiLM blocks of a stage: gb[n, B, 2C, T] fp32 and what the backward needs (see _MRFCondPath).
    parts = (speaker code [B, Cs], excitation [B, Ce, T]): the conditioning cat([code over time, excitation]) given as its
    two sources (c is None); the packed operand is then written from them directly."""
    n = len(w0s)
    if parts is not None:
        B, T = parts[1].shape[0], parts[1].shape[2]
        Cc = parts[0].shape[1] + parts[1].shape[1]
        c = parts[1]
    else:
        B, Cc, T = c.shape
    K = w0s[0].shape[2]
    C2 = w2s[0].shape[0]
    if K != 3 or any(tuple(w.shape) != (Cc, Cc, 3) for w in w0s) or any(tuple(w.shape) != (C2, Cc, 3) for w in w2s):
        raise RuntimeError("mrf_cond_path: unexpected cond_var geometry")
    lib = _lib.load()
    dev = c.device
    Cg = _ceil(Cc + 1, 16)              # per-block channel pitch; channel Cc is the constant-one channel
    C2p = _ceil(C2, 16)
    # operands
    cp = torch.empty(B, T, Cg, device=dev, dtype=torch.bfloat16)
    if parts is not None:
        _lib.check(lib.tdvc_cond_pack_cl(_p(parts[0]), _p(parts[1]), _p(cp), B, parts[0].shape[1], parts[1].shape[1], T, Cg, _st()),
                   "cond_pack_cl")
    else:
        _lib.check(lib.tdvc_pack_cl_bf16(_p(c), _p(cp), B, Cc, T, Cg, 0, PAD_ZEROS, 1.0, None, 0, Cg, Cc, None, _st()), "pack c")
    # cond_var.0 weights: stacked densely (pitch Cc, kernel with the weights as the M operand) when they fit, else at
    # the padded pitch Cg of the time-as-M kernels.  Blocks are packed in order: block j+1 overwrites the Cg - Cc
    # zero rows block j's pack wrote past its end.
    stacked = _STACKED_COND and Cg >= 64 and K * 128 * Cg * 2 + 2 * (256 + 8 * K) * 128 + 3 * 272 * 32 <= 220 * 1024
    pitch0 = Cc if stacked else Cg
    R0 = (n - 1) * pitch0 + Cg

    def pack_fwd():
        w0p = torch.empty(K, R0, Cg, device=dev, dtype=torch.bfloat16)
        w2p = torch.empty(K, n * C2p, Cg, device=dev, dtype=torch.bfloat16)
        b0p = torch.empty(R0, device=dev, dtype=torch.float32)
        b2p = torch.empty(n * C2p, device=dev, dtype=torch.float32)
        # ONE launch for the 2n weight blocks and 2n bias slices.  Stacked form (pitch0 = Cc < Cg): block j owns rows
        # [j*Cc, (j+1)*Cc); only the last block also writes the Cg - Cc zero rows behind it (no two jobs touch the same row)
        keep, jobs = [], []
        for j in range(n):
            w0, w2 = _c(w0s[j]), _c(w2s[j])
            keep += [w0, w2]
            rows0 = Cg if j == n - 1 else pitch0
            jobs.append(dict(src=w0, dst=w0p, kind=0, Cout=Cc, Cin=Cc, K=K, Rp=rows0, Qp=Cg, flip=0, R_total=R0, r_off=j * pitch0,
                             Q_total=Cg, q_off=0))
            jobs.append(dict(src=w2, dst=w2p, kind=0, Cout=C2, Cin=Cc, K=K, Rp=C2p, Qp=Cg, flip=0, R_total=n * C2p, r_off=j * C2p,
                             Q_total=Cg, q_off=0))
            jobs.append(dict(src=_c(b0s[j]), dst=b0p, dst_off=j * pitch0, kind=1, Cout=Cc, Rp=rows0))
            jobs.append(dict(src=_c(b2s[j]), dst=b2p, dst_off=j * C2p, kind=1, Cout=C2, Rp=C2p))
        _pack_jobs(jobs)
        return w0p, w2p, b0p, b2p

    # the generator runs several times per training iteration on the same weights: packed once per step scope
    w0p, w2p, b0p, b2p = _step_cached(("cond_fwd", stacked), list(w0s) + list(w2s) + list(b0s) + list(b2s), pack_fwd)
    # all cond_var.0 convs: packed bf16 output g1p[B, T, n*Cg] = leaky_relu(conv + bias)
    g1p = torch.empty(B, T, n * Cg, device=dev, dtype=torch.bfloat16)
    if stacked:
        _lib.check(lib.tdvc_conv1d_tc_fwd_stacked(_p(cp), _p(w0p), _p(b0p), _p(g1p), B, Cg, 0, Cg, T, T, K, 1, -1, R0, n, Cc,
                                                  ACT_LRELU, slope, T, n * Cg, 0, 0, Cg, _st()), "conv1d_tc_fwd_stacked")
    else:
        _tc_conv(xp=cp, wp=w0p, bias=b0p, B=B, Tp=T, Tout=T, K=K, dilation=1, t_off=-1, Cp_total=Cg, groups=1,
                 a_ch_off=0, a_ch_stride=0, Cinp_g=Cg, Cout_g=n * Cg, Coutp_g=n * Cg, bias_stride=0,
                 out_act=ACT_LRELU, out_slope=slope, out_packed=1, yp=g1p, tp_out=T, cp_out=n * Cg, out_halo=0,
                 out_ch_off=0, out_ch_stride=0)
    # all cond_var.2 convs, grouped: gb[n, B, 2C, T]
    gb = torch.empty(n, B, C2, T, device=dev, dtype=torch.float32)
    _tc_conv(xp=g1p, wp=w2p, bias=b2p, y=gb, B=B, Tp=T, Tout=T, K=K, dilation=1, t_off=-1, Cp_total=n * Cg, groups=n,
             a_ch_off=0, a_ch_stride=Cg, Cinp_g=Cg, Cout_g=C2, Coutp_g=C2p, bias_stride=C2p, out_act=ACT_NONE,
             out_slope=1.0, out_packed=0)
    dims = (n, B, Cc, T, K, C2, Cg, C2p, slope)
    return gb, dims, cp, g1p


def _cond_path_backward(dims, cp, g1p, w0s, w2s, has_b0, has_b2, dgbp, need_dc):
    """dgbp[B, T, n*C2p]: packed bf16 dL/d(gamma|beta) of every block.  Returns (dL/dc or None, flat per-block gradient
    list [dw0, db0, dw2, db2] * n)."""
    n, B, Cc, T, K, C2, Cg, C2p, slope = dims
    lib = _lib.load()
    dev = cp.device
    db2 = torch.empty(n, C2, device=dev, dtype=torch.float32)
    # cond_var.2 weight gradients (their bias gradients come out of the same GEMM through a tap of ones)
    dw2 = []
    if _USE_WGRAD2:
        # the n blocks' cond_var.2 weight (+ bias) gradients: ONE grouped launch
        dw2_all = torch.empty(n, C2, Cc, K, device=dev, dtype=torch.float32)
        wgrad2(dyp=dgbp, xp=g1p, B=B, Cdp=n * C2p, Tout=T, Cp=n * Cg, Tp=T, Cout=C2, Cin=Cc, K=K, dilation=1, ngroups=n,
               x_ch_stride=Cg, dy_ch_stride=C2p, t_off=[-1], dw=[dw2_all], db=[db2], dw_grp_stride=C2 * Cc * K,
               db_grp_stride=C2)
        dw2 = [dw2_all[j] for j in range(n)]
    else:
        ws, wz = _wgrad_ws(max(lib.tdvc_conv1d_tc_wgrad_ws(C2, Cc, K), lib.tdvc_conv1d_tc_wgrad_ws(n * Cg, Cc + 1, K)), dev)
        for j in range(n):
            g = torch.empty(C2, Cc, K, device=dev, dtype=torch.float32)
            _lib.check(lib.tdvc_conv1d_tc_wgrad(_p(dgbp), _p(g1p), _p(g), _p(ws), B, n * C2p, T, n * Cg, T, C2, Cc, K, 1, -1,
                                                j * Cg, j * C2p, _p(db2[j]), wz, _st()), "wgrad cond_var.2")
            dw2.append(g)

    # dL/dg1 (packed, LeakyReLU mask applied in the epilogue): grouped dgrad of cond_var.2
    def pack_bwd():
        w2tp = torch.empty(K, n * Cg, C2p, device=dev, dtype=torch.bfloat16)
        w0tp = torch.empty(K, Cg, n * Cg, device=dev, dtype=torch.bfloat16)
        jobs = []
        for j in range(n):
            # transposed, tap-reversed operands of the two data-gradient convs (one launch for all 2n blocks)
            jobs.append(dict(src=w2s[j], dst=w2tp, kind=0, Cout=C2, Cin=Cc, K=K, Rp=Cg, Qp=C2p, flip=1, R_total=n * Cg, r_off=j * Cg,
                             Q_total=C2p, q_off=0))
            jobs.append(dict(src=w0s[j], dst=w0tp, kind=0, Cout=Cc, Cin=Cc, K=K, Rp=Cg, Qp=Cg, flip=1, R_total=Cg, r_off=0,
                             Q_total=n * Cg, q_off=j * Cg))
        _pack_jobs(jobs)
        return w2tp, w0tp

    w2tp, w0tp = _step_cached(("cond_bwd",), list(w0s) + list(w2s), pack_bwd)
    dg1p = torch.empty(B, T, n * Cg, device=dev, dtype=torch.bfloat16)
    _tc_conv(xp=dgbp, wp=w2tp, B=B, Tp=T, Tout=T, K=K, dilation=1, t_off=-1, Cp_total=n * C2p, groups=n, a_ch_off=0,
             a_ch_stride=C2p, Cinp_g=C2p, Cout_g=Cg, Coutp_g=Cg, bias_stride=0, out_act=ACT_NONE, out_slope=1.0,
             out_packed=1, yp=dg1p, tp_out=T, cp_out=n * Cg, out_halo=0, out_ch_off=0, out_ch_stride=Cg,
             maskp=g1p, tm=T, cm=n * Cg, mask_halo=0, mask_ch_off=0, mask_ch_stride=Cg, mask_slope=slope)
    # cond_var.0 weight (+ bias, through the constant-one channel of cp) gradients: one GEMM for all blocks
    dw0_all = torch.empty(n * Cg, Cc + 1, K, device=dev, dtype=torch.float32)
    if _USE_WGRAD2:
        wgrad2(dyp=dg1p, xp=cp, B=B, Cdp=n * Cg, Tout=T, Cp=Cg, Tp=T, Cout=n * Cg, Cin=Cc + 1, K=K, dilation=1, t_off=[-1],
               dw=[dw0_all], db=[None])
    else:
        ws, wz = _wgrad_ws(max(lib.tdvc_conv1d_tc_wgrad_ws(C2, Cc, K), lib.tdvc_conv1d_tc_wgrad_ws(n * Cg, Cc + 1, K)), dev)
        _lib.check(lib.tdvc_conv1d_tc_wgrad(_p(dg1p), _p(cp), _p(dw0_all), _p(ws), B, n * Cg, T, Cg, T, n * Cg, Cc + 1, K, 1,
                                            -1, 0, 0, None, wz, _st()), "wgrad cond_var.0")
    # dL/dc: one conv over the n*Cg concatenated channels (sums the blocks' contributions in the GEMM)
    dc = None
    if need_dc:
        dc = torch.empty(B, Cc, T, device=dev, dtype=torch.float32)
        _tc_conv(xp=dg1p, wp=w0tp, y=dc, B=B, Tp=T, Tout=T, K=K, dilation=1, t_off=-1, Cp_total=n * Cg, groups=1,
                 a_ch_off=0, a_ch_stride=0, Cinp_g=n * Cg, Cout_g=Cc, Coutp_g=Cg, bias_stride=0, out_act=ACT_NONE,
                 out_slope=1.0, out_packed=0)
    grads = []
    for j in range(n):
        blk = dw0_all[j * Cg:j * Cg + Cc]
        grads.append(blk[:, :Cc, :].contiguous())
        grads.append(blk[:, Cc, (K - 1) // 2].contiguous() if has_b0[j] else None)
        grads.append(dw2[j])
        grads.append(db2[j] if has_b2[j] else None)
    return dc, grads


class _MRFCondPath(torch.autograd.Function):
    """(gamma|beta)_j = cond_var_j[2](leaky_relu(cond_var_j[0](c)))  for all n FiLM blocks of one MRF stage
    (model/generator.py:85-92,102), which all read the same conditioning tensor c[B, Cc, T]:

      forward   1 pack of c, ONE tcgen05 launch for the n `cond_var.0` convs (N = n*144) whose epilogue applies bias +
                LeakyReLU and writes the next conv's bf16 channels-last operand directly, ONE grouped launch for the n
                `cond_var.2` convs -> gb[n, B, 2C, T] fp32.
      backward  n packs of dL/dgb, one grouped dgrad launch whose epilogue applies the LeakyReLU mask and writes packed
                dL/dg1, the weight-gradient launches (cond_var.2: one grouped launch, bias gradients through a tap of ones;
                cond_var.0: one GEMM for all blocks, bias gradients through a constant-one input channel), and one
                dgrad launch over the concatenated 1296 channels, which also sums the n blocks' contributions to dL/dc.
    Intermediates never exist in fp32."""

    @staticmethod
    def forward(ctx, c, slope, *wb):
        n = len(wb) // 4
        w0s, b0s, w2s, b2s = wb[0::4], wb[1::4], wb[2::4], wb[3::4]
        _req(c, *wb)
        c = _c(c)
        gb, dims, cp, g1p = _cond_path_forward(c, slope, w0s, b0s, w2s, b2s)
        ctx.dims = dims
        ctx.has_b0 = [b is not None for b in b0s]
        ctx.has_b2 = [b is not None for b in b2s]
        ctx.save_for_backward(cp, g1p, *[_c(w) for w in w0s], *[_c(w) for w in w2s])
        return tuple(gb[j] for j in range(n))

    @staticmethod
    def backward(ctx, *dgb):
        n, B, Cc, T, K, C2, Cg, C2p, slope = ctx.dims
        saved = ctx.saved_tensors
        cp, g1p = saved[0], saved[1]
        w0s, w2s = saved[2:2 + n], saved[2 + n:2 + 2 * n]
        lib = _lib.load()
        # dL/dgb -> packed bf16
        dgbp = torch.empty(B, T, n * C2p, device=cp.device, dtype=torch.bfloat16)
        for j in range(n):
            if dgb[j] is None:
                dgbp[:, :, j * C2p:(j + 1) * C2p].zero_()
                continue
            d = _c(dgb[j])
            _lib.check(lib.tdvc_pack_cl_bf16(_p(d), _p(dgbp), B, C2, T, n * C2p, 0, PAD_ZEROS, 1.0, None, j * C2p, C2p,
                                             -1, None, _st()), "pack dgb")
        dc, grads = _cond_path_backward(ctx.dims, cp, g1p, w0s, w2s, ctx.has_b0, ctx.has_b2, dgbp, ctx.needs_input_grad[0])
        return (dc, None, *grads)


def mrf_cond_path(c, blocks_wb, slope=0.2):
    """blocks_wb: [(w0, b0, w2, b2), ...] effective (weight-normed) cond_var weights of the FiLM blocks sharing c.
    Returns the list of gamma|beta tensors [B, 2C, T], one per block."""
    flat = []
    for w0, b0, w2, b2 in blocks_wb:
        flat += [w0, b0, w2, b2]
    return list(_MRFCondPath.apply(c, float(slope), *flat))


def mrf_cond_path_eligible(Cc, C2, T) -> bool:
    return _PRECISION == "bf16" and Cc >= 16 and C2 >= 16 and _ceil(C2, 16) <= 256


# ----------------------------------------------------------------------------- the bf16-resident MRF stage

_MRF_CHAIN = os.environ.get("TDVC_MRF_CHAIN", "1") != "0"       # development switch: 0 = block-by-block path


def mrf_stage_eligible(C, T, kernel_sizes, dilations, has_cond, Cc=0) -> bool:
    """The whole-stage path needs bf16 mode, channel counts that are multiples of 16 (a thread's 16 accumulator columns
    are real channels), up to 4 kernel sizes of the same parity, and reflect halos shorter than the signal."""
    if _PRECISION != "bf16" or not _MRF_CHAIN or C % 16 or C > 256 or not (1 <= len(kernel_sizes) <= 4):
        return False
    kmax = max(kernel_sizes)
    if any(k % 2 == 0 or k < 1 for k in kernel_sizes) or any(d < 1 for d in dilations):
        return False
    if max(dilations) * (kmax - 1) // 2 >= T or 64 + (kmax - 1) * max(dilations) > 256:
        return False
    if has_cond and not mrf_cond_path_eligible(Cc, 2 * C, T):
        return False
    return True


def _chain_weights(ws_f32, bs, G, Cx, Kmax, ks):
    """Grouped operands of one depth, one launch: wp[Kmax][G*Cx][Cx] (forward; branch i's k_i taps centred, zeros around
    them), wtp[Kmax][G*Cx][Cx] (data gradient: rows = input channels, taps reversed) and the concatenated bias [G*Cx]."""
    lib = _lib.load()
    dev = ws_f32[0].device
    wp = torch.empty(Kmax, G * Cx, Cx, device=dev, dtype=torch.bfloat16)
    wtp = torch.empty(Kmax, G * Cx, Cx, device=dev, dtype=torch.bfloat16)
    bias = torch.empty(G * Cx, device=dev, dtype=torch.float32)
    keep = [_c(w) for w in ws_f32] + [_c(b) for b in bs]
    wptr = (C.c_void_p * G)(*[t.data_ptr() for t in keep[:G]])
    bptr = (C.c_void_p * G)(*[(t.data_ptr() if t is not None else None) for t in keep[G:]])
    karr = (C.c_int * G)(*[int(k) for k in ks])
    _lib.check(lib.tdvc_chain_pack(wptr, bptr, karr, G, Cx, Kmax, _p(wp), _p(wtp), _p(bias), _st()), "chain_pack")
    return wp, wtp, bias


class _MRFStage(torch.autograd.Function):
    """One MRFBlock (model/generator.py:175-194): G kernel-size branches of D chained FiLM residual blocks
    (model/generator.py:69-111), averaged -- as ONE autograd node whose tensors between the convolutions are bf16
    channels-last, the G branches side by side in the channel dimension, each depth ONE launch per convolution for all
    branches (group = branch, per-group tap count):

      forward   pack(x) -> per depth: [conv.1 + bias, FiLM, LeakyReLU -> packed]  [posconv.1 + bias + residual -> fp32
                residual stream + packed LeakyReLU copy with the next conv's reflect halo]  -> mean of the branches
      backward  per depth, last to first: posconv weight gradients (grouped), [posconv^T with the LeakyReLU mask and the
                FiLM backward in its epilogue -> packed dL/dh0, packed dL/d(gamma|beta)], conv.1 weight gradients (grouped),
                [conv.1^T over the padded rows with the LeakyReLU mask and the residual gradient added -> fp32 + packed],
                reflect fold; then the conditioning path's backward on the packed dL/d(gamma|beta).
    The fp32 NCW activations h0 / FiLM output / LeakyReLU inputs of the block-by-block path, their pack passes and the
    reflect-fold / mask passes do not exist here."""

    @staticmethod
    def forward(ctx, x, c, e, slope, cond_slope, ks, ds, *wb):
        G, D = len(ks), len(ds)
        has_cond = c is not None
        per = 8 if has_cond else 4
        assert len(wb) == per * G * D
        _req(x, c, e, *wb)
        x = _c(x)
        B, Cc_x, T = x.shape
        Cx = Cc_x
        lib = _lib.load()
        dev = x.device
        Kmax = max(ks)
        H = [(Kmax - 1) // 2 * d for d in ds]
        blk = lambda i, j: wb[per * (i * D + j): per * (i * D + j + 1)]
        # ---- conditioning path: gamma|beta of every block, block index i*D + j
        gb = cdims = cp = g1p = None
        ctx.cond_parts = None
        if has_cond:
            c = _c(c)
            w0s = [blk(i, j)[4] for i in range(G) for j in range(D)]
            b0s = [blk(i, j)[5] for i in range(G) for j in range(D)]
            w2s = [blk(i, j)[6] for i in range(G) for j in range(D)]
            b2s = [blk(i, j)[7] for i in range(G) for j in range(D)]
            if e is not None:
                # c = speaker code [B, Cs], e = excitation level [B, Ce, T]: the conditioning tensor is never built in fp32
                e = _c(e)
                if c.dim() != 2 or e.dim() != 3 or c.shape[0] != e.shape[0] or e.shape[2] != T:
                    raise RuntimeError(f"mrf_stage: conditioning parts {tuple(c.shape)}, {tuple(e.shape)} do not match x {tuple(x.shape)}")
                ctx.cond_parts = (c.shape[1], e.shape[1])
                gb, cdims, cp, g1p = _cond_path_forward(None, cond_slope, w0s, b0s, w2s, b2s, parts=(c, e))
            else:
                gb, cdims, cp, g1p = _cond_path_forward(c, cond_slope, w0s, b0s, w2s, b2s)
        gb_blk = B * 2 * Cx * T                                      # elements of one block's gamma|beta
        # ---- operands of the chain (once per step scope)
        wconv, wpos = [], []
        for j in range(D):
            cw = [blk(i, j)[0] for i in range(G)]
            cb = [blk(i, j)[1] for i in range(G)]
            pw = [blk(i, j)[2] for i in range(G)]
            pb = [blk(i, j)[3] for i in range(G)]
            wconv.append(_step_cached(("chain_conv", Kmax, tuple(ks)), cw + cb, lambda: _chain_weights(cw, cb, G, Cx, Kmax, ks)))
            wpos.append(_step_cached(("chain_pos",), pw + pb, lambda: _chain_weights(pw, pb, G, Cx, 1, [1] * G)))
        # ---- depth 0 operand: LeakyReLU(x) with the reflect halo, C channels shared by the G branches
        X = [torch.empty(B, T + 2 * H[0], Cx, device=dev, dtype=torch.bfloat16)]
        _lib.check(lib.tdvc_pack_cl_bf16(_p(x), _p(X[0]), B, Cx, T, Cx, H[0], PAD_REFLECT if H[0] > 0 else PAD_ZEROS, slope, None,
                                         0, 0, -1, None, _st()), "pack x")
        a1, h0 = [], []
        res, res_stride = x, 0
        xs = None
        needs_bwd = any(ctx.needs_input_grad)
        for j in range(D):
            a1j = torch.empty(B, T, G * Cx, device=dev, dtype=torch.bfloat16)
            # the un-modulated conv result is only read by the FiLM backward: not written for a forward-only call
            h0j = torch.empty(B, T, G * Cx, device=dev, dtype=torch.bfloat16) if (has_cond and needs_bwd) else None
            kw = dict(xp=X[j], wp=wconv[j][0], bias=wconv[j][2], B=B, Tp=T + 2 * H[j], Tout=T, K=Kmax, dilation=ds[j], t_off=0,
                      Cp_total=X[j].shape[2], groups=G, a_ch_off=0, a_ch_stride=(0 if j == 0 else Cx), Cinp_g=Cx, Cout_g=Cx,
                      Coutp_g=Cx, bias_stride=Cx, out_act=ACT_LRELU, out_slope=slope, out_packed=1, yp=a1j, tp_out=T,
                      cp_out=G * Cx, out_halo=0, out_ch_off=0, out_ch_stride=Cx, chain_mode=3, kg=ks)
            if has_cond:
                kw.update(gb=gb.data_ptr() + 4 * j * gb_blk, gb_grp_stride=D * gb_blk)
                if h0j is not None:
                    kw.update(yp2=h0j)
            _tc_conv(**kw)
            xs = torch.empty(G, B, Cx, T, device=dev, dtype=torch.float32)
            Hn = H[j + 1] if j + 1 < D else 0
            Xn = torch.empty(B, T + 2 * Hn, G * Cx, device=dev, dtype=torch.bfloat16) if j + 1 < D else None
            kw = dict(xp=a1j, wp=wpos[j][0], bias=wpos[j][2], residual=res, y=xs, B=B, Tp=T, Tout=T, K=1, dilation=1, t_off=0,
                      Cp_total=G * Cx, groups=G, a_ch_off=0, a_ch_stride=Cx, Cinp_g=Cx, Cout_g=Cx, Coutp_g=Cx, bias_stride=Cx,
                      out_act=ACT_NONE, out_slope=1.0, out_packed=0, chain_mode=4, res_grp_stride=res_stride,
                      y_grp_stride=B * Cx * T, y_b_stride=Cx * T, pk_slope=slope)
            if Xn is not None:
                kw.update(yp=Xn, tp_out=T + 2 * Hn, cp_out=G * Cx, out_halo=Hn, out_ch_off=0, out_ch_stride=Cx)
                X.append(Xn)
            _tc_conv(**kw)
            a1.append(a1j)
            h0.append(h0j)
            res, res_stride = xs, B * Cx * T
        y = torch.empty(B, Cx, T, device=dev, dtype=torch.float32)
        ptr = lambda i: C.c_void_p(xs.data_ptr() + 4 * i * B * Cx * T)
        if G <= 3:
            _lib.check(lib.tdvc_add3_scale(ptr(0), ptr(1) if G > 1 else None, ptr(2) if G > 2 else None, _p(y), y.numel(),
                                           1.0 / G, _st()), "mrf mean")
        else:
            tmp = torch.empty_like(y)
            _lib.check(lib.tdvc_add3_scale(ptr(0), ptr(1), ptr(2), _p(tmp), y.numel(), 1.0, _st()), "mrf sum")
            _lib.check(lib.tdvc_add3_scale(_p(tmp), ptr(3), None, _p(y), y.numel(), 1.0 / G, _st()), "mrf mean")
        ctx.cfg = (G, D, B, Cx, T, Kmax, tuple(ks), tuple(ds), tuple(H), slope, has_cond, per, cdims)
        ctx.has_bias = [t is not None for t in wb]
        saved = list(X) + a1 + ([t for t in h0] if has_cond else []) + ([gb, cp, g1p] if has_cond else [])
        saved += [w for pair in wconv for w in (pair[1],)] + [w for pair in wpos for w in (pair[1],)]
        if has_cond:
            saved += [_c(blk(i, j)[4]) for i in range(G) for j in range(D)] + [_c(blk(i, j)[6]) for i in range(G) for j in range(D)]
        ctx.save_for_backward(*saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        G, D, B, Cx, T, Kmax, ks, ds, H, slope, has_cond, per, cdims = ctx.cfg
        sv = list(ctx.saved_tensors)
        X, sv = sv[:D], sv[D:]
        a1, sv = sv[:D], sv[D:]
        h0 = [None] * D
        gb = cp = g1p = None
        if has_cond:
            h0, sv = sv[:D], sv[D:]
            (gb, cp, g1p), sv = sv[:3], sv[3:]
        wconvT, sv = sv[:D], sv[D:]
        wposT, sv = sv[:D], sv[D:]
        n = G * D
        w0s = w2s = None
        if has_cond:
            w0s, w2s = sv[:n], sv[n:2 * n]
        lib = _lib.load()
        dev = dy.device
        dy = _c(dy)
        need = ctx.needs_input_grad
        gb_blk = B * 2 * Cx * T
        C2p = 2 * Cx
        dgbp = torch.empty(B, T, n * C2p, device=dev, dtype=torch.bfloat16) if has_cond else None
        # gradient of the branch mean: the same dL/dy / G for every branch (group stride 0)
        d_out = torch.empty_like(dy)
        _lib.check(lib.tdvc_add3_scale(_p(dy), None, None, _p(d_out), dy.numel(), 1.0 / G, _st()), "mrf mean bwd")
        dyp = torch.empty(B, T, Cx, device=dev, dtype=torch.bfloat16)
        _lib.check(lib.tdvc_pack_cl_bf16(_p(d_out), _p(dyp), B, Cx, T, Cx, 0, PAD_ZEROS, 1.0, None, 0, 0, -1, None, _st()), "pack dy")
        grads = [None] * (per * n)
        gi = lambda i, j: per * (i * D + j)
        d_in = None
        for j in reversed(range(D)):
            top = j == D - 1
            # posconv.1 weight / bias gradients of the G branches
            dwp = [torch.empty(Cx, Cx, 1, device=dev, dtype=torch.float32) for _ in range(G)]
            dbp = [torch.empty(Cx, device=dev, dtype=torch.float32) if ctx.has_bias[gi(i, j) + 3] else None for i in range(G)]
            wgrad2(dyp=dyp, xp=a1[j], B=B, Cdp=dyp.shape[2], Tout=T, Cp=G * Cx, Tp=T, Cout=Cx, Cin=Cx, K=1, dilation=1, ngroups=G,
                   per_group=True, x_ch_stride=Cx, dy_ch_stride=(0 if top else Cx), kg=[1] * G, t_off=[0] * G, dw=dwp, db=dbp)
            # posconv^T (+ LeakyReLU mask + FiLM backward) -> packed dL/dh0 and packed dL/d(gamma|beta)
            dh0p = torch.empty(B, T, G * Cx, device=dev, dtype=torch.bfloat16)
            kw = dict(xp=dyp, wp=wposT[j], B=B, Tp=T, Tout=T, K=1, dilation=1, t_off=0, Cp_total=dyp.shape[2], groups=G,
                      a_ch_off=0, a_ch_stride=(0 if top else Cx), Cinp_g=Cx, Cout_g=Cx, Coutp_g=Cx, bias_stride=0,
                      out_act=ACT_NONE, out_slope=1.0, out_packed=1, yp=dh0p, tp_out=T, cp_out=G * Cx, out_halo=0, out_ch_off=0,
                      out_ch_stride=Cx, maskp=a1[j], tm=T, cm=G * Cx, mask_halo=0, mask_ch_off=0, mask_ch_stride=Cx,
                      mask_slope=slope, chain_mode=5)
            if has_cond:
                kw.update(gb=gb.data_ptr() + 4 * j * gb_blk, gb_grp_stride=D * gb_blk, auxp=h0[j], dgbp=dgbp, dgb_cp=n * C2p,
                          dgb_ch_off=j * C2p, dgb_ch_stride=D * C2p)
            _tc_conv(**kw)
            # conv.1 weight / bias gradients of the G branches (k_i taps each, one launch)
            dwc = [torch.empty(Cx, Cx, ks[i], device=dev, dtype=torch.float32) for i in range(G)]
            dbc = [torch.empty(Cx, device=dev, dtype=torch.float32) if ctx.has_bias[gi(i, j) + 1] else None for i in range(G)]
            wgrad2(dyp=dh0p, xp=X[j], B=B, Cdp=G * Cx, Tout=T, Cp=X[j].shape[2], Tp=T + 2 * H[j], Cout=Cx, Cin=Cx, K=Kmax,
                   dilation=ds[j], ngroups=G, per_group=True, x_ch_stride=(0 if j == 0 else Cx), dy_ch_stride=Cx, kg=list(ks),
                   t_off=[H[j] - ds[j] * (k - 1) // 2 for k in ks], dw=dwc, db=dbc)
            for i in range(G):
                grads[gi(i, j) + 0], grads[gi(i, j) + 1] = dwc[i], dbc[i]
                grads[gi(i, j) + 2], grads[gi(i, j) + 3] = dwp[i], dbp[i]
            # conv.1^T over the padded rows (+ LeakyReLU mask + residual gradient) -> dL/dx of this depth, fp32 + packed
            Lp = T + 2 * H[j]
            d_in = torch.empty(G, B, Cx, T, device=dev, dtype=torch.float32)
            dyp_prev = torch.empty(B, T, G * Cx, device=dev, dtype=torch.bfloat16) if j > 0 else None
            hb = torch.empty(G, B, Cx, max(1, 2 * H[j]), device=dev, dtype=torch.float32)
            kw = dict(xp=dh0p, wp=wconvT[j], B=B, Tp=T, Tout=Lp, K=Kmax, dilation=ds[j], t_off=-(Kmax - 1) * ds[j], Cp_total=G * Cx,
                      groups=G, a_ch_off=0, a_ch_stride=Cx, Cinp_g=Cx, Cout_g=Cx, Coutp_g=Cx, bias_stride=0, out_act=ACT_NONE,
                      out_slope=1.0, out_packed=0, y=d_in, y_grp_stride=B * Cx * T, y_b_stride=Cx * T, residual=d_out,
                      res_grp_stride=(0 if top else B * Cx * T), maskp=X[j], tm=Lp, cm=X[j].shape[2], mask_halo=0, mask_ch_off=0,
                      mask_ch_stride=(0 if j == 0 else Cx), mask_slope=slope, chain_mode=6, kg=ks, halo_buf=hb, halo=H[j],
                      t_valid=T)
            if dyp_prev is not None:
                kw.update(yp=dyp_prev, tp_out=T, cp_out=G * Cx, out_halo=0, out_ch_off=0, out_ch_stride=Cx)
            _tc_conv(**kw)
            if H[j] > 0:
                _lib.check(lib.tdvc_chain_fold(_p(hb), _p(d_in), _p(dyp_prev), G, B, Cx, T, H[j], B * Cx * T, Cx * T, G * Cx, 0, Cx,
                                               _st()), "chain_fold")
            d_out, dyp = d_in, dyp_prev
        dx = None
        if need[0]:
            dx = torch.empty(B, Cx, T, device=dev, dtype=torch.float32)
            ptr = lambda i: C.c_void_p(d_in.data_ptr() + 4 * i * B * Cx * T)
            if G <= 3:
                _lib.check(lib.tdvc_add3_scale(ptr(0), ptr(1) if G > 1 else None, ptr(2) if G > 2 else None, _p(dx), dx.numel(),
                                               1.0, _st()), "mrf dx")
            else:
                tmp = torch.empty_like(dx)
                _lib.check(lib.tdvc_add3_scale(ptr(0), ptr(1), ptr(2), _p(tmp), dx.numel(), 1.0, _st()), "mrf dx")
                _lib.check(lib.tdvc_add3_scale(_p(tmp), ptr(3), None, _p(dx), dx.numel(), 1.0, _st()), "mrf dx")
        dc = None
        if has_cond:
            has_b0 = [ctx.has_bias[gi(i, j) + 5] for i in range(G) for j in range(D)]
            has_b2 = [ctx.has_bias[gi(i, j) + 7] for i in range(G) for j in range(D)]
            dc, cg = _cond_path_backward(cdims, cp, g1p, w0s, w2s, has_b0, has_b2, dgbp, need[1] or need[2])
            for b_ in range(n):
                grads[per * b_ + 4: per * b_ + 8] = cg[4 * b_: 4 * b_ + 4]
        de = None
        if ctx.cond_parts is not None and dc is not None:
            # conditioning given as (speaker code, excitation): time-sum / split of dL/dc (what cond_concat's backward does)
            Cs, Ce = ctx.cond_parts
            dfull = dc
            dc = torch.empty(B, Cs, device=dev, dtype=torch.float32)
            de = torch.empty(B, Ce, T, device=dev, dtype=torch.float32) if need[2] else None
            _lib.check(lib.tdvc_cond_concat_bwd(_p(dfull), _p(dc), _p(de), B, Cs, Ce, T, 1, _st()), "cond_concat_bwd")
        return (dx, dc, de, None, None, None, None, *grads)


class CondParts:
    """The decoder's conditioning cat([c.unsqueeze(2).repeat(1, 1, T), e], dim=1) (model/generator.py:387-399) kept as its two
    sources -- speaker code c[B, Cs], excitation level e[B, Ce, T] -- so that the bf16 whole-stage path can write its packed
    operand from them directly; .tensor() builds (once) the fp32 tensor every other consumer gets."""

    def __init__(self, c, e):
        self.c, self.e = c, e
        self._full = None
        self.shape = (e.shape[0], c.shape[1] + e.shape[1], e.shape[2])
        self.ndim = 3

    def tensor(self):
        if self._full is None:
            self._full = cond_concat(self.c, self.e)
        return self._full


def cond_parts(c, e):
    """cond_concat(c, e), lazily: a CondParts where the whole-stage path can consume the parts, else the tensor."""
    if _PRECISION == "bf16" and _MRF_CHAIN and e.is_cuda and c.dim() == 2 and e.dim() == 3:
        return CondParts(c, e)
    return cond_concat(c, e)


def mrf_stage(x, c, blocks, kernel_sizes, dilations, slope=0.2, cond_slope=0.2):
    """blocks[i][j] = (conv_w, conv_b, pos_w, pos_b[, cv0_w, cv0_b, cv2_w, cv2_b]): effective (weight-normed) weights of the
    FiLM block of kernel size i / dilation j.  c: the conditioning tensor [B, Cc, T], a CondParts, or None."""
    flat = []
    for row in blocks:
        for blk in row:
            flat += list(blk)
    e = None
    if isinstance(c, CondParts):
        c, e = c.c, c.e
    return _MRFStage.apply(x, c, e, float(slope), float(cond_slope), tuple(int(k) for k in kernel_sizes),
                           tuple(int(d) for d in dilations), *flat)


# ----------------------------------------------------------------------------- fused FiLM + posconv

class _FilmPosconvTC(torch.autograd.Function):
    """out = conv1x1(leaky_relu(h0 * (1 + gamma) + beta)) + bias + x      (model/generator.py:104-109)

    forward : ONE pass applies FiLM + LeakyReLU and writes the bf16 channels-last operand (no fp32 FiLM output, no
              separate pack), then the tcgen05 1x1 conv with the residual add in its epilogue.
    backward: dL/dout is packed once (also the bias gradient); the data-gradient conv applies the LeakyReLU mask in
              its epilogue straight from the packed activation (no staging pass), the FiLM backward kernel yields
              dL/dh0 and dL/d(gamma|beta)."""

    @staticmethod
    def forward(ctx, h0, gb, w, bias, x_res, slope):
        _req(h0, gb, w, bias, x_res)
        h0, gb, w, bias, x_res = _c(h0), _c(gb), _c(w), _c(bias), _c(x_res)
        B, Cc, T = h0.shape
        Cout = w.shape[0]
        if w.shape[1] != Cc or w.shape[2] != 1:
            raise RuntimeError("film_posconv: expects a 1x1 convolution")
        if gb is not None and tuple(gb.shape) != (B, 2 * Cc, T):
            raise RuntimeError(f"film: gamma/beta tensor {tuple(gb.shape)} does not match {tuple(h0.shape)}")
        lib = _lib.load()
        Cp, Coutp = _cp(Cc), _ceil(Cout, 16)
        a1p = torch.empty(B, T, Cp, device=h0.device, dtype=torch.bfloat16)
        _lib.check(lib.tdvc_pack_cl_bf16(_p(h0), _p(a1p), B, Cc, T, Cp, 0, PAD_ZEROS, slope, None, 0, 0, -1, _p(gb), _st()),
                   "film_pack")
        wp = _pack_w(w, Coutp, Cp, False)
        y = torch.empty(B, Cout, T, device=h0.device, dtype=torch.float32)
        _lib.check(lib.tdvc_conv1d_tc_fwd(_p(a1p), _p(wp), _p(bias), None, _p(x_res), _p(y), B, Cp, T, Cout, Coutp, T, 1, 1, 0,
                                          ACT_NONE, 1.0, _st()), "posconv_tc_fwd")
        ctx.slope = slope
        ctx.has_bias = bias is not None
        ctx.save_for_backward(h0, gb, w, a1p)
        return y

    @staticmethod
    def backward(ctx, dy):
        h0, gb, w, a1p = ctx.saved_tensors
        lib = _lib.load()
        dy = _c(dy)
        B, Cc, T = h0.shape
        Cout = w.shape[0]
        Cp = a1p.shape[2]
        Cdp = _cp(Cout)
        db = torch.empty(Cout, device=dy.device, dtype=torch.float32) if (ctx.has_bias and ctx.needs_input_grad[3]) else None
        db_from_wgrad = db is not None and ctx.needs_input_grad[2]      # the wgrad GEMM yields it through a tap of ones
        dyp = _pack_act(dy, Cdp, 0, PAD_ZEROS, 1.0, cache=False, chan_sum=None if db_from_wgrad else db)
        dw = None
        if ctx.needs_input_grad[2]:
            dw = torch.empty_like(w)
            if _USE_WGRAD2:
                wgrad2(dyp=dyp, xp=a1p, B=B, Cdp=Cdp, Tout=T, Cp=Cp, Tp=T, Cout=Cout, Cin=Cc, K=1, dilation=1, t_off=[0],
                       dw=[dw], db=[db if db_from_wgrad else None])
            else:
                ws, wz = _wgrad_ws(lib.tdvc_conv1d_tc_wgrad_ws(Cout, Cc, 1), dy.device)
                _lib.check(lib.tdvc_conv1d_tc_wgrad(_p(dyp), _p(a1p), _p(dw), _p(ws), B, Cdp, T, Cp, T, Cout, Cc, 1, 1, 0, 0, 0,
                                                    _p(db) if db_from_wgrad else None, wz, _st()), "posconv_tc_wgrad")
        dh0 = dgb = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            Cinp16 = _ceil(Cc, 16)
            wtp = _pack_w(w, Cinp16, Cdp, True)
            dh1 = torch.empty(B, Cc, T, device=dy.device, dtype=torch.float32)
            # LeakyReLU derivative from the sign of the packed activation, in the dgrad epilogue
            _tc_conv(xp=dyp, wp=wtp, y=dh1, B=B, Tp=T, Tout=T, K=1, dilation=1, t_off=0, Cp_total=Cdp, groups=1, a_ch_off=0,
                     a_ch_stride=0, Cinp_g=Cdp, Cout_g=Cc, Coutp_g=Cinp16, bias_stride=0, out_act=ACT_NONE, out_slope=1.0,
                     out_packed=0, maskp=a1p, tm=T, cm=Cp, mask_halo=0, mask_ch_off=0, mask_ch_stride=0,
                     mask_slope=ctx.slope)
            if gb is not None:
                dh0 = torch.empty_like(h0)
                dgb = torch.empty_like(gb)
                _lib.check(lib.tdvc_film_bwd(_p(dh1), _p(h0), _p(gb), _p(dh0), _p(dgb), B, Cc, T, _st()), "film_bwd")
            else:
                dh0 = dh1
        dres = dy if ctx.needs_input_grad[4] else None
        return dh0, dgb, dw, db, dres, None


def film_posconv(h0, gb, weight, bias, x_res, slope=0.2):
    return _FilmPosconvTC.apply(h0, gb, weight, bias, x_res, float(slope))


def film_posconv_eligible(C) -> bool:
    return (_PRECISION == "bf16" and C % 16 == 0 and tc_eligible(C, C, 1, 1)
            and os.environ.get("TDVC_NO_FILM_FUSION") != "1")
