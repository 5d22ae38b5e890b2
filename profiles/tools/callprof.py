"""GPU time of every C-ABI call of one eager G+D step, keyed by entry point + integer arguments (i.e. by shape) --
development aid: which SHAPES of which kernels the step's time sits in, with bytes/flops left to the reader.
Each call is bracketed by CUDA events on its stream (eager mode, PDL off so calls do not overlap).
usage: TDVC_PDL=0 python profiles/tools/callprof.py [bf16|fp32] [name-filter]"""
import collections
import ctypes as C
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402

from tdvc import _lib, ops  # noqa: E402
from tdvc.optim import FusedAdamW  # noqa: E402
from tdvc.train_step import TrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ops.set_precision(prec)
real = _lib.load()
records = []
enabled = [False]


class Proxy:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if not name.startswith("tdvc_") or name in ("tdvc_last_error", "tdvc_version", "tdvc_launch_count",
                                                    "tdvc_device_is_sm100", "tdvc_conv1d_bwd_data_ws", "tdvc_conv1d_tc_wgrad_ws"):
            return fn

        def wrapped(*args):
            if not enabled[0]:
                return fn(*args)
            key = [name]
            for a in args:
                if isinstance(a, (int, float)) and not isinstance(a, bool):
                    key.append(a if isinstance(a, int) and abs(a) < (1 << 31) else ("p" if isinstance(a, int) else round(a, 3)))
                elif hasattr(a, "_obj"):        # byref(struct)
                    s = a._obj
                    key.append(tuple((f, getattr(s, f)) for f, t in s._fields_ if t in (C.c_int32,) and getattr(s, f) != 0))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            records.append((tuple(key), e0, e1))
            return rc
        return wrapped


_lib._lib = Proxy()
dev = torch.device("cuda", 0)
G, D = bench.build_models(dev)
oG = FusedAdamW(G.parameters(), 1e-4, (0.8, 0.99)).use_grad_bank()
oD = FusedAdamW(D.parameters(), 1e-4, (0.8, 0.99)).use_grad_bank()
ts = TrainStep(G, D, bench.TRAIN, oG, oD, 100)
batch, _ = bench.to_device(bench.synth_batch(16, 8960, 100, 1234), dev)
for _ in range(2):
    ts.step(batch)
torch.cuda.synchronize()
enabled[0] = True
ts.step(batch)
torch.cuda.synchronize()
enabled[0] = False
agg = collections.defaultdict(lambda: [0, 0.0])
byname = collections.defaultdict(lambda: [0, 0.0])
for key, e0, e1 in records:
    ms = e0.elapsed_time(e1)
    agg[key][0] += 1; agg[key][1] += ms
    byname[key[0]][0] += 1; byname[key[0]][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"# {len(records)} calls, {tot:.2f} ms between events")
for k, (n, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms n={n:5d} avg={ms / n * 1e3:7.1f} us  {k}")
print("# by shape")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if flt and flt not in k[0]:
        continue
    print(f"{ms:8.3f} ms n={n:4d} avg={ms / n * 1e3:7.1f} us  {k[0]} {k[1:]}")
