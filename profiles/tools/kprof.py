"""Per-kernel GPU time of one G+D step via torch.profiler (CUPTI) -- a development aid for choosing what
to optimise next; the judged evidence is the ncu launch list (profiles/*launches*).
usage: python profiles/tools/kprof.py [fp32|bf16] [out.txt]"""
import collections
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (sets sys.path for the package)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from tdvc import ops  # noqa: E402
from tdvc.optim import FusedAdamW  # noqa: E402
from tdvc.train_step import TrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ops.set_precision(prec)
dev = torch.device("cuda", 0)
G, D = bench.build_models(dev)
oG = FusedAdamW(G.parameters(), 1e-4, (0.8, 0.99))
oD = FusedAdamW(D.parameters(), 1e-4, (0.8, 0.99))
ts = TrainStep(G, D, bench.TRAIN, oG, oD, 100)
batch, _ = bench.to_device(bench.synth_batch(16, 8960, 100, 1234), dev)
for _ in range(2):
    ts.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)
        agg[name][0] += 1
        agg[name][1] += ev.device_time
        tot += ev.device_time
lines = [f"# precision={prec}: one G+D step, kernel time total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} GPU activities"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append(f"{t / tot * 100:6.2f}%  {t / 1e3:9.3f} ms  n={n:5d}  avg={t / n:8.1f} us  {k[:100]}")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
