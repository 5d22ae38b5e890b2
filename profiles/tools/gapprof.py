"""Where the idle time of one CUDA-graph replay of the G+D step sits -- development aid.
Traces one replay with CUPTI (torch.profiler), sorts GPU activities by start time and reports, per kernel family,
its busy time and the idle gap that FOLLOWS / PRECEDES it.
usage: python profiles/tools/gapprof.py [bf16|fp32]"""
import collections
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from tdvc import ops  # noqa: E402
from tdvc.optim import FusedAdamW  # noqa: E402
from tdvc.train_step import GraphedTrainStep, TrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ops.set_precision(prec)
dev = torch.device("cuda", 0)
G, D = bench.build_models(dev)
oG = FusedAdamW(G.parameters(), 1e-4, (0.8, 0.99))
oD = FusedAdamW(D.parameters(), 1e-4, (0.8, 0.99))
ts = TrainStep(G, D, bench.TRAIN, oG, oD, 100)
batch, _ = bench.to_device(bench.synth_batch(16, 8960, 100, 1234), dev)
gs = GraphedTrainStep(ts, batch, warmup=2)
for _ in range(2):
    gs.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = collections.Counter(); cnt = collections.Counter(); gap_after = collections.Counter(); gap_before = collections.Counter()
end = None
prev = None
tot_busy = tot_gap = 0.0
hist = collections.Counter()
for e in evs:
    name = re.sub(r"\(.*", "", e.name)[:70]
    d = e.time_range.end - e.time_range.start
    busy[name] += d; cnt[name] += 1; tot_busy += d
    if end is not None:
        g = max(0.0, e.time_range.start - end)
        tot_gap += g
        gap_after[prev] += g
        gap_before[name] += g
        hist[min(int(g), 20)] += 1
    end = max(end or 0, e.time_range.end)
    prev = name
print(f"# one replay: span {(t1 - t0) / 1e3:.2f} ms, busy {tot_busy / 1e3:.2f} ms, idle {tot_gap / 1e3:.2f} ms over {len(evs)} activities")
print("# gap histogram (us -> count):", dict(sorted(hist.items())))
print("# family: n, busy ms, idle ms AFTER it, idle ms BEFORE it")
for k, b in sorted(busy.items(), key=lambda kv: -(kv[1] + gap_after[kv[0]]))[:40]:
    print(f"n={cnt[k]:5d} busy={b / 1e3:7.3f} after={gap_after[k] / 1e3:7.3f} before={gap_before[k] / 1e3:7.3f}  {k}")
