import os, sys, collections
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench, torch
from torch.profiler import profile, ProfilerActivity
from tdvc import ops
from tdvc.optim import FusedAdamW
from tdvc.train_step import TrainStep
ops.set_precision("bf16")
dev = torch.device("cuda", 0)
G, D = bench.build_models(dev)
oG = FusedAdamW(G.parameters(), 1e-4, (0.8, 0.99)); oD = FusedAdamW(D.parameters(), 1e-4, (0.8, 0.99))
ts = TrainStep(G, D, bench.TRAIN, oG, oD, 100)
batch, _ = bench.to_device(bench.synth_batch(16, 8960, 100, 1234), dev)
for _ in range(2): ts.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU], with_stack=True) as prof:
    ts.step(batch)
    torch.cuda.synchronize()
cnt = collections.Counter()
for ev in prof.events():
    if ev.name in ("aten::zero_", "aten::fill_", "aten::zeros", "aten::zeros_like"):
        st = [s for s in (ev.stack or []) if "td-vc-gan_b200" in s or "bench" in s or "torch/autograd" in s][:3]
        cnt[(ev.name, tuple(st))] += 1
for (name, st), n in cnt.most_common(12):
    print(n, name, " | ".join(s.split("/")[-1][:80] for s in st))
