"""Which Python lines launch the torch (non-tdvc) kernels of one G+D step -- development aid.
Runs the eager step exactly as GraphedTrainStep captures it (grad-bank optimisers) under torch.profiler with stacks and
prints, per aten op that reaches the GPU, the innermost repo frame that called it.
usage: python profiles/tools/torchsrc.py [bf16|fp32]"""
import collections
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from tdvc import ops  # noqa: E402
from tdvc.optim import FusedAdamW  # noqa: E402
from tdvc.train_step import TrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ops.set_precision(prec)
dev = torch.device("cuda", 0)
G, D = bench.build_models(dev)
oG = FusedAdamW(G.parameters(), 1e-4, (0.8, 0.99)).use_grad_bank()
oD = FusedAdamW(D.parameters(), 1e-4, (0.8, 0.99)).use_grad_bank()
ts = TrainStep(G, D, bench.TRAIN, oG, oD, 100)
batch, _ = bench.to_device(bench.synth_batch(16, 8960, 100, 1234), dev)
for _ in range(2):
    ts.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    ts.step(batch)
    torch.cuda.synchronize()
agg = collections.Counter()
tim = collections.Counter()
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CPU or not ev.name.startswith("aten::"):
        continue
    if not ev.kernels:                                   # only ops that launched something themselves
        continue
    frame = "?"
    for fr in (ev.stack or []):
        if "/root/repo" in fr or "td-vc-gan_b200" in fr or "bench.py" in fr:
            frame = fr.replace(REPO + "/", "")
            break
    else:
        if ev.stack:
            frame = ev.stack[0][-90:]
    key = (ev.name, frame[:110])
    agg[key] += len(ev.kernels)
    tim[key] += sum(k.duration for k in ev.kernels)
print(f"# torch-launched kernels in one step: {sum(agg.values())}, {sum(tim.values()) / 1e3:.2f} ms")
for key, n in agg.most_common(45):
    print(f"n={n:5d} {tim[key] / 1e3:7.3f} ms  {key[0]:28s} {key[1]}")
