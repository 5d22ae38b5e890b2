"""Runs the dominant convolution (FiLM cond_var.0: 136->136, k=3, B=16, T=8960) alone, for ncu --set full.
usage: python profiles/tools/prof_dominant.py [fp32|bf16]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402,F401
import torch  # noqa: E402

from tdvc import ops  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ops.set_precision(prec)
dev = torch.device("cuda", 0)
B, C, T = 16, 136, 8960
x = torch.randn(B, C, T, device=dev)
w = torch.randn(C, C, 3, device=dev) * 0.05
b = torch.zeros(C, device=dev)
with torch.no_grad():
    for _ in range(6):
        y = ops.conv1d(x, w, b, padding=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s), torch.no_grad():
    ops.conv1d(x, w, b, padding=1)
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(10):
            y = ops.conv1d(x, w, b, padding=1)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{prec}: {ms * 1e3:.1f} us per conv (graph replay, incl. pack_w + conv), {2 * B * T * C * C * 3 / ms / 1e9:.1f} TFLOP/s")
