"""Runs the dominant launch (the 9 FiLM cond_var.0 convs of the full-rate decoder stage, one tcgen05 launch in
bf16 mode) alone -- the same code path bench.py's `roofline` leg times -- for `ncu --set full`.
usage: python profiles/tools/prof_dominant.py [fp32|bf16]"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402

from tdvc import ops  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ops.set_precision(prec)
dev = torch.device("cuda", 0)
args = argparse.Namespace(precision=prec)
r = bench.dominant_kernel_roofline(dev, 16, 8960, bench.peaks(), args)
print(r)
