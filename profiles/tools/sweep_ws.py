"""Development sweep: main-loop throughput of the weight-stationary tcgen05 conv vs N tile width.
usage: TDVC_TC_DEBUG=32 python profiles/tools/sweep_ws.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402,F401
import torch  # noqa: E402

from tdvc import ops  # noqa: E402
from tdvc._lib import ACT_LRELU  # noqa: E402

dev = torch.device("cuda", 0)
B, T, K = 16, 8960, 3
for Cg, ntile, bn in ((144, 9, 144), (144, 9, 128), (144, 5, 256), (128, 9, 128), (128, 5, 256), (144, 18, 64), (144, 8, 160)):
    N = ntile * bn
    cp = (torch.randn(B, T, Cg, device=dev) * 0.5).to(torch.bfloat16)
    wp = (torch.randn(K, N, Cg, device=dev) * 0.05).to(torch.bfloat16)
    out = torch.empty(B, T, N, device=dev, dtype=torch.bfloat16)

    def launch():
        ops._tc_conv(xp=cp, wp=wp, B=B, Tp=T, Tout=T, K=K, dilation=1, t_off=-1, Cp_total=Cg, groups=ntile, a_ch_off=0,
                     a_ch_stride=0, Cinp_g=Cg, Cout_g=bn, Coutp_g=bn, bias_stride=0, out_act=ACT_LRELU, out_slope=0.2,
                     out_packed=1, yp=out, tp_out=T, cp_out=N, out_halo=0, out_ch_off=0, out_ch_stride=bn)
    launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * T * Cg * N * K
    print(f"Cin={Cg} N={ntile}x{bn}: {ms * 1e3:7.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (padded flops)")
