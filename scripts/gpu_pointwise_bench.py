"""1x1x1 convolution kernels at the UNETR-128 decoder's shapes (batch 16, 128^3 voxels, channels-last bf16): CUDA-event time
and algorithmic bytes / time against the measured HBM copy peak.   python scripts/gpu_pointwise_bench.py"""
import json, sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
dev = "cuda"
def ev(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
N, S = 16, 128
V = N * S ** 3
for Ci, Co in [(32, 16), (16, 32), (16, 4), (4, 16)]:
    x = torch.randn(N, S, S, S, Ci, device=dev).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    dy = torch.randn(N, S, S, S, Co, device=dev).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    w = torch.randn(Co, Ci, device=dev) * 0.1
    b = torch.randn(Co, device=dev)
    t_f = ev(lambda: ops.pointwise_conv(x, w, b))
    t_w = ev(lambda: ops.pointwise_conv_wgrad(x, dy, with_bias=True))
    bytes_f = 2.0 * V * (Ci + Co)
    print(f"{Ci:2d}->{Co:2d} at 16 x 128^3: forward {t_f:7.1f} us = {bytes_f/t_f/1e3:5.0f} GB/s ({bytes_f/t_f/1e3/peak:.2f} of peak); "
          f"weight+bias gradient {t_w:7.1f} us = {bytes_f/t_w/1e3:5.0f} GB/s ({bytes_f/t_w/1e3/peak:.2f} of peak)")
