"""InstanceNorm + residual + LeakyReLU kernels at the UNETR-128 decoder shapes: CUDA-event time per call and algorithmic
bytes / time against the measured HBM copy peak (MEASURED_PEAKS.json).   python scripts/gpu_inorm_bench.py"""
import json, sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    peak = 6551.7
dev = "cuda"
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"HBM copy peak used: {peak:.0f} GB/s; tensors are 0.27-1.07 GB each (larger than the 126 MB L2)")
for (N, C, S) in [(16, 16, 128), (16, 32, 64), (16, 64, 32), (16, 128, 16)]:
    mk = lambda: torch.randn(N, S, S, S, C, device=dev).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    a, b, dy = mk(), mk(), mk()
    T = a.numel() * 2
    sa, sb = ops.inorm_stats(a), ops.inorm_stats(b)
    y = ops.inorm_apply(a, sa, b, sb, 0.01)
    rows = [("stats(a)", lambda: ops.inorm_stats(a), 1),
            ("apply lrelu(IN(a))", lambda: ops.inorm_apply(a, sa, None, None, 0.01), 2),
            ("apply lrelu(IN(a)+b)", lambda: ops.inorm_apply(a, sa, b, None, 0.01), 3),
            ("apply lrelu(IN(a)+IN(b))", lambda: ops.inorm_apply(a, sa, b, sb, 0.01), 3),
            ("bwd of lrelu(IN(a))", lambda: ops.inorm_bwd(dy, y, a, sa, None, None, 0.01), 3 + 3 + 1),
            ("bwd of lrelu(IN(a)+b)", lambda: ops.inorm_bwd(dy, y, a, sa, b, None, 0.01), 3 + 3 + 2),
            ("bwd of lrelu(IN(a)+IN(b))", lambda: ops.inorm_bwd(dy, y, a, sa, b, sb, 0.01), 4 + 4 + 2)]
    for name, fn, passes in rows:
        us = ev(fn)
        gbps = passes * T / us / 1e3
        print(f"[N={N} C={C:3d} S={S}^3, {T/1e9:.2f} GB/tensor] {name:28s} {us:8.1f} us  {gbps:7.0f} GB/s algorithmic ({passes} tensor passes) = {gbps/peak:.2f} of peak")
