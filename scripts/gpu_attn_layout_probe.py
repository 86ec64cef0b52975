"""Is attention bound by the 128-byte-row strided layout of the packed qkv projection?  Same kernels on
[B,N,3,H,hd] (rows 4608 B apart) and on head-major [B,H,N,hd] (rows contiguous).
timeout 120 python scripts/gpu_attn_layout_probe.py"""
import sys, torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops

B, N, H, hd = 256, 197, 12, 64


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for layout in ("packed [B,N,3,H,hd]", "head-major [B,H,N,hd]"):
    if layout.startswith("packed"):
        qkv = torch.randn(B, N, 3, H, hd, device="cuda").bfloat16()
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        o_buf = torch.empty(B, N, H, hd, device="cuda", dtype=torch.bfloat16)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]
    else:
        q, k, v = [torch.randn(B, H, N, hd, device="cuda").bfloat16().transpose(1, 2) for _ in range(3)]
        dq, dk, dv = [torch.empty(B, H, N, hd, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(3)]
    o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
    do = torch.randn_like(o)
    tf = timeit(lambda: ops.attention_fwd(q, k, v, hd ** -0.5))
    tb = timeit(lambda: ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5, dq=dq, dk=dk, dv=dv))
    print(f"{layout:24s}: fwd {tf:6.1f} us   bwd {tb:6.1f} us   (o / dO stay [B,N,H,hd])", flush=True)
