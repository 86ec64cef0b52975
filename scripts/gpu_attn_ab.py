"""A/B of the shipped attention forward kernel against an experimental entry point with the same signature.
    python scripts/gpu_attn_ab.py ucf_debug_attention_fwd_split"""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
lib = L.lib()
alt_name = sys.argv[1]
alt = getattr(lib, alt_name)
alt.restype = ctypes.c_int
alt.argtypes = lib.ucf_attention_fwd.argtypes
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(fn, q, k, v, o, lse, scale):
    B, Nq, H, hd = q.shape
    rc = fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, H, Nq, k.shape[1], hd,
            *ops._bnhd(q), *ops._bnhd(k), *ops._bnhd(v), *ops._bnhd(o), float(scale), ops._stream())
    assert rc == 0, lib.ucf_last_error()

def timeit(f, n=10):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3

for (B, N, H, hd) in [(256, 197, 12, 64), (16, 1024, 12, 64), (4, 4096, 12, 64), (64, 197, 16, 32), (4, 4096, 24, 32)]:
    qkv = torch.randn(B, N, 3, H, hd, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o1 = torch.empty(B, N, H, hd, dtype=torch.bfloat16, device="cuda"); o2 = torch.empty_like(o1)
    l1 = torch.empty(B, H, N, device="cuda"); l2 = torch.empty_like(l1)
    scale = hd ** -0.5
    run(lib.ucf_attention_fwd, q, k, v, o1, l1, scale); run(alt, q, k, v, o2, l2, scale)
    torch.cuda.synchronize()
    err = (o1.float() - o2.float()).abs().max().item(); lerr = (l1 - l2).abs().max().item()
    ta = timeit(lambda: run(lib.ucf_attention_fwd, q, k, v, o1, l1, scale))
    tb = timeit(lambda: run(alt, q, k, v, o2, l2, scale))
    ta2 = timeit(lambda: run(lib.ucf_attention_fwd, q, k, v, o1, l1, scale))
    print(f"B{B} N{N} H{H} hd{hd}: shipped {ta:7.1f} / {ta2:7.1f} us   {alt_name} {tb:7.1f} us   ({tb / min(ta, ta2):.3f}x)   max|do| {err:.2e} max|dlse| {lerr:.2e}", flush=True)
