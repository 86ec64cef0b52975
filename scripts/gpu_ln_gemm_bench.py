"""LayerNorm -> QKV: separate kernels (LN forward, then GEMM) against the fused path (statistics pass + EPI_LN GEMM).
    python scripts/gpu_ln_gemm_bench.py"""
import sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, functional as UF, _lib as L
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(f, n=10):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

for (M, D, N, tag) in [(50432, 768, 2304, "ViT-B/16 batch 256 qkv"), (12800, 1024, 3072, "ViT-L encoder qkv, 49 tokens x 256"),
                       (8192, 768, 2304, "UNETR-128 batch 16 qkv")]:
    x = (torch.randn(M, D, device=dev) + 0.5).to(torch.bfloat16)
    g, b = torch.randn(D, device=dev) * 0.2 + 1, torch.randn(D, device=dev) * 0.1
    w, bias = torch.randn(N, D, device=dev) * 0.03, torch.randn(N, device=dev) * 0.1
    w16 = w.to(torch.bfloat16)
    wg, colsum, bfold = UF.fold_layernorm(w, bias, g, b)
    def unfused():
        h, _, _ = ops.layernorm_fwd(x, g, b, 1e-6)
        return ops.gemm(h, w16, M=M, N=N, K=D, bias=bias)
    def fused():
        mean, rstd = ops.layernorm_stats(x, 1e-6)
        return ops.ln_gemm(x, wg, bfold, colsum, mean, rstd)
    ya, yb = unfused().float(), fused().float()
    rel = ((ya - yb).norm() / ya.norm()).item()
    ta, tb = timeit(unfused), timeit(fused)
    t_ln = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-6)); t_st = timeit(lambda: ops.layernorm_stats(x, 1e-6))
    print(f"{tag}: LN + GEMM {ta:7.1f} us   stats + fused GEMM {tb:7.1f} us  ({tb / ta:.3f}x)   LN alone {t_ln:.1f} us, stats alone {t_st:.1f} us   "
          f"rel-L2 fused vs unfused {rel:.2e}", flush=True)
