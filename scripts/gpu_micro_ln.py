"""LayerNorm fwd/bwd timing on the ViT-B block shape, inputs cycled through > L2 worth of buffers.
timeout 120 python scripts/gpu_micro_ln.py"""
import sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops

R, D, NB = 50432, 768, 6          # 6 x 77 MB per tensor kind > 126 MB L2
xs = [torch.randn(R, D, device="cuda").bfloat16() for _ in range(NB)]
dys = [torch.randn(R, D, device="cuda").bfloat16() for _ in range(NB)]
g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
y, mean, rstd = ops.layernorm_fwd(xs[0], g, b, 1e-6)
dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")


def timeit(f, n=30):
    for i in range(3):
        f(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


t = timeit(lambda i: ops.layernorm_fwd(xs[i % NB], g, b, 1e-6))
print(f"layernorm fwd: {t:6.1f} us   {2 * R * D * 2 / t / 1e6:6.2f} TB/s (1 read + 1 write)")
t = timeit(lambda i: ops.layernorm_bwd(dys[i % NB], xs[i % NB], g, mean, rstd, dres=dys[(i + 1) % NB], dgamma=dg, dbeta=db))
print(f"layernorm bwd: {t:6.1f} us   {4 * R * D * 2 / t / 1e6:6.2f} TB/s (3 reads + 1 write)")
t = timeit(lambda i: xs[i % NB].copy_(dys[i % NB]))
print(f"torch copy   : {t:6.1f} us   {2 * R * D * 2 / t / 1e6:6.2f} TB/s")
