"""Can a bandwidth-bound kernel hide under a tensor-bound wgrad GEMM when launched on a second stream?
timeout 120 python scripts/gpu_overlap_probe.py"""
import sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L

dev = "cuda"
M, D, Hd = 50432, 768, 3072
bf = lambda t: t.to(torch.bfloat16)
dz = bf(torch.randn(M, Hd, device=dev)); x = bf(torch.randn(M, D, device=dev))
dw = torch.zeros(Hd, D, device=dev); db = torch.zeros(Hd, device=dev)
dy = bf(torch.randn(M, D, device=dev)); xs = bf(torch.randn(M, D, device=dev)); g = torch.randn(D, device=dev)
y, mean, rstd = ops.layernorm_fwd(xs, g, g, 1e-6)
dg = torch.zeros(D, device=dev); dbb = torch.zeros(D, device=dev)
side = torch.cuda.Stream()

wgrad = lambda: ops.gemm(dz, x, M=Hd, N=D, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw, splits=8, bias_grad=db)
lnb = lambda: ops.layernorm_bwd(dy, xs, g, mean, rstd, dres=dy, dgamma=dg, dbeta=dbb)
lnf = lambda: ops.layernorm_fwd(xs, g, g, 1e-6)


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


def both(small):
    def f():
        ev = torch.cuda.Event(); ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            small()
            done = torch.cuda.Event(); done.record(side)
        wgrad()
        torch.cuda.current_stream().wait_event(done)
    return f


print(f"wgrad alone      : {timeit(wgrad):7.1f} us")
print(f"ln bwd alone     : {timeit(lnb):7.1f} us")
print(f"ln fwd alone     : {timeit(lnf):7.1f} us")
print(f"wgrad then ln bwd: {timeit(lambda: (wgrad(), lnb())):7.1f} us (serial)")
print(f"wgrad || ln bwd  : {timeit(both(lnb)):7.1f} us (two streams)")
print(f"wgrad || ln fwd  : {timeit(both(lnf)):7.1f} us (two streams)")
