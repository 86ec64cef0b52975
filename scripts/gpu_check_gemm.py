"""GPU bring-up check for the GEMM / LayerNorm / elementwise kernels against torch fp32 math.
Run on the B200 box:  timeout 300 python scripts/gpu_check_gemm.py"""
import sys, time
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L

torch.manual_seed(0)
dev = "cuda"
fails = 0


def report(name, got, ref, tol):
    global fails
    got = got.float(); ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    ok = err <= tol * scale and torch.isfinite(got).all().item()
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e}", flush=True)
    if not ok:
        fails += 1


def bf(x):
    return x.to(torch.bfloat16)


def run_gemm_cases():
    shapes = [(128, 128, 64), (128, 256, 128), (256, 256, 256), (200, 192, 192), (1000, 576, 192),
              (394 * 128 // 16, 768, 768), (333, 1000, 768), (4096, 3072, 768), (4096, 768, 3072), (77, 40, 72)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=dev) * 0.5
        b = torch.randn(N, K, device=dev) * 0.5
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        ab, bb, rb = bf(a), bf(b), bf(res)
        ref = ab.float() @ bb.float().t()
        for tn in (128, 256, 512):
            tag = f"M{M} N{N} K{K} tn{tn}"
            try:
                out = ops.gemm(ab, bb, M=M, N=N, K=K, bias=bias, tile_n=tn)
                report(f"fwd bias      {tag}", out, ref + bias, 1e-2)
                out = ops.gemm(ab, bb, M=M, N=N, K=K, bias=None, tile_n=tn)
                report(f"fwd nobias    {tag}", out, ref, 1e-2)
                out = ops.gemm(ab, bb, M=M, N=N, K=K, bias=bias, aux=rb, epilogue=L.EPI_BIAS_RESIDUAL, tile_n=tn)
                report(f"fwd residual  {tag}", out, ref + bias + rb.float(), 1e-2)
                out, z = ops.gemm(ab, bb, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_GELU_AUX, tile_n=tn)
                report(f"fwd gelu z    {tag}", z, ref + bias, 1e-2)
                report(f"fwd gelu u    {tag}", out, torch.nn.functional.gelu(z.float()), 1e-2)
                # dgrad: dX[M,K] = dY[M,N] * W[N,K]  -> A = dY (K-major over N), B = W stored [N(red), K(out)] MN-major
                dy = bf(torch.randn(M, N, device=dev) * 0.5)
                dref = dy.float() @ bb.float()
                out = ops.gemm(dy, bb, M=M, N=K, K=N, b_mn=True, tile_n=tn)
                report(f"dgrad         {tag}", out, dref, 1e-2)
                zz = bf(torch.randn(M, K, device=dev))
                out = ops.gemm(dy, bb, M=M, N=K, K=N, b_mn=True, aux=zz, epilogue=L.EPI_DGELU, tile_n=tn)
                zf = zz.float().requires_grad_(True)
                g = torch.autograd.grad(torch.nn.functional.gelu(zf).sum(), zf)[0]
                report(f"dgrad dgelu   {tag}", out, dref * g, 1e-2)
                # wgrad: dW[N,K] = dY^T[N,M] * X[M,K]: A = dY stored [M(red), N] MN-major, B = X stored [M(red), K] MN-major
                wref = dy.float().t() @ ab.float()
                for splits in (1, 4):
                    dw = torch.zeros(N, K, device=dev)
                    ops.gemm(dy, ab, M=N, N=K, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw, splits=splits, tile_n=tn)
                    report(f"wgrad s{splits}      {tag}", dw, wref, 2e-3)
                dw = torch.ones(N, K, device=dev)
                ops.gemm(dy, ab, M=N, N=K, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw, splits=2, tile_n=tn)
                report(f"wgrad acc     {tag}", dw, wref + 1.0, 2e-3)
                torch.cuda.synchronize()
            except Exception as e:  # noqa
                print(f"EXC  {tag}: {e}", flush=True)
                global fails
                fails += 1
                raise


def run_ln_cases():
    for rows, D in [(7, 64), (1000, 192), (50432, 768), (3000, 1024), (513, 512), (100, 2048)]:
        x = torch.randn(rows, D, device=dev) * 2 + 0.5
        g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
        for xdt in (torch.float32, torch.bfloat16):
            xx = x.to(xdt)
            y, mean, rstd = ops.layernorm_fwd(xx, g, b, 1e-6)
            ref = torch.nn.functional.layer_norm(xx.float(), (D,), g, b, 1e-6)
            report(f"ln fwd rows{rows} D{D} {xdt}", y, ref, 1e-2)
            report(f"ln mean rows{rows} D{D}", mean, xx.float().mean(-1), 1e-4)
        xb = bf(x)
        y, mean, rstd = ops.layernorm_fwd(xb, g, b, 1e-6)
        dy = bf(torch.randn(rows, D, device=dev))
        dres = bf(torch.randn(rows, D, device=dev))
        dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
        dx = ops.layernorm_bwd(dy, xb, g, mean, rstd, dres=dres, dgamma=dg, dbeta=db)
        xf = xb.float().requires_grad_(True); gf = g.clone().requires_grad_(True); bfp = b.clone().requires_grad_(True)
        out = torch.nn.functional.layer_norm(xf, (D,), gf, bfp, 1e-6)
        out.backward(dy.float())
        report(f"ln bwd dx rows{rows} D{D}", dx, xf.grad + dres.float(), 1e-2)
        report(f"ln bwd dgamma rows{rows} D{D}", dg, gf.grad, 1e-3)
        report(f"ln bwd dbeta rows{rows} D{D}", db, bfp.grad, 1e-3)


def run_misc():
    x = torch.randn(1000003, device=dev)
    report("cast f32->bf16", ops.cast_to_bf16(x[:1000000].contiguous()), x[:1000000].to(torch.bfloat16), 0)
    xb = bf(torch.randn(5000, 776, device=dev))
    report("colsum", ops.colsum(xb), xb.float().sum(0), 1e-4)
    report("colsum strided", ops.colsum(xb[:, 8:520]), xb[:, 8:520].float().sum(0), 1e-4)
    img = torch.randn(3, 3, 32, 48, device=dev)
    p = 16
    ref = img.reshape(3, 3, 2, p, 3, p).permute(0, 2, 4, 1, 3, 5).reshape(3 * 6, 3 * p * p)
    report("patchify 2d", ops.patchify(img, p), ref.to(torch.bfloat16), 0)
    vol = torch.randn(2, 2, 16, 32, 16, device=dev)
    p = 8
    ref = vol.reshape(2, 2, 2, p, 4, p, 2, p).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(2 * 16, 2 * p ** 3)
    report("patchify 3d", ops.patchify(vol, p), ref.to(torch.bfloat16), 0)


def _time(f, iters=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_epilogues():
    """The ViT-B/16 block's epilogue-carrying GEMMs (auto tile selection), batch 256 x 197 tokens."""
    print("---- epilogue GEMMs (CUDA events, 20 iters)")
    M, D, Hd = 50432, 768, 3072
    x = bf(torch.randn(M, D, device=dev)); w1 = bf(torch.randn(Hd, D, device=dev) * 0.03); b1 = torch.randn(Hd, device=dev)
    h = bf(torch.randn(M, Hd, device=dev)); w2 = bf(torch.randn(D, Hd, device=dev) * 0.03); b2 = torch.randn(D, device=dev)
    dy = bf(torch.randn(M, D, device=dev)); dz = bf(torch.randn(M, Hd, device=dev))
    res = bf(torch.randn(M, D, device=dev))
    dw1 = torch.zeros(Hd, D, device=dev); db1 = torch.zeros(Hd, device=dev)
    dw2 = torch.zeros(D, Hd, device=dev); db2 = torch.zeros(D, device=dev)
    cases = [
        ("fc1 plain bias      ", 2 * M * Hd * D, lambda: ops.gemm(x, w1, M=M, N=Hd, K=D, bias=b1)),
        ("fc1 bias+GELU(+aux) ", 2 * M * Hd * D, lambda: ops.gemm(x, w1, M=M, N=Hd, K=D, bias=b1, epilogue=L.EPI_BIAS_GELU_AUX)),
        ("fc2 bias+residual   ", 2 * M * Hd * D, lambda: ops.gemm(h, w2, M=M, N=D, K=Hd, bias=b2, aux=res, epilogue=L.EPI_BIAS_RESIDUAL)),
        ("fc2 dgrad plain     ", 2 * M * Hd * D, lambda: ops.gemm(dy, w2, M=M, N=Hd, K=D, b_mn=True)),
        ("fc2 dgrad * GELU'   ", 2 * M * Hd * D, lambda: ops.gemm(dy, w2, M=M, N=Hd, K=D, b_mn=True, aux=h, epilogue=L.EPI_DGELU)),
        ("fc1 wgrad           ", 2 * M * Hd * D, lambda: ops.gemm(dz, x, M=Hd, N=D, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw1, splits=8)),
        ("fc1 wgrad + bias grad", 2 * M * Hd * D, lambda: ops.gemm(dz, x, M=Hd, N=D, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw1, splits=8, bias_grad=db1)),
        ("fc2 wgrad + bias grad", 2 * M * Hd * D, lambda: ops.gemm(dy, h, M=D, N=Hd, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw2, splits=8, bias_grad=db2)),
    ]
    for name, fl, f in cases:
        ms = _time(f)
        print(f"{name}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)


def bench_power_scaling():
    """Same kernel on fewer SM pairs: time x clusters stays constant when the SMs are the limit and drops
    when the whole-chip power budget is (fewer active SMs -> each runs faster)."""
    print("---- time x clusters for the pair kernels (constant => per-SM bound; falling => chip power/bandwidth bound)")
    lib = L.lib()
    M, D, Hd = 50432, 768, 3072
    x = bf(torch.randn(M, D, device=dev)); w1 = bf(torch.randn(Hd, D, device=dev) * 0.03); b1 = torch.randn(Hd, device=dev)
    cases = [("fc1 plain", lambda: ops.gemm(x, w1, M=M, N=Hd, K=D, bias=b1)),
             ("fc1 gelu ", lambda: ops.gemm(x, w1, M=M, N=Hd, K=D, bias=b1, epilogue=L.EPI_BIAS_GELU_AUX))]
    for name, f in cases:
        for ncl in (74, 56, 37, 18, 8):
            lib.ucf_debug_set_gemm_max_clusters(ncl)
            ms = _time(f, iters=10)
            print(f"{name} clusters={ncl:3d}: {ms*1e3:8.1f} us   time x clusters = {ms*1e3*ncl:9.0f} us", flush=True)
    lib.ucf_debug_set_gemm_max_clusters(0)


def bench_gemm():
    print("---- GEMM throughput (CUDA events, 20 iters, inputs > L2 where stated)")
    for (M, N, K, kind) in [(50432, 2304, 768, "fwd"), (50432, 768, 768, "fwd"), (50432, 3072, 768, "fwd"),
                            (50432, 768, 3072, "fwd"), (50432, 768, 3072, "dgrad"), (50432, 3072, 768, "dgrad"),
                            (3072, 768, 50432, "wgrad"), (768, 3072, 50432, "wgrad"), (2304, 768, 50432, "wgrad"),
                            (8192, 8192, 8192, "fwd")]:
        for tn in (256, 512):
            if kind == "fwd":
                a = bf(torch.randn(M, K, device=dev)); b = bf(torch.randn(N, K, device=dev))
                f = lambda: ops.gemm(a, b, M=M, N=N, K=K, tile_n=tn)
            elif kind == "dgrad":
                a = bf(torch.randn(M, K, device=dev)); b = bf(torch.randn(K, N, device=dev))
                f = lambda: ops.gemm(a, b, M=M, N=N, K=K, b_mn=True, tile_n=tn)
            else:
                a = bf(torch.randn(K, M, device=dev)); b = bf(torch.randn(K, N, device=dev))
                o = torch.zeros(M, N, device=dev)
                spl = 8
                f = lambda: ops.gemm(a, b, M=M, N=N, K=K, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=o, splits=spl, tile_n=tn)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"gemm {kind:5s} M{M} N{N} K{K} tn{tn}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
    # cuBLAS reference point
    a = bf(torch.randn(50432, 768, device=dev)); b = bf(torch.randn(3072, 768, device=dev))
    for _ in range(3): a @ b.t()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): a @ b.t()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"cuBLAS fwd M50432 N3072 K768: {ms*1e3:8.1f} us {2*50432*3072*768/ms/1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "abi", L.lib().ucf_abi_version())
    which = sys.argv[1:] or ["misc", "ln", "gemm", "bench"]
    if "misc" in which: run_misc()
    if "ln" in which: run_ln_cases()
    if "gemm" in which: run_gemm_cases()
    if "bench" in which and fails == 0: bench_gemm()
    if "epi" in which and fails == 0: bench_epilogues()
    if "power" in which: bench_power_scaling()
    print("FAILS", fails)
    sys.exit(1 if fails else 0)
