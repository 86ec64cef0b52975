"""GPU bring-up check for the attention kernels vs fp32 torch math.  timeout 300 python scripts/gpu_check_attn.py"""
import sys, math
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


def ref_attn(q, k, v, scale):
    # q,k,v [B,N,H,hd] fp32
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    p = s.softmax(-1)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v)
    lse = torch.logsumexp(s, -1)
    return o, lse


def case(B, N, H, hd, packed=True, do_bwd=True):
    D = H * hd
    scale = hd ** -0.5
    if packed:
        qkv = (torch.randn(B, N, 3, H, hd, device=dev) * 1.0).to(torch.bfloat16)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    else:
        q, k, v = [(torch.randn(B, N, H, hd, device=dev)).to(torch.bfloat16) for _ in range(3)]
    tag = f"B{B} N{N} H{H} hd{hd} {'packed' if packed else 'split'}"
    o, lse = ops.attention_fwd(q, k, v, scale)
    torch.cuda.synchronize()
    qf, kf, vf = [t.float().detach().requires_grad_(True) for t in (q, k, v)]
    oref, lref = ref_attn(qf, kf, vf, scale)
    report(f"attn fwd o   {tag}", o, oref, 1.5e-2)
    report(f"attn fwd lse {tag}", lse, lref, 1e-3)
    if do_bwd:
        do = (torch.randn(B, N, H, hd, device=dev) * 0.5).to(torch.bfloat16)
        if packed:
            dqkv = torch.empty_like(qkv)
            dq, dk, dv = ops.attention_bwd(q, k, v, o, do, lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
        else:
            dq, dk, dv = ops.attention_bwd(q, k, v, o, do, lse, scale)
        torch.cuda.synchronize()
        oref.backward(do.float())
        report(f"attn bwd dq  {tag}", dq, qf.grad, 2e-2)
        report(f"attn bwd dk  {tag}", dk, kf.grad, 2e-2)
        report(f"attn bwd dv  {tag}", dv, vf.grad, 2e-2)


def bench(B, N, H, hd):
    scale = hd ** -0.5
    qkv = torch.randn(B, N, 3, H, hd, device=dev).to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, scale)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    f1 = lambda: ops.attention_fwd(q, k, v, scale)
    f2 = lambda: ops.attention_bwd(q, k, v, o, do, lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
    for name, f, mult in (("fwd", f1, 4), ("bwd", f2, 10)):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = mult * B * H * N * N * hd
        print(f"attn {name} B{B} N{N} H{H} hd{hd}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
    # torch SDPA for reference
    qq, kk, vv = [t.permute(0, 2, 1, 3).contiguous() for t in (q, k, v)]
    for _ in range(3): torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"torch SDPA fwd same shape: {ms*1e3:8.1f} us  {4*B*H*N*N*hd/ms/1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage in ("fwd", "all"):
        for (B, N, H, hd) in [(1, 128, 1, 64), (2, 197, 3, 64), (1, 256, 2, 64), (2, 50, 2, 64), (1, 512, 2, 64), (1, 1000, 1, 64),
                              (2, 197, 2, 32), (1, 384, 2, 32)]:
            case(B, N, H, hd, packed=True, do_bwd=False)
        case(2, 197, 3, 64, packed=False, do_bwd=False)
    if stage in ("bwd", "all"):
        for (B, N, H, hd) in [(1, 128, 1, 64), (2, 197, 3, 64), (1, 256, 2, 64), (2, 50, 2, 64), (1, 512, 2, 64), (1, 1000, 1, 64),
                              (2, 197, 2, 32), (1, 384, 2, 32)]:
            case(B, N, H, hd, packed=True, do_bwd=True)
        case(2, 197, 3, 64, packed=False, do_bwd=True)
    if stage in ("bench", "all") and fails == 0:
        bench(256, 197, 12, 64)
        bench(16, 1024, 12, 64)
        bench(4, 4096, 12, 64)
    print("FAILS", fails)
    sys.exit(1 if fails else 0)
