"""Run ONE hot kernel a few times so that ncu can capture it in isolation.
python scripts/gpu_one_kernel.py {fc1_gelu|fc2_dgelu|fc1_plain|attn_fwd|attn_bwd|ln_fwd|ln_bwd|wgrad|sap_gather|var_attn|inorm|canny|conv_wgrad} [B N H hd]"""
import sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "fc1_gelu"
dev = "cuda"
M, D, Hd = 50432, 768, 3072
bf = lambda t: t.to(torch.bfloat16)
if which in ("fc1_gelu", "fc1_plain"):
    x = bf(torch.randn(M, D, device=dev)); w = bf(torch.randn(Hd, D, device=dev) * 0.03); b = torch.randn(Hd, device=dev)
    f = (lambda: ops.gemm(x, w, M=M, N=Hd, K=D, bias=b, epilogue=L.EPI_BIAS_GELU_AUX)) if which == "fc1_gelu" else \
        (lambda: ops.gemm(x, w, M=M, N=Hd, K=D, bias=b))
elif which == "fc2_dgelu":
    dy = bf(torch.randn(M, D, device=dev)); w = bf(torch.randn(D, Hd, device=dev) * 0.03); z = bf(torch.randn(M, Hd, device=dev))
    f = lambda: ops.gemm(dy, w, M=M, N=Hd, K=D, b_mn=True, aux=z, epilogue=L.EPI_DGELU)
elif which == "wgrad":
    dz = bf(torch.randn(M, Hd, device=dev)); x = bf(torch.randn(M, D, device=dev))
    dw = torch.zeros(Hd, D, device=dev); db = torch.zeros(Hd, device=dev)
    f = lambda: ops.gemm(dz, x, M=Hd, N=D, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw, splits=8, bias_grad=db)
elif which == "sap_gather":
    import numpy as np
    from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree
    import bench_workloads
    g = torch.Generator().manual_seed(0)
    edge = bench_workloads.Sap4096._edge_map(g)
    img = torch.randint(0, 256, (4096, 4096, 3), generator=g, dtype=torch.uint8).to(dev)
    qdt = FixedQuadTree(edge, 4096, device=dev)
    f = lambda: qdt.serialize_device(img, size=(16, 16, 3))
elif which == "var_attn":
    rows, V, H, hd = 16 * 512, 4, 12, 64          # UNETR 128^3 / patch 16, batch 16, 4 variables
    q = bf(torch.randn(1, 1, H, hd, device=dev)); kv = bf(torch.randn(rows, V, 2, H, hd, device=dev))
    o, lse = ops.var_attention_fwd(q, kv, hd ** -0.5)
    do = torch.randn_like(o)
    def f():
        ops.var_attention_fwd(q, kv, hd ** -0.5)
        ops.var_attention_bwd(q, kv, o, do, lse, hd ** -0.5)
elif which in ("attn_fwd", "attn_bwd"):
    B, N, H, hd = [int(a) for a in sys.argv[2:6]] if len(sys.argv) >= 6 else (256, 197, 12, 64)
    qkv = bf(torch.randn(B, N, 3, H, hd, device=dev))
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
    do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
    f = (lambda: ops.attention_fwd(q, k, v, hd ** -0.5)) if which == "attn_fwd" else \
        (lambda: ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2]))
elif which in ("ln_fwd", "ln_bwd"):
    x = bf(torch.randn(M, D, device=dev)); g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
    dy = bf(torch.randn(M, D, device=dev)); dg = torch.zeros(D, device=dev); dbb = torch.zeros(D, device=dev)
    f = (lambda: ops.layernorm_fwd(x, g, b, 1e-6)) if which == "ln_fwd" else \
        (lambda: ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dy, dgamma=dg, dbeta=dbb))
elif which == "inorm":
    # UNETR-128 decoder2 / encoder1 block bodies: batch 16, 16 channels, 128^3 voxels, channels-last bf16 (1.07 GB per tensor)
    from ucf_vit_b200 import functional as UF
    N, C, S = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (16, 16, 128)
    mk = lambda: bf(torch.randn(N, S, S, S, C, device=dev)).permute(0, 4, 1, 2, 3)
    a, b, dy = mk(), mk(), mk()
    sa, sb = ops.inorm_stats(a), ops.inorm_stats(b)
    y = ops.inorm_apply(a, sa, b, sb, 0.01)
    def f():
        ops.inorm_stats(a)
        ops.inorm_apply(a, sa, None, None, 0.01)
        ops.inorm_apply(a, sa, b, sb, 0.01)
        ops.inorm_bwd(dy, y, a, sa, None, None, 0.01)
        ops.inorm_bwd(dy, y, a, sa, b, sb, 0.01)
elif which == "canny":
    import numpy as np
    sys.path.insert(0, "tests")
    from test_gpu_canny import _scene
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    img = torch.as_tensor(_scene(n, np.random.default_rng(n))).to(dev)
    def f():
        ops.canny_u8(ops.gaussian_blur_u8(img, 5), 60, 110)
        ops.gaussian_blur_u8(img, 3)
elif which == "conv_wgrad":
    N, Ci, Co, S = (int(a) for a in sys.argv[2:6]) if len(sys.argv) >= 6 else (16, 32, 16, 128)
    x = bf(torch.randn(N, S, S, S, Ci, device=dev)).permute(0, 4, 1, 2, 3)
    dy = bf(torch.randn(N, S, S, S, Co, device=dev)).permute(0, 4, 1, 2, 3)
    f = lambda: ops.conv3d_wgrad(x, dy)
else:
    raise SystemExit("unknown kernel " + which)
for _ in range(4):
    f()
torch.cuda.synchronize()
print("ok", which)
