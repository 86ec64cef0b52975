"""clock64 timeline of CTA 0's first items of the attention forward kernels (profiling aid).
timeout 60 python scripts/gpu_attn_fwd_timeline.py [B N]      (N <= 256: short-key kernel; else general kernel)"""
import sys, ctypes, torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
B, N, H, hd = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 197, 12, 64
qkv = torch.randn(B, N, 3, H, hd, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
tl = torch.zeros(256, dtype=torch.int64, device="cuda")
lib = L.lib()
lib.ucf_debug_set_attn_fwd_timeline.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attention_fwd(q, k, v, hd ** -0.5)
lib.ucf_debug_set_attn_fwd_timeline(tl.data_ptr())
ops.attention_fwd(q, k, v, hd ** -0.5)
torch.cuda.synchronize()
lib.ucf_debug_set_attn_fwd_timeline(None)
if N <= 256:
    t = tl.cpu()[:128].view(8, 16)
    t0 = t[0, 8].item()
    names = {8: "mma:S issue", 7: "wg:item start", 0: "wg:S visible", 1: "pass1 done", 2: "P0 published", 9: "mma:PV0 issue", 11: "P1 in regs",
             3: "PV0 retired seen", 4: "P1 published", 10: "mma:PV1 issue", 5: "PV1 retired seen", 6: "O stored"}
    for k_ in range(4):
        for g in range(2):
            print(f"item{k_} wg{g}: " + "  ".join(f"{nm}={t[k_ * 2 + g, i].item() - t0}" for i, nm in names.items()))
else:
    t = tl.cpu()[:64].view(8, 8)
    t0 = t[0, 0].item()
    names = ["mma:S issue", "mma:PV issue", "wg:S visible", "wg:pass1 done", "wg:P written"]
    for g in range(2):
        for i in range(4):
            print(f"wg{g} tile{i}", " ".join(f"{names[j]}={t[g * 4 + i, j].item() - t0}" for j in range(5)))
