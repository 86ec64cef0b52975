import sys, ctypes, torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
B, N, H, hd = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 197, 12, 64
qkv = torch.randn(B, N, 3, H, hd, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
tl = torch.zeros(64, dtype=torch.int64, device="cuda")
lib = L.lib()
lib.ucf_debug_set_attn_fwd_timeline.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attention_fwd(q, k, v, hd ** -0.5)
lib.ucf_debug_set_attn_fwd_timeline(tl.data_ptr())
ops.attention_fwd(q, k, v, hd ** -0.5)
torch.cuda.synchronize()
lib.ucf_debug_set_attn_fwd_timeline(None)
t = tl.cpu().view(8, 8)
t0 = t[0, 0].item()
names = ["mma:S issue", "mma:PV issue", "wg:S visible", "wg:pass1 done", "wg:P written"]
for g in range(2):
    for i in range(4):
        print(f"wg{g} tile{i}", " ".join(f"{names[j]}={t[g * 4 + i, j].item() - t0}" for j in range(5)))
