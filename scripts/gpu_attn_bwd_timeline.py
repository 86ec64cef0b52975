import sys, ctypes, torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
B, N, H, hd = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 197, 12, 64
qkv = torch.randn(B, N, 3, H, hd, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
do = torch.randn_like(o)
tl = torch.zeros(64, dtype=torch.int64, device="cuda")
lib = L.lib()
lib.ucf_debug_set_attn_bwd_timeline.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5)
lib.ucf_debug_set_attn_bwd_timeline(tl.data_ptr())
ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5)
torch.cuda.synchronize()
lib.ucf_debug_set_attn_bwd_timeline(None)
t = tl.cpu().view(8, 8)[:4, :6]
t0 = t[0, 0].item()
names = ["mma:S/dP issue", "mma:PdS ready->issue dV/dK/dQ", "cmp:S/dP visible", "cmp:P/dS written", "cmp:dQ visible", "cmp:dQ to TMA"]
for i in range(4):
    print("tile", i, " ".join(f"{names[j]}={t[i, j].item() - t0}" for j in range(6)))
