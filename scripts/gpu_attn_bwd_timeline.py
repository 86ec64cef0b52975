"""clock64 timeline of CTA 0's first 8 tiles of the attention backward kernel (profiling aid).
timeout 60 python scripts/gpu_attn_bwd_timeline.py [B N]"""
import sys, ctypes, torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
B, N, H, hd = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 197, 12, 64
qkv = torch.randn(B, N, 3, H, hd, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
do = torch.randn_like(o)
tl = torch.zeros(256, dtype=torch.int64, device="cuda")
lib = L.lib()
lib.ucf_debug_set_attn_bwd_timeline.argtypes = [ctypes.c_void_p]
for _ in range(2):
    ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5)
lib.ucf_debug_set_attn_bwd_timeline(tl.data_ptr())
ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5)
torch.cuda.synchronize()
lib.ucf_debug_set_attn_bwd_timeline(None)
t = tl.cpu().view(8, 32)
t0 = t[0, 7].item()
names = {7: "mma:sdp enter", 8: "kv_full", 9: "qdo_full", 0: "sdp_empty->issue", 1: "mma:pds_full(+dkv_empty)", 11: "dq_empty",
         12: "mma:tile issued", 2: "cmp:sdp_full", 13: "cmp:math done", 14: "cmp:p_free", 3: "cmp:pds arrive",
         4: "cmp:dq_full", 5: "cmp:dq flushed"}
for i in range(8):
    print(f"tile {i}: " + "  ".join(f"{nm}={t[i, k].item() - t0 if t[i, k].item() else None}" for k, nm in names.items()))
    print("        sdp_full seen per warp:", [t[i, 24 + w].item() - t0 for w in range(8)])
    print("        pds arrive per warp:   ", [t[i, 16 + w].item() - t0 for w in range(8)])
