"""clock64 timeline of CTA 0 of the single-sweep attention forward kernel (first 24 key tiles per warpgroup).
    python scripts/gpu_attn_fwd2_timeline.py [B N H hd]"""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
lib = L.lib()
lib.ucf_debug_set_attn_fwd_timeline.argtypes = [ctypes.c_void_p]
B, N, H, hd = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (4, 4096, 12, 64)
qkv = torch.randn(B, N, 3, H, hd, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
for _ in range(3):
    ops.attention_fwd(q, k, v, hd ** -0.5)
tl = torch.zeros(24 * 2 * 8, dtype=torch.int64, device="cuda")
lib.ucf_debug_set_attn_fwd_timeline(tl.data_ptr())
ops.attention_fwd(q, k, v, hd ** -0.5)
torch.cuda.synchronize()
lib.ucf_debug_set_attn_fwd_timeline(None)
t = tl.cpu().view(24, 2, 8)
t0 = int(t[0, 0, 0])
names = ["S seen", "sweep done", "P arrived", "mma: P seen", "mma: PV+S issued", "epi start", "epi done"]
print(f"B{B} N{N} H{H} hd{hd}; cycles relative to warpgroup 0's first S")
for i in range(24):
    for g in range(2):
        row = [int(x) - t0 if int(x) else None for x in t[i, g, :7]]
        print(f"tile {i:2d} wg{g}: " + "  ".join(f"{n}={r}" for n, r in zip(names, row) if r is not None))
for g in range(2):
    d = [int(t[i + 1, g, 0] - t[i, g, 0]) for i in range(2, 20)]
    s = [int(t[i, g, 1] - t[i, g, 0]) for i in range(2, 20)]
    w = [int(t[i + 1, g, 0] - t[i, g, 2]) for i in range(2, 20)]
    print(f"wg{g}: tile period {sum(d)/len(d):.0f}  sweep {sum(s)/len(s):.0f}  P arrived -> next S seen {sum(w)/len(w):.0f}")
