import sys, torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L
B, N, H, hd = 256, 197, 12, 64
qkv = torch.randn(B, N, 3, H, hd, device="cuda").bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
lib = L.lib()
for st in (0, 1500, 3000, 4000, 4500, 5000, 6000, 8000):
    lib.ucf_debug_set_attn_fwd_stagger(st)
    for _ in range(3): ops.attention_fwd(q, k, v, hd ** -0.5)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.attention_fwd(q, k, v, hd ** -0.5)
    e1.record(); torch.cuda.synchronize()
    print(f"stagger {st}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
