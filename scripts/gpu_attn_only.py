"""Run only the attention kernels a few times (ncu target).  python scripts/gpu_attn_only.py fwd|bwd B N H hd [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops
which = sys.argv[1]
B, N, H, hd = [int(a) for a in sys.argv[2:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
qkv = torch.randn(B, N, 3, H, hd, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
for _ in range(reps):
    if which == "fwd":
        ops.attention_fwd(q, k, v, hd ** -0.5)
    else:
        ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
torch.cuda.synchronize()
print("done")
