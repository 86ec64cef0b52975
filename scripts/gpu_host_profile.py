"""Where the host time of one ViT-B/16 training step goes (cProfile over a few steps; the GPU runs behind).
    python scripts/gpu_host_profile.py [batch]"""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
import bench
from ucf_vit_b200.simple.arch import VIT
from ucf_vit_b200.utils.fused_attn import FusedAttn
from ucf_vit_b200.utils.misc import configure_optimizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
model = VIT(**bench.CFG, mlp_ratio=4, class_token=True, twoD=True, default_vars=["r", "g", "b"], FusedAttn_option=FusedAttn.FLASH).to(dev).train()
opt = configure_optimizer(model, 1e-4, 0.9, 0.95, 1e-5, fused="ucf")
x = torch.rand(B, 3, 224, 224, device=dev) * 255
y = torch.randint(0, 1000, (B,), device=dev)
lossf = torch.nn.CrossEntropyLoss()
def step():
    logits = model(x, ["r", "g", "b"])
    loss = lossf(logits.float(), y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
for name, fn in (("forward", lambda: model(x, ["r", "g", "b"])),):
    torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); dt = time.perf_counter() - t
    print(f"{name}: host {dt*1e3:.2f} ms")
torch.cuda.synchronize()
t = time.perf_counter(); logits = model(x, ["r", "g", "b"]); t1 = time.perf_counter(); loss = lossf(logits.float(), y); t2 = time.perf_counter()
opt.zero_grad(set_to_none=True); t3 = time.perf_counter(); loss.backward(); t4 = time.perf_counter(); opt.step(); t5 = time.perf_counter()
print(f"fwd {1e3*(t1-t):.2f}  loss {1e3*(t2-t1):.2f}  zero_grad {1e3*(t3-t2):.2f}  bwd {1e3*(t4-t3):.2f}  opt {1e3*(t5-t4):.2f} ms (host)")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
    torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
