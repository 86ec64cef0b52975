"""Device-time table (torch.profiler, CUDA activities) of one bench workload's training step.
   python scripts/gpu_kernel_profile_wl.py CONFIG [steps] [--bf16-decoder | --fp32-decoder]"""
import argparse, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
import bench_workloads
name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2
wl = bench_workloads.registry()[name]
args = argparse.Namespace(optimizer="ucf", bf16_allreduce=False, steps=n, bf16_decoder="--bf16-decoder" in sys.argv, fp32_decoder="--fp32-decoder" in sys.argv,
                          cuda_graph=False, eager=True, fp32_pixels=False, batch=None)
if "--cudnn-benchmark" in sys.argv:
    torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
if wl.uses_fsdp:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
wl.build(dev, 1, 0, 0, args)
if "--channels-last" in sys.argv:
    m = wl.model
    m.to(memory_format=torch.channels_last_3d)
    enc1_fwd = m.encoder1.forward
    m.encoder1.forward = lambda x: enc1_fwd(x.contiguous(memory_format=torch.channels_last_3d))
batch = tuple(t.to(dev) for t in wl.host_batch(0))
for _ in range(3):
    wl.step(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(n):
        wl.step(*batch)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(e.key, e.device_time_total / n, e.count / n) for e in ka if e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"{name}: {tot/1e3:.2f} ms of device time per step over {sum(r[2] for r in rows):.0f} launches")
for k, t, c in rows[:45]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% x{c:5.0f}  {k[:150]}")
print("--- by operator (shapes)")
ops_ = [(e.key, str(e.input_shapes)[:120], e.device_time_total / n, e.count / n) for e in prof.key_averages(group_by_input_shape=True)
        if e.device_time_total > 0 and any(w in e.key.lower() for w in ("conv", "norm", "leaky", "copy", "contiguous", "cat", "to_copy", "fill", "zero"))]
ops_.sort(key=lambda r: -r[2])
for k, s, t, c in ops_[:40]:
    print(f"{t/1e3:9.3f} ms x{c:4.0f} {k[:40]:40s} {s}")
