"""Where the DDP step's collective time goes: torch-profiler CUDA timeline of the vit_b16 workload on N GPUs (rank 0 reports).
For every NCCL kernel: total time, and the part of it during which no compute kernel of this process was running on the
device (the exposed tail).   torchrun --nproc-per-node N scripts/gpu_ddp_timeline.py [config] [steps]"""
import argparse, os, sys
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
import bench_workloads
name = sys.argv[1] if len(sys.argv) > 1 else "vit_b16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=dev)
wl = bench_workloads.registry()[name]
args = argparse.Namespace(optimizer="ucf", bf16_allreduce=True, fp32_allreduce=False, steps=steps, cuda_graph=False, eager=True,
                          fp32_pixels=False, batch=None, bf16_decoder=False, fp32_decoder=False)
wl.build(dev, world, local, rank, args)
batch = tuple(t.to(dev) for t in wl.host_batch(rank))
for _ in range(4):
    wl.step(*batch)
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        wl.step(*batch)
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start
          and "Memcpy" not in e.name and "Memset" not in e.name]
    nccl = [e for e in ev if "nccl" in e.name.lower()]
    comp = sorted([(e.time_range.start, e.time_range.end) for e in ev if "nccl" not in e.name.lower()])
    merged = []
    for s, t in comp:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], t)
        else:
            merged.append([s, t])
    def covered(s, t):
        c = 0
        for a, b in merged:
            if b <= s: continue
            if a >= t: break
            c += min(b, t) - max(a, s)
        return c
    t0, t1 = min(e.time_range.start for e in ev), max(e.time_range.end for e in ev)
    print(f"{name} on {world} GPUs, {steps} steps: {1e-3*(t1-t0)/steps:.2f} ms per step on the device timeline; compute kernels busy "
          f"{1e-3*sum(b-a for a,b in merged)/steps:.2f} ms per step")
    by = {}
    for e in nccl:
        d = e.time_range.end - e.time_range.start
        x = d - covered(e.time_range.start, e.time_range.end)
        k = by.setdefault(e.name[:90], [0, 0.0, 0.0]); k[0] += 1; k[1] += d; k[2] += x
    for k, (n, d, x) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k}: {n/steps:.0f} launches per step, {1e-3*d/steps:.2f} ms per step in flight, {1e-3*x/steps:.2f} ms of it with no compute kernel running (exposed)")
    last = sorted(nccl, key=lambda e: e.time_range.end)
    per_step_tail = []
    # the exposed tail of each step: NCCL time after the step's last backward compute kernel and before the optimizer kernels
    print("  largest single exposed stretches (us):", sorted((round(e.time_range.end - e.time_range.start - covered(e.time_range.start, e.time_range.end)) for e in nccl), reverse=True)[:8])
dist.barrier(); dist.destroy_process_group()
