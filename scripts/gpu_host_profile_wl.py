"""cProfile of the host side of one bench workload's training step.   python scripts/gpu_host_profile_wl.py CONFIG [steps]"""
import argparse, cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, ".")
import bench_workloads
name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = bench_workloads.registry()[name]
args = argparse.Namespace(optimizer="ucf", bf16_allreduce=False, steps=n)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
if wl.uses_fsdp:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
wl.build(dev, 1, 0, 0, args)
batch = tuple(t.to(dev) for t in wl.host_batch(0))
for _ in range(3):
    wl.step(*batch)
torch.cuda.synchronize()
t = time.perf_counter(); wl.step(*batch); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"{name}: host enqueue {1e3*(t1-t):.2f} ms, step incl. GPU {1e3*(t2-t):.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(n):
    wl.step(*batch); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(40)
