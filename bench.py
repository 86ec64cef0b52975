#!/usr/bin/env python
"""Benchmarks of the B200-native ViT training hot path, one workload per BASELINE.json config.

    python bench.py [--config NAME] [--gpus N] [--steps K] [--warmup W] [--impl reference]

Default (--config vit_b16) is the headline: BASELINE.json configs[1], the ViT-B/16 ImageNet-shape classification
training step on N B200s, data-parallel, through the reference's module API backed by this package's sm_100a kernels.
Other configs (bench_workloads.py): vit_tiny (configs[0]), mae_vitl_fsdp (configs[2]), unetr_128 (configs[3]),
sap_4096_L4096 / sap_4096_L1024 and diffusion_fsdp (configs[4]).

One "step" = forward + loss + backward + AdamW update on one synthetic batch per GPU, bf16 tensor-core math with fp32
accumulation and fp32 master weights.  Prints ONE JSON line (see DESIGN.md "Measurement").

  value     whole-job samples/s, inputs resident in HBM, K steps timed with CUDA events between
            barrier + synchronize pairs, max over ranks
  e2e       same metric through the public module API from pinned HOST buffers: every step
            copies its batch host->device (prefetched on a side stream) and reads the loss back
  roofline  dominant kernel = the tcgen05 GEMM family; achieved = algorithmic FLOPs of every
            GEMM launch / summed per-launch CUDA-event durations inside the timed region
  cpu_baseline / --impl reference
            the oracle's CPU restatement of the reference path (fp32, all host threads) on a
            bounded sample of the same workload
"""
import argparse
import json
import os
import statistics
import subprocess
import threading
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench_workloads  # noqa: E402

METRIC = "ViT train imgs/s at 1/2/4/8 B200; attn+MLP TFLOP/s vs bf16 peak"
# kept for scripts that import them (the headline workload)
CFG = dict(img_size=[224, 224], patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12)
PER_GPU_BATCH = 256


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1
    rate, cores, sec, sample = wl.cpu_rate(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": wl.unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.workload, "name": wl.name, "sample": f"batch {wl.cpu_batch} per step on the host CPU"},
        "cpu_baseline": {"value": rate, "unit": wl.unit, "cores": cores, "kind": wl.cpu_kind, "sample": sample},
        "e2e": {"value": rate, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, board power and throttle reasons sampled DURING the timed region.

    Primary source: NVML in a background thread of this process (pynvml; ~5 ms period -- the main thread
    sits in cudaStreamSynchronize with the GIL released).  Fallback: an `nvidia-smi -lms` child.  The sampler
    is started before warm-up; `mark()` / `stop()` bracket the timed region and only samples taken between
    them are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/ucf_clocks_{os.getpid()}.csv"
        self.samples, self.t_mark, self.thread, self.stop_flag, self.nvml = [], None, None, None, None

    def _nvml_loop(self):
        nv, h = self.nvml, self.handle
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                mask = get_reasons(h)
                self.samples.append((time.time(), float(sm), pw, [k for k, b in bits.items() if mask & b]))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.nvml, self.stop_flag = nv, threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu=timestamp,{self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark(self):
        self.t_mark = time.time()

    def stop(self):
        t_end = time.time()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        t0 = self.t_mark or 0.0
        sm, mx, reasons, pw = [], [], set(), []
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            for ts, c, w, rs in self.samples:
                if t0 <= ts <= t_end:
                    sm.append(c); pw.append(w); reasons.update(rs)
            mx = [self.max_sm]
            out["source"] = "nvml"
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            self.f.close()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            import datetime
            for ln in open(self.path):
                parts = [q.strip() for q in ln.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if not (t0 - 0.05 <= ts <= t_end + 0.05):
                        continue
                    sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            try:
                os.remove(self.path)
            except OSError:
                pass
            out["source"] = "nvidia-smi"
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
        return out


class GemmTimer:
    """Per-launch CUDA-event timing of the GEMM family, taken inside the library on the launching stream (the launches
    come from ucf_block_fwd / ucf_block_bwd in C++ as well as from the Python op wrappers)."""

    def __init__(self):
        import ctypes
        from ucf_vit_b200 import _lib
        self.lib, self.ct, self.enabled_ = _lib.lib(), ctypes, False
        self.lib.ucf_debug_gemm_timing.argtypes = [ctypes.c_int]
        self.lib.ucf_debug_gemm_timing_summary.argtypes = [ctypes.c_void_p] * 3

    def install(self):
        pass

    @property
    def enabled(self):
        return self.enabled_

    @enabled.setter
    def enabled(self, v):
        self.enabled_ = bool(v)
        self.lib.ucf_debug_gemm_timing(int(bool(v)))

    def summary(self):
        ct = self.ct
        fl, ms, n = ct.c_double(0), ct.c_double(0), ct.c_longlong(0)
        self.lib.ucf_debug_gemm_timing_summary(ct.byref(fl), ct.byref(ms), ct.byref(n))
        return fl.value, ms.value, n.value


def run_gpu_arm(args, wl):
    import torch.distributed as dist
    from ucf_vit_b200 import _lib
    from ucf_vit_b200.dataloaders.prefetch import DevicePrefetcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    need_pg = world > 1 or wl.uses_fsdp
    if need_pg:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        if world == 1:
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        else:
            dist.init_process_group("nccl", device_id=dev)

    if getattr(wl, "default_cuda_graph", False) and world == 1 and not args.eager:
        args.cuda_graph = True
    wl.build(dev, world, local, rank, args)
    host = wl.host_batch(rank)
    resident = tuple(t.to(dev) for t in host)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if getattr(wl, "flush_l2", False) else None

    graphed = None
    if args.cuda_graph:
        if world > 1 or wl.uses_fsdp:
            raise SystemExit("--cuda-graph: single-GPU, non-FSDP workloads only")
        from ucf_vit_b200.utils.graph import GraphedTrainStep
        graphed = GraphedTrainStep(wl.step, resident, warmup=3)

    def step(batch):
        if flush is not None:
            flush.zero_()
        return graphed(*batch) if graphed is not None else wl.step(*batch)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timer = GemmTimer()

    # ---------------- resident-input arm
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step(resident)
    fence()
    # host cost of enqueueing ONE (untimed) step into an empty launch queue: is the step launch-bound?
    t_host0 = time.perf_counter()
    step(resident)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    fence()
    n0 = _lib.launch_count()
    timer.enabled = rank == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        loss = step(resident)
    e1.record()
    fence()
    timer.enabled = False
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    clocks = sampler.stop()
    if graphed is not None:
        # a replayed graph makes no library calls: count the kernels of one eager step, and time the GEMM family there
        n1 = _lib.launch_count()
        timer.enabled = rank == 0
        for _ in range(args.steps):
            wl.step(*resident)
        torch.cuda.synchronize()
        timer.enabled = False
        launches = _lib.launch_count() - n1
    final_loss = float(loss.item())

    # ---------------- end-to-end arm: pinned host batch -> device every step, loss read back every step
    # host batches (pinned) -> DevicePrefetcher (the package's loader hand-off: H2D of batch k+1 on a side stream
    # under step k) -> step -> loss read back
    # The loss of EVERY step is copied device -> host (4 bytes into pinned memory, asynchronously behind the step) and read on
    # the host one step later, so the read of step k does not drain the launch queue before step k + 1 is enqueued (a
    # blocking `.item()` per step idles the GPU for the host's whole inter-step latency, which grows with 8 ranks on one box).
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(n):
        seen = []
        k = 0
        for batch in DevicePrefetcher((host for _ in range(n)), dev):
            l = step(tuple(batch))
            loss_host[k & 1].copy_(l.detach().reshape(1).float(), non_blocking=True)
            loss_ev[k & 1].record()
            if k > 0:
                loss_ev[(k - 1) & 1].synchronize()
                seen.append(float(loss_host[(k - 1) & 1][0]))      # device -> host read of step k - 1's result
            k += 1
        if k > 0:
            loss_ev[(k - 1) & 1].synchronize()
            seen.append(float(loss_host[(k - 1) & 1][0]))
        assert len(seen) == n
        return seen

    e2e_steps(2)
    fence()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_steps(args.steps)
    f1.record()
    fence()
    ms_e2e = f0.elapsed_time(f1)

    # ---------------- reduce over ranks (max time)
    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    extras = wl.extras(dev, args) if rank == 0 else {}
    if rank != 0:
        if need_pg:
            dist.barrier()
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    samples = wl.batch * world
    value = samples / (ms_step * 1e-3)
    e2e_value = samples / (ms_e2e / args.steps * 1e-3)
    f_sample, f_blocks = wl.flops()
    gemm_fl, gemm_ms, gemm_n = timer.summary()
    peak, peak_src = measured_peaks()
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if wl.name == "vit_b16" and os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl.workload, "name": wl.name, "global_batch": samples, "parallelism": wl.parallelism(world),
                   "l2_policy": wl.l2_policy, "cuda_graph": bool(args.cuda_graph),
                   "optimizer": ("AdamW, ucf_adamw_multi kernel" if args.optimizer == "ucf" else "AdamW, torch kernel") +
                                ", fp32 master weights", "loss_final": final_loss},
        "model_tflops": value * f_sample / 1e12,
        "attn_mlp_tflops": value * f_blocks / 1e12,
        "attn_mlp_frac_of_peak": value * f_blocks / 1e12 / (peak * world),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "power_w_max": clocks.get("power_w_max")},
        "e2e": {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": int(h2d_bytes) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all layouts/epilogues)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "launches_timed": gemm_n,
                     "share_of_step": gemm_ms / ms_total if ms_total > 0 else None,
                     "timed_in": "eager pass after the graph-replay region" if args.cuda_graph else "the timed region"},
    }
    if extras:
        line["extras"] = extras
        if wl.name == "unetr_128":       # the decoder's FLOPs are known only after the counter ran
            f_sample, _ = wl.flops()
            line["model_tflops"] = value * f_sample / 1e12
    if world == 1 and not args.no_cpu_baseline:
        try:
            rate, cores, sec, sample = wl.cpu_rate(2 if wl.name.startswith("vit") else 1, 1 if wl.name.startswith("vit") else 0)
            line["cpu_baseline"] = {"value": rate, "unit": wl.unit, "cores": cores, "kind": wl.cpu_kind, "sample": sample + " on the GPU box host"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
    print(json.dumps(line), flush=True)
    if need_pg:
        dist.barrier()
        dist.destroy_process_group()


def main():
    reg = bench_workloads.registry()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="vit_b16", choices=sorted(reg))
    ap.add_argument("--batch", type=int, default=0, help="override the workload's per-GPU batch (the line's config.workload says so)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="vit_tiny: do not use the CUDA-graph step it defaults to on one GPU")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="capture the whole training step in a CUDA graph (utils/graph.py) and replay it; the GEMM roofline "
                         "is then taken from an eager pass after the timed region")
    ap.add_argument("--bf16-decoder", action="store_true",
                    help="unetr_128: PyTorch decoder under bf16 autocast (PyTorch InstanceNorm / LeakyReLU kernels; for comparison)")
    ap.add_argument("--fp32-decoder", action="store_true",
                    help="unetr_128: PyTorch / cuDNN decoder in fp32, the reference's arithmetic (default: channels-last bf16 decoder "
                         "-- cuDNN convolutions under autocast with this package's fused InstanceNorm + residual + LeakyReLU "
                         "kernels between them, UNETR.use_fused_decoder)")
    ap.add_argument("--device-edges", action="store_true",
                    help="sap_4096_*: run the edge detector (ucf_gaussian_blur_u8 + ucf_canny_u8) of every image inside the step, on "
                         "a side stream one batch ahead, instead of building the trees from fixed synthetic edge maps")
    ap.add_argument("--fp32-pixels", action="store_true", help="vit configs: host batches as fp32 pixels (round-1 form) instead of uint8")
    ap.add_argument("--fp32-allreduce", action="store_true",
                    help="DDP gradient all-reduce in fp32 (round-1 form); default: bf16 (torch's bf16_compress_hook -- the "
                         "reference's FSDP drivers likewise reduce in bf16, MixedPrecision(reduce_dtype=bf16))")
    ap.add_argument("--optimizer", default="ucf", choices=["torch", "ucf"],
                    help="AdamW update: torch's kernel or this package's ucf_adamw_multi")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = reg[args.config]
    if args.batch > 0 and args.batch != wl.batch:
        wl.workload += f" [per-GPU batch overridden: {args.batch} instead of {wl.batch}]"
        wl.batch = args.batch
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
