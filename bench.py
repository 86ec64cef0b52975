#!/usr/bin/env python
"""Headline benchmark: ViT-B/16 ImageNet-shape classification training step (BASELINE.json
configs[1]) on N B200s, data-parallel, through the reference's module API backed by this
package's sm_100a kernels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = forward + cross-entropy + backward + AdamW update on one synthetic batch of
256 images per GPU (224x224x3), bf16 tensor-core math with fp32 accumulation and fp32 master
weights.  Prints ONE JSON line (see DESIGN.md "Measurement").

  value     whole-job images/s, inputs resident in HBM, K steps timed with CUDA events between
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

WORKLOAD = "ViT-B/16 ImageNet-shape classification train step (fwd+CE+bwd+AdamW), 224x224x3, batch 256/GPU"
CFG = dict(img_size=[224, 224], patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12)
PER_GPU_BATCH = 256
METRIC = "ViT train imgs/s at 1/2/4/8 B200; attn+MLP TFLOP/s vs bf16 peak"


def flops_per_image_train(cfg=CFG):
    """F_block = 24 N D^2 + 4 N^2 D, F_pe = 2 L K D, train = 3 x forward."""
    D, depth = cfg["embed_dim"], cfg["depth"]
    L = (cfg["img_size"][0] // cfg["patch_size"]) ** 2
    N = L + 1
    K = cfg["in_chans"] * cfg["patch_size"] ** 2
    f_block = 24 * N * D * D + 4 * N * N * D
    f_model = 2 * L * K * D + depth * f_block + 2 * D * cfg["num_classes"]
    return 3 * f_model, 3 * depth * f_block


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(batch, steps, warmup):
    """imgs/s of the reference algorithm (oracle port, fp32, torch CPU kernels, all threads)."""
    from oracle import vit_ref as R
    from ucf_vit_b200.simple import arch as A      # used ONLY to obtain reference-shaped initial weights on CPU
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    shapes_model = A.VIT(**CFG, mlp_ratio=4, class_token=True, twoD=True, default_vars=["r", "g", "b"])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in shapes_model.state_dict().items()
          if not k.startswith("token_embeds.")}
    del shapes_model
    cfg = dict(CFG)
    opt = torch.optim.AdamW([v for v in sd.values()], lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(batch, 3, 224, 224, generator=g) * 255.0
    y = torch.randint(0, CFG["num_classes"], (batch,), generator=g)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        logits = R.vit_forward(x, sd, cfg)
        loss = torch.nn.functional.cross_entropy(logits, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch / statistics.median(times), cores, statistics.median(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16
    steps = max(1, min(args.steps, 3))
    warm = 1
    rate, cores, sec = cpu_reference_step_rate(batch, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"batch {batch} per step on the host CPU"},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} timed + {warm} warm-up steps of batch {batch} (same model, fp32, oracle/vit_ref.py + torch autograd + AdamW)"},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    """Wraps ops.gemm: brackets every launch with CUDA events on the launching (current) stream."""

    def __init__(self):
        from ucf_vit_b200 import ops
        self.ops, self.orig, self.records, self.enabled = ops, ops.gemm, [], False

    def install(self):
        def timed(a, b, *, M, N, K, **kw):
            if not self.enabled:
                return self.orig(a, b, M=M, N=N, K=K, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = self.orig(a, b, M=M, N=N, K=K, **kw)
            e1.record()
            self.records.append((2.0 * M * N * K, e0, e1))
            return r
        self.ops.gemm = timed

    def summary(self):
        fl = sum(r[0] for r in self.records)
        ms = sum(r[1].elapsed_time(r[2]) for r in self.records)
        return fl, ms, len(self.records)


def run_gpu_arm(args):
    import torch.distributed as dist
    from ucf_vit_b200 import _lib
    from ucf_vit_b200.simple.arch import VIT
    from ucf_vit_b200.utils.fused_attn import FusedAttn
    from ucf_vit_b200.utils.misc import configure_optimizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = VIT(**CFG, mlp_ratio=4, class_token=True, twoD=True, default_vars=["r", "g", "b"],
                FusedAttn_option=FusedAttn.FLASH).to(dev).train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=64)
    opt = configure_optimizer(model, 1e-4, 0.9, 0.95, 1e-5, fused="ucf" if args.optimizer == "ucf" else True)
    variables = ["r", "g", "b"]
    B = PER_GPU_BATCH
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = (torch.rand(B, 3, 224, 224, generator=g) * 255.0).pin_memory()      # raw 0..255 pixels as float (catsdogs path)
    y_host = torch.randint(0, CFG["num_classes"], (B,), generator=g).pin_memory()
    x_dev = x_host.to(dev)
    y_dev = y_host.to(dev)
    lossf = torch.nn.CrossEntropyLoss()

    def step(x, y):
        logits = net(x, variables)
        loss = lossf(logits.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timer = GemmTimer()
    if rank == 0:
        timer.install()

    # ---------------- resident-input arm
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    fence()
    # host cost of enqueueing ONE (untimed) step into an empty launch queue: is the step launch-bound?
    t_host0 = time.perf_counter()
    step(x_dev, y_dev)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    fence()
    n0 = _lib.launch_count()
    timer.enabled = rank == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        loss = step(x_dev, y_dev)
    e1.record()
    fence()
    timer.enabled = False
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    clocks = sampler.stop()
    final_loss = float(loss.item())

    # ---------------- end-to-end arm: pinned host batch -> device every step, loss read back every step
    # host batches (pinned) -> DevicePrefetcher (the package's loader hand-off: H2D of batch k+1 on a side stream
    # under step k) -> step -> loss read back
    from ucf_vit_b200.dataloaders.prefetch import DevicePrefetcher

    def e2e_steps(n):
        for x, y in DevicePrefetcher(((x_host, y_host) for _ in range(n)), dev):
            l = step(x, y)
            _ = l.item()                      # device -> host read of the step's result

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
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    imgs = B * world
    value = imgs / (ms_step * 1e-3)
    e2e_value = imgs / (ms_e2e / args.steps * 1e-3)
    f_img, f_blocks = flops_per_image_train()
    gemm_fl, gemm_ms, gemm_n = timer.summary()
    peak, peak_src = measured_peaks()
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": imgs, "parallelism": f"dp{world}",
                   "l2_policy": "per-step inputs (154 MB) and activations (>10 GB) exceed the 126 MB L2",
                   "optimizer": ("AdamW, ucf_adamw_multi kernel" if args.optimizer == "ucf" else "AdamW, torch fused kernel") +
                                ", fp32 master weights", "loss_final": final_loss},
        "model_tflops": value * f_img / 1e12,
        "attn_mlp_tflops": value * f_blocks / 1e12,
        "attn_mlp_frac_of_peak": value * f_blocks / 1e12 / (peak * world),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "power_w_max": clocks.get("power_w_max")},
        "e2e": {"value": e2e_value, "unit": "images/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 8) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all layouts/epilogues)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "launches_timed": gemm_n,
                     "share_of_step": gemm_ms / ms_total if ms_total > 0 else None},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            rate, cores, sec = cpu_reference_step_rate(16, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "2 timed + 1 warm-up steps of batch 16 of the same model (fp32, oracle/vit_ref.py + torch autograd + AdamW) on the GPU box host"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default="ucf", choices=["torch", "ucf"],
                    help="AdamW update: torch's fused CUDA kernel or this package's ucf_adamw_multi")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
