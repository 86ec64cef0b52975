"""SM clock / board power while one kernel loops for ~2 s (is a kernel power-capped or pipeline-bound?).
timeout 120 python scripts/gpu_power_probe.py"""
import subprocess, sys, time, threading
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops, _lib as L

dev = "cuda"
M, D, Hd = 50432, 768, 3072
bf = lambda t: t.to(torch.bfloat16)
x = bf(torch.randn(M, D, device=dev)); w1 = bf(torch.randn(Hd, D, device=dev) * 0.03); b1 = torch.randn(Hd, device=dev)
dy = bf(torch.randn(M, D, device=dev)); w2 = bf(torch.randn(D, Hd, device=dev) * 0.03); z = bf(torch.randn(M, Hd, device=dev))
B, N, H, hd = 256, 197, 12, 64
qkv = bf(torch.randn(B, N, 3, H, hd, device=dev))
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
g = torch.randn(D, device=dev)
cases = {
    "fc1 plain": lambda: ops.gemm(x, w1, M=M, N=Hd, K=D, bias=b1),
    "fc1 gelu": lambda: ops.gemm(x, w1, M=M, N=Hd, K=D, bias=b1, epilogue=L.EPI_BIAS_GELU_AUX),
    "fc2 dgelu": lambda: ops.gemm(dy, w2, M=M, N=Hd, K=D, b_mn=True, aux=z, epilogue=L.EPI_DGELU),
    "cublas fc1": lambda: torch.matmul(x, w1.t()),
    "attn fwd": lambda: ops.attention_fwd(q, k, v, hd ** -0.5),
    "attn bwd": lambda: ops.attention_bwd(q, k, v, o, do, lse, hd ** -0.5, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2]),
    "ln fwd": lambda: ops.layernorm_fwd(x, g, g, 1e-6),
}


def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits",
                            "-i", "0"], capture_output=True, text=True)
        out.append(r.stdout.strip())


for name, f in cases.items():
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out)); th.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n = 0
    e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(50):
            f()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    us = e0.elapsed_time(e1) * 1e3 / n
    clk = [float(s.split(",")[0]) for s in out if s]
    pw = [float(s.split(",")[1]) for s in out if s]
    print(f"{name:12s} {us:8.1f} us/iter  sm_mhz median {sorted(clk)[len(clk)//2] if clk else None} min {min(clk) if clk else None}  "
          f"power median {sorted(pw)[len(pw)//2] if pw else None} W max {max(pw) if pw else None} ({len(clk)} samples) reasons {out[-1].split(',')[2] if out else None}", flush=True)
    time.sleep(1.0)
