"""ucf_conv3d_wgrad against cuDNN's weight-gradient kernel (torch convolution_backward, cudnn.benchmark on) at the UNETR-128
decoder's 3x3x3 layers (batch 16, channels-last bf16).   python scripts/gpu_conv_wgrad_bench.py"""
import sys
import torch
sys.path.insert(0, ".")
from ucf_vit_b200 import ops
torch.backends.cudnn.benchmark = True
dev = "cuda"
def ev(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for (N, Ci, Co, S) in [(16, 16, 16, 128), (16, 32, 16, 128), (16, 32, 32, 64), (16, 64, 32, 64)]:
    x = torch.randn(N, S, S, S, Ci, device=dev).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    dy = torch.randn(N, S, S, S, Co, device=dev).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    w = torch.randn(Co, Ci, 3, 3, 3, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    flops = 2.0 * N * S ** 3 * 27 * Ci * Co
    t_lib = ev(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1, [False, True, False]))
    t_ucf = ev(lambda: ops.conv3d_wgrad(x, dy))
    ref = torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1, [False, True, False])[1].float()
    got = ops.conv3d_wgrad(x, dy)
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"N={N} {Ci}->{Co} @ {S}^3: cuDNN wgrad {t_lib:.2f} ms ({flops/t_lib/1e9:.0f} TFLOP/s), ucf_conv3d_wgrad {t_ucf:.2f} ms ({flops/t_ucf/1e9:.0f} TFLOP/s), "
          f"{t_lib/t_ucf:.2f}x; rel-L2 against the library's bf16 result {rel:.1e}")
