"""GPU: bandwidth of the loss / optimizer kernels next to the PyTorch formulation.

    python scripts/gpu_train_side_bench.py

AdamW on the ViT-B/16 parameter set (86 M fp32 elements, 28 B of traffic each), the reconstruction loss at the
MAE ViT-L/16 shape (B 256, 196 tokens x 768 values + a 3x224x224 image per sample)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucf_vit_b200.utils import metrics, misc  # noqa: E402
from ucf_vit_b200.utils.optim import FusedAdamW  # noqa: E402


def timeit(fn, iters=20, warm=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                                   # evict L2 between iterations
        torch.cuda._sleep(4_000_000)                    # ~2 ms of device idle: the host runs ahead, so launch cost is hidden
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3                            # us


def main():
    torch.manual_seed(0)
    shapes = [(2304, 768), (2304,), (768, 768), (768,), (3072, 768), (3072,), (768, 3072), (768,), (768,), (768,),
              (768,), (768,)] * 12 + [(768, 768), (768,), (1, 197, 768), (1, 1, 768), (1000, 768), (1000,), (768,), (768,)]
    n = sum(int(torch.Size(s).numel()) for s in shapes)
    for name, mk in (("ucf_adamw_multi", lambda ps: FusedAdamW(ps, lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5)),
                     ("torch fused    ", lambda ps: torch.optim.AdamW(ps, lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5, fused=True))):
        ps = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.02) for s in shapes]
        for p in ps:
            p.grad = torch.randn_like(p)
        opt = mk(ps)
        us = timeit(opt.step)
        print(f"AdamW {name}: {len(ps)} tensors, {n / 1e6:.1f} M elements  {us:8.1f} us  {n * 28 / us / 1e6:6.2f} TB/s "
              f"(4 reads + 3 writes of fp32)")
    B, p = 256, 16
    data = torch.randn(B, 3, 224, 224, device="cuda")
    mask = (torch.rand(B, 196, device="cuda") < 0.75).float()
    for dt in (torch.float32, torch.bfloat16):
        pred = torch.randn(B, 196, 768, device="cuda").to(dt).requires_grad_(True)
        e = pred.element_size()
        nel = pred.numel()
        for mname, mk in (("masked", mask), ("full  ", None)):
            frac = float(mask.mean()) if mk is not None else 1.0
            us_f = timeit(lambda: metrics.patch_mse(pred.detach(), data, p, True, mk))
            loss = metrics.patch_mse(pred, data, p, True, mk)
            us_b = timeit(lambda: torch.autograd.grad(loss, pred, retain_graph=True))
            by_f = nel * frac * (e + 4)
            by_b = nel * frac * (e + 4) + nel * e
            print(f"patch_mse {mname} pred {str(dt)[6:]:8s}: fwd {us_f:7.1f} us {by_f / us_f / 1e6:5.2f} TB/s   "
                  f"bwd {us_b:7.1f} us {by_b / us_b / 1e6:5.2f} TB/s")

            def torch_form():
                tgt = misc.patchify(data, p, True)
                return metrics.masked_mse(pred, tgt.to(dt), mk) if mk is not None else torch.nn.functional.mse_loss(pred, tgt.to(dt))
            us_tf = timeit(lambda: torch_form())
            lt = torch_form()
            us_tb = timeit(lambda: torch.autograd.grad(lt, pred, retain_graph=True))
            print(f"   PyTorch formulation (patchify + {'masked_mse' if mk is not None else 'mse_loss'}): fwd {us_tf:7.1f} us   bwd {us_tb:7.1f} us")


if __name__ == "__main__":
    main()
