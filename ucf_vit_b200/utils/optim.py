"""AdamW whose update runs in this package's multi-tensor CUDA kernel.

The reference builds `torch.optim.AdamW` over two parameter groups (utils/misc.py:58-84) and calls
`optimizer.step(); optimizer.zero_grad(); scheduler.step()` once per batch
(training_scripts/train_class_simple.py:355-357).  `FusedAdamW` IS a `torch.optim.AdamW`: same
constructor, same `param_groups`, same per-parameter state (`step`, `exp_avg`, `exp_avg_sq`), so LR
schedulers and `state_dict()` / `load_state_dict()` checkpoints are interchangeable with the stock
optimizer.  Only `step()` differs: every group is updated by `ucf_adamw_multi` launches (24 tensors per
launch) instead of per-tensor or foreach kernels.  CUDA fp32 parameters only; anything else raises.
"""
from collections import defaultdict

import torch

from .. import ops


class FusedAdamW(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False):
        if amsgrad:
            raise NotImplementedError("FusedAdamW: amsgrad is not implemented (the reference never enables it)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                         maximize=maximize, foreach=False, capturable=False, differentiable=False, fused=False)

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)          # host scalar, as the stock optimizer keeps it
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif st["step"].is_cuda:                                           # state loaded from a fused-optimizer checkpoint
            st["step"] = st["step"].detach().to("cpu", torch.float32)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group.get("amsgrad", False):
                raise NotImplementedError("FusedAdamW: amsgrad is not implemented")
            beta1, beta2 = group["betas"]
            by_step = defaultdict(lambda: ([], [], [], []))
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if not p.is_cuda or p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW updates fp32 CUDA parameters with fp32 gradients only "
                                       f"(got {p.dtype} on {p.device}, grad {g.dtype}); there is no CPU fallback")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters")
                st = self._state_of(p)
                st["step"] += 1
                ps, gs, ms, vs = by_step[int(st["step"].item())]
                ps.append(p)
                gs.append(g if g.is_contiguous() else g.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            for step, (ps, gs, ms, vs) in by_step.items():
                # large tensors first: the launches of 24 then hold tensors of similar size
                order = sorted(range(len(ps)), key=lambda i: -ps[i].numel())
                pick = lambda xs: [xs[i] for i in order]
                ops.adamw_multi(pick(ps), pick(gs), pick(ms), pick(vs), lr=float(group["lr"]), beta1=beta1,
                                beta2=beta2, eps=group["eps"], weight_decay=group["weight_decay"], step=step,
                                maximize=group.get("maximize", False))
        return loss
