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


def _grad_like(p, g):
    """The gradient in the parameter's own dense layout (the kernel walks both storages in step)."""
    if g.stride() == p.stride() or (p.is_contiguous() and g.is_contiguous()):
        return g
    out = torch.empty_like(p, memory_format=torch.preserve_format)
    out.copy_(g)
    return out


class FusedAdamW(torch.optim.AdamW):
    """`capturable=True`: the step count and every group's learning rate live in fp32 CUDA scalars (`group["lr"]` becomes
    a tensor; torch's LR schedulers `fill_` tensor learning rates in place), the kernel derives its bias corrections from
    them on the device, and `step()` performs no host read -- so a whole training step can be captured in a CUDA graph
    (utils/graph.py).  All parameters then share ONE step counter (they must all receive a gradient at every step)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, capturable=False):
        if amsgrad:
            raise NotImplementedError("FusedAdamW: amsgrad is not implemented (the reference never enables it)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                         maximize=maximize, foreach=False, capturable=False, differentiable=False, fused=False)
        self._ucf_capturable = bool(capturable)
        self._step_dev = None
        self._tables = {}

    def _make_capturable(self):
        dev = None
        for group in self.param_groups:
            for p in group["params"]:
                dev = p.device
                break
            if dev is not None:
                break
        if dev is None or dev.type != "cuda":
            raise RuntimeError("FusedAdamW(capturable=True) needs CUDA parameters")
        steps = [float(self.state[p]["step"]) for g in self.param_groups for p in g["params"] if "step" in self.state[p]]
        self._step_dev = torch.full((1,), max(steps) if steps else 0.0, dtype=torch.float32, device=dev)
        for group in self.param_groups:
            if not torch.is_tensor(group["lr"]):
                group["lr"] = torch.tensor(float(group["lr"]), dtype=torch.float32, device=dev)
            elif not group["lr"].is_cuda:
                group["lr"] = group["lr"].to(dev, torch.float32)

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)          # host scalar, as the stock optimizer keeps it
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif st["step"].is_cuda and not self._ucf_capturable:              # state loaded from a fused-optimizer checkpoint
            st["step"] = st["step"].detach().to("cpu", torch.float32)
        return st

    def _check(self, p, g):
        if g.is_sparse:
            raise RuntimeError("FusedAdamW does not support sparse gradients")
        if not p.is_cuda or p.dtype != torch.float32 or g.dtype != torch.float32:
            raise RuntimeError("FusedAdamW updates fp32 CUDA parameters with fp32 gradients only "
                               f"(got {p.dtype} on {p.device}, grad {g.dtype}); there is no CPU fallback")
        if not (p.is_contiguous() or ops._is_dense(p)):
            raise RuntimeError("FusedAdamW needs dense parameters (contiguous or a permuted-dense layout such as channels_last)")

    @torch.no_grad()
    def _step_capturable(self):
        if self._step_dev is None:
            self._make_capturable()
        self._step_dev.add_(1.0)
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            ps, gs, ms, vs = [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                self._check(p, g)
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = self._step_dev[0]           # a view: every parameter shares the device counter
                ps.append(p)
                gs.append(_grad_like(p, g))
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            if ps:
                order = sorted(range(len(ps)), key=lambda i: -ps[i].numel())
                pick = lambda xs: [xs[i] for i in order]
                ops.adamw_multi(pick(ps), pick(gs), pick(ms), pick(vs), lr=group["lr"], beta1=beta1, beta2=beta2,
                                eps=group["eps"], weight_decay=group["weight_decay"], step=self._step_dev,
                                maximize=group.get("maximize", False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._ucf_capturable:
            self._step_capturable()
            return loss
        for group in self.param_groups:
            if group.get("amsgrad", False):
                raise NotImplementedError("FusedAdamW: amsgrad is not implemented")
            beta1, beta2 = group["betas"]
            by_step = defaultdict(lambda: ([], [], [], []))
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                self._check(p, g)
                st = self._state_of(p)
                st["step"] += 1
                ps, gs, ms, vs = by_step[int(st["step"].item())]
                ps.append(p)
                gs.append(_grad_like(p, g))
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
