"""Per-iteration linear warm-up followed by cosine annealing.

Same class name, constructor and LR sequence as
/root/reference/src/UCF_VIT/utils/lr_scheduler.py:12-94 (host scalar math, no kernel)."""
import math
import warnings
from typing import List

from torch.optim import Optimizer
from torch.optim.lr_scheduler import _LRScheduler


class LinearWarmupCosineAnnealingLR(_LRScheduler):
    def __init__(self, optimizer: Optimizer, warmup_epochs: int, max_epochs: int, warmup_start_lr: float = 0.0,
                 eta_min: float = 0.0, last_epoch: int = -1) -> None:
        self.warmup_epochs, self.max_epochs = warmup_epochs, max_epochs
        self.warmup_start_lr, self.eta_min = warmup_start_lr, eta_min
        super().__init__(optimizer, last_epoch)

    def get_lr(self) -> List[float]:
        """Chainable (recursive) form: each value derives from the group's current lr."""
        if not self._get_lr_called_within_step:
            warnings.warn("To get the last learning rate computed by the scheduler, please use `get_last_lr()`.",
                          UserWarning)
        e, w, T = self.last_epoch, self.warmup_epochs, self.max_epochs
        groups = self.optimizer.param_groups
        if e == w:
            return self.base_lrs
        if e == 0:
            return [self.warmup_start_lr] * len(self.base_lrs)
        if e < w:
            return [g["lr"] + (b - self.warmup_start_lr) / (w - 1) for b, g in zip(self.base_lrs, groups)]
        span = T - w
        if (e - 1 - T) % (2 * span) == 0:
            return [g["lr"] + (b - self.eta_min) * (1 - math.cos(math.pi / span)) / 2
                    for b, g in zip(self.base_lrs, groups)]
        num = 1 + math.cos(math.pi * (e - w) / span)
        den = 1 + math.cos(math.pi * (e - w - 1) / span)
        return [num / den * (g["lr"] - self.eta_min) + self.eta_min for g in groups]

    def _get_closed_form_lr(self) -> List[float]:
        e, w, T = self.last_epoch, self.warmup_epochs, self.max_epochs
        if e < w:
            return [self.warmup_start_lr + e * (b - self.warmup_start_lr) / max(1, w - 1) for b in self.base_lrs]
        return [self.eta_min + 0.5 * (b - self.eta_min) * (1 + math.cos(math.pi * (e - w) / (T - w)))
                for b in self.base_lrs]
