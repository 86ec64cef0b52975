"""DDPM noise schedule.

Host-side scalars only: the diffusion trainer indexes `beta[t]` / `alpha[t]` on the CPU and sends the
gathered values to the device with the batch, so nothing here touches the CUDA library.
Behaviour follows /root/reference/src/UCF_VIT/ddpm/ddpm.py:4-13: a linear beta ramp 1e-4 → 0.02 over
`num_time_steps` and `alpha[t] = prod_{s<=t} (1 - beta[s])` (the cumulative "alpha-bar").
"""
from typing import Tuple

import torch
from torch import Tensor, nn

BETA_START = 1e-4
BETA_END = 0.02


def linear_noise_schedule(steps: int) -> Tuple[Tensor, Tensor]:
    """(beta, alpha_bar) of a `steps`-long linear schedule, both fp32 CPU tensors without grad."""
    with torch.no_grad():
        beta = torch.linspace(BETA_START, BETA_END, steps)
        alpha_bar = (1.0 - beta).cumprod(0)
    return beta, alpha_bar


class DDPM_Scheduler(nn.Module):
    """`scheduler(t) -> (beta[t], alpha_bar[t])`; `t` is an int or an index tensor.

    `beta` and `alpha` are plain attributes (not buffers), as in the reference, so they stay on the
    host when the module is moved and do not appear in the state dict."""

    def __init__(self, num_time_steps: int = 1000):
        super().__init__()
        self.num_time_steps = int(num_time_steps)
        self.beta, self.alpha = linear_noise_schedule(self.num_time_steps)

    def forward(self, t):
        b = self.beta[t]
        a = self.alpha[t]
        return b, a
