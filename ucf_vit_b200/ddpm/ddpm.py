"""DDPM noise schedule (host scalars), /root/reference/src/UCF_VIT/ddpm/ddpm.py:4-13."""
import torch
import torch.nn as nn


class DDPM_Scheduler(nn.Module):
    def __init__(self, num_time_steps: int = 1000):
        super().__init__()
        self.num_time_steps = num_time_steps
        self.beta = torch.linspace(1e-4, 0.02, num_time_steps, requires_grad=False)
        self.alpha = torch.cumprod(1 - self.beta, dim=0).requires_grad_(False)

    def forward(self, t):
        return self.beta[t], self.alpha[t]
