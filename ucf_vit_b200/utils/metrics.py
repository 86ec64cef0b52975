"""Losses next to the hot path (/root/reference/src/UCF_VIT/utils/metrics.py:11-17,95-121).
Tiny element-wise reductions; they stay in PyTorch."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def masked_mse(pred, y, mask):
    per_token = ((pred - y) ** 2).mean(dim=-1)
    return (per_token * mask).sum() / mask.sum()


class DiceBLoss(nn.Module):
    def __init__(self, weight=0.5, num_class=2, size_average=True):
        super().__init__()
        self.weight = weight
        self.num_class = num_class

    def forward(self, inputs, targets, smooth=1, act=True):
        if act:
            inputs = torch.sigmoid(inputs)
        pred = torch.flatten(inputs[:, 1:, :, :])
        true = torch.flatten(targets[:, 1:, :, :])
        inter = (pred * true).sum()
        dice = 1 - (2. * inter + smooth) / (pred.sum() + true.sum() + smooth)
        bce = F.binary_cross_entropy(pred, true, reduction='mean')
        return self.weight * bce + (1 - self.weight) * dice
