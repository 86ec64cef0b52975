"""UNETR convolutional decoder blocks (PyTorch / cuDNN; off the sm_100a kernel path).

The reference builds its UNETR decoder from MONAI blocks
(`/root/reference/src/UCF_VIT/simple/arch.py:33-34,808-940`,
`/root/reference/src/UCF_VIT/simple/building_blocks.py:26,271-279`).  MONAI (pinned
`monai==1.4.0`, `/root/reference/Docker/Dockerfile:6`) is not installed in this image and its
source is not under /root/reference, so the blocks are restated here from MONAI 1.4's published
behaviour, keeping every sub-module name so reference checkpoints load with ``strict=True``
(e.g. ``decoder5.transp_conv.conv.weight``, ``encoder1.layer.conv1.conv.weight``,
``out.conv.conv.bias``).  Parity against a real MONAI install is UNPINNED.

Two execution modes:
* default -- PyTorch / cuDNN in the parameter dtype, the reference's arithmetic;
* ``ndhwc_bf16`` (``UNETR.use_fused_decoder()``) -- channels-last bf16 activations end to end: the convolutions stay cuDNN
  calls (under autocast, fp32 accumulate), everything between them -- InstanceNorm, residual add, LeakyReLU, forward and
  backward -- is one ``ucf_inorm_*`` kernel sequence per block body (``csrc/instnorm.cu``) instead of PyTorch's batch-norm
  kernels, layout copies and element-wise passes (60 % of the bf16 decoder step before).
"""
from typing import Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from .. import functional as UF
from .. import ops

__all__ = [
    "get_conv_layer", "UnetResBlock", "UnetBasicBlock", "UnetrBasicBlock",
    "UnetrPrUpBlock", "UnetrUpBlock", "UnetOutBlock",
]

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}
_INORM = {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}


def _same_padding(kernel_size, stride):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = (k - s + 1) / 2
    if np.min(p) < 0:
        raise AssertionError("padding value should not be negative, please change the kernel size and/or stride.")
    p = tuple(int(v) for v in p)
    return p if len(p) > 1 else p[0]


def _output_padding(kernel_size, stride, padding):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = np.atleast_1d(padding)
    op = 2 * p + s - k
    if np.min(op) < 0:
        raise AssertionError("out_padding value should not be negative, please change the kernel size and/or stride.")
    op = tuple(int(v) for v in op)
    return op if len(op) > 1 else op[0]


class _Convolution(nn.Sequential):
    """`monai.networks.blocks.Convolution` reduced to the conv-only form UNETR uses:
    an ``nn.Sequential`` whose single child is named ``conv``."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, bias, is_transposed):
        super().__init__()
        padding = _same_padding(kernel_size, stride)
        if is_transposed:
            conv = _CONVT[spatial_dims](
                in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                output_padding=_output_padding(kernel_size, stride, padding), bias=bias)
        else:
            conv = _CONV[spatial_dims](
                in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)
        self.add_module("conv", conv)

    # set by UNETR.use_fused_decoder(): 3x3x3 layers take their weight gradient from ucf_conv3d_wgrad, 1x1x1 layers run on
    # ucf_pointwise_conv* in all three directions
    ndhwc_bf16 = False

    def forward(self, x):
        c = self.conv
        if (self.ndhwc_bf16 and isinstance(c, nn.Conv3d) and x.is_cuda and x.dim() == 5 and c.bias is None and c.groups == 1
                and c.kernel_size == (3, 3, 3) and c.stride == (1, 1, 1) and c.padding == (1, 1, 1) and c.dilation == (1, 1, 1)
                and c.in_channels * c.out_channels <= 1024       # measured ahead of cuDNN there (profiles/r02_conv_wgrad.log), behind at 64 -> 32
                and ops.conv3d_wgrad_supported(max(c.in_channels, 16), c.out_channels, *x.shape[2:])):   # < 16 inputs: zero-padded
            return UF.conv3x3x3(x.to(torch.bfloat16), c.weight)
        if (self.ndhwc_bf16 and isinstance(c, (nn.Conv2d, nn.Conv3d)) and x.is_cuda and c.groups == 1
                and all(k == 1 for k in c.kernel_size) and all(v == 1 for v in c.stride) and all(v == 0 for v in c.padding)
                and ops.pointwise_conv_supported(c.in_channels, c.out_channels)):
            return UF.conv1x1x1(x.to(torch.bfloat16), c.weight, c.bias)
        return super().forward(x)


def get_conv_layer(spatial_dims: int, in_channels: int, out_channels: int,
                   kernel_size: Union[Sequence[int], int] = 3, stride: Union[Sequence[int], int] = 1,
                   act=None, norm=None, dropout=None, bias: bool = False, conv_only: bool = True,
                   is_transposed: bool = False):
    # every call site in the reference passes act=None/norm=None (or conv_only=True), for which
    # MONAI's Convolution adds no ADN sub-module.
    if not conv_only and (act is not None or norm is not None or dropout is not None):
        raise NotImplementedError("ADN variants are not used by UNETR")
    return _Convolution(spatial_dims, in_channels, out_channels, kernel_size, stride, bias, is_transposed)


def _norm(norm_name, spatial_dims, channels):
    name = norm_name[0] if isinstance(norm_name, (tuple, list)) else norm_name
    if str(name).lower() == "instance":
        return _INORM[spatial_dims](channels)       # affine=False: no parameters
    if str(name).lower() == "batch":
        return {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[spatial_dims](channels)
    raise NotImplementedError(f"norm {norm_name!r}")


def _fused_norm_args(norm, lrelu):
    """(eps, slope) for the fused InstanceNorm + LeakyReLU kernels; anything else has no fused form."""
    if not isinstance(norm, tuple(_INORM.values())) or norm.affine or norm.track_running_stats:
        raise NotImplementedError("ndhwc_bf16 decoder: only parameter-free InstanceNorm (norm_name='instance') is fused")
    return norm.eps, lrelu.negative_slope


class UnetResBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name,
                 act_name=None, dropout=None):
        super().__init__()
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size, stride, conv_only=False)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size, 1, conv_only=False)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = _norm(norm_name, spatial_dims, out_channels)
        self.norm2 = _norm(norm_name, spatial_dims, out_channels)
        self.downsample = in_channels != out_channels
        if not np.all(np.atleast_1d(stride) == 1):
            self.downsample = True
        if self.downsample:
            self.conv3 = get_conv_layer(spatial_dims, in_channels, out_channels, 1, stride, conv_only=False)
            self.norm3 = _norm(norm_name, spatial_dims, out_channels)

    ndhwc_bf16 = False

    def _forward_fused(self, inp):
        eps, slope = _fused_norm_args(self.norm1, self.lrelu)
        inp = UF.channels_last(inp.to(torch.bfloat16))
        out = UF.instance_norm_act(self.conv1(inp), negative_slope=slope, eps=eps)
        out = self.conv2(out)
        if hasattr(self, "conv3"):
            return UF.instance_norm_act(out, self.conv3(inp), norm_residual=True, negative_slope=slope, eps=eps)
        return UF.instance_norm_act(out, inp, norm_residual=False, negative_slope=slope, eps=eps)

    def forward(self, inp):
        if self.ndhwc_bf16:
            return self._forward_fused(inp)
        residual = inp
        out = self.lrelu(self.norm1(self.conv1(inp)))
        out = self.norm2(self.conv2(out))
        if hasattr(self, "conv3"):
            residual = self.norm3(self.conv3(residual))
        out = out + residual
        return self.lrelu(out)


class UnetBasicBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name,
                 act_name=None, dropout=None):
        super().__init__()
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size, stride, conv_only=False)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size, 1, conv_only=False)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = _norm(norm_name, spatial_dims, out_channels)
        self.norm2 = _norm(norm_name, spatial_dims, out_channels)

    ndhwc_bf16 = False

    def forward(self, inp):
        if self.ndhwc_bf16:
            eps, slope = _fused_norm_args(self.norm1, self.lrelu)
            out = UF.instance_norm_act(self.conv1(UF.channels_last(inp.to(torch.bfloat16))), negative_slope=slope, eps=eps)
            return UF.instance_norm_act(self.conv2(out), negative_slope=slope, eps=eps)
        out = self.lrelu(self.norm1(self.conv1(inp)))
        return self.lrelu(self.norm2(self.conv2(out)))


class UnetrBasicBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, res_block=False):
        super().__init__()
        blk = UnetResBlock if res_block else UnetBasicBlock
        self.layer = blk(spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name)

    def forward(self, inp):
        return self.layer(inp)


class UnetrPrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, num_layer, kernel_size, stride,
                 upsample_kernel_size, norm_name, conv_block=False, res_block=False):
        super().__init__()
        up = upsample_kernel_size
        self.transp_conv_init = get_conv_layer(spatial_dims, in_channels, out_channels, up, up,
                                               conv_only=True, is_transposed=True)
        if conv_block:
            blk = UnetResBlock if res_block else UnetBasicBlock
            self.blocks = nn.ModuleList([
                nn.Sequential(
                    get_conv_layer(spatial_dims, out_channels, out_channels, up, up, conv_only=True, is_transposed=True),
                    blk(spatial_dims, out_channels, out_channels, kernel_size, stride, norm_name),
                ) for _ in range(num_layer)])
        else:
            self.blocks = nn.ModuleList([
                get_conv_layer(spatial_dims, out_channels, out_channels, up, up, conv_only=True, is_transposed=True)
                for _ in range(num_layer)])

    def forward(self, x):
        x = self.transp_conv_init(x)
        for blk in self.blocks:
            x = blk(x)
        return x


class UnetrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size,
                 norm_name, res_block=False):
        super().__init__()
        up = upsample_kernel_size
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels, up, up,
                                          conv_only=True, is_transposed=True)
        blk = UnetResBlock if res_block else UnetBasicBlock
        self.conv_block = blk(spatial_dims, out_channels + out_channels, out_channels, kernel_size, 1, norm_name)

    def forward(self, inp, skip):
        out = self.transp_conv(inp)
        out = torch.cat((out, skip), dim=1)
        return self.conv_block(out)


class UnetOutBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        self.conv = get_conv_layer(spatial_dims, in_channels, out_channels, 1, 1, bias=True, conv_only=False)

    def forward(self, inp):
        return self.conv(inp)
