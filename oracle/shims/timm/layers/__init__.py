"""timm.layers subset: only what UCF_VIT/{simple,fsdp}/building_blocks.py imports."""
from typing import Callable, Optional, Type, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .helpers import to_2tuple, to_3tuple  # noqa: F401

LayerType = Union[str, Callable, Type[nn.Module]]


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    # timm.layers.trunc_normal_ == torch.nn.init.trunc_normal_ (absolute cut-offs a, b)
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def drop_path(x, drop_prob: float = 0., training: bool = False, scale_by_keep: bool = True):
    if drop_prob == 0. or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
    if keep_prob > 0.0 and scale_by_keep:
        random_tensor.div_(keep_prob)
    return x * random_tensor


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training, self.scale_by_keep)


class PatchDropout(nn.Module):
    def __init__(self, prob: float = 0.5, num_prefix_tokens: int = 1, ordered: bool = False,
                 return_indices: bool = False):
        super().__init__()
        assert 0 <= prob < 1.
        self.prob = prob
        self.num_prefix_tokens = num_prefix_tokens

    def forward(self, x):
        if not self.training or self.prob == 0.:
            return x
        raise NotImplementedError("PatchDropout>0 is never configured by the reference drivers")


class AttentionPoolLatent(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("AttentionPoolLatent is imported but never constructed by the reference")


def resample_patch_embed(*a, **k):
    raise NotImplementedError


def resample_abs_pos_embed(*a, **k):
    raise NotImplementedError


def get_act_layer(name=None):
    if name is None:
        return None
    if not isinstance(name, str):
        return name
    return {'gelu': nn.GELU, 'relu': nn.ReLU, 'silu': nn.SiLU}[name.lower()]


def get_norm_layer(norm_layer=None):
    if norm_layer is None:
        return None
    if not isinstance(norm_layer, str):
        return norm_layer
    return {'layernorm': nn.LayerNorm}[norm_layer.lower()]


def use_fused_attn(experimental: bool = False) -> bool:
    import os
    if not hasattr(F, 'scaled_dot_product_attention'):
        return False
    return int(os.environ.get('TIMM_FUSED_ATTN', '1')) > 0
