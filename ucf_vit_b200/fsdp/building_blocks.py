"""`UCF_VIT.fsdp.building_blocks` surface (/root/reference/src/UCF_VIT/fsdp/building_blocks.py):
the same blocks with the tensor-parallel constructor arguments accepted.  North-star scope is
DP / FSDP only, so `tensor_par_size` must be 1; the process group is
ignored.  `Block` is the class FSDP's auto-wrap policy and activation checkpointing key on
(/root/reference/training_scripts/train_masked_fsdp.py:361-366,393)."""
import torch.nn as nn

from ..simple import building_blocks as _s
from ..simple.building_blocks import (DropPath, EmbeddingDenseLayer, LayerScale, MyUnetBlock,  # noqa: F401
                                      trunc_normal_, to_2tuple, to_3tuple)
from ..utils.fused_attn import FusedAttn


def _check_tp(tensor_par_size):
    if tensor_par_size not in (1, None):
        raise NotImplementedError(
            f"tensor_par_size={tensor_par_size}: tensor (Hybrid-OP) parallelism is out of scope for the B200 build; "
            "use data-parallel / FSDP with tensor_par_size=1")


class PatchEmbed(_s.PatchEmbed):
    pass


class Mlp(_s.Mlp):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                 bias=True, drop=0., use_conv=False, tensor_par_size=1, tensor_par_group=None):
        _check_tp(tensor_par_size)
        super().__init__(in_features, hidden_features, out_features, act_layer, norm_layer, bias, drop, use_conv)


class Attention(_s.Attention):
    def __init__(self, dim, fused_attn=FusedAttn.NONE, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.,
                 proj_drop=0., norm_layer=nn.LayerNorm, tensor_par_size=1, tensor_par_group=None):
        _check_tp(tensor_par_size)
        super().__init__(dim, fused_attn, num_heads, qkv_bias, qk_norm, attn_drop, proj_drop, norm_layer)


class Block(_s.Block):
    def __init__(self, dim, num_heads, fused_attn=FusedAttn.NONE, mlp_ratio=4., qkv_bias=False, qk_norm=False,
                 proj_drop=0., attn_drop=0., init_values=None, drop_path=0., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, mlp_layer=_s.Mlp, tensor_par_size=1, tensor_par_group=None):
        _check_tp(tensor_par_size)
        if mlp_layer is Mlp:          # the fsdp Mlp only adds TP kwargs; keep the fused fast path
            mlp_layer = _s.Mlp
        super().__init__(dim, num_heads, fused_attn, mlp_ratio, qkv_bias, qk_norm, proj_drop, attn_drop,
                         init_values, drop_path, act_layer, norm_layer, mlp_layer)


class VariableMapping_Attention(_s.VariableMapping_Attention):
    def __init__(self, dim, fused_attn=FusedAttn.NONE, num_heads=8, qkv_bias=False, qk_norm=False, proj_bias=True,
                 attn_drop=0., proj_drop=0., norm_layer=nn.LayerNorm, tensor_par_size=1, tensor_par_group=None):
        _check_tp(tensor_par_size)
        super().__init__(dim, fused_attn, num_heads, qkv_bias, qk_norm, proj_bias, attn_drop, proj_drop, norm_layer)
