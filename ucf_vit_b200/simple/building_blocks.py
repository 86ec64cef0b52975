"""B200-native building blocks with the reference's module API.

Class names, constructor arguments, sub-module names (=> state_dict keys) and forward signatures
mirror /root/reference/src/UCF_VIT/simple/building_blocks.py (PatchEmbed :30-92, Mlp :94-129,
Attention :131-192, Block :194-239, MyUnetBlock :241-284, EmbeddingDenseLayer :286-299,
VariableMapping_Attention :301-373) so reference checkpoints load with strict=True and the
training scripts' `embed_layer= / block_fn= / mlp_layer=` injection keeps working.

The nn.Linear / nn.LayerNorm / nn.Conv children are PARAMETER CONTAINERS only: forward never
calls them.  All device math goes through `ucf_vit_b200.functional` (hand-written sm_100a
kernels behind the C ABI); there is no CPU or library fallback -- a CPU tensor raises.
"""
from functools import partial
from typing import Callable, Optional

import torch
import torch.nn as nn

from .. import functional as UF
from .. import ops
from ..utils.fused_attn import FusedAttn
from ..utils.layers import DropPath, LayerScale, to_2tuple, to_3tuple, trunc_normal_  # noqa: F401 (re-exported)
from ..utils.unetr_blocks import get_conv_layer

LayerType = object  # typing alias kept for import compatibility


def _dropout_active(m: nn.Module) -> bool:
    return isinstance(m, nn.Dropout) and m.p > 0.0 and m.training


def _uniform(t):
    assert all(v == t[0] for v in t), f"non-uniform patch size {t} is not supported by the sm_100a patch kernel"
    return int(t[0])


class PatchEmbed(nn.Module):
    """2-D / 3-D image -> patch tokens.  Conv(k = s = p) is a GEMM over non-overlapping patches:
    one bandwidth-bound cast+patchify pass lays patches out as bf16 rows, then the tcgen05 GEMM
    applies `proj.weight.view(D, -1)` with the bias fused in its epilogue."""

    def __init__(self, img_size: Optional[int] = 224, patch_size: int = 16, in_chans: int = 3,
                 embed_dim: int = 768, twoD: Optional[bool] = True, norm_layer: Optional[Callable] = None,
                 bias: bool = True, sqrt_len_method: bool = False):
        super().__init__()
        self.twoD = twoD
        self.sqrt_len_method = sqrt_len_method
        tup = to_2tuple if twoD else to_3tuple
        self.patch_size = tup(patch_size)
        if img_size is None:
            self.img_size = self.grid_size = self.num_patches = None
        else:
            self.img_size = tup(img_size)
            self.grid_size = tuple(s // p for s, p in zip(self.img_size, self.patch_size))
            n = 1
            for g in self.grid_size:
                n *= g
            self.num_patches = n
        conv = nn.Conv2d if twoD else nn.Conv3d
        self.proj = conv(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        nsp = 2 if self.twoD else 3
        assert x.dim() == nsp + 2, f"expected a {nsp + 2}-D input, got {tuple(x.shape)}"
        if self.img_size is not None and not self.sqrt_len_method:
            names = ("height", "width", "width")
            for i in range(nsp):
                assert x.shape[2 + i] == self.img_size[i], \
                    f"Input {names[i]} ({x.shape[2 + i]}) doesn't match model ({self.img_size[i]})."
        p = _uniform(self.patch_size)
        D = self.proj.out_channels
        B = x.shape[0]
        rows = UF.patchify(x, p)                                   # [B*L, K8] bf16, K8 = C*p^d rounded up to 8
        w = self.proj.weight.view(D, -1)
        if rows.shape[1] != w.shape[1]:                            # K not a multiple of 8 (e.g. C = 1, p = 2): zero columns
            w = nn.functional.pad(w, (0, rows.shape[1] - w.shape[1]))
        y = UF.linear(rows, w, self.proj.bias)
        y = y.view(B, -1, D)
        if not isinstance(self.norm, nn.Identity):
            y = _apply_norm(self.norm, y)
        return y


def _apply_norm(norm: nn.Module, x):
    if isinstance(norm, nn.LayerNorm):
        assert norm.elementwise_affine or norm.weight is None
        return UF.layer_norm(x, norm.weight, norm.bias, norm.eps)
    if isinstance(norm, nn.Identity):
        return x
    raise NotImplementedError(f"norm layer {type(norm).__name__} has no sm_100a kernel in this package")


class Mlp(nn.Module):
    """fc1 -> act -> fc2.  With the default exact-erf GELU the activation (and the pre-activation
    needed by backward) is produced inside fc1's GEMM epilogue and GELU' inside fc2's dgrad epilogue."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0., use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        bias = to_2tuple(bias)
        drop_probs = to_2tuple(drop)
        if use_conv:
            raise NotImplementedError("Mlp(use_conv=True) is not used by any reference model")
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop_probs[0])
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])
        self.drop2 = nn.Dropout(drop_probs[1])

    def _fusable(self):
        return (isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none"
                and isinstance(self.norm, nn.Identity) and not _dropout_active(self.drop1)
                and not _dropout_active(self.drop2))

    def forward(self, x, residual=None):
        if self._fusable():
            return UF.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual)
        h = UF.linear(x, self.fc1.weight, self.fc1.bias)
        h = self.drop1(self.act(h))
        h = _apply_norm(self.norm, h)
        y = UF.linear(h, self.fc2.weight, self.fc2.bias)
        y = self.drop2(y)
        return y if residual is None else y + residual


def _checked_head_dim(dim, num_heads):
    """dim // num_heads, refused at construction when no attention kernel serves it: the tcgen05 kernels run heads of 32 and 64
    columns (narrower heads, e.g. the 36 of configs/basic_ct, are zero-padded to those); wider heads (ViT-H's 80, 128) would
    need more tensor-memory columns than the backward kernel has."""
    hd = dim // num_heads
    if hd > 64:
        raise ValueError(f"ucf_vit_b200: head_dim = {dim} // {num_heads} = {hd} is not supported (attention kernels serve "
                         "head_dim <= 64; use more heads)")
    return hd


class Attention(nn.Module):
    """Multi-head self-attention.  Every `FusedAttn` member runs the same tcgen05 flash-attention
    kernel, reading q/k/v in place from the packed qkv projection."""

    def __init__(self, dim: int, fused_attn: FusedAttn = FusedAttn.NONE, num_heads: int = 8,
                 qkv_bias: bool = False, qk_norm: bool = False, attn_drop: float = 0., proj_drop: float = 0.,
                 norm_layer: nn.Module = nn.LayerNorm) -> None:
        super().__init__()
        assert dim % num_heads == 0, 'dim should be divisible by num_heads'
        self.num_heads = num_heads
        self.head_dim = _checked_head_dim(dim, num_heads)
        self.scale = self.head_dim ** -0.5
        self.fused_attn = fused_attn
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x: torch.Tensor, residual=None) -> torch.Tensor:
        B, N, C = x.shape
        if _dropout_active(self.attn_drop):
            raise NotImplementedError("attn_drop > 0 in training mode is not implemented by the fused kernel "
                                      "(every reference config uses 0)")
        H, hd = self.num_heads, self.head_dim
        qkv = UF.linear(x, self.qkv.weight, self.qkv.bias).view(B, N, 3, H, hd)
        if isinstance(self.q_norm, nn.Identity) and isinstance(self.k_norm, nn.Identity):
            o = UF.attention_packed(qkv, self.scale)
        else:
            q = _apply_norm(self.q_norm, qkv[:, :, 0].contiguous())
            k = _apply_norm(self.k_norm, qkv[:, :, 1].contiguous())
            o = UF.attention(q, k, qkv[:, :, 2], self.scale).reshape(B, N, C)
        if _dropout_active(self.proj_drop):
            y = self.proj_drop(UF.linear(o, self.proj.weight, self.proj.bias))
            return y if residual is None else y + residual
        return UF.linear(o, self.proj.weight, self.proj.bias, residual)


class Block(nn.Module):
    """Pre-norm transformer block.  In the configuration all reference drivers use (no qk_norm /
    LayerScale / DropPath / dropout) the whole block -- forward and a hand-scheduled backward --
    is one autograd node (`functional.fused_block`)."""

    def __init__(self, dim: int, num_heads: int, fused_attn: FusedAttn = FusedAttn.NONE, mlp_ratio: float = 4.,
                 qkv_bias: bool = False, qk_norm: bool = False, proj_drop: float = 0., attn_drop: float = 0.,
                 init_values: Optional[float] = None, drop_path: float = 0., act_layer: nn.Module = nn.GELU,
                 norm_layer: nn.Module = nn.LayerNorm, mlp_layer: nn.Module = Mlp) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, fused_attn=fused_attn, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm,
                              attn_drop=attn_drop, proj_drop=proj_drop, norm_layer=norm_layer)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=proj_drop)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def _plain(self):
        a, m = self.attn, self.mlp
        idn = nn.Identity
        return (isinstance(self.ls1, idn) and isinstance(self.ls2, idn)
                and (isinstance(self.drop_path1, idn) or not self.training or self.drop_path1.drop_prob == 0.)
                and (isinstance(self.drop_path2, idn) or not self.training or self.drop_path2.drop_prob == 0.)
                and isinstance(self.norm1, nn.LayerNorm) and isinstance(self.norm2, nn.LayerNorm))

    def _fully_fused(self):
        a, m = self.attn, self.mlp
        return (self._plain() and type(m) is Mlp and m._fusable() and m.fc1.bias is not None
                and a.head_dim in (32, 64) and isinstance(a.q_norm, nn.Identity) and isinstance(a.k_norm, nn.Identity)
                and not _dropout_active(a.attn_drop) and not _dropout_active(a.proj_drop))

    def train(self, mode: bool = True):
        self._ln_fold = None            # folded LayerNorm1 -> QKV constants belong to one set of eval-time weights
        return super().train(mode)

    def _folded_qkv(self):
        """(Wg, colsum, b_folded) of functional.fold_layernorm for norm1 -> attn.qkv, cached while the module stays in eval
        mode and the four tensors keep their versions and storage (load_state_dict / in-place edits re-fold)."""
        a = self.attn
        src = (a.qkv.weight, a.qkv.bias, self.norm1.weight, self.norm1.bias)
        key = tuple((t.data_ptr(), t._version) for t in src if t is not None)
        cached = getattr(self, "_ln_fold", None)
        if cached is None or cached[0] != key:
            cached = (key, UF.fold_layernorm(*src))
            self._ln_fold = cached
        return cached[1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = UF.to_bf16(x)
        if (not self.training and not torch.is_grad_enabled() and x.dim() == 3 and self._fully_fused()
                and ops.ln_gemm_supported(x.shape[0] * x.shape[1], 3 * x.shape[2], x.shape[2])):
            # forward-only execution: LayerNorm1 is folded into the QKV projection (no normalised copy of x is written)
            a, m = self.attn, self.mlp
            return UF.block_forward_nograd(x, self._folded_qkv(), self.norm1.eps, a.proj.weight, a.proj.bias,
                                           self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias,
                                           m.fc2.weight, m.fc2.bias, a.num_heads, self.norm2.eps)
        if self._fully_fused():
            a, m = self.attn, self.mlp
            return UF.fused_block(x, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias,
                                  a.proj.weight, a.proj.bias, self.norm2.weight, self.norm2.bias,
                                  m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias,
                                  a.num_heads, self.norm1.eps, self.norm2.eps)
        if self._plain() and type(self.mlp) is Mlp:
            x = self.attn(_apply_norm(self.norm1, x), residual=x)
            return self.mlp(_apply_norm(self.norm2, x), residual=x)
        x = x + self.drop_path1(self.ls1(self.attn(_apply_norm(self.norm1, x))))
        x = x + self.drop_path2(self.ls2(self.mlp(_apply_norm(self.norm2, x))))
        return x


class MyUnetBlock(nn.Module):
    """Transposed-conv upsampling step of the UNETR decoder (cuDNN; SURVEY.md §8f 'next')."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, upsample_kernel_size: int,
                 res_block: bool = False) -> None:
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels,
                                          kernel_size=upsample_kernel_size, stride=upsample_kernel_size,
                                          conv_only=True, is_transposed=True)

    def forward(self, inp):
        return self.transp_conv(inp)


class EmbeddingDenseLayer(nn.Module):
    """Time-step embedding MLP of DiffusionVIT: [B, C] only, negligible work -> plain torch."""

    def __init__(self, c_in: int, c_out: int, dropout_prob: float):
        super().__init__()
        self.linear1 = nn.Linear(c_in, c_out)
        self.linear2 = nn.Linear(c_out, c_out)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(p=dropout_prob)

    def forward(self, x):
        return self.linear2(self.dropout(self.relu(self.linear1(x))))


class VariableMapping_Attention(nn.Module):
    """Cross-attention that folds V per-variable token streams into `N_a` aggregated streams
    (channel aggregation).  q/kv/proj run on the tcgen05 GEMM; the tiny (N_a x V) attention per
    token runs in a bandwidth-bound kernel (`ucf_var_attention_*`)."""

    def __init__(self, dim: int, fused_attn: FusedAttn = FusedAttn.NONE, num_heads: int = 8, qkv_bias: bool = False,
                 qk_norm: bool = False, proj_bias: bool = True, attn_drop: float = 0., proj_drop: float = 0.,
                 norm_layer: nn.Module = nn.LayerNorm) -> None:
        super().__init__()
        assert dim % num_heads == 0, 'dim should be divisible by num_heads'
        self.num_heads = num_heads
        self.head_dim = _checked_head_dim(dim, num_heads)
        self.scale = self.head_dim ** -0.5
        self.fused_attn = fused_attn
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.q_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, var_query: torch.Tensor, x: torch.Tensor, shared_query: bool = False) -> torch.Tensor:
        """var_query [B', N_a, C] (or [1, N_a, C] with shared_query=True, which skips the B'-fold
        redundant q projection the reference performs), x [B', N_i, C] -> [B', N_a, C]."""
        if _dropout_active(self.attn_drop):
            raise NotImplementedError("attn_drop > 0 is not implemented by the fused kernel")
        if not (isinstance(self.q_norm, nn.Identity) and isinstance(self.k_norm, nn.Identity)):
            raise NotImplementedError("qk_norm in VariableMapping_Attention is not used by the reference")
        Bp, N_i, C = x.shape
        N_a = var_query.size(1)
        H, hd = self.num_heads, self.head_dim
        q = UF.linear(var_query, self.q.weight, self.q.bias).view(-1, N_a, H, hd)
        kv = UF.linear(x, self.kv.weight, self.kv.bias).view(Bp, N_i, 2, H, hd)
        o = UF.var_attention(q, kv, self.scale)                # [B', N_a, H, hd]
        y = UF.linear(o.view(Bp, N_a, C), self.proj.weight, self.proj.bias)
        return self.proj_drop(y) if _dropout_active(self.proj_drop) else y
