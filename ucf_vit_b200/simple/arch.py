"""VIT / SAP / MAE / UNETR / DiffusionVIT with the reference's Python API, running on the sm_100a
kernels of this package.

API contract mirrored from /root/reference/src/UCF_VIT/simple/arch.py (VIT :101-489, SAP :491-536,
MAE :538-755, UNETR :757-1113, DiffusionVIT :1115-1283): same constructor keyword arguments (incl.
the `embed_layer / block_fn / mlp_layer` injection points), same attribute and sub-module names
(=> identical state_dict keys, incl. the `token_embeds.*` aliases of `patch_embed.*`), same forward
signatures and return types.  The per-class token-embedding code the reference repeats four times
lives once in `VIT._embed_tokens`.

Documented deviations:
  A1  `sqrt_len_method=` is passed correctly to the patch embedder (the reference's simple-mode typo
      raises TypeError for use_varemb=True);
  A3  DiffusionVIT calls `_pos_embed(x, None)` (the reference omits the argument and raises);
  A7  all per-variable patch embedders alias ONE module; they are applied to all variables in one
      batched GEMM instead of V separate convolutions;
  A14 the diffusion time table stays on the device (no per-step D2H sync / H2D upload).
Activations are bf16 between kernels; heads that stay in PyTorch run in the parameter dtype.
"""
from functools import partial
from typing import Callable, List, Optional, Tuple, Type, Union

import numpy as np
import torch
import torch.nn as nn

from .. import functional as UF
from .. import ops
from ..utils.fused_attn import FusedAttn
from ..utils.layers import named_apply, trunc_normal_
from ..utils.pos_embed import (SinusoidalEmbeddings, get_1d_sincos_pos_embed_from_grid,
                               get_2d_sincos_pos_embed, get_3d_sincos_pos_embed)
from ..utils.unetr_blocks import UnetOutBlock, UnetrBasicBlock, UnetrPrUpBlock, UnetrUpBlock
from .building_blocks import (Block, DropPath, EmbeddingDenseLayer, Mlp, MyUnetBlock, PatchEmbed,  # noqa: F401
                              VariableMapping_Attention, _apply_norm)

try:
    from typing import Literal
except ImportError:  # pragma: no cover
    from typing_extensions import Literal


def feature_take_indices(num_features: int, indices: Optional[Union[int, List[int]]] = None,
                         as_set: bool = False) -> Tuple[List[int], int]:
    """None -> all, int n -> last n, sequence -> those (negative = from the end)."""
    if indices is None:
        indices = num_features
    if isinstance(indices, int):
        assert 0 < indices <= num_features, f'last-n ({indices}) is out of range (1 to {num_features})'
        take = [num_features - indices + i for i in range(indices)]
    else:
        take = []
        for i in indices:
            idx = num_features + i if i < 0 else i
            assert 0 <= idx < num_features, f'feature index {idx} is out of range (0 to {num_features - 1})'
            take.append(idx)
    return (set(take) if as_set else take), max(take)


def init_weights_vit_timm(module: nn.Module, name: str = '') -> None:
    """timm's original ViT init: every nn.Linear ~ trunc_normal(std=.02), zero bias."""
    if isinstance(module, nn.Linear):
        trunc_normal_(module.weight, std=.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif hasattr(module, 'init_weights'):
        module.init_weights()


def get_init_weights_vit(head_bias: float = 0.0) -> Callable:
    return init_weights_vit_timm


def global_pool_nlc(x: torch.Tensor, num_prefix_tokens: int = 1):
    return x[:, 0] if num_prefix_tokens == 1 else x[:, num_prefix_tokens:]


def _adaptive_embed(k_in: int, dim: int) -> nn.Sequential:
    """LN(K) -> Linear(K, D) -> LN(D) for pre-gathered adaptive patches (reference :282-289)."""
    return nn.Sequential(nn.LayerNorm(k_in), nn.Linear(k_in, dim), nn.LayerNorm(dim))


def _run_adaptive_embed(seq: nn.Sequential, x):
    ln0, lin, ln1 = seq[0], seq[1], seq[2]
    h = UF.layer_norm(x.contiguous(), ln0.weight, ln0.bias, ln0.eps)
    h = UF.linear(h, lin.weight, lin.bias)
    return UF.layer_norm(h, ln1.weight, ln1.bias, ln1.eps)


def _pos_dep_mlp(in_features: int, dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_features=in_features, out_features=dim), nn.GELU())


def _module_dtype(mod: nn.Module):
    """dtype the module's weights currently have.  Inside torch FSDP (use_orig_params=False) `weight` is a plain
    tensor view of the unsharded flat parameter during forward -- `mod.parameters()` is empty there -- and under
    MixedPrecision it is bf16 while the activations handed over may still be fp32."""
    for m in mod.modules():
        w = getattr(m, "weight", None)
        if torch.is_tensor(w) and w.is_floating_point():
            return w.dtype
    p = next(mod.parameters(), None)
    return None if p is None else p.dtype


def _torch_head(mod: nn.Module, x):
    """Heads the north-star leaves in PyTorch: run in the parameter dtype."""
    dt = _module_dtype(mod)
    return mod(x if dt is None else x.to(dt))


class VIT(nn.Module):
    def __init__(
            self,
            img_size: Union[int, Tuple[int, int], Tuple[int, int, int]] = 224,
            patch_size: Union[int, Tuple[int, int], Tuple[int, int, int]] = 16,
            in_chans: int = 3,
            num_classes: Optional[int] = None,
            embed_dim: int = 768,
            depth: int = 12,
            num_heads: int = 12,
            mlp_ratio: float = 4.,
            qkv_bias: bool = True,
            qk_norm: bool = False,
            init_values: Optional[float] = None,
            class_token: bool = True,
            pos_embed: str = 'learn',
            drop_rate: float = 0.,
            pos_drop_rate: float = 0.,
            patch_drop_rate: float = 0.,
            proj_drop_rate: float = 0.,
            attn_drop_rate: float = 0.,
            drop_path_rate: float = 0.,
            weight_init: Literal['skip', ''] = '',
            embed_layer: Callable = PatchEmbed,
            norm_layer=None,
            act_layer=None,
            block_fn: Type[nn.Module] = Block,
            mlp_layer: Type[nn.Module] = Mlp,
            twoD: Optional[bool] = True,
            adaptive_patching: Optional[bool] = False,
            fixed_length: Optional[int] = 4096,
            default_vars: List = None,
            single_channel: bool = False,
            use_varemb: bool = False,
            FusedAttn_option=FusedAttn.NONE,
            use_adaptive_pos_emb: bool = False,
            sqrt_len_method: bool = False,
    ) -> None:
        super().__init__()
        assert pos_embed in ('', 'none', 'learn')
        if isinstance(norm_layer, str) or isinstance(act_layer, str):
            raise NotImplementedError("string layer names need timm; pass a callable")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        if patch_drop_rate > 0:
            raise NotImplementedError("patch_drop_rate > 0 is never configured by the reference drivers")

        # ---- bookkeeping attributes (names are API: drivers and subclasses read them)
        self.norm_layer, self.act_layer, self.mlp_layer, self.block_fn = norm_layer, act_layer, mlp_layer, block_fn
        self.num_classes, self.embed_dim, self.depth, self.num_heads = num_classes, embed_dim, depth, num_heads
        self.num_prefix_tokens = 1 if class_token else 0
        self.in_chans, self.patch_size, self.img_size, self.twoD = in_chans, patch_size, img_size, twoD
        self.qkv_bias, self.qk_norm, self.init_values = qkv_bias, qk_norm, init_values
        self.drop_path_rate, self.proj_drop_rate, self.attn_drop_rate = drop_path_rate, proj_drop_rate, attn_drop_rate
        self.adaptive_patching, self.fixed_length = adaptive_patching, fixed_length
        self.default_vars, self.single_channel, self.use_varemb = default_vars, single_channel, use_varemb
        self.aggregated_variables = 1
        self.class_token = class_token
        self.FusedAttn_option = FusedAttn_option
        self.use_adaptive_pos_emb, self.sqrt_len_method = use_adaptive_pos_emb, sqrt_len_method

        # ---- token embedding: conv patch embed, or LN-Linear-LN over pre-gathered adaptive patches
        self._seq_tokens = adaptive_patching and not sqrt_len_method
        nsp = 2 if twoD else 3
        self.patch_dim_woc = patch_size ** nsp
        self.patch_dim = in_chans * self.patch_dim_woc
        n_vars = len(default_vars) if default_vars is not None else 0
        if self._seq_tokens:
            num_patches = fixed_length
        else:
            self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size,
                                           in_chans=1 if use_varemb else in_chans, embed_dim=embed_dim,
                                           twoD=twoD, sqrt_len_method=sqrt_len_method)
            num_patches = self.patch_embed.num_patches
            self.grid_size = self.patch_embed.grid_size
        self.num_patches = num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if class_token else None
        self.embed_len = num_patches + self.num_prefix_tokens
        self.pos_embed = None if (not pos_embed or pos_embed == 'none') else \
            nn.Parameter(torch.randn(1, self.embed_len, embed_dim) * .02)
        self.pos_drop = nn.Dropout(p=pos_drop_rate)
        self.patch_drop = nn.Identity()

        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[
            block_fn(dim=embed_dim, num_heads=num_heads, fused_attn=FusedAttn_option, mlp_ratio=mlp_ratio,
                     qkv_bias=qkv_bias, qk_norm=qk_norm, init_values=init_values, proj_drop=proj_drop_rate,
                     attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, act_layer=act_layer,
                     mlp_layer=mlp_layer)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head_drop = nn.Dropout(drop_rate)
        if num_classes is None:
            self.head = None
        else:
            self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        if self._seq_tokens:
            self.token_embeds = (nn.ModuleList([_adaptive_embed(self.patch_dim_woc, embed_dim) for _ in range(n_vars)])
                                 if use_varemb else _adaptive_embed(self.patch_dim, embed_dim))
        else:
            # every variable shares the single patch embedder (aliases in the state_dict)
            self.token_embeds = nn.ModuleList([self.patch_embed] * n_vars) if use_varemb else self.patch_embed

        if use_varemb:
            self.var_embed, self.var_map = self.create_var_embedding(embed_dim)
            if single_channel or n_vars == 1:
                self.var_query = self.var_agg = None
            else:
                self.var_query = nn.Parameter(torch.zeros(1, self.aggregated_variables, embed_dim), requires_grad=True)
                self.var_agg = VariableMapping_Attention(embed_dim, fused_attn=FusedAttn_option,
                                                         num_heads=num_heads, qkv_bias=False)
        if use_adaptive_pos_emb:
            self.adaptive_pos_dep_emb = _pos_dep_mlp(3 if twoD else 4, embed_dim)

        self._var_id_cache = {}
        if weight_init != 'skip':
            self.init_weights('')

    # ------------------------------------------------------------------ initialisation
    def _grid_counts(self):
        return [int(s / self.patch_size) for s in self.img_size]

    def _fill_sincos(self, param, cls_token):
        dim = param.shape[-1]
        g = self._grid_counts()
        table = (get_2d_sincos_pos_embed(dim, g[0], g[1], cls_token=cls_token) if self.twoD
                 else get_3d_sincos_pos_embed(dim, g[0], g[1], g[2], cls_token=cls_token))
        param.data.copy_(torch.from_numpy(table).float().unsqueeze(0))

    def _init_embedders(self):
        if self.cls_token is not None:
            nn.init.normal_(self.cls_token, std=1e-6)
        if not self.adaptive_patching:
            embs = list(self.token_embeds) if self.use_varemb else [self.token_embeds]
            for e in embs:
                w = e.proj.weight.data
                trunc_normal_(w.view([w.shape[0], -1]), std=0.02)
        if self.use_varemb:
            ve = get_1d_sincos_pos_embed_from_grid(self.var_embed.shape[-1], np.arange(len(self.default_vars)))
            self.var_embed.data.copy_(torch.from_numpy(ve).float().unsqueeze(0))

    def init_weights(self, mode: str = '') -> None:
        if (not self.adaptive_patching or self.sqrt_len_method) and self.pos_embed is not None:
            self._fill_sincos(self.pos_embed, self.class_token)
        self._init_embedders()
        named_apply(get_init_weights_vit(0.), self)

    def _init_weights_decoder_family(self) -> None:
        """MAE / DiffusionVIT variant (reference :603-651, :1167-1215): grids without cls row,
        decoder table too, and only for the non-adaptive case."""
        if not self.adaptive_patching:
            if self.pos_embed is not None:
                self._fill_sincos(self.pos_embed, False)
            if getattr(self, 'decoder_pos_embed', None) is not None:
                self._fill_sincos(self.decoder_pos_embed, False)
        self._init_embedders()
        named_apply(get_init_weights_vit(0.), self)

    # ------------------------------------------------------------------ variables
    def create_var_embedding(self, dim):
        var_map = {var: idx for idx, var in enumerate(self.default_vars)}
        return nn.Parameter(torch.zeros(1, len(self.default_vars), dim), requires_grad=True), var_map

    def get_var_ids(self, vars, device):
        key = (tuple(vars), str(device))
        ids = self._var_id_cache.get(key)
        if ids is None:
            ids = torch.from_numpy(np.array([self.var_map[v] for v in vars])).to(device)
            self._var_id_cache[key] = ids
        return ids

    def get_var_emb(self, var_emb, vars):
        return var_emb[:, self.get_var_ids(vars, var_emb.device), :]

    def aggregate_variables(self, x: torch.Tensor):
        """x: [B, V, L, D] -> [B, L, D]: the learned query attends over the V variable tokens of
        each location.  The query is projected once (shared), not B*L times."""
        b, v, l, d = x.shape
        x = x.permute(0, 2, 1, 3).reshape(b * l, v, d)
        x = self.var_agg(self.var_query, x, shared_query=True)        # [B*L, V~, D]
        if self.aggregated_variables > 1:
            return x.view(b, l, self.aggregated_variables, d).permute(0, 2, 1, 3)
        return x.view(b, l, d)

    # ------------------------------------------------------------------ forward pieces
    def _embed_one(self, embedder, x):
        if isinstance(embedder, nn.Sequential):
            return _run_adaptive_embed(embedder, x)
        return embedder(x)

    def _embed_tokens(self, x: torch.Tensor, variables, seq_tokens: Optional[bool] = None) -> torch.Tensor:
        """Input -> [B, L, D] tokens (variable embedding + aggregation included)."""
        seq_tokens = self._seq_tokens if seq_tokens is None else seq_tokens
        if not self.use_varemb:
            if seq_tokens:
                b, c, s, p = x.shape                                   # 'b c s p -> b s (p c)'
                x = x.permute(0, 2, 3, 1).reshape(b, s, p * c)
            return self._embed_one(self.token_embeds, x)

        variables = tuple(variables) if isinstance(variables, list) else variables
        var_ids = [self.var_map[v] for v in variables]                 # host ints: no device sync
        var_embed = self.get_var_emb(self.var_embed, variables)        # [1, V, D]
        if self.single_channel:
            inp = torch.squeeze(x) if self.adaptive_patching else x
            tok = self._embed_one(self.token_embeds[var_ids[0]], inp)
            return UF.add_bcast(tok, var_embed.unsqueeze(2).squeeze(1))
        V = len(var_ids)
        shared = all(self.token_embeds[i] is self.token_embeds[0] for i in var_ids)
        if shared and not self.adaptive_patching:
            b = x.shape[0]
            tok = self.token_embeds[0](x.reshape(b * V, 1, *x.shape[2:]))          # one GEMM for all V
            tok = tok.view(b, V, tok.shape[1], tok.shape[2])
        else:
            toks = []
            for i in range(V):
                xi = x[:, i:i + 1]
                toks.append(self._embed_one(self.token_embeds[var_ids[i]],
                                            torch.squeeze(xi) if self.adaptive_patching else xi))
            tok = torch.stack(toks, dim=1)
        tok = UF.add_bcast(tok, var_embed.unsqueeze(2))
        return self.aggregate_variables(tok)

    def _pos_embed(self, x: torch.Tensor, seq_ps) -> torch.Tensor:
        """cls-token concat + (learned | sin-cos | size/position-dependent) embedding add, one fused
        pass.  As in the reference, nothing is added (and no cls token is prepended) when the model
        has no pos_embed parameter."""
        if self.pos_embed is None:
            return x.view(x.shape[0], -1, x.shape[-1])
        adaptive = self.use_adaptive_pos_emb
        pe = _torch_head(self.adaptive_pos_dep_emb, seq_ps) if adaptive else self.pos_embed
        if x.dtype != torch.bfloat16 or not x.is_cuda or x.shape[-1] % 8:
            raise RuntimeError("ucf_vit_b200: tokens must be bf16 CUDA tensors (no CPU fallback)")
        prefix = self.cls_token.reshape(1, -1) if self.cls_token is not None else None
        # the learned table has a row for the cls token; the adaptive embedding has not (the
        # reference concatenates a zero row there, arch.py:377-390)
        x = UF.assemble_tokens(x, prefix, pe, pos_has_prefix=not adaptive)
        return self.pos_drop(x)

    def _final_norm(self, x):
        return _apply_norm(self.norm, x)

    def forward_features(self, x: torch.Tensor, variables, seq_ps) -> torch.Tensor:
        x = self._embed_tokens(x, variables)
        x = self._pos_embed(x, seq_ps)
        x = self.blocks(x)
        return self._final_norm(x)

    def pool(self, x: torch.Tensor) -> torch.Tensor:
        return global_pool_nlc(x, num_prefix_tokens=self.num_prefix_tokens)

    def forward_head(self, x: torch.Tensor) -> torch.Tensor:
        x = self.head_drop(self.pool(x))
        h = self.head
        # a plain Linear head whose pitches satisfy TMA alignment reuses the block GEMM ("the
        # Linears can reuse the a5/a7 GEMM kernel for free"); anything else stays in PyTorch
        if isinstance(h, nn.Linear) and x.is_cuda and h.in_features % 8 == 0 and h.out_features % 8 == 0:
            return UF.linear(x, h.weight, h.bias)
        return _torch_head(h, x)

    def forward(self, x: torch.Tensor, variables, seq_ps=None) -> torch.Tensor:
        return self.forward_head(self.forward_features(x, variables, seq_ps))


# ---------------------------------------------------------------------------------------------
class SAP(VIT):
    """Segmentation over adaptively-patched sequences: ViT encoder + transposed-conv neck."""

    def __init__(self, *args, **kwargs):
        self.sqrt_len = kwargs.pop('sqrt_len', '')
        super().__init__(*args, **kwargs)
        self.head = None
        p = self.patch_size
        if self.twoD:
            self.neck = nn.Sequential(nn.ConvTranspose2d(self.embed_dim, 256, kernel_size=(p, p), stride=(p, p), bias=False))
            self.mask_header = nn.Sequential(nn.Conv2d(256, self.num_classes, 1))
        else:
            self.neck = nn.Sequential(nn.ConvTranspose3d(self.embed_dim, 256, kernel_size=(p, p, p), stride=(p, p, p), bias=False))
            self.mask_header = nn.Sequential(nn.Conv3d(256, self.num_classes, 1))
        self.init_weights('')

    def mask_head(self, x: torch.Tensor):
        s = self.sqrt_len
        b, L, D = x.shape
        p = self.patch_size
        wn, wh, bh = self.neck[0].weight, self.mask_header[0].weight, self.mask_header[0].bias
        C, M = wh.shape[0], wh.shape[1]
        nd = 2 if self.twoD else 3
        P = p ** nd
        # ConvTranspose(k = s = p, no bias) followed by the 1x1 mask header (reference simple/arch.py:520-536) has no
        # non-linearity in between: the two linear maps are composed into ONE [D -> C * p^d] projection per token, which
        # runs on the tcgen05 GEMM.  Same function (up to fp reassociation), ~M/C = 64x fewer FLOPs than the 256-channel
        # transposed convolution and no [B, 256, (s p)^d] fp32 intermediate (1 GB per 1024^2 image).  Autograd splits the
        # composite weight's gradient back onto neck.weight and mask_header.weight through the einsum.
        w2 = torch.einsum('dmk,cm->ckd', wn.reshape(D, M, P), wh.reshape(C, M)).reshape(C * P, D)
        b2 = None if bh is None else bh.repeat_interleave(P)
        y = UF.linear(x.reshape(b * L, D), w2, b2)
        if self.twoD:
            y = y.view(b, s, s, C, p, p).permute(0, 3, 1, 4, 2, 5).reshape(b, C, s * p, s * p)
        else:
            y = y.view(b, s, s, s, C, p, p, p).permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(b, C, s * p, s * p, s * p)
        return y.to(wh.dtype)

    def forward_head(self, x: torch.Tensor) -> torch.Tensor:
        return self.mask_head(self.pool(x))


# ---------------------------------------------------------------------------------------------
class _DecoderMixin:
    """Light transformer decoder shared by MAE and DiffusionVIT."""

    def _build_decoder(self, with_mask_token: bool, adaptive_pos: bool):
        if self.linear_decoder:
            self.decoder_pred = nn.Linear(self.embed_dim, self.patch_dim)
            if with_mask_token:
                self.mask_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
            self.decoder_pos_embed = None
            return
        dd = self.decoder_embed_dim
        self.decoder_pred = nn.Linear(dd, self.patch_dim)
        if with_mask_token:
            self.mask_token = nn.Parameter(torch.zeros(1, 1, dd))
        self.decoder_embed = nn.Linear(self.embed_dim, dd)
        self.decoder_norm = nn.LayerNorm(dd)
        if adaptive_pos:
            self.decoder_pos_embed = None
        elif self.adaptive_patching:
            self.decoder_pos_embed = nn.Parameter(torch.randn(1, self.num_patches, dd) * .02)
        else:
            self.decoder_pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, dd))
        dpr = [v.item() for v in torch.linspace(0, self.drop_path_rate, self.decoder_depth)]
        self.decoder_blocks = nn.Sequential(*[
            self.block_fn(dim=dd, num_heads=self.decoder_num_heads, fused_attn=self.FusedAttn_option,
                          mlp_ratio=self.mlp_ratio_decoder, qkv_bias=self.qkv_bias, qk_norm=self.qk_norm,
                          init_values=self.init_values, proj_drop=self.proj_drop_rate, attn_drop=self.attn_drop_rate,
                          drop_path=dpr[i], norm_layer=self.norm_layer, act_layer=self.act_layer,
                          mlp_layer=self.mlp_layer)
            for i in range(self.decoder_depth)])
        if adaptive_pos:
            self.decoder_adaptive_pos_dep_emb = _pos_dep_mlp(3 if self.twoD else 4, dd)

    def _decode(self, x, pos):
        if pos is not None:
            x = UF.assemble_tokens(x, None, pos, True)
        x = self.decoder_blocks(x)
        x = UF.layer_norm(x, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm.eps)
        return UF.linear(x, self.decoder_pred.weight, self.decoder_pred.bias)


class MAE(_DecoderMixin, VIT):
    """Masked auto-encoder: random token drop, encoder on the kept tokens, light decoder."""

    def __init__(self, *args, **kwargs):
        for k in ('mask_ratio', 'linear_decoder', 'decoder_depth', 'decoder_embed_dim', 'decoder_num_heads',
                  'mlp_ratio_decoder'):
            setattr(self, k, kwargs.pop(k, ''))
        super().__init__(*args, **kwargs)
        self.head = None
        self._build_decoder(with_mask_token=True, adaptive_pos=self.use_adaptive_pos_emb)
        self.init_weights('')

    def init_weights(self, mode: str = '') -> None:
        self._init_weights_decoder_family()

    def random_masking(self, sequence, noise=None):
        """Keep the int(L*(1-r)) lowest-noise tokens per sample.  Returns (kept, mask, ids_restore);
        mask is 1 for removed tokens, in original order.  One kernel ranks the noise (both permutations and the
        mask), one gathers the kept rows."""
        if sequence.dim() != 3:
            raise ValueError(f"random_masking expects [B, L, D] tokens, got {tuple(sequence.shape)}")
        batch_size, seq_length, dim = sequence.shape
        len_keep = int(seq_length * (1 - self.mask_ratio))
        if noise is None:
            noise = torch.rand(batch_size, seq_length, device=sequence.device)
        ids_shuffle, ids_restore, mask = ops.mask_plan(noise.to(torch.float32).contiguous(), len_keep)
        kept = UF.gather_tokens(sequence, ids_shuffle[:, :len_keep].contiguous())
        return kept, mask, ids_restore

    def mask_head(self, x: torch.Tensor, ids_restore, seq_ps):
        if not self.linear_decoder:
            x = UF.linear(x, self.decoder_embed.weight, self.decoder_embed.bias)
        # cat(x, mask tokens) -> gather(ids_restore) (-> + decoder position embedding): one pass
        if self.linear_decoder:
            full = UF.gather_tokens(x, ids_restore, fill=self.mask_token, complete=True)
            return UF.linear(full, self.decoder_pred.weight, self.decoder_pred.bias)
        pos = (_torch_head(self.decoder_adaptive_pos_dep_emb, seq_ps) if self.use_adaptive_pos_emb
               else self.decoder_pos_embed)
        full = UF.gather_tokens(x, ids_restore, fill=self.mask_token, pos=pos, complete=True)
        return self._decode(full, None)

    def forward_features(self, x: torch.Tensor, variables, seq_ps, noise=None):
        # NB: like the reference (:735-736) every adaptive input is a pre-gathered sequence here
        x = self._embed_tokens(x, variables, seq_tokens=self.adaptive_patching)
        x = self._pos_embed(x, seq_ps)
        x, mask, ids_restore = self.random_masking(x, noise)
        x = self.blocks(x.contiguous())
        return self._final_norm(x), mask, ids_restore

    def forward_head(self, x: torch.Tensor, ids_restore, seq_ps):
        return self.mask_head(self.pool(x), ids_restore, seq_ps)

    def forward(self, x: torch.Tensor, variables, seq_ps=None):
        x, mask, ids_restore = self.forward_features(x, variables, seq_ps)
        return self.forward_head(x, ids_restore, seq_ps), mask


# ---------------------------------------------------------------------------------------------
class UNETR(VIT):
    """ViT encoder + UNETR convolutional decoder.  The encoder (patch embed, variable aggregation,
    blocks) runs on this package's kernels; the conv decoder stays on cuDNN."""

    def __init__(self, *args, **kwargs):
        for k in ('linear_decoder', 'feature_size', 'skip_connection', 'sqrt_len'):
            setattr(self, k, kwargs.pop(k, ''))
        super().__init__(*args, **kwargs)
        self.head = None
        nsp = 2 if self.twoD else 3
        self.feat_size = ((self.sqrt_len,) * nsp if self.adaptive_patching
                          else tuple(int(self.img_size[i] / self.patch_size) for i in range(nsp)))
        fs, ed = self.feature_size, self.embed_dim
        if self.linear_decoder:
            self.mlp_head = nn.Linear(ed, self.num_classes)
            self.upsample = nn.Upsample(scale_factor=self.patch_size, mode='trilinear', align_corners=True)
        else:
            full_res = self.feat_size[0] * 16 == self.img_size[0]
            if self.skip_connection:
                inc = self.depth // 4
                self.skip_indices = [(i + 1) * inc for i in range(3)]
                self.encoder1 = UnetrBasicBlock(spatial_dims=nsp, in_channels=self.in_chans, out_channels=fs,
                                                kernel_size=3, stride=1, norm_name="instance", res_block=True)
                for name, mult, layers in (("encoder2", 2, 2), ("encoder3", 4, 1), ("encoder4", 8, 0)):
                    setattr(self, name, UnetrPrUpBlock(spatial_dims=nsp, in_channels=ed, out_channels=fs * mult,
                                                       num_layer=layers, kernel_size=3, stride=1,
                                                       upsample_kernel_size=2, norm_name="instance",
                                                       conv_block=True, res_block=True))
                chain = (("decoder5", ed, fs * 8, 2), ("decoder4", fs * 8, fs * 4, 2), ("decoder3", fs * 4, fs * 2, 2),
                         ("decoder2", fs * 2, fs, 2 if full_res else 1))
                for name, cin, cout, up in chain:
                    setattr(self, name, UnetrUpBlock(spatial_dims=nsp, in_channels=cin, out_channels=cout,
                                                     kernel_size=3, upsample_kernel_size=up, norm_name="instance",
                                                     res_block=True))
            else:
                for name, cin, cout in (("decoder5", ed, fs * 8), ("decoder4", fs * 8, fs * 4),
                                        ("decoder3", fs * 4, fs * 2), ("decoder2", fs * 2, fs)):
                    setattr(self, name, MyUnetBlock(spatial_dims=nsp, in_channels=cin, out_channels=cout,
                                                    upsample_kernel_size=2, res_block=True))
            self.out = UnetOutBlock(spatial_dims=nsp, in_channels=fs, out_channels=self.num_classes)
            if not full_res:
                self.upsample = nn.Upsample(size=self.img_size, mode='trilinear', align_corners=True)
        self.init_weights('')

    def proj_feat(self, x, hidden_size, feat_size):
        x = x.view(x.size(0), *feat_size, hidden_size)
        return x.permute(0, 3, 1, 2) if self.twoD else x.permute(0, 4, 1, 2, 3)

    def _conv_dtype(self):
        return self.out.conv.conv.weight.dtype

    def unetr_head(self, x: torch.Tensor, intermediates, enc1):
        if self.linear_decoder and not self.skip_connection:
            x = _torch_head(self.mlp_head, x)
            g = self.grid_size
            x = x.view(x.shape[0], *g, x.shape[-1])
            x = x.permute(0, 3, 1, 2) if self.twoD else x.permute(0, 4, 1, 2, 3)
            return self.upsample(x)
        cd = self._conv_dtype()
        feat = lambda t: self.proj_feat(t.to(cd), self.embed_dim, self.feat_size)  # noqa: E731
        resize = self.feat_size[0] * 16 != self.img_size[0]
        if not self.skip_connection:
            out = self.decoder2(self.decoder3(self.decoder4(self.decoder5(feat(x)))))
            if resize:
                out = self.upsample(out)
            return self.out(out)
        n = len(intermediates)
        dec3 = self.decoder5(feat(x), self.encoder4(feat(intermediates[n - 1])))
        dec2 = self.decoder4(dec3, self.encoder3(feat(intermediates[n - 2])))
        dec1 = self.decoder3(dec2, self.encoder2(feat(intermediates[n - 3])))
        if resize:
            dec1 = self.upsample(dec1)
        return self.out(self.decoder2(dec1, enc1))

    def forward_intermediates(self, x: torch.Tensor, variables, seq_ps,
                              indices: Optional[Union[int, List[int]]] = None, return_prefix_tokens: bool = False,
                              norm: bool = False, stop_early: bool = False, intermediates_only: bool = False):
        take, max_index = feature_take_indices(len(self.blocks), indices)
        x = self._embed_tokens(x, variables)
        x = self._pos_embed(x, seq_ps)
        blocks = self.blocks if not stop_early else self.blocks[:max_index + 1]
        inter = []
        for i, blk in enumerate(blocks):
            x = blk(x)
            if i in take:
                inter.append(self._final_norm(x) if norm else x)
        if self.num_prefix_tokens:
            prefix = [y[:, :self.num_prefix_tokens] for y in inter]
            inter = [y[:, self.num_prefix_tokens:] for y in inter]
            if return_prefix_tokens:
                inter = list(zip(inter, prefix))
        if intermediates_only:
            return inter
        return self._final_norm(x), inter

    def forward_head(self, x: torch.Tensor, intermediates, enc1):
        return self.unetr_head(self.pool(x), intermediates, enc1)

    # Convolutional decoder precision.  None (default): cuDNN in the parameter dtype, like the reference.  torch.bfloat16:
    # the decoder's convolutions run under autocast in bf16 (fp32 accumulate; InstanceNorm statistics stay fp32) -- not the
    # reference's arithmetic, opt-in for throughput (bench.py --config unetr_128 states which one it timed).
    conv_autocast_dtype = None

    def use_fused_decoder(self, on: bool = True):
        """Channels-last bf16 decoder: cuDNN convolutions under bf16 autocast (fp32 accumulate) with this package's fused
        InstanceNorm + residual + LeakyReLU kernels between them (`utils/unetr_blocks.py`, `csrc/instnorm.cu`)."""
        self.conv_autocast_dtype = torch.bfloat16 if on else None
        for m in self.modules():
            if hasattr(m, "ndhwc_bf16"):
                m.ndhwc_bf16 = bool(on)
        return self

    def _conv_ctx(self):
        import contextlib
        if self.conv_autocast_dtype is None:
            return contextlib.nullcontext()
        return torch.autocast("cuda", dtype=self.conv_autocast_dtype)

    def forward(self, x: torch.Tensor, variables, seq_ps=None, x_seq=None) -> torch.Tensor:
        tokens_in = x_seq if self.adaptive_patching else x
        if self.skip_connection:
            with self._conv_ctx():
                enc1 = _torch_head(self.encoder1, x)
            feats, inter = self.forward_intermediates(tokens_in, variables, seq_ps, indices=self.skip_indices)
            with self._conv_ctx():
                return self.forward_head(feats, inter, enc1)
        feats = self.forward_features(tokens_in, variables, seq_ps)
        with self._conv_ctx():
            return self.forward_head(feats, None, None)


# ---------------------------------------------------------------------------------------------
class DiffusionVIT(_DecoderMixin, VIT):
    """Noise-prediction ViT: encoder conditioned on a sinusoidal time-step embedding + decoder."""

    def __init__(self, *args, **kwargs):
        for k in ('linear_decoder', 'decoder_depth', 'decoder_embed_dim', 'decoder_num_heads', 'mlp_ratio_decoder',
                  'time_steps'):
            setattr(self, k, kwargs.pop(k, ''))
        super().__init__(*args, **kwargs)
        self.head = None
        self.temporalEmbeddings = SinusoidalEmbeddings(time_steps=self.time_steps, embed_dim=self.embed_dim)
        self.timeEmbeddingMap = EmbeddingDenseLayer(self.embed_dim, self.embed_dim, 0.5)
        self._build_decoder(with_mask_token=False, adaptive_pos=False)
        self.init_weights('')

    def init_weights(self, mode: str = '') -> None:
        self._init_weights_decoder_family()

    def forward_features(self, x: torch.Tensor, t, variables) -> torch.Tensor:
        x = self._embed_tokens(x, variables, seq_tokens=self.adaptive_patching)
        x = self._pos_embed(x, None)
        temb = self.temporalEmbeddings(x, t)
        temb = _torch_head(self.timeEmbeddingMap, temb)[:, None, :]
        x = UF.add_bcast(x, temb)
        x = self.blocks(x)
        return self._final_norm(x)

    def forward_head(self, x: torch.Tensor):
        x = self.pool(x)
        if self.linear_decoder:
            return UF.linear(x, self.decoder_pred.weight, self.decoder_pred.bias)
        x = UF.linear(x, self.decoder_embed.weight, self.decoder_embed.bias)
        return self._decode(x, self.decoder_pos_embed)

    def forward(self, x: torch.Tensor, t, variables) -> torch.Tensor:
        return self.forward_head(self.forward_features(x, t, variables))
