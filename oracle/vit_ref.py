"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

CPU fp32 restatement of the reference's hot path as pure functions over a state_dict (plain
`torch` tensor math; no nn.Module, no autocast, no fused SDPA), so that the arithmetic each
sm_100a kernel must reproduce is written down once, explicitly.  Gradients come from torch
autograd over these functions.  Allowed importers: tests/, __graft_entry__.smoke(), and the
`cpu_baseline` / `--impl reference` legs of bench.py.

Pinned against the reference itself: `oracle/gen_golden.py` imports the unmodified modules from
/root/reference/src (behind oracle/shims) in the build container, runs them on seeded inputs and
asserts this restatement matches to fp32 round-off before writing tests/golden/*.npz;
tests/test_oracle_golden.py re-checks the restatement against those committed vectors.
The reference ships no tests or golden vectors of its own.

Every function cites the reference lines it restates (paths relative to
/root/reference/src/UCF_VIT/).
"""
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# building blocks                                     simple/building_blocks.py
# ------------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps):
    """nn.LayerNorm over the last dim (biased variance)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    if w is not None:
        y = y * w
    if b is not None:
        y = y + b
    return y


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def gelu_erf(x):
    """nn.GELU() default (exact erf form), building_blocks.py:102,116."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def patch_embed(x, w, b, p: int):
    """PatchEmbed.forward, building_blocks.py:77-92: Conv{2,3}d(k=s=p) then flatten(2).transpose(1,2).
    Restated as an explicit patch gather + matmul; K order is (c, p0, p1[, p2])."""
    B, C = x.shape[:2]
    G = [s // p for s in x.shape[2:]]
    if x.dim() == 4:
        rows = x.reshape(B, C, G[0], p, G[1], p).permute(0, 2, 4, 1, 3, 5).reshape(B, G[0] * G[1], C * p * p)
    else:
        rows = x.reshape(B, C, G[0], p, G[1], p, G[2], p).permute(0, 2, 4, 6, 1, 3, 5, 7)
        rows = rows.reshape(B, G[0] * G[1] * G[2], C * p ** 3)
    return linear(rows, w.reshape(w.shape[0], -1), b)


def attention(x, sd: SD, pre: str, H: int):
    """Attention.forward, building_blocks.py:157-192 (FusedAttn.NONE branch :181-187, which the
    FLASH / CK / DEFAULT branches equal mathematically); qk_norm off, dropout 0."""
    B, N, C = x.shape
    hd = C // H
    qkv = linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias")).reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return linear(o, sd[pre + "proj.weight"], sd.get(pre + "proj.bias"))


def mlp(x, sd: SD, pre: str):
    """Mlp.forward, building_blocks.py:122-129 (drop 0, norm Identity)."""
    h = gelu_erf(linear(x, sd[pre + "fc1.weight"], sd.get(pre + "fc1.bias")))
    return linear(h, sd[pre + "fc2.weight"], sd.get(pre + "fc2.bias"))


def block(x, sd: SD, pre: str, H: int, eps: float = 1e-6):
    """Block.forward, building_blocks.py:236-239 (LayerScale / DropPath identity)."""
    x = x + attention(layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps), sd, pre + "attn.", H)
    x = x + mlp(layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps), sd, pre + "mlp.")
    return x


def blocks(x, sd: SD, pre: str, depth: int, H: int, eps: float = 1e-6, take: Optional[Sequence[int]] = None):
    inter = []
    for i in range(depth):
        x = block(x, sd, f"{pre}{i}.", H, eps)
        if take is not None and i in take:
            inter.append(x)
    return (x, inter) if take is not None else x


def var_aggregate(x, sd: SD, H: int):
    """VIT.aggregate_variables, simple/arch.py:414-432 + VariableMapping_Attention.forward,
    building_blocks.py:321-373.  x: [B, V, L, D] -> [B, L, D]."""
    B, V, L, D = x.shape
    hd = D // H
    t = x.permute(0, 2, 1, 3).reshape(B * L, V, D)
    q = linear(sd["var_query"].expand(B * L, -1, -1), sd["var_agg.q.weight"], sd.get("var_agg.q.bias"))
    q = q.reshape(B * L, 1, H, hd).permute(0, 2, 1, 3)
    kv = linear(t, sd["var_agg.kv.weight"], sd.get("var_agg.kv.bias")).reshape(B * L, V, 2, H, hd).permute(2, 0, 3, 1, 4)
    att = ((q * hd ** -0.5) @ kv[0].transpose(-2, -1)).softmax(dim=-1)
    o = (att @ kv[1]).transpose(1, 2).reshape(B * L, 1, D)
    o = linear(o, sd["var_agg.proj.weight"], sd.get("var_agg.proj.bias"))
    return o.reshape(B, L, D)


# ------------------------------------------------------------------------------------------------
# model classes                                                     simple/arch.py
# ------------------------------------------------------------------------------------------------
def embed_tokens(x, sd: SD, cfg: dict, var_ids: Optional[List[int]] = None):
    """Token embedding part of VIT.forward_features, simple/arch.py:434-469 (non-single-channel)."""
    p, H = cfg["patch_size"], cfg["num_heads"]
    if cfg.get("use_varemb", False):
        V = x.shape[1]
        toks = [patch_embed(x[:, i:i + 1], sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], p) for i in range(V)]
        t = torch.stack(toks, dim=1) + sd["var_embed"][:, var_ids].unsqueeze(2)
        return var_aggregate(t, sd, H)
    if cfg.get("seq_tokens", False):
        # adaptive, pre-gathered: 'b c s p -> b s (p c)', LN(K) -> Linear -> LN(D)  (arch.py:282-289,466-467)
        b, c, s, pp = x.shape
        t = x.permute(0, 2, 3, 1).reshape(b, s, pp * c)
        t = layer_norm(t, sd["token_embeds.0.weight"], sd["token_embeds.0.bias"], 1e-5)
        t = linear(t, sd["token_embeds.1.weight"], sd["token_embeds.1.bias"])
        return layer_norm(t, sd["token_embeds.2.weight"], sd["token_embeds.2.bias"], 1e-5)
    return patch_embed(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], p)


def pos_embed(x, sd: SD, cfg: dict, seq_ps=None):
    """VIT._pos_embed, simple/arch.py:367-393."""
    if "pos_embed" not in sd:
        return x
    if cfg.get("use_adaptive_pos_emb", False):
        pe = gelu_erf(linear(seq_ps, sd["adaptive_pos_dep_emb.0.weight"], sd["adaptive_pos_dep_emb.0.bias"]))
    else:
        pe = sd["pos_embed"]
    if "cls_token" in sd:
        x = torch.cat([sd["cls_token"].expand(x.shape[0], -1, -1), x], dim=1)
        if cfg.get("use_adaptive_pos_emb", False):
            pe = torch.cat([torch.zeros(x.shape[0], 1, x.shape[-1]), pe], dim=1)
    return x + pe


def vit_features(x, sd: SD, cfg: dict, var_ids=None, seq_ps=None, take=None):
    t = pos_embed(embed_tokens(x, sd, cfg, var_ids), sd, cfg, seq_ps)
    out = blocks(t, sd, "blocks.", cfg["depth"], cfg["num_heads"], 1e-6, take)
    if take is not None:
        t, inter = out
        return layer_norm(t, sd["norm.weight"], sd["norm.bias"], 1e-6), inter
    return layer_norm(out, sd["norm.weight"], sd["norm.bias"], 1e-6)


def vit_forward(x, sd: SD, cfg: dict, var_ids=None, seq_ps=None):
    """VIT.forward, simple/arch.py:434-489: features -> cls pool -> head."""
    f = vit_features(x, sd, cfg, var_ids, seq_ps)
    pooled = f[:, 0] if "cls_token" in sd else f
    return linear(pooled, sd["head.weight"], sd["head.bias"])


def mae_forward(x, sd: SD, cfg: dict, noise, var_ids=None, seq_ps=None):
    """MAE.forward, simple/arch.py:663-755 with the masking noise made explicit."""
    t = pos_embed(embed_tokens(x, sd, cfg, var_ids), sd, cfg, seq_ps)
    B, L, D = t.shape
    keep = int(L * (1 - cfg["mask_ratio"]))
    ids_shuffle = torch.argsort(noise, dim=1)
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    t = torch.gather(t, 1, ids_shuffle[:, :keep].unsqueeze(-1).repeat(1, 1, D))
    mask = torch.ones(B, L)
    mask[:, :keep] = 0
    mask = torch.gather(mask, 1, ids_restore)
    t = blocks(t, sd, "blocks.", cfg["depth"], cfg["num_heads"])
    t = layer_norm(t, sd["norm.weight"], sd["norm.bias"], 1e-6)
    t = linear(t, sd["decoder_embed.weight"], sd["decoder_embed.bias"])
    full = torch.cat([t, sd["mask_token"].repeat(B, L - keep, 1)], dim=1)
    full = torch.gather(full, 1, ids_restore.unsqueeze(-1).repeat(1, 1, t.shape[2]))
    full = full + sd["decoder_pos_embed"]
    full = blocks(full, sd, "decoder_blocks.", cfg["decoder_depth"], cfg["decoder_num_heads"])
    full = layer_norm(full, sd["decoder_norm.weight"], sd["decoder_norm.bias"], 1e-5)
    return linear(full, sd["decoder_pred.weight"], sd["decoder_pred.bias"]), mask


def diffusion_forward(x, t_steps, sd: SD, cfg: dict, time_table):
    """DiffusionVIT.forward, simple/arch.py:1217-1283 in eval mode (the 0.5 dropout of the time
    MLP is the identity), with `_pos_embed(x, None)`."""
    t = pos_embed(embed_tokens(x, sd, cfg), sd, cfg)
    temb = time_table[t_steps]
    temb = linear(torch.relu(linear(temb, sd["timeEmbeddingMap.linear1.weight"], sd["timeEmbeddingMap.linear1.bias"])),
                  sd["timeEmbeddingMap.linear2.weight"], sd["timeEmbeddingMap.linear2.bias"])
    t = t + temb[:, None, :]
    t = blocks(t, sd, "blocks.", cfg["depth"], cfg["num_heads"])
    t = layer_norm(t, sd["norm.weight"], sd["norm.bias"], 1e-6)
    t = linear(t, sd["decoder_embed.weight"], sd["decoder_embed.bias"]) + sd["decoder_pos_embed"]
    t = blocks(t, sd, "decoder_blocks.", cfg["decoder_depth"], cfg["decoder_num_heads"])
    t = layer_norm(t, sd["decoder_norm.weight"], sd["decoder_norm.bias"], 1e-5)
    return linear(t, sd["decoder_pred.weight"], sd["decoder_pred.bias"])


def sap_forward(x, sd: SD, cfg: dict, seq_ps):
    """SAP.forward, simple/arch.py:491-536: ViT features -> (sqrt_len x sqrt_len) map ->
    ConvTranspose(k=s=p) neck -> 1x1 conv."""
    f = vit_features(x, sd, cfg, None, seq_ps)
    s = cfg["sqrt_len"]
    B, _, C = f.shape
    m = f.reshape(B, s, s, C).permute(0, 3, 1, 2)
    m = F.conv_transpose2d(m, sd["neck.0.weight"], None, stride=cfg["patch_size"])
    return F.conv2d(m, sd["mask_header.0.weight"], sd["mask_header.0.bias"])


# ---- UNETR decoder (MONAI blocks restated; see ucf_vit_b200/utils/unetr_blocks.py, parity unpinned)
def _conv(x, w, b=None, stride=1, pad=0):
    return (F.conv3d if x.dim() == 5 else F.conv2d)(x, w, b, stride=stride, padding=pad)


def _convT(x, w, stride):
    return (F.conv_transpose3d if x.dim() == 5 else F.conv_transpose2d)(x, w, None, stride=stride)


def _inorm(x):
    return F.instance_norm(x, eps=1e-5)


def _res_block(x, sd: SD, pre: str):
    """UnetResBlock: conv3-IN-lrelu-conv3-IN (+ 1x1 conv-IN on the skip when channels change)."""
    out = F.leaky_relu(_inorm(_conv(x, sd[pre + "conv1.conv.weight"], pad=1)), 0.01)
    out = _inorm(_conv(out, sd[pre + "conv2.conv.weight"], pad=1))
    res = x
    if pre + "conv3.conv.weight" in sd:
        res = _inorm(_conv(x, sd[pre + "conv3.conv.weight"]))
    return F.leaky_relu(out + res, 0.01)


def _pr_up(x, sd: SD, pre: str, n_layer: int):
    x = _convT(x, sd[pre + "transp_conv_init.conv.weight"], 2)
    for i in range(n_layer):
        x = _convT(x, sd[f"{pre}blocks.{i}.0.conv.weight"], 2)
        x = _res_block(x, sd, f"{pre}blocks.{i}.1.")
    return x


def _up(x, skip, sd: SD, pre: str, stride: int):
    x = _convT(x, sd[pre + "transp_conv.conv.weight"], stride)
    return _res_block(torch.cat([x, skip], dim=1), sd, pre + "conv_block.")


def unetr_forward(x, sd: SD, cfg: dict, var_ids=None):
    """UNETR.forward (skip_connection=True, non-adaptive), simple/arch.py:960-1113."""
    depth = cfg["depth"]
    inc = depth // 4
    take = [(i + 1) * inc for i in range(3)]
    f, inter = vit_features(x, sd, cfg, var_ids, None, take)
    return unetr_decode(x, f, inter, sd, cfg)


def unetr_decode(x, f, inter, sd: SD, cfg: dict):
    """The convolutional half of UNETR.forward (simple/arch.py:1040-1113) on given encoder features `f` (final, normed)
    and `inter` (the three skip features): lets a test feed the decoder exactly the features another encoder produced."""
    nsp = x.dim() - 2
    g = [s // cfg["patch_size"] for s in x.shape[2:]]

    def feat(t):
        t = t.reshape(t.shape[0], *g, t.shape[-1])
        return t.permute(0, 3, 1, 2) if nsp == 2 else t.permute(0, 4, 1, 2, 3)

    enc1 = _res_block(x, sd, "encoder1.layer.")
    dec3 = _up(feat(f), _pr_up(feat(inter[2]), sd, "encoder4.", 0), sd, "decoder5.", 2)
    dec2 = _up(dec3, _pr_up(feat(inter[1]), sd, "encoder3.", 1), sd, "decoder4.", 2)
    dec1 = _up(dec2, _pr_up(feat(inter[0]), sd, "encoder2.", 2), sd, "decoder3.", 2)
    full_res = g[0] * 16 == x.shape[2]
    if not full_res:
        dec1 = F.interpolate(dec1, size=tuple(x.shape[2:]), mode="trilinear" if nsp == 3 else "bilinear", align_corners=True)
    out = _up(dec1, enc1, sd, "decoder2.", 2 if full_res else 1)
    return _conv(out, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])


# ------------------------------------------------------------------------------------------------
# losses / targets next to the path                                   utils/misc.py, utils/metrics.py
# ------------------------------------------------------------------------------------------------
def patchify_target(data, p: int, twoD: bool):
    """utils/misc.py:14-33 (channel-fastest inside a patch)."""
    n, c = data.shape[:2]
    g = [s // p for s in data.shape[2:]]
    if twoD:
        return data.reshape(n, c, g[0], p, g[1], p).permute(0, 2, 4, 3, 5, 1).reshape(n, g[0] * g[1], p * p * c)
    t = data.reshape(n, c, g[0], p, g[1], p, g[2], p).permute(0, 2, 4, 6, 3, 5, 7, 1)
    return t.reshape(n, g[0] * g[1] * g[2], p ** 3 * c)


def masked_mse(pred, y, mask):
    """utils/metrics.py:11-17."""
    return ((((pred - y) ** 2).mean(dim=-1)) * mask).sum() / mask.sum()
