"""Fixed sin-cos position / variable / time-step embeddings (init-time numpy, host only).

Same function names, argument order and float64 results as
/root/reference/src/UCF_VIT/utils/pos_embed.py:16-102 so that models initialised here are
bit-identical to the reference's at construction.
"""
import math

import numpy as np
import torch
import torch.nn as nn


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    """pos: any array of positions (flattened to M) -> (M, embed_dim) = [sin(pos*w) | cos(pos*w)]."""
    assert embed_dim % 2 == 0
    half = embed_dim // 2
    omega = 1.0 / 10000 ** (np.arange(half, dtype=float) / (embed_dim / 2.0))
    ang = np.einsum("m,d->md", np.asarray(pos).reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim, grid):
    assert embed_dim % 2 == 0
    first = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    second = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    return np.concatenate([first, second], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size_h, grid_size_w, cls_token=False):
    """(H*W [+1], D): row-major over (h, w); the first D/2 channels encode the w index, the last
    D/2 the h index (meshgrid is called with w first, as in MAE)."""
    gw, gh = np.meshgrid(np.arange(grid_size_w, dtype=np.float32), np.arange(grid_size_h, dtype=np.float32))
    grid = np.stack([gw, gh], axis=0).reshape([2, 1, grid_size_h, grid_size_w])
    emb = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


def get_3d_sincos_pos_embed(embed_dim, grid_size_h, grid_size_w, grid_size_d, cls_token=False):
    """(H*W*Z, D): row-major over (h, w, d); channel thirds encode h, w, d.  `cls_token` is
    accepted and ignored, exactly like the reference."""
    assert embed_dim % 3 == 0
    third = embed_dim // 3
    e_d = get_1d_sincos_pos_embed_from_grid(third, np.arange(grid_size_d))
    e_w = get_1d_sincos_pos_embed_from_grid(third, np.arange(grid_size_w))
    e_h = get_1d_sincos_pos_embed_from_grid(third, np.arange(grid_size_h))
    e_d = np.tile(e_d, (grid_size_h * grid_size_w, 1))
    e_w = np.tile(np.repeat(e_w, grid_size_d, axis=0), (grid_size_h, 1))
    e_h = np.repeat(e_h, grid_size_w * grid_size_d, axis=0)
    return np.concatenate((e_h, e_w, e_d), axis=1)


class SinusoidalEmbeddings(nn.Module):
    """Diffusion time-step table [time_steps, embed_dim]; even channels sin, odd channels cos.
    Kept as a plain attribute (not a buffer) so state_dict keys match the reference; the table is
    moved to the activation's device once and cached (the reference re-uploads it every step)."""

    def __init__(self, time_steps: int, embed_dim: int):
        super().__init__()
        position = torch.arange(time_steps).unsqueeze(1).float()
        div = torch.exp(torch.arange(0, embed_dim, 2).float() * -(math.log(10000.0) / embed_dim))
        table = torch.zeros(time_steps, embed_dim, requires_grad=False)
        table[:, 0::2] = torch.sin(position * div)
        table[:, 1::2] = torch.cos(position * div)
        self.embeddings = table

    def forward(self, x, t):
        if self.embeddings.device != x.device:
            self.embeddings = self.embeddings.to(x.device)
        return self.embeddings[t.to(x.device)]


def interpolate_pos_embed(model, checkpoint_model, new_size=(64, 128)):
    key = "net.pos_embed"
    if key not in checkpoint_model:
        return
    pe = checkpoint_model[key]
    dim, n_old = pe.shape[-1], pe.shape[-2]
    ratio = 2
    old_h = int((n_old // ratio) ** 0.5)
    old = (old_h, ratio * old_h)
    new = (new_size[0] // model.patch_size, new_size[1] // model.patch_size)
    if old[0] != new[0]:
        print("Interpolate PEs from %dx%d to %dx%d" % (old[0], old[1], new[0], new[1]))
        tok = pe.reshape(-1, old[0], old[1], dim).permute(0, 3, 1, 2)
        tok = torch.nn.functional.interpolate(tok, size=new, mode="bicubic", align_corners=False)
        checkpoint_model[key] = tok.permute(0, 2, 3, 1).flatten(1, 2)


def interpolate_channel_embed(checkpoint_model, new_len):
    key = "net.channel_embed"
    if key in checkpoint_model and new_len <= checkpoint_model[key].shape[1]:
        checkpoint_model[key] = checkpoint_model[key][:, :new_len]
