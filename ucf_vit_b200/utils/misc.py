"""Host-side helpers of /root/reference/src/UCF_VIT/utils/misc.py that sit next to the hot path:
patchify / unpatchify (MAE & diffusion targets, :14-56), optimizer / scheduler factories (:58-96).
Pure index shuffles and host logic -- no kernel."""
import torch

from .lr_scheduler import LinearWarmupCosineAnnealingLR


def patchify(data, patch_size, twoD):
    """[N,C,X,Y(,Z)] -> [N, L, p^d * C] with the channel FASTEST inside a patch (pixel-major)."""
    n, c = data.shape[:2]
    p = patch_size
    g = [s // p for s in data.shape[2:]]
    if twoD:
        t = data.reshape(n, c, g[0], p, g[1], p).permute(0, 2, 4, 3, 5, 1)
        return t.reshape(n, g[0] * g[1], p * p * c)
    t = data.reshape(n, c, g[0], p, g[1], p, g[2], p).permute(0, 2, 4, 6, 3, 5, 7, 1)
    return t.reshape(n, g[0] * g[1] * g[2], p ** 3 * c)


def unpatchify(patchified_pixel_values, data, patch_size, twoD):
    """Inverse of `patchify`; `data` only provides the target shape."""
    p = patch_size
    n = patchified_pixel_values.shape[0]
    c = data.shape[1]
    g = [s // p for s in data.shape[2:]]
    if twoD:
        t = patchified_pixel_values.reshape(n, g[0], g[1], p, p, c).permute(0, 5, 1, 3, 2, 4)
        return t.reshape(n, c, g[0] * p, g[1] * p)
    t = patchified_pixel_values.reshape(n, g[0], g[1], g[2], p, p, p, c).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return t.reshape(n, c, g[0] * p, g[1] * p, g[2] * p)


def configure_optimizer(model, lr, beta_1, beta_2, weight_decay, fused=None):
    """AdamW with two groups: weight decay everywhere except var/pos/time embeddings.

    `fused`: None = stock torch.optim.AdamW defaults (what the reference builds), True / False = torch's own
    `fused` flag, "ucf" = this package's `FusedAdamW` (same state layout, multi-tensor CUDA update), "ucf_capturable" = the same with
    device-resident step / learning rate so the training step can be captured in a CUDA graph (utils/graph.py)."""
    decay, no_decay = [], []
    for name, prm in model.named_parameters():
        (no_decay if ("var_embed" in name or "pos_embed" in name or "time_pos_embed" in name) else decay).append(prm)
    groups = [
        {"params": decay, "lr": lr, "betas": (beta_1, beta_2), "weight_decay": weight_decay},
        {"params": no_decay, "lr": lr, "betas": (beta_1, beta_2), "weight_decay": 0},
    ]
    if fused in ("ucf", "ucf_capturable"):
        from .optim import FusedAdamW
        return FusedAdamW(groups, capturable=(fused == "ucf_capturable"))
    kw = {} if fused is None else {"fused": fused}
    return torch.optim.AdamW(groups, **kw)


def configure_scheduler(optimizer, warmup_steps, max_steps, warmup_start_lr, eta_min):
    return LinearWarmupCosineAnnealingLR(optimizer, warmup_steps, max_steps, warmup_start_lr, eta_min)


def interpolate_pos_embed_adaptive(model, checkpoint_model, new_size=127):
    """Resample the learned position tables of a checkpoint to `new_size` tokens, in place in the state dict
    (reference utils/misc.py:98-127: 1-D linear interpolation along the token axis, align_corners=False).
    `model` is unused, as in the reference."""
    for key in ("pos_embed", "decoder_pos_embed"):
        table = checkpoint_model.get(key)
        if table is None or table.shape[-2] == new_size:
            continue
        n_tok, width = table.shape[-2], table.shape[-1]
        as_channels = table.reshape(-1, n_tok, width).transpose(1, 2)           # [1, D, L]: tokens last for interpolate
        checkpoint_model[key] = torch.nn.functional.interpolate(as_channels, size=new_size, mode="linear",
                                                                align_corners=False).transpose(1, 2)


def is_power_of_two(n):
    """utils/misc.py:553-554."""
    return n != 0 and (n & (n - 1)) == 0


# ---------------------------------------------------------------------------------------------
# process groups (reference: utils/misc.py:129-238).  Rank layout: tensor-parallel fastest, then
# sequence-parallel, then data-parallel; the data-parallel ranks are further cut into contiguous
# FSDP groups of `fsdp_size` and strided "simple DDP" groups across them.
# ---------------------------------------------------------------------------------------------
def par_group_layout(data_par_size, tensor_par_size, seq_par_size, fsdp_size, simple_ddp_size):
    """Pure function: the rank lists of every group, in the order the groups must be created
    (identical on all ranks).  Returns a list of (kind, ranks)."""
    tp, sp, dp = tensor_par_size, seq_par_size, data_par_size
    out = []
    for i in range(dp * sp):
        out.append(("tensor", list(range(i * tp, (i + 1) * tp))))
    for t in range(dp):
        for i in range(tp):
            out.append(("seq", [t * tp * sp + i + j * tp for j in range(sp)]))
    for i in range(tp * sp):
        ranks = [i + j * tp * sp for j in range(dp)]
        for k in range(simple_ddp_size):
            out.append(("fsdp", ranks[k * fsdp_size:(k + 1) * fsdp_size]))
        for k in range(fsdp_size):
            out.append(("simple_ddp", ranks[k:len(ranks):fsdp_size]))
        out.append(("ddp", ranks))
    for i in range(tp):
        out.append(("data_seq_ort", [i + tp * j for j in range(dp * sp)]))
    return out


def init_par_groups(world_rank, data_par_size, tensor_par_size, seq_par_size, fsdp_size, simple_ddp_size):
    """-> (seq_par_group, ddp_group, tensor_par_group, data_seq_ort_group, fsdp_group, simple_ddp_group),
    the tuple the reference's FSDP drivers unpack (train_masked_fsdp.py:272)."""
    import torch.distributed as dist
    mine = {}
    for kind, ranks in par_group_layout(data_par_size, tensor_par_size, seq_par_size, fsdp_size, simple_ddp_size):
        group = dist.new_group(ranks)
        if world_rank in ranks:
            mine[kind] = group
    return (mine.get("seq"), mine.get("ddp"), mine.get("tensor"), mine.get("data_seq_ort"), mine.get("fsdp"),
            mine.get("simple_ddp"))
