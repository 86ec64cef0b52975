"""Shared by the CPU and GPU tests: rebuild the golden cases (tests/golden/*.npz, written by
oracle/gen_golden.py) for the oracle and for the product package."""
import glob
import os

import numpy as np
import torch

from oracle import fixtures as fx
from oracle import vit_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                     if not os.path.basename(p).startswith(("host_", "sap_tree")))
VARS3 = ["r", "g", "b"]


def load(name):
    cfg, shapes, arrays = fx.load_case(os.path.join(GOLDEN, name + ".npz"))
    base = {"vit_cls_hd64": 1, "vit_tokens_hd32": 1, "vit_adaptive_seq": 2, "mae_hd64_dec32": 3, "diffusion_eval": 4,
            "sap_2d": 5, "unetr_3d_var2": 6}[name]
    sd = fx.det_state_dict(shapes, base)
    return cfg, shapes, arrays, sd


def inputs(cfg, arrays):
    """Deterministic inputs of a case, exactly as gen_golden.py built them."""
    k, B = cfg["kind"], cfg["batch"]
    d = {}
    if k == "vit":
        img = cfg["img_size"][0]
        d["x"] = fx.det_tensor((B, 3, img, img), cfg["x_seed"])
        d["y"] = torch.from_numpy(arrays["labels"])
    elif k == "vit_seq":
        d["x"] = fx.det_tensor((B, cfg["in_chans"], cfg["fixed_length"], cfg["patch_size"] ** 2), cfg["x_seed"])
        d["y"] = torch.from_numpy(arrays["labels"])
    elif k == "mae":
        img, p = cfg["img_size"][0], cfg["patch_size"]
        d["x"] = fx.det_tensor((B, 3, img, img), cfg["x_seed"])
        d["noise"] = torch.from_numpy(arrays["noise"])
        d["target"] = R.patchify_target(d["x"], p, True)
    elif k == "diffusion":
        img, p = cfg["img_size"][0], cfg["patch_size"]
        d["x"] = fx.det_tensor((B, 3, img, img), cfg["x_seed"])
        d["t"] = torch.from_numpy(arrays["t"])
        d["table"] = torch.from_numpy(arrays["time_table"])
        d["target"] = R.patchify_target(fx.det_tensor((B, 3, img, img), 43), p, True)
    elif k == "sap":
        side = cfg["patch_size"] * cfg["sqrt_len"]
        d["x"] = fx.det_tensor((B, 3, side, side), cfg["x_seed"])
        d["seq_ps"] = fx.det_tensor((B, cfg["fixed_length"], 3), cfg["ps_seed"]).abs() * 4
        d["target"] = fx.det_tensor((B, 4, side, side), 53)
    elif k == "unetr":
        img = cfg["img_size"][0]
        d["x"] = fx.det_tensor((B, cfg["in_chans"], img, img, img), cfg["x_seed"]).abs()
        d["target"] = fx.det_tensor((B, cfg["num_classes"], img, img, img), cfg["y_seed"])
    return d


def run_oracle(cfg, sd, inp):
    """-> (outputs dict, loss) in fp32 on the CPU."""
    k = cfg["kind"]
    if k == "vit":
        feats = R.vit_features(inp["x"], sd, cfg)
        o = R.linear(feats[:, 0] if cfg["class_token"] else feats, sd["head.weight"], sd["head.bias"])
        loss = torch.nn.functional.cross_entropy(o, inp["y"]) if cfg["class_token"] else (o ** 2).mean()
        return {"features": feats, "logits": o}, loss
    if k == "vit_seq":
        o = R.vit_forward(inp["x"], sd, cfg)
        return {"logits": o}, torch.nn.functional.cross_entropy(o, inp["y"])
    if k == "mae":
        pred, mask = R.mae_forward(inp["x"], sd, cfg, inp["noise"])
        return {"pred": pred, "mask": mask}, R.masked_mse(pred, inp["target"], mask)
    if k == "diffusion":
        o = R.diffusion_forward(inp["x"], inp["t"], sd, cfg, inp["table"])
        return {"pred": o}, torch.nn.functional.mse_loss(o, inp["target"])
    if k == "sap":
        o = R.sap_forward(inp["x"], sd, cfg, inp["seq_ps"])
        return {"mask_logits": o}, ((o - inp["target"]) ** 2).mean()
    if k == "unetr":
        o = R.unetr_forward(inp["x"], sd, cfg, var_ids=[0, 1])
        return ({"seg_logits_slice": o[:, :, ::8, ::8, ::8].contiguous(), "seg_mean": o.mean().reshape(1)},
                ((o - inp["target"]) ** 2).mean())
    raise KeyError(k)


def build_product(cfg):
    """The product-package model for a golden case (constructor args mirror gen_golden.py)."""
    from ucf_vit_b200.simple import arch as A
    from ucf_vit_b200.utils.fused_attn import FusedAttn
    k = cfg["kind"]
    c = cfg
    if k == "vit":
        return A.VIT(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=3, num_classes=c["num_classes"],
                     embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], mlp_ratio=4,
                     class_token=c["class_token"], twoD=True, default_vars=VARS3, FusedAttn_option=FusedAttn.FLASH)
    if k == "vit_seq":
        return A.VIT(img_size=[16, 16], patch_size=c["patch_size"], in_chans=c["in_chans"], num_classes=3,
                     embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], class_token=True, twoD=True,
                     adaptive_patching=True, fixed_length=c["fixed_length"], default_vars=VARS3)
    if k == "mae":
        return A.MAE(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=3, embed_dim=c["embed_dim"],
                     depth=c["depth"], num_heads=c["num_heads"], decoder_embed_dim=c["decoder_embed_dim"],
                     decoder_depth=c["decoder_depth"], decoder_num_heads=c["decoder_num_heads"], mlp_ratio=4,
                     mlp_ratio_decoder=4, mask_ratio=c["mask_ratio"], linear_decoder=False, class_token=False,
                     weight_init="skip", twoD=True, default_vars=VARS3, adaptive_patching=False)
    if k == "diffusion":
        return A.DiffusionVIT(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=3, embed_dim=c["embed_dim"],
                              depth=c["depth"], num_heads=c["num_heads"], decoder_embed_dim=c["decoder_embed_dim"],
                              decoder_depth=c["decoder_depth"], decoder_num_heads=c["decoder_num_heads"], mlp_ratio=4,
                              mlp_ratio_decoder=4, linear_decoder=False, class_token=False, weight_init="skip",
                              twoD=True, default_vars=VARS3, time_steps=c["time_steps"])
    if k == "sap":
        side = c["patch_size"] * c["sqrt_len"]
        return A.SAP(img_size=[side, side], patch_size=c["patch_size"], in_chans=3, num_classes=4,
                     embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], twoD=True,
                     default_vars=VARS3, adaptive_patching=True, fixed_length=c["fixed_length"], sqrt_len=c["sqrt_len"],
                     sqrt_len_method=True, use_adaptive_pos_emb=True, class_token=False)
    if k == "unetr":
        return A.UNETR(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=c["in_chans"],
                       num_classes=c["num_classes"], embed_dim=c["embed_dim"], depth=c["depth"],
                       num_heads=c["num_heads"], twoD=False, use_varemb=True, default_vars=["v0", "v1"],
                       feature_size=c["feature_size"], skip_connection=True, linear_decoder=False, class_token=False,
                       weight_init="skip")
    raise KeyError(k)


def run_product(cfg, model, inp):
    """Same outputs / loss through the product's public module API (inputs already on the device)."""
    k = cfg["kind"]
    if k == "vit":
        feats = model.forward_features(inp["x"], VARS3, None)
        o = model.forward_head(feats)
        loss = torch.nn.functional.cross_entropy(o.float(), inp["y"]) if cfg["class_token"] else (o.float() ** 2).mean()
        return {"features": feats, "logits": o}, loss
    if k == "vit_seq":
        o = model(inp["x"], VARS3)
        return {"logits": o}, torch.nn.functional.cross_entropy(o.float(), inp["y"])
    if k == "mae":
        feats, mask, ids = model.forward_features(inp["x"], VARS3, None, noise=inp["noise"])
        pred = model.forward_head(feats, ids, None)
        # the product's own loss: patchify folded into the masked MSE kernel (the target is never materialised)
        from ucf_vit_b200.utils.metrics import patch_mse
        return {"pred": pred, "mask": mask}, patch_mse(pred, inp["x"], cfg["patch_size"], True, mask)
    if k == "diffusion":
        o = model(inp["x"], inp["t"], VARS3)
        return {"pred": o}, torch.nn.functional.mse_loss(o.float(), inp["target"])
    if k == "sap":
        o = model(inp["x"], VARS3, inp["seq_ps"])
        return {"mask_logits": o}, ((o.float() - inp["target"]) ** 2).mean()
    if k == "unetr":
        o = model(inp["x"], ["v0", "v1"]).float()
        return ({"seg_logits_slice": o[:, :, ::8, ::8, ::8].contiguous(), "seg_mean": o.mean().reshape(1)},
                ((o - inp["target"]) ** 2).mean())
    raise KeyError(k)
