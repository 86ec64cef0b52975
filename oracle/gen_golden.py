"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Golden-vector generator (runs in the BUILD container only).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Imports the UNMODIFIED reference from /root/reference/src behind oracle/shims (timm / monai /
xformers / matplotlib / nibabel / torchdata stand-ins), applies the two
constructor/forward fixes of SURVEY.md Appendix A as monkeypatches (A1: `sqrt_len_meth` kwarg
typo, A3: DiffusionVIT `_pos_embed(x)`), runs each model class on deterministic inputs in fp32 on
the CPU, asserts that the restatement in oracle/vit_ref.py agrees with the reference to fp32
round-off on outputs, loss and EVERY parameter gradient, and stores compact fixtures.
/root/reference does not exist on the GPU box; only the committed fixtures travel.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from oracle import vit_ref as R  # noqa: E402

import UCF_VIT.simple.arch as ref_arch  # noqa: E402
import UCF_VIT.simple.building_blocks as ref_bb  # noqa: E402
from UCF_VIT.utils.fused_attn import FusedAttn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


# ---- Appendix A monkeypatches (never edit /root/reference) ------------------------------------
def _embed_layer_A1(**kw):
    if "sqrt_len_meth" in kw:                       # A1: simple/arch.py:217 typo
        kw["sqrt_len_method"] = kw.pop("sqrt_len_meth")
    return ref_bb.PatchEmbed(**kw)


_orig_pos = ref_arch.VIT._pos_embed
ref_arch.DiffusionVIT._pos_embed = lambda self, x, seq_ps=None: _orig_pos(self, x, seq_ps)   # A3


def _close(name, a, b, tol=2e-5):
    a, b = a.detach().double(), b.detach().double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err <= tol * max(1.0, ref), f"{name}: restatement deviates from the reference: {err:.3e} (ref max {ref:.3e})"
    return err


def _finish(name, cfg, model, sd, run_ref, run_oracle, extra_arrays):
    """run_ref(model) / run_oracle(sd_with_grad) -> (outputs dict, loss)."""
    model.load_state_dict(sd, strict=True)
    model.train(cfg.get("train", True))
    outs_ref, loss_ref = run_ref(model)
    model.zero_grad()
    loss_ref.backward()
    g_ref = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}

    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs_o, loss_o = run_oracle(sdg)
    loss_o.backward()
    worst = 0.0
    for k in outs_ref:
        worst = max(worst, _close(f"{name}.{k}", outs_o[k], outs_ref[k]))
    _close(f"{name}.loss", loss_o, loss_ref)
    named = dict(model.named_parameters())
    for k, g in g_ref.items():
        go = sdg[k].grad
        assert go is not None, f"{name}: oracle produced no grad for {k}"
        worst = max(worst, _close(f"{name}.grad.{k}", go, g, 5e-5))
    arrays = dict(extra_arrays)
    for k, v in outs_ref.items():
        arrays["out." + k] = v.detach().numpy()
    arrays["loss"] = np.float64(loss_ref.item())
    arrays["sd_checksum"] = np.float64(fx.sd_checksum(sd))
    gkeys = sorted(g_ref)
    arrays["grad_keys"] = np.array(gkeys)
    arrays["grad_norms"] = np.array([g_ref[k].double().norm().item() for k in gkeys])
    for k in gkeys:
        if g_ref[k].numel() <= 4096:
            arrays["grad." + k] = g_ref[k].numpy()
    # parameters that exist but receive no gradient in this configuration (e.g. unused pos_embed)
    arrays["nograd_keys"] = np.array(sorted(set(named) - set(g_ref)) or [""])
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    fx.save_case(os.path.join(OUT, name + ".npz"), cfg, shapes, arrays)
    size = os.path.getsize(os.path.join(OUT, name + ".npz"))
    print(f"[golden] {name}: oracle==reference (max abs dev {worst:.2e}); {len(sd)} tensors; fixture {size/1024:.0f} KiB")


def _shapes(model):
    return {k: tuple(v.shape) for k, v in model.state_dict().items()}


def case_vit(name, D, H, depth, class_token, classes, img=32, p=8, B=2):
    cfg = dict(kind="vit", img_size=[img, img], patch_size=p, in_chans=3, num_classes=classes, embed_dim=D, depth=depth,
               num_heads=H, class_token=class_token, batch=B, x_seed=11, y_seed=12)
    m = ref_arch.VIT(img_size=[img, img], patch_size=p, in_chans=3, num_classes=classes, embed_dim=D, depth=depth,
                     num_heads=H, mlp_ratio=4, class_token=class_token, twoD=True, default_vars=["r", "g", "b"],
                     FusedAttn_option=FusedAttn.NONE)
    sd = fx.det_state_dict(_shapes(m), 1)
    x = fx.det_tensor((B, 3, img, img), cfg["x_seed"])
    if class_token:
        y = torch.randint(0, classes, (B,), generator=torch.Generator().manual_seed(cfg["y_seed"]))
        lossf = lambda o: torch.nn.functional.cross_entropy(o, y)   # noqa: E731
    else:
        lossf = lambda o: (o ** 2).mean()   # noqa: E731
        y = torch.zeros(1, dtype=torch.long)

    def run_ref(model):
        feats = model.forward_features(x, ["r", "g", "b"], None)
        o = model.forward_head(feats)
        return {"features": feats, "logits": o}, lossf(o)

    def run_oracle(s):
        feats = R.vit_features(x, s, cfg)
        pooled = feats[:, 0] if class_token else feats
        o = R.linear(pooled, s["head.weight"], s["head.bias"])
        return {"features": feats, "logits": o}, lossf(o)

    _finish(name, cfg, m, sd, run_ref, run_oracle, {"labels": y.numpy()})


def case_vit_seq(name):
    """adaptive_patching without sqrt_len_method: pre-gathered (B,C,L,p^2) -> LN-Linear-LN."""
    D, H, depth, p, L, B, C = 64, 2, 2, 4, 10, 2, 3
    cfg = dict(kind="vit_seq", patch_size=p, in_chans=C, num_classes=3, embed_dim=D, depth=depth, num_heads=H,
               fixed_length=L, class_token=True, seq_tokens=True, batch=B, x_seed=21, y_seed=22)
    m = ref_arch.VIT(img_size=[16, 16], patch_size=p, in_chans=C, num_classes=3, embed_dim=D, depth=depth, num_heads=H,
                     class_token=True, twoD=True, adaptive_patching=True, fixed_length=L, default_vars=["r", "g", "b"])
    sd = fx.det_state_dict(_shapes(m), 2)
    x = fx.det_tensor((B, C, L, p * p), cfg["x_seed"])
    y = torch.randint(0, 3, (B,), generator=torch.Generator().manual_seed(cfg["y_seed"]))

    def run_ref(model):
        o = model(x, ["r", "g", "b"])
        return {"logits": o}, torch.nn.functional.cross_entropy(o, y)

    def run_oracle(s):
        o = R.vit_forward(x, s, cfg)
        return {"logits": o}, torch.nn.functional.cross_entropy(o, y)

    _finish(name, cfg, m, sd, run_ref, run_oracle, {"labels": y.numpy()})


def case_mae(name):
    D, H, depth, dD, dH, dd, img, p, B = 128, 2, 2, 64, 2, 1, 32, 8, 2
    cfg = dict(kind="mae", img_size=[img, img], patch_size=p, in_chans=3, embed_dim=D, depth=depth, num_heads=H,
               decoder_embed_dim=dD, decoder_depth=dd, decoder_num_heads=dH, mask_ratio=0.75, class_token=False,
               batch=B, x_seed=31, noise_seed=32)
    m = ref_arch.MAE(img_size=[img, img], patch_size=p, in_chans=3, embed_dim=D, depth=depth, num_heads=H,
                     decoder_embed_dim=dD, decoder_depth=dd, decoder_num_heads=dH, mlp_ratio=4, mlp_ratio_decoder=4,
                     mask_ratio=0.75, linear_decoder=False, class_token=False, weight_init="skip", twoD=True,
                     default_vars=["r", "g", "b"], adaptive_patching=False)
    sd = fx.det_state_dict(_shapes(m), 3)
    x = fx.det_tensor((B, 3, img, img), cfg["x_seed"])
    noise = torch.rand(B, (img // p) ** 2, generator=torch.Generator().manual_seed(cfg["noise_seed"]))
    target = R.patchify_target(x, p, True)

    def run_ref(model):
        orig = model.random_masking
        model.random_masking = lambda seq, n=None: orig(seq, noise)
        pred, mask = model(x, ["r", "g", "b"])
        model.random_masking = orig
        return {"pred": pred, "mask": mask}, R.masked_mse(pred, target, mask)

    def run_oracle(s):
        pred, mask = R.mae_forward(x, s, cfg, noise)
        return {"pred": pred, "mask": mask}, R.masked_mse(pred, target, mask)

    _finish(name, cfg, m, sd, run_ref, run_oracle, {"noise": noise.numpy()})


def case_diffusion(name):
    D, H, depth, dD, dH, dd, img, p, B, T = 64, 2, 2, 64, 2, 1, 32, 8, 2, 50
    cfg = dict(kind="diffusion", img_size=[img, img], patch_size=p, in_chans=3, embed_dim=D, depth=depth, num_heads=H,
               decoder_embed_dim=dD, decoder_depth=dd, decoder_num_heads=dH, class_token=False, time_steps=T,
               batch=B, x_seed=41, t_seed=42, train=False)
    m = ref_arch.DiffusionVIT(img_size=[img, img], patch_size=p, in_chans=3, embed_dim=D, depth=depth, num_heads=H,
                              decoder_embed_dim=dD, decoder_depth=dd, decoder_num_heads=dH, mlp_ratio=4,
                              mlp_ratio_decoder=4, linear_decoder=False, class_token=False, weight_init="skip",
                              twoD=True, default_vars=["r", "g", "b"], time_steps=T)
    sd = fx.det_state_dict(_shapes(m), 4)
    x = fx.det_tensor((B, 3, img, img), cfg["x_seed"])
    t = torch.randint(0, T, (B,), generator=torch.Generator().manual_seed(cfg["t_seed"]))
    target = R.patchify_target(fx.det_tensor((B, 3, img, img), 43), p, True)
    table = m.temporalEmbeddings.embeddings

    def run_ref(model):
        o = model(x, t, ["r", "g", "b"])
        return {"pred": o}, torch.nn.functional.mse_loss(o, target)

    def run_oracle(s):
        o = R.diffusion_forward(x, t, s, cfg, table)
        return {"pred": o}, torch.nn.functional.mse_loss(o, target)

    _finish(name, cfg, m, sd, run_ref, run_oracle, {"t": t.numpy(), "time_table": table.numpy()})


def case_sap(name):
    D, H, depth, p, s, B = 64, 2, 2, 8, 4, 2
    L = s * s
    cfg = dict(kind="sap", patch_size=p, in_chans=3, num_classes=4, embed_dim=D, depth=depth, num_heads=H,
               fixed_length=L, sqrt_len=s, class_token=False, use_adaptive_pos_emb=True, batch=B, x_seed=51, ps_seed=52)
    m = ref_arch.SAP(img_size=[p * s, p * s], patch_size=p, in_chans=3, num_classes=4, embed_dim=D, depth=depth,
                     num_heads=H, twoD=True, default_vars=["r", "g", "b"], adaptive_patching=True, fixed_length=L,
                     sqrt_len=s, sqrt_len_method=True, use_adaptive_pos_emb=True, class_token=False)
    sd = fx.det_state_dict(_shapes(m), 5)
    x = fx.det_tensor((B, 3, p * s, p * s), cfg["x_seed"])
    seq_ps = fx.det_tensor((B, L, 3), cfg["ps_seed"]).abs() * 4
    tgt = fx.det_tensor((B, 4, p * s, p * s), 53)

    def run_ref(model):
        o = model(x, ["r", "g", "b"], seq_ps)
        return {"mask_logits": o}, ((o - tgt) ** 2).mean()

    def run_oracle(sdd):
        o = R.sap_forward(x, sdd, cfg, seq_ps)
        return {"mask_logits": o}, ((o - tgt) ** 2).mean()

    _finish(name, cfg, m, sd, run_ref, run_oracle, {})


def case_unetr(name):
    """3-D, patch 16, two variables -> shared patch embed + variable aggregation + conv decoder."""
    D, H, depth, p, img, V, B, fs, ncls = 96, 3, 4, 16, 64, 2, 1, 4, 3
    vars_ = ["v0", "v1"]
    cfg = dict(kind="unetr", img_size=[img] * 3, patch_size=p, in_chans=V, num_classes=ncls, embed_dim=D, depth=depth,
               num_heads=H, use_varemb=True, feature_size=fs, class_token=False, batch=B, x_seed=61, y_seed=62)
    m = ref_arch.UNETR(img_size=[img] * 3, patch_size=p, in_chans=V, num_classes=ncls, embed_dim=D, depth=depth,
                       num_heads=H, twoD=False, use_varemb=True, default_vars=vars_, feature_size=fs,
                       skip_connection=True, linear_decoder=False, class_token=False, weight_init="skip",
                       embed_layer=_embed_layer_A1)
    sd = fx.det_state_dict(_shapes(m), 6)
    x = fx.det_tensor((B, V, img, img, img), cfg["x_seed"]).abs()
    tgt = fx.det_tensor((B, ncls, img, img, img), cfg["y_seed"])

    def run_ref(model):
        o = model(x, vars_)
        return {"seg_logits_slice": o[:, :, ::8, ::8, ::8].contiguous(), "seg_mean": o.mean().reshape(1)}, ((o - tgt) ** 2).mean()

    def run_oracle(s):
        o = R.unetr_forward(x, s, cfg, var_ids=[0, 1])
        return {"seg_logits_slice": o[:, :, ::8, ::8, ::8].contiguous(), "seg_mean": o.mean().reshape(1)}, ((o - tgt) ** 2).mean()

    _finish(name, cfg, m, sd, run_ref, run_oracle, {})


def host_logic():
    """Init tables, LR schedule and target (un)patchify: pure host logic the product restates."""
    from UCF_VIT.utils import pos_embed as rp
    from UCF_VIT.utils.lr_scheduler import LinearWarmupCosineAnnealingLR
    from UCF_VIT.utils import misc as rm
    arrays = {
        "pe2d_16_4x6": rp.get_2d_sincos_pos_embed(16, 4, 6, cls_token=False),
        "pe2d_16_3x3_cls": rp.get_2d_sincos_pos_embed(16, 3, 3, cls_token=True),
        "pe3d_12_2x3x2": rp.get_3d_sincos_pos_embed(12, 2, 3, 2),
        "pe1d_8": rp.get_1d_sincos_pos_embed_from_grid(8, np.arange(5)),
        "time_table_10x8": rp.SinusoidalEmbeddings(10, 8).embeddings.numpy(),
    }
    prm = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([prm], lr=1e-4)
    sch = LinearWarmupCosineAnnealingLR(opt, 5, 20, 1e-8, 1e-8)
    lrs = []
    for _ in range(30):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    arrays["lr_w5_m20"] = np.array(lrs)
    x2 = fx.det_tensor((2, 3, 8, 12), 71)
    x3 = fx.det_tensor((1, 2, 4, 8, 4), 72)
    arrays["patchify2d_p4"] = rm.patchify(x2, 4, True).numpy()
    arrays["patchify3d_p4"] = rm.patchify(x3, 4, False).numpy()
    assert torch.equal(rm.unpatchify(rm.patchify(x2, 4, True), x2, 4, True), x2)
    assert torch.equal(rm.unpatchify(rm.patchify(x3, 4, False), x3, 4, False), x3)
    ck = {"pos_embed": fx.det_tensor((1, 10, 8), 73), "decoder_pos_embed": fx.det_tensor((1, 7, 4), 74),
          "other": torch.zeros(2)}
    rm.interpolate_pos_embed_adaptive(None, ck, new_size=7)        # pos_embed 10 -> 7, decoder table already 7
    arrays["interp_pos_10to7"] = ck["pos_embed"].numpy()
    arrays["interp_dec_unchanged"] = ck["decoder_pos_embed"].numpy()
    ck2 = {"decoder_pos_embed": fx.det_tensor((1, 5, 4), 75)}
    rm.interpolate_pos_embed_adaptive(None, ck2, new_size=12)
    arrays["interp_dec_5to12"] = ck2["decoder_pos_embed"].numpy()
    arrays["pow2"] = np.array([int(bool(rm.is_power_of_two(n))) for n in range(0, 70)])
    fx.save_case(os.path.join(OUT, "host_logic.npz"), {"kind": "host"}, {}, arrays)
    print("[golden] host_logic: pos-embed tables, LR schedule, patchify targets, pos-embed interpolation")


if __name__ == "__main__":
    case_vit("vit_cls_hd64", D=128, H=2, depth=2, class_token=True, classes=5)
    case_vit("vit_tokens_hd32", D=64, H=2, depth=2, class_token=False, classes=3)
    case_vit_seq("vit_adaptive_seq")
    case_mae("mae_hd64_dec32")
    case_diffusion("diffusion_eval")
    case_sap("sap_2d")
    case_unetr("unetr_3d_var2")
    host_logic()
