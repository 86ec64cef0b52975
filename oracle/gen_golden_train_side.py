"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Golden vectors for the loss / optimizer side of a training step
(runs in the BUILD container only; /root/reference does not exist on the GPU box).

    python oracle/gen_golden_train_side.py        # writes tests/golden/host_train_side.npz

Runs the UNMODIFIED reference helpers -- `configure_optimizer`, `configure_scheduler`, `patchify`
(/root/reference/src/UCF_VIT/utils/misc.py) and `masked_mse` (utils/metrics.py) -- on deterministic
inputs, checks the restatement in oracle/train_side_ref.py against them and stores the results.
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
from einops import rearrange  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from oracle import train_side_ref as T  # noqa: E402

from UCF_VIT.utils import misc as ref_misc  # noqa: E402
from UCF_VIT.utils.metrics import DiceBLoss as RefDiceBLoss  # noqa: E402
from UCF_VIT.utils.metrics import masked_mse as ref_masked_mse  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "host_train_side.npz")

# name -> shape of the toy parameter set: both no-decay keys of misc.py:62 and odd sizes (vector tail)
PARAM_SHAPES = {"blocks.0.attn.qkv.weight": (24, 8), "blocks.0.attn.qkv.bias": (24,), "pos_embed": (1, 5, 8),
                "var_embed": (1, 3, 8), "head.weight": (3, 7), "norm.bias": (13,)}
HYPER = dict(lr=3e-3, beta_1=0.9, beta_2=0.95, weight_decay=0.05)
SCHED = dict(warmup_steps=2, max_steps=6, warmup_start_lr=1e-5, eta_min=1e-4)
STEPS = 6


def grad_of(name, step):
    return fx.det_tensor(PARAM_SHAPES[name], 900 + 17 * step + sorted(PARAM_SHAPES).index(name), scale=0.5)


def optimizer_case(arrays):
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for i, (n, shp) in enumerate(PARAM_SHAPES.items()):
                self.register_parameter(n.replace(".", "_"), torch.nn.Parameter(fx.det_tensor(shp, 800 + i)))

    toy = Toy()
    names = list(PARAM_SHAPES)
    opt = ref_misc.configure_optimizer(toy, HYPER["lr"], HYPER["beta_1"], HYPER["beta_2"], HYPER["weight_decay"])
    sch = ref_misc.configure_scheduler(opt, SCHED["warmup_steps"], SCHED["max_steps"], SCHED["warmup_start_lr"],
                                       SCHED["eta_min"])
    # the module registers names with '_' for '.', which keeps the 'pos_embed' / 'var_embed' substrings
    decay, no_decay = T.decay_groups([n.replace(".", "_") for n in names])
    assert [id(p) for p in opt.param_groups[1]["params"]] == [id(getattr(toy, n)) for n in no_decay]
    assert opt.param_groups[1]["weight_decay"] == 0 and opt.param_groups[0]["weight_decay"] == HYPER["weight_decay"]
    state = {n: (getattr(toy, n.replace(".", "_")).detach().numpy().copy(), np.zeros(PARAM_SHAPES[n], np.float32),
                 np.zeros(PARAM_SHAPES[n], np.float32)) for n in names}
    lrs = []
    for step in range(1, STEPS + 1):
        lr = opt.param_groups[0]["lr"]
        lrs.append(lr)
        for n in names:
            getattr(toy, n.replace(".", "_")).grad = grad_of(n, step)
        opt.step()
        opt.zero_grad()
        sch.step()
        for n in names:
            wd = 0.0 if n.replace(".", "_") in no_decay else HYPER["weight_decay"]
            p, m, v = state[n]
            state[n] = T.adamw_step(p, grad_of(n, step).numpy(), m, v, step, lr, HYPER["beta_1"], HYPER["beta_2"], 1e-8, wd)
            ref = getattr(toy, n.replace(".", "_")).detach().numpy()
            err = np.abs(state[n][0] - ref).max()
            assert err <= 2e-7 * max(1.0, np.abs(ref).max()), (n, step, err)
        if step in (1, STEPS):
            for n in names:
                arrays[f"adamw_step{step}:{n}"] = getattr(toy, n.replace(".", "_")).detach().numpy().copy()
    arrays["adamw_lrs"] = np.array(lrs)
    for n in names:
        st = opt.state[getattr(toy, n.replace(".", "_"))]
        arrays[f"adamw_m:{n}"] = st["exp_avg"].numpy().copy()
        arrays[f"adamw_v:{n}"] = st["exp_avg_sq"].numpy().copy()
    print("[golden] optimizer: restatement within 2e-7 of torch.optim.AdamW as configured by the reference,",
          STEPS, "steps")


def loss_case(arrays, tag, data_shape, p, twoD, seed):
    data = fx.det_tensor(data_shape, seed)
    target = ref_misc.patchify(data, p, twoD)
    assert np.array_equal(T.patchify_np(data.numpy(), p, twoD), target.numpy())
    pred = fx.det_tensor(tuple(target.shape), seed + 1).requires_grad_(True)
    mask = (fx.det_tensor(tuple(target.shape[:2]), seed + 2) > 0).float()
    assert 0 < mask.sum() < mask.numel()
    lm = ref_masked_mse(pred, target, mask)
    gm, = torch.autograd.grad(lm, pred)
    lf = torch.nn.MSELoss()(pred, target)
    gf, = torch.autograd.grad(lf, pred)
    for (l, g, mk) in ((lm, gm, mask.numpy()), (lf, gf, None)):
        lo, go = T.mse_loss_and_grad(pred.detach().numpy(), target.numpy(), mk)
        assert abs(lo - l.item()) <= 1e-6 * abs(l.item()) and np.abs(go - g.numpy()).max() <= 1e-6 * np.abs(g.numpy()).max()
    arrays[f"{tag}:mask"] = mask.numpy()
    arrays[f"{tag}:loss_masked"] = np.float64(lm.item())
    arrays[f"{tag}:loss_full"] = np.float64(lf.item())
    arrays[f"{tag}:grad_masked"] = gm.numpy()
    arrays[f"{tag}:grad_full"] = gf.numpy()


def adaptive_case(arrays, seed):
    seq = fx.det_tensor((2, 3, 6, 16), seed)                  # b c s p
    target = rearrange(seq, "b c s p -> b s (p c)")           # train_masked_fsdp.py:42
    assert np.array_equal(T.adaptive_target_np(seq.numpy()), target.numpy())
    pred = fx.det_tensor(tuple(target.shape), seed + 1).requires_grad_(True)
    l = torch.nn.MSELoss()(pred, target)
    g, = torch.autograd.grad(l, pred)
    arrays["adaptive:loss_full"] = np.float64(l.item())
    arrays["adaptive:grad_full"] = g.numpy()


def dice_case(arrays, tag, shape, seed, weight, smooth, act):
    x = fx.det_tensor(shape, seed, scale=3.0)
    if not act:
        x = torch.sigmoid(x)
    x.requires_grad_(True)
    t = (fx.det_tensor(shape, seed + 1) > 0.3).float()
    loss = RefDiceBLoss(weight=weight, num_class=shape[1])(x, t, smooth=smooth, act=act)
    g, = torch.autograd.grad(loss, x)
    lo, go = T.dice_bce_loss_and_grad(x.detach().numpy(), t.numpy(), weight, smooth, act)
    assert abs(lo - loss.item()) <= 2e-6 * abs(lo), (lo, loss.item())
    assert np.abs(go - g.numpy()).max() <= 2e-6 * np.abs(g.numpy()).max()
    arrays[f"{tag}:loss"] = np.float64(loss.item())
    arrays[f"{tag}:grad"] = g.numpy()


DICE_CASES = {"dice_2c": [[2, 2, 8, 12], 940, 0.5, 1.0, True], "dice_3c_w03": [[1, 3, 6, 10], 950, 0.3, 2.0, True],
              "dice_probs": [[2, 2, 4, 4], 960, 0.5, 1.0, False]}

if __name__ == "__main__":
    arrays = {}
    for tag, (shape, seed, w, sm, act) in DICE_CASES.items():
        dice_case(arrays, tag, tuple(shape), seed, w, sm, act)
    optimizer_case(arrays)
    loss_case(arrays, "mse2d", (2, 3, 8, 12), 4, True, 910)
    loss_case(arrays, "mse3d", (1, 2, 4, 8, 4), 4, False, 920)
    adaptive_case(arrays, 930)
    cfg = {"kind": "host", "param_shapes": {k: list(v) for k, v in PARAM_SHAPES.items()}, "hyper": HYPER,
           "sched": SCHED, "steps": STEPS, "eps": 1e-8,
           "loss_cases": {"mse2d": [[2, 3, 8, 12], 4, True, 910], "mse3d": [[1, 2, 4, 8, 4], 4, False, 920]},
           "adaptive_seed": 930, "dice_cases": DICE_CASES}
    fx.save_case(OUT, cfg, {}, arrays)
    print("[golden] host_train_side: AdamW trajectory, masked / full MSE on patchified targets (2-D, 3-D, adaptive), DiceBLoss")
