"""Loss and optimizer either side of the transformer path.

CPU (`-m "not gpu"`): the oracle restatement (oracle/train_side_ref.py) against the vectors the unmodified
reference produced (tests/golden/host_train_side.npz, oracle/gen_golden_train_side.py), and the host logic of
`FusedAdamW` / `patch_mse` (grouping, state layout, loud failure without CUDA).
GPU (`-m gpu`): the CUDA kernels through the C ABI against the oracle and the golden vectors.
Tolerances: AdamW fp32 state 2e-6 relative to max|ref| (fp32 round-off of a re-ordered update); loss 2e-6
relative (fp32 accumulation per thread, double across threads); fp32 gradients 2e-6, bf16 gradients 1e-2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import train_side_ref as T
from tests import _cases as C
from ucf_vit_b200 import _lib as L
from ucf_vit_b200.utils import metrics, misc
from ucf_vit_b200.utils.optim import FusedAdamW

CFG, _, G = fx.load_case(os.path.join(C.GOLDEN, "host_train_side.npz"))
SHAPES = {k: tuple(v) for k, v in CFG["param_shapes"].items()}
NAMES = list(SHAPES)
HYP, SCH = CFG["hyper"], CFG["sched"]
gpu = pytest.mark.gpu


def _grad(name, step):
    return fx.det_tensor(SHAPES[name], 900 + 17 * step + sorted(SHAPES).index(name), scale=0.5)


def _toy(device="cpu"):
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for i, (n, shp) in enumerate(SHAPES.items()):
                self.register_parameter(n.replace(".", "_"), torch.nn.Parameter(fx.det_tensor(shp, 800 + i)))
    return Toy().to(device)


def _close(got, ref, tol):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max()
    assert err <= tol * max(np.abs(ref).max(), 1e-30), (err, np.abs(ref).max())


# ---------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_adamw_follows_the_reference_trajectory():
    no_decay = set(T.decay_groups([n.replace(".", "_") for n in NAMES])[1])
    assert no_decay == {"pos_embed", "var_embed"}
    st = {n: (fx.det_tensor(SHAPES[n], 800 + i).numpy(), np.zeros(SHAPES[n], np.float32), np.zeros(SHAPES[n], np.float32))
          for i, n in enumerate(NAMES)}
    for step in range(1, CFG["steps"] + 1):
        for n in NAMES:
            wd = 0.0 if n.replace(".", "_") in no_decay else HYP["weight_decay"]
            st[n] = T.adamw_step(*([st[n][0], _grad(n, step).numpy()] + list(st[n][1:])), step, G["adamw_lrs"][step - 1],
                                 HYP["beta_1"], HYP["beta_2"], CFG["eps"], wd)
        if step in (1, CFG["steps"]):
            for n in NAMES:
                _close(st[n][0], G[f"adamw_step{step}:{n}"], 3e-7)
    for n in NAMES:
        _close(st[n][1], G[f"adamw_m:{n}"], 3e-7)
        _close(st[n][2], G[f"adamw_v:{n}"], 3e-7)


@pytest.mark.parametrize("tag", ["mse2d", "mse3d"])
def test_oracle_patch_losses_match_the_reference(tag):
    shape, p, twoD, seed = CFG["loss_cases"][tag]
    data = fx.det_tensor(tuple(shape), seed).numpy()
    tgt = T.patchify_np(data, p, twoD)
    assert np.array_equal(tgt, misc.patchify(torch.from_numpy(data), p, twoD).numpy())
    pred = fx.det_tensor(tgt.shape, seed + 1).numpy()
    for key, mk in (("masked", G[f"{tag}:mask"]), ("full", None)):
        lo, go = T.mse_loss_and_grad(pred, tgt, mk)
        assert abs(lo - float(G[f"{tag}:loss_{key}"])) <= 1e-6 * abs(lo)
        _close(go, G[f"{tag}:grad_{key}"], 1e-6)


def test_oracle_adaptive_target():
    seq = fx.det_tensor((2, 3, 6, 16), CFG["adaptive_seed"]).numpy()
    tgt = T.adaptive_target_np(seq)
    pred = fx.det_tensor(tgt.shape, CFG["adaptive_seed"] + 1).numpy()
    lo, go = T.mse_loss_and_grad(pred, tgt)
    assert abs(lo - float(G["adaptive:loss_full"])) <= 1e-6 * abs(lo)
    _close(go, G["adaptive:grad_full"], 1e-6)


# ------------------------------------------------------------------------------------ host logic (CPU)
def test_fused_adamw_is_a_drop_in_for_the_stock_optimizer_state():
    toy = _toy()
    opt = misc.configure_optimizer(toy, HYP["lr"], HYP["beta_1"], HYP["beta_2"], HYP["weight_decay"], fused="ucf")
    assert isinstance(opt, FusedAdamW) and isinstance(opt, torch.optim.AdamW)
    assert [g["weight_decay"] for g in opt.param_groups] == [HYP["weight_decay"], 0]
    assert [id(p) for p in opt.param_groups[1]["params"]] == [id(toy.pos_embed), id(toy.var_embed)]
    # a checkpoint written by the stock optimizer loads, and the scheduler drives the same lr field
    stock = misc.configure_optimizer(_toy(), HYP["lr"], HYP["beta_1"], HYP["beta_2"], HYP["weight_decay"])
    for grp in stock.param_groups:
        for p in grp["params"]:
            p.grad = torch.ones_like(p)
    stock.step()
    opt.load_state_dict(stock.state_dict())
    st = opt.state[toy.pos_embed]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 1.0
    sch = misc.configure_scheduler(opt, SCH["warmup_steps"], SCH["max_steps"], SCH["warmup_start_lr"], SCH["eta_min"])
    assert opt.param_groups[0]["lr"] == pytest.approx(G["adamw_lrs"][0], rel=1e-12)
    del sch
    with pytest.raises(NotImplementedError):
        FusedAdamW(toy.parameters(), amsgrad=True)


def test_train_side_has_no_cpu_fallback():
    toy = _toy()
    opt = FusedAdamW(toy.parameters(), lr=1e-3)
    for p in toy.parameters():
        p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    pred, img = torch.zeros(1, 4, 48), torch.zeros(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        metrics.patch_mse(pred, img, 4, True)
    with pytest.raises(ValueError):
        metrics.patch_mse(pred, torch.zeros(1, 3, 8, 9), 4, True)
    with pytest.raises(NotImplementedError):
        metrics.patch_mse(pred, img.requires_grad_(True), 4, True)


def test_train_side_entry_points_reject_bad_arguments_without_a_device():
    lib = L.lib()
    one = (torch.zeros(4).data_ptr(),)
    import ctypes
    tbl = (ctypes.c_void_p * 1)(*one)
    cnt = (ctypes.c_longlong * 1)(4)
    assert lib.ucf_adamw_multi(1, tbl, tbl, tbl, tbl, cnt, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 0, None) == -1
    assert b"step must be >= 1" in lib.ucf_last_error()
    assert lib.ucf_adamw_multi(1, tbl, tbl, tbl, tbl, cnt, 1e-3, 1.0, 0.999, 1e-8, 0.0, 1, 0, None) == -1
    assert lib.ucf_adamw_multi(0, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 0, None) == 0
    assert lib.ucf_patch_mse_fwd(1, 0, 1, 0, None, 0, 3, 2, 2, 1, 4, 4, 1, 1, 1, None) == -1
    assert b"must be positive" in lib.ucf_last_error()
    assert lib.ucf_patch_mse_fwd(1, 5, 1, 0, None, 1, 3, 2, 2, 1, 4, 4, 1, 1, 1, None) == -1
    assert b"f32 or bf16" in lib.ucf_last_error()
    assert lib.ucf_patch_mse_bwd(1, 0, 1, 0, None, None, None, 1, 3, 2, 2, 1, 4, 4, 1, 1, None) == -1
    assert b"null pointer" in lib.ucf_last_error()


# ------------------------------------------------------------------------------------------ CUDA (GPU)
@gpu
def test_adamw_kernel_follows_the_reference_trajectory():
    toy = _toy("cuda")
    opt = misc.configure_optimizer(toy, HYP["lr"], HYP["beta_1"], HYP["beta_2"], HYP["weight_decay"], fused="ucf")
    sch = misc.configure_scheduler(opt, SCH["warmup_steps"], SCH["max_steps"], SCH["warmup_start_lr"], SCH["eta_min"])
    n0 = L.launch_count()
    for step in range(1, CFG["steps"] + 1):
        assert opt.param_groups[0]["lr"] == pytest.approx(G["adamw_lrs"][step - 1], rel=1e-12)
        for n in NAMES:
            getattr(toy, n.replace(".", "_")).grad = _grad(n, step).cuda()
        opt.step()
        opt.zero_grad()
        sch.step()
        if step in (1, CFG["steps"]):
            for n in NAMES:
                _close(getattr(toy, n.replace(".", "_")).detach().cpu().numpy(), G[f"adamw_step{step}:{n}"], 2e-6)
    assert L.launch_count() - n0 == 2 * CFG["steps"]          # one launch per parameter group per step
    for n in NAMES:
        st = opt.state[getattr(toy, n.replace(".", "_"))]
        _close(st["exp_avg"].cpu().numpy(), G[f"adamw_m:{n}"], 2e-6)
        _close(st["exp_avg_sq"].cpu().numpy(), G[f"adamw_v:{n}"], 2e-6)
        assert float(st["step"]) == CFG["steps"]


@gpu
@pytest.mark.parametrize("maximize", [False, True])
def test_adamw_kernel_many_ragged_and_unaligned_tensors(maximize):
    """60 tensors (three launches of 24, largest first), sizes 1 .. 3 M elements; the small ones sit at
    4-byte-only alignment, so the first launch takes the float4 path (with 1..3 trailing elements) and the
    later ones the scalar path.  Checked against the oracle's restatement of torch's update, four steps."""
    torch.manual_seed(3)
    sizes = [1, 2, 3, 5, 7, 31, 257, 1000, 4099, 65537, 3_000_001] + [int(x) for x in torch.randint(1, 50000, (49,))]
    pool = torch.randn(sum(sizes) + 5 * len(sizes) + 8, device="cuda")
    ps, off = [], 0
    for i, n in enumerate(sizes):
        off = (off + 3) // 4 * 4 + (1 if (n < 1000 and i % 2) else 0)    # 16-byte aligned, or 4 bytes past it
        ps.append(torch.nn.Parameter(pool[off:off + n].detach()))
        off += n
    big = sorted(ps, key=lambda t: -t.numel())[:24]
    assert all(t.data_ptr() % 16 == 0 for t in big) and any(t.data_ptr() % 16 for t in ps)
    ref = [(p.detach().cpu().numpy().copy(), np.zeros(p.numel(), np.float32), np.zeros(p.numel(), np.float32)) for p in ps]
    opt = FusedAdamW(ps, lr=2e-3, betas=(0.8, 0.9), eps=1e-6, weight_decay=0.1, maximize=maximize)
    for step in range(1, 5):
        gs = [torch.randn(p.numel(), device="cuda") * (0.1 + i % 3) for i, p in enumerate(ps)]
        for p, g in zip(ps, gs):
            p.grad = g
        opt.step()
        ref = [T.adamw_step(r[0], (-g if maximize else g).cpu().numpy(), r[1], r[2], step, 2e-3, 0.8, 0.9, 1e-6, 0.1)
               for r, g in zip(ref, gs)]
    for p, r in zip(ps, ref):
        _close(p.detach().cpu().numpy(), r[0], 2e-6)
        _close(opt.state[p]["exp_avg_sq"].cpu().numpy(), r[2], 2e-6)


@gpu
def test_adamw_kernel_matches_torch_fused_at_benchmark_size():
    """ViT-B sized parameter set (86 M elements): our update and torch's own fused CUDA AdamW agree."""
    torch.manual_seed(5)
    shapes = [(2304, 768), (2304,), (768, 768), (768,), (3072, 768), (3072,), (768, 3072), (768,), (768,), (768,)] * 12
    shapes += [(768, 768), (1, 197, 768), (1000, 768), (1000,)]
    a = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.02) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa = FusedAdamW(a, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    ob = torch.optim.AdamW(b, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    for _ in range(3):
        for p, q in zip(a, b):
            p.grad = torch.randn_like(p)
            q.grad = p.grad.clone()
        oa.step()
        ob.step()
    for p, q in zip(a, b):
        _close(p.detach().cpu().numpy(), q.detach().cpu().numpy(), 2e-6)


def _loss_inputs(tag):
    shape, p, twoD, seed = CFG["loss_cases"][tag]
    data = fx.det_tensor(tuple(shape), seed)
    L_, D_ = G[f"{tag}:grad_full"].shape[1:]
    pred = fx.det_tensor((shape[0], L_, D_), seed + 1)
    return data, pred, torch.from_numpy(G[f"{tag}:mask"]), p, twoD


@gpu
@pytest.mark.parametrize("tag", ["mse2d", "mse3d"])
def test_patch_mse_matches_the_reference_vectors(tag):
    data, pred, mask, p, twoD = _loss_inputs(tag)
    for key, mk in (("masked", mask.cuda()), ("full", None)):
        pr = pred.cuda().requires_grad_(True)
        loss = metrics.patch_mse(pr, data.cuda(), p, twoD, mk)
        (loss * 1.0).backward()
        assert abs(loss.item() - float(G[f"{tag}:loss_{key}"])) <= 2e-6 * abs(loss.item())
        _close(pr.grad.cpu().numpy(), G[f"{tag}:grad_{key}"], 2e-6)


@gpu
def test_adaptive_patch_mse_matches_the_reference_vectors():
    seq = fx.det_tensor((2, 3, 6, 16), CFG["adaptive_seed"]).cuda()
    pr = fx.det_tensor((2, 6, 48), CFG["adaptive_seed"] + 1).cuda().requires_grad_(True)
    loss = metrics.adaptive_patch_mse(pr, seq)
    loss.backward()
    assert abs(loss.item() - float(G["adaptive:loss_full"])) <= 2e-6 * abs(loss.item())
    _close(pr.grad.cpu().numpy(), G["adaptive:grad_full"], 2e-6)


@gpu
@pytest.mark.parametrize("B,C,sp,p,dt_pred,dt_img", [
    (3, 3, (32, 48), 16, torch.float32, torch.float32),      # MAE 2-D, p = 16
    (2, 1, (24, 24), 8, torch.bfloat16, torch.float32),      # 64-pixel patches: idle lanes, bf16 prediction
    (2, 5, (40, 24), 8, torch.float32, torch.bfloat16),      # bf16 image (FSDP mixed precision feeds bf16 data)
    (1, 2, (32, 16, 32), 16, torch.float32, torch.float32),  # 3-D, 4096-pixel patches (16 pixels per thread)
    (2, 1, (12, 20, 12), 4, torch.bfloat16, torch.bfloat16),
    (1, 1, (20, 20), 20, torch.float32, torch.float32),      # 400 pixels = 100 quads per token
    (2, 2, (18, 30), 6, torch.float32, torch.float32),       # p % 4 != 0: scalar path, 36 pixels per token
    (1, 3, (12, 12, 18), 6, torch.bfloat16, torch.float32),  # scalar path in 3-D (216 pixels)
    (2, 4, (64, 32), 32, torch.bfloat16, torch.bfloat16),    # 1024 pixels, four channels
])
def test_patch_mse_against_the_oracle(B, C, sp, p, dt_pred, dt_img):
    torch.manual_seed(B * 100 + C)
    twoD = len(sp) == 2
    data = torch.randn(B, C, *sp).to(dt_img)
    tgt = T.patchify_np(data.float().numpy(), p, twoD)
    pred = torch.randn(*tgt.shape).to(dt_pred)
    mask = (torch.rand(tgt.shape[:2]) < 0.75).float()
    mask[0, 0] = 1.0
    tol = 2e-6 if dt_pred == torch.float32 else 1e-2
    for mk in (mask, None):
        lo, go = T.mse_loss_and_grad(pred.float().numpy(), tgt, None if mk is None else mk.numpy())
        pr = pred.cuda().requires_grad_(True)
        loss = metrics.patch_mse(pr, data.cuda(), p, twoD, None if mk is None else mk.cuda())
        (3.0 * loss).backward()                        # the incoming gradient is a device scalar, not assumed 1
        assert loss.dtype == torch.float32 and pr.grad.dtype == dt_pred
        assert abs(loss.item() - lo) <= 2e-6 * abs(lo)
        _close(pr.grad.float().cpu().numpy(), 3.0 * go, tol)
        if mk is not None and (mk == 0).any():         # visible tokens get exact zeros
            assert pr.grad[mk.cuda() == 0].abs().max().item() == 0.0


@gpu
def test_patch_mse_unaligned_tensors_take_the_scalar_path():
    """A prediction that starts 4 bytes past a 16-byte boundary cannot use vector loads: same numbers."""
    torch.manual_seed(21)
    data = torch.randn(2, 3, 32, 32, device="cuda")
    store = torch.randn(2 * 4 * 768 + 1, device="cuda")
    pred = store[1:].view(2, 4, 768)
    assert pred.data_ptr() % 16 == 4
    mask = torch.tensor([[1, 0, 1, 1], [0, 1, 1, 0]], device="cuda", dtype=torch.float32)
    pa, pb = pred.clone().requires_grad_(True), pred.detach().requires_grad_(True)
    la, lb = metrics.patch_mse(pa, data, 16, True, mask), metrics.patch_mse(pb, data, 16, True, mask)
    la.backward()
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-6 * abs(la.item())
    _close(pb.grad.cpu().numpy(), pa.grad.cpu().numpy(), 1e-6)
    lo, go = T.mse_loss_and_grad(pred.cpu().numpy(), T.patchify_np(data.cpu().numpy(), 16, True), mask.cpu().numpy())
    assert abs(la.item() - lo) <= 2e-6 * abs(lo)
    _close(pa.grad.cpu().numpy(), go, 2e-6)


@gpu
def test_patch_mse_properties_at_mae_size():
    """MAE ViT-L/16 pre-training shape (BASELINE configs[2]: B 256, 196 tokens of 768 values): size-independent
    properties instead of a CPU reference, plus agreement with the PyTorch formulation on the device."""
    torch.manual_seed(11)
    B, p = 256, 16
    data = torch.randn(B, 3, 224, 224, device="cuda")
    tgt = misc.patchify(data, p, True)
    mask = (torch.rand(B, 196, device="cuda") < 0.75).float()
    # (1) the patchified image itself has zero loss and zero gradient
    pr = tgt.clone().requires_grad_(True)
    loss = metrics.patch_mse(pr, data, p, True, mask)
    loss.backward()
    assert loss.item() == 0.0 and pr.grad.abs().max().item() == 0.0
    # (2) a constant offset d on every element gives exactly d^2 (masked and full)
    for mk in (mask, None):
        assert abs(metrics.patch_mse(tgt + 0.5, data, p, True, mk).item() - 0.25) <= 1e-6
    # (3) same value and gradient as the reference's formulation evaluated by PyTorch on the device
    pred = torch.randn_like(tgt)
    pa, pb = pred.clone().requires_grad_(True), pred.clone().requires_grad_(True)
    la = metrics.patch_mse(pa, data, p, True, mask)
    lb = metrics.masked_mse(pb, tgt, mask)
    la.backward()
    lb.backward()
    assert abs(la.item() - lb.item()) <= 2e-6 * abs(lb.item())
    _close(pa.grad.cpu().numpy(), pb.grad.cpu().numpy(), 2e-6)
    # (4) bit-reproducible
    assert metrics.patch_mse(pred, data, p, True, mask).item() == la.item()


# ----------------------------------------------------------------------------------------- DiceBLoss
def _dice_inputs(tag):
    shape, seed, w, sm, act = CFG["dice_cases"][tag]
    x = fx.det_tensor(tuple(shape), seed, scale=3.0)
    if not act:
        x = torch.sigmoid(x)
    t = (fx.det_tensor(tuple(shape), seed + 1) > 0.3).float()
    return x, t, w, sm, act


@pytest.mark.parametrize("tag", ["dice_2c", "dice_3c_w03", "dice_probs"])
def test_oracle_dice_bce_matches_the_reference(tag):
    x, t, w, sm, act = _dice_inputs(tag)
    lo, go = T.dice_bce_loss_and_grad(x.numpy(), t.numpy(), w, sm, act)
    assert abs(lo - float(G[f"{tag}:loss"])) <= 2e-6 * abs(lo)
    _close(go, G[f"{tag}:grad"], 2e-6)


def test_dice_loss_has_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        metrics.DiceBLoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    lib = L.lib()
    assert lib.ucf_dice_bce_fwd(1, 0, 1, 0, 2, 1, 16, 0.5, 1.0, 1, 1, 1, None) == -1 and b"C >= 2" in lib.ucf_last_error()
    assert lib.ucf_dice_bce_bwd(1, 0, 1, 0, None, None, 2, 2, 16, 0.5, 1, None, None) == -1


@gpu
@pytest.mark.parametrize("tag", ["dice_2c", "dice_3c_w03", "dice_probs"])
def test_dice_bce_matches_the_reference_vectors(tag):
    x, t, w, sm, act = _dice_inputs(tag)
    xg = x.cuda().requires_grad_(True)
    loss = metrics.DiceBLoss(weight=w, num_class=x.shape[1])(xg, t.cuda(), smooth=sm, act=act)
    loss.backward()
    assert abs(loss.item() - float(G[f"{tag}:loss"])) <= 4e-6 * abs(loss.item())
    _close(xg.grad.cpu().numpy(), G[f"{tag}:grad"], 4e-6)
    assert xg.grad[:, 0].abs().max().item() == 0.0


@gpu
@pytest.mark.parametrize("shape,dt,dtt", [((2, 2, 64, 64), torch.float32, torch.float32),
                                          ((3, 4, 33, 17), torch.float32, torch.float32),      # plane % 4 != 0: scalar path
                                          ((2, 2, 32, 32), torch.bfloat16, torch.float32),
                                          ((1, 3, 16, 24), torch.bfloat16, torch.bfloat16),
                                          ((2, 2, 8, 8, 8), torch.float32, torch.float32)])     # 3-D volumes flatten the same way
def test_dice_bce_against_the_oracle(shape, dt, dtt):
    torch.manual_seed(len(shape) + shape[-1])
    x = (torch.randn(*shape) * 2).to(dt)                    # |logits| < ~9: fp32 sigmoid is not yet saturated (float64 oracle)
    t = (torch.rand(*shape) < 0.4).to(dtt)
    lo, go = T.dice_bce_loss_and_grad(x.float().numpy(), t.float().numpy(), 0.5, 1.0, True)
    xg = x.cuda().requires_grad_(True)
    loss = metrics.DiceBLoss()(xg, t.cuda())
    (2.0 * loss).backward()
    assert loss.dtype == torch.float32 and xg.grad.dtype == dt
    assert abs(loss.item() - lo) <= 1e-5 * abs(lo)
    _close(xg.grad.float().cpu().numpy(), 2.0 * go, 2e-5 if dt == torch.float32 else 1e-2)


@gpu
def test_dice_bce_at_sap_size_matches_pytorch_on_the_device():
    """SAP segmentation shape (1024 x 1024 masks, two classes): same value and gradient as the reference's
    formulation evaluated by PyTorch ops on the device; bit-reproducible."""
    torch.manual_seed(17)
    x = torch.randn(4, 2, 1024, 1024, device="cuda") * 3
    t = (torch.rand(4, 2, 1024, 1024, device="cuda") < 0.3).float()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    la = metrics.DiceBLoss()(xa, t)
    la.backward()
    p = torch.sigmoid(xb)
    pred, true = torch.flatten(p[:, 1:]), torch.flatten(t[:, 1:])
    inter = (pred * true).sum()
    lb = 0.5 * torch.nn.functional.binary_cross_entropy(pred, true) + 0.5 * (1 - (2 * inter + 1) / (pred.sum() + true.sum() + 1))
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())
    _close(xa.grad.cpu().numpy(), xb.grad.cpu().numpy(), 2e-5)
    assert metrics.DiceBLoss()(x, t).item() == la.item()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dt,tdt", [((2, 4, 16, 16, 16), torch.float32, torch.uint8), ((3, 2, 40, 33), torch.float32, torch.int64),
                                           ((2, 5, 9, 7, 11), torch.bfloat16, torch.float32), ((1, 8, 64, 64, 64), torch.float32, torch.uint8)])
def test_dice_ce_kernels_match_the_torch_formulation(shape, dt, tdt):
    """ucf_dice_ce_fwd / _bwd against the PyTorch formulation of the same loss (DiceCELoss.forward_torch, fp32)."""
    from ucf_vit_b200.utils.metrics import DiceCELoss
    g = torch.Generator().manual_seed(sum(shape))
    logits = (torch.randn(shape, generator=g) * 2).cuda()
    target = torch.randint(0, shape[1], (shape[0], 1) + shape[2:], generator=g).to(tdt).cuda()
    for squared in (True, False):
        lossf = DiceCELoss(squared_pred=squared, smooth_nr=0.0, smooth_dr=1e-6)
        a = logits.to(dt).detach().requires_grad_(True)
        b = a.detach().float().requires_grad_(True)
        la = lossf(a, target)
        lb = lossf.forward_torch(b, target)
        (la * 1.7).backward()
        (lb * 1.7).backward()
        assert abs(la.item() - lb.item()) <= 2e-5 * abs(lb.item()) + 1e-6, (la.item(), lb.item())
        tol = 2e-5 if dt == torch.float32 else 1e-2
        err = (a.grad.float() - b.grad).abs().max().item()
        assert err <= tol * b.grad.abs().max().item() + 1e-9, err
        # channels-last logits (what the channels-last UNETR decoder emits) are read in place: same loss, gradient in that layout
        mf = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
        c = logits.to(dt).contiguous(memory_format=mf).detach().requires_grad_(True)
        lc = lossf(c, target)
        (lc * 1.7).backward()
        assert abs(lc.item() - la.item()) <= 1e-6 * abs(la.item()) + 1e-7, (lc.item(), la.item())
        assert c.grad.stride() == c.stride()
        assert (c.grad.float() - a.grad.float()).abs().max().item() <= 1e-6 * a.grad.float().abs().max().item() + 1e-12
