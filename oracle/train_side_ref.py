"""ORACLE -- TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

CPU restatement of the two steps either side of the transformer path in a training step
(SURVEY.md §8f ranks 2 and 3), pinned to the unmodified reference by tests/golden/host_train_side.npz
(written by oracle/gen_golden_train_side.py):

  * optimizer: the reference builds torch.optim.AdamW over two groups
    (/root/reference/src/UCF_VIT/utils/misc.py:58-84) and steps it once per batch
    (/root/reference/training_scripts/train_class_simple.py:355-357).  The arithmetic lives in the
    third-party dependency torch (`torch/optim/adam.py::_single_tensor_adam`, decoupled weight decay);
    `adamw_step` restates it in numpy fp32, operation by operation.
  * loss: `masked_mse(output, patchify(data), mask)` / `nn.MSELoss()` / the adaptive
    `rearrange(seq, 'b c s p -> b s (p c)')` target
    (/root/reference/training_scripts/train_masked_fsdp.py:40-62, utils/metrics.py:11-17, utils/misc.py:14-33).
"""
import numpy as np

NO_DECAY_KEYS = ("var_embed", "pos_embed", "time_pos_embed")      # utils/misc.py:62


def decay_groups(names):
    """utils/misc.py:59-66: (decay, no_decay) name lists, order preserved."""
    decay, no_decay = [], []
    for n in names:
        (no_decay if any(k in n for k in NO_DECAY_KEYS) else decay).append(n)
    return decay, no_decay


def adamw_step(p, g, m, v, step, lr, beta1, beta2, eps, weight_decay):
    """One update of fp32 arrays (returns new p, m, v); `step` is the 1-based count after the update."""
    f = np.float32
    p = p.astype(f) * f(1 - lr * weight_decay)                         # param.mul_(1 - lr * wd)
    m = m.astype(f) + (g.astype(f) - m.astype(f)) * f(1 - beta1)       # exp_avg.lerp_(grad, 1 - beta1)
    v = v.astype(f) * f(beta2) + f(1 - beta2) * g.astype(f) * g.astype(f)
    bc1 = 1 - beta1 ** step
    bc2_sqrt = (1 - beta2 ** step) ** 0.5
    denom = np.sqrt(v) / f(bc2_sqrt) + f(eps)
    p = p + f(-(lr / bc1)) * (m / denom)                               # addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(f), m.astype(f), v.astype(f)


def patchify_np(data, p, twoD):
    """utils/misc.py:14-33 on a numpy array: [N, C, X, Y(, Z)] -> [N, L, p^d * C], channel fastest."""
    n, c = data.shape[:2]
    g = [s // p for s in data.shape[2:]]
    if twoD:
        return data.reshape(n, c, g[0], p, g[1], p).transpose(0, 2, 4, 3, 5, 1).reshape(n, g[0] * g[1], p * p * c)
    t = data.reshape(n, c, g[0], p, g[1], p, g[2], p).transpose(0, 2, 4, 6, 3, 5, 7, 1)
    return t.reshape(n, g[0] * g[1] * g[2], p ** 3 * c)


def adaptive_target_np(seq):
    """einops 'b c s p -> b s (p c)' (train_masked_fsdp.py:42)."""
    b, c, s, p = seq.shape
    return seq.transpose(0, 2, 3, 1).reshape(b, s, p * c)


def mse_loss_and_grad(pred, target, mask=None):
    """Loss and d loss / d pred in float64.  mask None: nn.MSELoss (mean over every element);
    else utils/metrics.py:11-17: sum_tokens(mask * mean_d (pred - y)^2) / sum(mask)."""
    d = pred.astype(np.float64) - target.astype(np.float64)
    if mask is None:
        return (d * d).mean(), 2.0 * d / d.size
    w = mask.astype(np.float64)
    den = w.sum() * d.shape[-1]
    return ((d * d).sum(-1) * w).sum() / den, 2.0 * d * w[..., None] / den


def dice_bce_loss_and_grad(logits, targets, weight=0.5, smooth=1.0, act=True):
    """utils/metrics.py:95-121 in float64: loss and d loss / d logits (channel 0 gets zeros).  BCE follows
    torch.nn.functional.binary_cross_entropy: logs clamped at -100, gradient (p - t) / max(p (1 - p), 1e-12)."""
    x = logits.astype(np.float64)
    t = targets.astype(np.float64)[:, 1:]
    s_all = 1.0 / (1.0 + np.exp(-x)) if act else x
    s = s_all[:, 1:]
    inter, den = (s * t).sum(), s.sum() + t.sum() + smooth
    with np.errstate(divide="ignore"):
        bce = -(t * np.maximum(np.log(s), -100.0) + (1.0 - t) * np.maximum(np.log(1.0 - s), -100.0))
    loss = weight * bce.mean() + (1.0 - weight) * (1.0 - (2.0 * inter + smooth) / den)
    ds = weight / s.size * (s - t) / np.maximum(s * (1.0 - s), 1e-12) \
        + (1.0 - weight) * (-2.0 * t / den + (2.0 * inter + smooth) / den ** 2)
    grad = np.zeros_like(x)
    grad[:, 1:] = ds * (s * (1.0 - s) if act else 1.0)
    return loss, grad
