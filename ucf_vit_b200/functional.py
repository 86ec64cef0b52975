"""Autograd layer over the C-ABI ops: every Function's forward AND backward run hand-written
sm_100a kernels (ucf_vit_b200/csrc).  Activations are bf16, accumulation fp32, parameter
gradients are produced in the parameter's dtype (fp32 masters get fp32 grads straight from the
tensor-memory accumulators through the TMA reduce-add epilogue).

Reference semantics: /root/reference/src/UCF_VIT/simple/building_blocks.py
  Mlp.forward :122-129, Attention.forward :157-192, Block.forward :236-239.
"""
import ctypes
import math

import torch

from . import _lib as L
from . import ops

BF16 = torch.bfloat16


# ---------------------------------------------------------------------------------------------
# parameter staging: fp32 master -> bf16 compute copy
# ---------------------------------------------------------------------------------------------
def bf16_param(p: torch.Tensor) -> torch.Tensor:
    """bf16, contiguous, detached compute copy of a parameter (no copy if it already is bf16).

    Made afresh on every forward call and handed to that call's backward through `ctx` -- never cached
    across calls.  A cache keyed on the tensor's version counter is NOT safe: torch's fused (CUDA) AdamW /
    SGD update parameters without bumping `_version` (and so does any `p.data` mutation), so a cached copy
    silently goes stale after the first optimizer step.  Cost of doing it right: one pass over the weights
    per forward (0.5 GB for ViT-B, ~0.2 ms)."""
    q = p.detach()
    if q.dtype == BF16:
        return q if q.is_contiguous() else q.contiguous()
    if not q.is_contiguous():
        q = q.contiguous()
    return ops.cast_to_bf16(q) if q.dtype == torch.float32 else q.to(BF16)


def bf16_params(*ps):
    """bf16_param for several parameters; fp32 contiguous ones share one cast launch."""
    if all(p.dtype == torch.float32 and p.is_contiguous() for p in ps) and 1 < len(ps) <= 8:
        return ops.cast_to_bf16_multi([p.detach() for p in ps])
    return [bf16_param(p) for p in ps]


def _carve(flat, shapes, sizes):
    out, off = [], 0
    for s, n in zip(shapes, sizes):
        out.append(None if s is None else flat[off:off + math.prod(s)].view(s))
        off += n
    return out


def _zeros_f32(dev, *shapes, also_as=None):
    """One zero-filled fp32 allocation carved into 256-byte aligned views (one fill launch instead of
    one per gradient; the reduce-add epilogues accumulate into them).  A shape of None yields None.
    `also_as=dtype`: returns (views, finish) where `finish()` casts the WHOLE slab to `dtype` in one launch and
    returns the same views of the copy (bf16 parameters under FSDP MixedPrecision: 1 cast instead of 12)."""
    sizes = [0 if s is None else -(-math.prod(s) // 64) * 64 for s in shapes]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    views = _carve(flat, shapes, sizes)
    if also_as is None:
        return views
    return views, (lambda: _carve(flat.to(also_as), shapes, sizes))


_SM_COUNT = {}


def _wgrad_splits(n_out, k_in, M, dev):
    """Split-K factor of the weight-gradient GEMM dW[n_out, k_in] = dY^T X over M tokens.

    The persistent kernels hand tile x split work units to SMs (CTA-pair kernel: 256x256 tiles over
    SMs/2 clusters) in rounds, so the factor is picked to fill whole rounds: 36 tiles x 5 splits on 74
    clusters is 2.43 rounds = 81 % busy, 36 x 4 is 1.95 rounds = 97 %."""
    kb = (M + 63) // 64
    if kb < 16:
        return 1
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    sms = _SM_COUNT.get(idx)
    if sms is None:
        sms = _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    pair = n_out >= 512 and 512 <= k_in <= 4096          # mirrors the auto tile selection in ucf_gemm_bf16
    if pair:
        tiles, workers = ((n_out + 255) // 256) * ((k_in + 255) // 256), max(1, sms // 2)
    else:
        tiles, workers = ((n_out + 127) // 128) * ((k_in + 255) // 256), sms
    best, best_eff = 1, 0.0
    cands = []
    for s_ in range(1, 17):
        if kb // s_ < 8:
            break
        per = -(-kb // s_)
        units = tiles * (-(-kb // per))                   # the launcher drops empty splits the same way
        eff = units / (-(-units // workers) * workers)
        cands.append((s_, units, eff))
        best_eff = max(best_eff, eff)
    for s_, units, eff in cands:                          # fewest splits within 0.5 % of the best fill, two rounds if possible
        if eff >= best_eff - 0.005 and units >= 2 * workers:
            return s_
    for s_, units, eff in cands:
        if eff >= best_eff - 0.005:
            return s_
    return best


def _wgrad(dy2, x2, n_out, k_in, like: torch.Tensor, bias_like=None, out=None, bias_out=None):
    """dW[n_out,k_in] = dy2^T x2 (fp32, split-K TMA reduce-add) in `like`'s dtype, and -- fused in the
    same kernel from the dY tiles it stages anyway -- the bias gradient db[n_out] = colsum(dy2).
    `out` / `bias_out` are zeroed fp32 buffers to accumulate into (allocated here when None).
    Returns (dW, db); db is None when `bias_like` is None."""
    M = dy2.shape[0]
    dw = out if out is not None else torch.zeros((n_out, k_in), dtype=torch.float32, device=dy2.device)
    db = None
    if bias_like is not None:
        db = bias_out if bias_out is not None else torch.zeros((n_out,), dtype=torch.float32, device=dy2.device)
    splits = _wgrad_splits(n_out, k_in, M, dy2.device)
    ops.gemm(dy2, x2, M=n_out, N=k_in, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw, splits=splits,
             bias_grad=db)
    if like.dtype != torch.float32:
        dw = dw.to(like.dtype)
    if db is not None and bias_like.dtype != torch.float32:
        db = db.to(bias_like.dtype)
    return dw, db


def _bgrad(dy2, like):
    if like is None:
        return None
    db = ops.colsum(dy2)
    return db if like.dtype == torch.float32 else db.to(like.dtype)


def _as_bf16_2d(x):
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    if x2.dtype != BF16:
        x2 = ops.cast_to_bf16(x2.contiguous()) if x2.dtype == torch.float32 else x2.to(BF16)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    return x2


# ---------------------------------------------------------------------------------------------
# dtype boundary
# ---------------------------------------------------------------------------------------------
class _ToBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.src_dtype = x.dtype
        return ops.cast_to_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.cast_to_f32(g.contiguous()) if ctx.src_dtype == torch.float32 else g.to(ctx.src_dtype)


def to_bf16(x):
    if x.dtype == BF16:
        return x
    if x.dtype != torch.float32:
        x = x.float()
    return _ToBF16.apply(x)


# ---------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        xs = x if x.is_contiguous() else x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(xs, weight, bias, eps)
        if xs.dtype != BF16:      # backward kernel reads bf16 activations
            xs = ops.cast_to_bf16(xs)
        ctx.save_for_backward(xs, weight, mean, rstd)
        ctx.has_bias = bias is not None
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, weight, mean, rstd = ctx.saved_tensors
        D = xs.shape[-1]
        dy = dy if dy.is_contiguous() else dy.contiguous()
        dg = torch.zeros(D, dtype=torch.float32, device=dy.device) if weight is not None else None
        db = torch.zeros(D, dtype=torch.float32, device=dy.device) if ctx.has_bias else None
        dx = ops.layernorm_bwd(dy, xs, weight, mean, rstd, dgamma=dg, dbeta=db)
        if ctx.x_dtype != BF16:
            dx = dx.to(ctx.x_dtype)
        if dg is not None and weight.dtype != torch.float32:
            dg = dg.to(weight.dtype)
            db = db.to(weight.dtype) if db is not None else None
        return dx, dg, db, None


def layer_norm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm over the last dim; returns bf16."""
    return _LayerNormFn.apply(x, weight, bias, eps)


# ---------------------------------------------------------------------------------------------
# Linear (+bias, +residual)
# ---------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        shp = x.shape
        N, K = weight.shape
        x2 = _as_bf16_2d(x)
        M = x2.shape[0]
        w = bf16_param(weight)
        if residual is not None:
            r2 = _as_bf16_2d(residual)
            y = ops.gemm(x2, w, M=M, N=N, K=K, bias=bias, aux=r2, epilogue=L.EPI_BIAS_RESIDUAL)
        else:
            y = ops.gemm(x2, w, M=M, N=N, K=K, bias=bias)
        ctx.save_for_backward(x2, weight, bias)
        ctx.w16 = w                      # this call's bf16 weight copy, reused by its backward
        ctx.has_res = residual is not None
        ctx.x_dtype = x.dtype
        ctx.shp = shp
        return y.view(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, bias = ctx.saved_tensors
        N, K = weight.shape
        dy2 = _as_bf16_2d(dy)
        M = dy2.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy2, ctx.w16, M=M, N=K, K=N, b_mn=True).view(ctx.shp)
            if ctx.x_dtype != BF16:
                dx = dx.to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            dw, db = _wgrad(dy2, x2, N, K, weight, bias if (bias is not None and ctx.needs_input_grad[2]) else None)
        elif bias is not None and ctx.needs_input_grad[2]:
            db = _bgrad(dy2, bias)
        dres = dy if ctx.has_res else None
        return dx, dw, db, dres


def linear(x, weight, bias=None, residual=None):
    """y = x W^T + b (+ residual); bf16 out.  weight [out, in] as nn.Linear stores it."""
    return _LinearFn.apply(x, weight, bias, residual)


# ---------------------------------------------------------------------------------------------
# MLP: fc1 -> GELU(erf) -> fc2 (+residual), GELU fused in fc1's epilogue, GELU' in fc2-dgrad's
# ---------------------------------------------------------------------------------------------
class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        shp = x.shape
        Hd, K = w1.shape
        Nout = w2.shape[0]
        x2 = _as_bf16_2d(x)
        M = x2.shape[0]
        w1h, w2h = bf16_param(w1), bf16_param(w2)
        u, z = ops.gemm(x2, w1h, M=M, N=Hd, K=K, bias=b1, epilogue=L.EPI_BIAS_GELU_AUX)
        if residual is not None:
            y = ops.gemm(u, w2h, M=M, N=Nout, K=Hd, bias=b2, aux=_as_bf16_2d(residual),
                         epilogue=L.EPI_BIAS_RESIDUAL)
        else:
            y = ops.gemm(u, w2h, M=M, N=Nout, K=Hd, bias=b2)
        ctx.save_for_backward(x2, z, u, w1, b1, w2, b2)
        ctx.w16 = (w1h, w2h)
        ctx.has_res = residual is not None
        ctx.shp = shp
        ctx.x_dtype = x.dtype
        return y.view(*shp[:-1], Nout)

    @staticmethod
    def backward(ctx, dy):
        x2, z, u, w1, b1, w2, b2 = ctx.saved_tensors
        Hd, K = w1.shape
        Nout = w2.shape[0]
        dy2 = _as_bf16_2d(dy)
        M = dy2.shape[0]
        dw2, db2 = _wgrad(dy2, u, Nout, Hd, w2, b2)
        w1h, w2h = ctx.w16
        dz = ops.gemm(dy2, w2h, M=M, N=Hd, K=Nout, b_mn=True, aux=z, epilogue=L.EPI_DGELU)
        dw1, db1 = _wgrad(dz, x2, Hd, K, w1, b1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dz, w1h, M=M, N=K, K=Hd, b_mn=True).view(ctx.shp)
            if ctx.x_dtype != BF16:
                dx = dx.to(ctx.x_dtype)
        return dx, dw1, db1, dw2, db2, (dy if ctx.has_res else None)


def mlp(x, w1, b1, w2, b2, residual=None):
    return _MlpFn.apply(x, w1, b1, w2, b2, residual)


# ---------------------------------------------------------------------------------------------
# attention core on a packed qkv projection
# ---------------------------------------------------------------------------------------------
class _AttnPackedFn(torch.autograd.Function):
    """qkv: [B, N, 3, H, hd] bf16 (the raw output of Attention.qkv) -> o: [B, N, H*hd] bf16."""

    @staticmethod
    def forward(ctx, qkv, scale):
        B, N, _, H, hd = qkv.shape
        o, lse = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.scale = scale
        return o.view(B, N, H * hd)

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse = ctx.saved_tensors
        B, N, _, H, hd = qkv.shape
        d_o = d_o.contiguous().view(B, N, H, hd)
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, d_o, lse, ctx.scale,
                          dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
        return dqkv, None


def _padded_head_dim(hd):
    """Head width the tcgen05 attention kernels run a head of `hd` columns at (zero columns appended to q, k and v
    change neither the scores nor the output columns that are kept)."""
    if hd in (32, 64):
        return hd
    if hd < 32:
        return 32
    if hd < 64:
        return 64
    raise NotImplementedError(f"attention: head_dim {hd} > 64 has no kernel (32 and 64 run natively, narrower heads "
                              "zero-padded; e.g. ViT-H's 80 is not supported)")


def attention_packed(qkv, scale):
    """qkv [B, N, 3, H, hd] -> [B, N, H*hd].  head_dim 36 (decoder_embed_dim 576 / 16 heads in the reference's
    configs/basic_ct MAE and diffusion YAMLs) and other widths below 64 run zero-padded to 32 / 64."""
    hd = qkv.shape[-1]
    hp = _padded_head_dim(hd)
    if hp == hd:
        return _AttnPackedFn.apply(qkv, scale)
    B, N, _, H, _ = qkv.shape
    o = _AttnPackedFn.apply(torch.nn.functional.pad(qkv, (0, hp - hd)), scale)
    return o.view(B, N, H, hp)[..., :hd].reshape(B, N, H * hd)


class _AttnFn(torch.autograd.Function):
    """q [B,Nq,H,hd], k/v [B,Nk,H,hd] (bf16, last dim contiguous) -> o [B,Nq,H,hd]."""

    @staticmethod
    def forward(ctx, q, k, v, scale):
        o, lse = ops.attention_fwd(q, k, v, scale)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = ops.attention_bwd(q, k, v, o, d_o.contiguous(), lse, ctx.scale)
        return dq, dk, dv, None


def attention(q, k, v, scale):
    hd = q.shape[-1]
    hp = _padded_head_dim(hd)
    if hp == hd:
        return _AttnFn.apply(q, k, v, scale)
    pad = torch.nn.functional.pad
    return _AttnFn.apply(pad(q, (0, hp - hd)), pad(k, (0, hp - hd)), pad(v, (0, hp - hd)), scale)[..., :hd]


# ---------------------------------------------------------------------------------------------
# whole pre-norm transformer block with a hand-scheduled backward
# ---------------------------------------------------------------------------------------------
class _BlockFn(torch.autograd.Function):
    """x + proj(attn(LN1(x))) then + fc2(gelu(fc1(LN2(.))))  (Block.forward, building_blocks.py:236-239)
    for the configuration every reference driver uses: no qk_norm, no LayerScale, drop rates 0.

    Fusions: bias / bias+GELU(+pre-activation) / bias+residual in GEMM epilogues; GELU' in the
    fc2-dgrad epilogue; residual-gradient add inside the LayerNorm backward kernel; attention reads
    the packed qkv projection in place and its backward writes the packed dqkv in place."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                num_heads, eps1, eps2):
        B, N, D = x.shape
        M = B * N
        H = num_heads
        Hd = fc1_w.shape[0]
        dev = x.device
        x2 = x.reshape(M, D)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        # one bf16 slab for everything backward re-reads, one fp32 slab for the statistics, one for the weights
        sizes = (M * D, 3 * M * D, M * D, M * D, M * D, M * Hd, M * Hd)
        slab = torch.empty(sum(sizes), dtype=BF16, device=dev)
        h1, qkv, o, x1, h2, z, u = torch.split(slab, sizes)
        Mp = -(-M // 64) * 64
        stats = torch.empty(4 * Mp + B * H * N, dtype=torch.float32, device=dev)
        y = torch.empty((B, N, D), dtype=BF16, device=dev)
        ws, masters, w16 = (qkv_w, proj_w, fc1_w, fc2_w), [], []
        cast = [w for w in ws if w.dtype != BF16]
        if cast:
            if any(w.dtype != torch.float32 or not w.is_contiguous() for w in cast):
                raise TypeError("fused_block: weights must be contiguous fp32 or bf16 tensors")
            wslab = torch.empty(sum(-(-w.numel() // 64) * 64 for w in cast), dtype=BF16, device=dev)
            off = 0
        for w in ws:
            if w.dtype == BF16:
                w16.append(w.detach() if w.is_contiguous() else w.detach().contiguous())
                masters.append(None)
            else:
                w16.append(wslab[off:off + w.numel()].view(w.shape))
                masters.append(w.data_ptr())
                off += -(-w.numel() // 64) * 64
        ptr = lambda t: None if t is None else t.data_ptr()
        biases = (qkv_b, proj_b, fc1_b, fc2_b)
        bdt = next((ops._dt(b) for b in biases if b is not None), 0)
        prm = L.BlockParams(B, N, D, H, Hd, float(eps1), float(eps2), ops._dt(n1w), bdt,
                            ptr(n1w), ptr(n1b), ptr(n2w), ptr(n2b), ptr(qkv_b), ptr(proj_b), ptr(fc1_b), ptr(fc2_b),
                            *[w.data_ptr() for w in w16], *masters)
        acts = L.BlockActs(x2.data_ptr(), h1.data_ptr(), qkv.data_ptr(), o.data_ptr(), x1.data_ptr(), h2.data_ptr(),
                           z.data_ptr(), u.data_ptr(), y.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * Mp,
                           stats.data_ptr() + 8 * Mp, stats.data_ptr() + 12 * Mp, stats.data_ptr() + 16 * Mp)
        ops._require_cuda(x2, n1w, qkv_w, proj_w, fc1_w, fc2_w)
        L.check(L.lib().ucf_block_fwd(ctypes.byref(prm), ctypes.byref(acts), ops._stream()), "block_fwd")
        ctx.w16 = w16                    # this call's bf16 weight copies, reused by its backward
        ctx.eps = (float(eps1), float(eps2))
        ctx.save_for_backward(x2, slab, stats, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
        ctx.dims = (B, N, D, H, Hd)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x2, slab, stats, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b) = ctx.saved_tensors
        B, N, D, H, Hd = ctx.dims
        M = B * N
        dev = dy.device
        dy2 = dy.reshape(M, D)
        if dy2.dtype != BF16:
            dy2 = dy2.to(BF16)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        f32 = torch.float32

        def cast_like(g, like):
            return g if (g is None or like.dtype == f32) else g.to(like.dtype)

        def opt(b, shape):
            return shape if b is not None else None

        plist = (n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
        pdt = {t.dtype for t in plist if t is not None}
        one_cast = len(pdt) == 1 and f32 not in pdt            # e.g. every parameter bf16 (FSDP MixedPrecision)
        grads, finish = _zeros_f32(dev, (D,), opt(n1b, (D,)), (3 * D, D), opt(qkv_b, (3 * D,)), (D, D), opt(proj_b, (D,)),
                                   (D,), opt(n2b, (D,)), (Hd, D), opt(fc1_b, (Hd,)), (D, Hd), opt(fc2_b, (D,)),
                                   also_as=next(iter(pdt)) if one_cast else f32)
        wsa = max(Hd, 3 * D)
        ws = torch.empty(M * (wsa + 2 * D), dtype=BF16, device=dev)
        dx = torch.empty((B, N, D), dtype=BF16, device=dev)
        nlse = B * H * N
        f32ws = torch.empty(nlse + (M * D if N > 256 else 0), dtype=f32, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        g = L.BlockGrads(dy2.data_ptr(), dx.data_ptr(), *[ptr(t) for t in grads],
                         ws.data_ptr(), ws.data_ptr() + 2 * M * wsa, ws.data_ptr() + 2 * M * (wsa + D),
                         (f32ws.data_ptr() + 4 * nlse) if N > 256 else None, f32ws.data_ptr())
        # the descriptors are rebuilt from the saved tensors: under activation checkpointing these are the RECOMPUTED
        # buffers, not the ones the original forward call wrote
        h1, qkv, o, x1, h2, z, u = torch.split(slab, (M * D, 3 * M * D, M * D, M * D, M * D, M * Hd, M * Hd))
        Mp = -(-M // 64) * 64
        biases = (qkv_b, proj_b, fc1_b, fc2_b)
        bdt = next((ops._dt(b) for b in biases if b is not None), 0)
        prm = L.BlockParams(B, N, D, H, Hd, ctx.eps[0], ctx.eps[1], ops._dt(n1w), bdt,
                            ptr(n1w), ptr(n1b), ptr(n2w), ptr(n2b), ptr(qkv_b), ptr(proj_b), ptr(fc1_b), ptr(fc2_b),
                            *[w.data_ptr() for w in ctx.w16], None, None, None, None)
        acts = L.BlockActs(x2.data_ptr(), h1.data_ptr(), qkv.data_ptr(), o.data_ptr(), x1.data_ptr(), h2.data_ptr(),
                           z.data_ptr(), u.data_ptr(), None, stats.data_ptr(), stats.data_ptr() + 4 * Mp,
                           stats.data_ptr() + 8 * Mp, stats.data_ptr() + 12 * Mp, stats.data_ptr() + 16 * Mp)
        L.check(L.lib().ucf_block_bwd(ctypes.byref(prm), ctypes.byref(acts), ctypes.byref(g), ops._stream()), "block_bwd")
        if one_cast:
            grads = finish()
        (d_n1w, d_n1b, d_qkv_w, d_qkv_b, d_proj_w, d_proj_b, d_n2w, d_n2b, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b) = grads
        return (dx, cast_like(d_n1w, n1w), cast_like(d_n1b, n1b),
                cast_like(d_qkv_w, qkv_w), cast_like(d_qkv_b, qkv_b), cast_like(d_proj_w, proj_w), cast_like(d_proj_b, proj_b),
                cast_like(d_n2w, n2w), cast_like(d_n2b, n2b),
                cast_like(d_fc1_w, fc1_w), cast_like(d_fc1_b, fc1_b), cast_like(d_fc2_w, fc2_w), cast_like(d_fc2_b, fc2_b),
                None, None, None)


# ---------------------------------------------------------------------------------------------
# forward-only block: LayerNorm folded into the QKV projection (no normalised activations are written)
# ---------------------------------------------------------------------------------------------
def fold_layernorm(weight, bias, ln_weight, ln_bias):
    """Constants of `LN(x) W^T + b == rstd * (x Wg^T - mean * colsum) + b_folded` (ops.ln_gemm):
    Wg = W diag(gamma) rounded to bf16, colsum = row sums of THAT rounded matrix (so the mean term cancels exactly what
    the tensor cores accumulated), b_folded = W beta + b in fp32."""
    w32 = weight.detach().float()
    g = ln_weight.detach().float() if ln_weight is not None else torch.ones(w32.shape[1], device=w32.device)
    wg = (w32 * g[None, :]).to(BF16).contiguous()
    colsum = wg.float().sum(1).contiguous()
    bf = torch.zeros(w32.shape[0], device=w32.device) if bias is None else bias.detach().float().clone()
    if ln_bias is not None:
        bf = bf + w32 @ ln_bias.detach().float()
    return wg, colsum, bf.contiguous()


@torch.no_grad()
def block_forward_nograd(x, folded, n1_eps, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b, num_heads, eps2):
    """Block.forward when no gradient is wanted (eval / inference): statistics pass + ONE GEMM for LayerNorm1 + QKV
    (`ucf_layernorm_stats` + `ucf_ln_gemm`), nothing saved for backward.  `folded` = fold_layernorm(qkv.weight, qkv.bias,
    norm1.weight, norm1.bias)."""
    B, N, D = x.shape
    M, H = B * N, num_heads
    hd = D // H
    x2 = x.reshape(M, D)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    wg, colsum, bfold = folded
    mean, rstd = ops.layernorm_stats(x2, n1_eps)
    qkv5 = ops.ln_gemm(x2, wg, bfold, colsum, mean, rstd).view(B, N, 3, H, hd)
    o, _ = ops.attention_fwd(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], hd ** -0.5)
    wp, w1h, w2h = bf16_params(proj_w, fc1_w, fc2_w)
    x1 = ops.gemm(o.view(M, D), wp, M=M, N=D, K=D, bias=proj_b, aux=x2, epilogue=L.EPI_BIAS_RESIDUAL)
    h2, _, _ = ops.layernorm_fwd(x1, n2w, n2b, eps2)
    Hd = fc1_w.shape[0]
    u, _ = ops.gemm(h2, w1h, M=M, N=Hd, K=D, bias=fc1_b, epilogue=L.EPI_BIAS_GELU_AUX)
    y = ops.gemm(u, w2h, M=M, N=D, K=Hd, bias=fc2_b, aux=x1, epilogue=L.EPI_BIAS_RESIDUAL)
    return y.view(B, N, D)


def fused_block(x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                num_heads, eps1, eps2):
    return _BlockFn.apply(x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                          num_heads, eps1, eps2)


# ---------------------------------------------------------------------------------------------
# variable-aggregation cross-attention core (Nq = N_a, Nk = V per spatial token)
# ---------------------------------------------------------------------------------------------
class _VarAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, kv, scale):
        q = q.contiguous()
        kv = kv.contiguous()
        o, lse = ops.var_attention_fwd(q, kv, scale)
        ctx.save_for_backward(q, kv, o, lse)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, kv, o, lse = ctx.saved_tensors
        dq_acc, dkv = ops.var_attention_bwd(q, kv, o, d_o, lse, ctx.scale)
        return dq_acc.to(q.dtype), dkv, None


def var_attention(q, kv, scale):
    """q [Bq, N_a, H, hd] (Bq == rows, or 1 = one query shared by all rows), kv [rows, V, 2, H, hd]."""
    return _VarAttnFn.apply(q, kv, scale)


# ---------------------------------------------------------------------------------------------
# cast + patchify (Conv(k = s = p) input -> GEMM rows)
# ---------------------------------------------------------------------------------------------
class _PatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        ctx.shape, ctx.p, ctx.dt = x.shape, p, x.dtype
        if x.dtype not in (torch.float32, BF16, torch.uint8):
            x = x.float()
        return ops.patchify(x.contiguous(), p)

    @staticmethod
    def backward(ctx, g):
        # fold back (only reached when the image itself requires grad, which no driver does)
        shp, p = ctx.shape, ctx.p
        B, C = shp[:2]
        G = [s // p for s in shp[2:]]
        nd = len(G)
        g = g[:, :C * p ** nd]                                   # drop the K pad columns
        if nd == 2:
            core = g.reshape(B, G[0], G[1], C, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, C, G[0] * p, G[1] * p)
        else:
            core = g.reshape(B, G[0], G[1], G[2], C, p, p, p).permute(0, 4, 1, 5, 2, 6, 3, 7).reshape(
                B, C, G[0] * p, G[1] * p, G[2] * p)
        if tuple(core.shape) != tuple(shp):                      # cropped border pixels get no gradient
            dx = core.new_zeros(shp)
            dx[(slice(None), slice(None)) + tuple(slice(0, gi * p) for gi in G)] = core
            core = dx
        return core.to(ctx.dt), None


def patchify(x, p):
    """[B,C,H,W(,Z)] -> bf16 [B*L, K8], K = C*p^d ordered (c, p0, p1(, p2)) like conv.weight.view(D,-1), zero-padded to
    K8 = ceil(K/8)*8 columns; pixels past the last whole patch are ignored (like the strided convolution)."""
    if x.requires_grad:
        return _PatchifyFn.apply(x, p)
    with torch.no_grad():
        return _PatchifyFn.apply(x, p)


# ---------------------------------------------------------------------------------------------
# class-token concat + position-embedding add in one pass
# ---------------------------------------------------------------------------------------------
class _AssembleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, prefix, pos, pos_has_prefix):
        tok_c = tok if tok.is_contiguous() else tok.contiguous()
        pdt = None
        for t in (prefix, pos):
            if t is not None:
                pdt = t.dtype if pdt is None else pdt
        pf = None if prefix is None else prefix.reshape(-1, prefix.shape[-1]).to(pdt).contiguous()
        ps = None if pos is None else pos.to(pdt).contiguous()
        ctx.P = 0 if pf is None else pf.shape[0]
        ctx.meta = (None if prefix is None else (prefix.shape, prefix.dtype), None if pos is None else (pos.shape, pos.dtype),
                    pos_has_prefix)
        return ops.assemble_tokens(tok_c, pf, ps, pos_has_prefix)

    @staticmethod
    def backward(ctx, g):
        P = ctx.P
        pmeta, smeta, has_prefix = ctx.meta
        d_tok = g[:, P:, :] if ctx.needs_input_grad[0] else None
        d_prefix = d_pos = None
        if pmeta is not None and ctx.needs_input_grad[1]:
            d_prefix = g[:, :P, :].sum(0, dtype=torch.float32).reshape(pmeta[0]).to(pmeta[1])
        if smeta is not None and ctx.needs_input_grad[2]:
            gp = g if has_prefix else g[:, P:, :]
            shp, dt = smeta
            if len(shp) == 3 and shp[0] == g.shape[0] and shp[0] != 1:
                d_pos = gp.to(dt)                                      # per-sample embedding
            else:
                d_pos = gp.sum(0, dtype=torch.float32).reshape(shp).to(dt)
        return d_tok, d_prefix, d_pos, None


class _AddBcastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, e):
        ctx.meta = (tuple(x.shape), tuple(e.shape), e.dtype)
        return ops.add_bcast(x if x.is_contiguous() else x.contiguous(), e if e.is_contiguous() else e.contiguous())

    @staticmethod
    def backward(ctx, g):
        xs, es, edt = ctx.meta
        d_e = None
        if ctx.needs_input_grad[1]:
            pad = (1,) * (len(xs) - len(es)) + es
            dims = [i for i in range(len(xs)) if pad[i] == 1 and xs[i] != 1]
            d_e = (g.sum(dim=dims, keepdim=True, dtype=torch.float32) if dims else g.float()).reshape(es).to(edt)
        return (g if ctx.needs_input_grad[0] else None), d_e


def add_bcast(x, e):
    """bf16 tokens + a broadcast embedding (variable embedding over [B, V, L, D], time embedding over [B, N, D])
    in one pass (ucf_add_bcast); e keeps its own dtype (fp32 parameters are not cast first)."""
    return _AddBcastFn.apply(x, e)


class _GatherTokensFn(torch.autograd.Function):
    """out[b, i] = (idx[b, i] < Ls ? src[b, idx[b, i]] : fill) + pos  (ucf_gather_tokens / ucf_scatter_tokens).
    `complete`: every row of src is named exactly once by idx (idx is a permutation padded with fill rows), so
    the backward scatter needs no zero-fill."""

    @staticmethod
    def forward(ctx, src, idx, fill, pos, complete):
        src_c = src if src.is_contiguous() else src.contiguous()
        idx_c = idx if idx.is_contiguous() else idx.contiguous()
        pdt = None
        for t in (fill, pos):
            if t is not None:
                pdt = t.dtype if pdt is None else pdt
        fl = None if fill is None else fill.reshape(-1).to(pdt).contiguous()
        ps = None if pos is None else pos.to(pdt).contiguous()
        ctx.save_for_backward(idx_c)
        ctx.meta = (src.shape[1], None if fill is None else (fill.shape, fill.dtype),
                    None if pos is None else (pos.shape, pos.dtype), bool(complete))
        return ops.gather_tokens(src_c, idx_c, fl, ps)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        Ls, fmeta, pmeta, complete = ctx.meta
        g = g if g.is_contiguous() else g.contiguous()
        need_src = ctx.needs_input_grad[0]
        need_fill = fmeta is not None and ctx.needs_input_grad[2]
        d_src = d_fill = d_pos = None
        if need_src or need_fill:
            d_src, df = ops.scatter_tokens(g, idx, Ls, need_src=need_src, need_fill=need_fill, zero_first=not complete)
            if need_fill:
                d_fill = df.reshape(fmeta[0]).to(fmeta[1])
        if pmeta is not None and ctx.needs_input_grad[3]:
            shp, dt = pmeta
            if len(shp) == 3 and shp[0] == g.shape[0] and shp[0] != 1:
                d_pos = g.to(dt)                                       # per-sample embedding
            else:
                d_pos = g.sum(0, dtype=torch.float32).reshape(shp).to(dt)
        return d_src, None, d_fill, d_pos, None


def gather_tokens(src, idx, fill=None, pos=None, complete=False):
    """Row gather with an optional fill row for out-of-range indices and an optional embedding add:
    MAE.random_masking's kept-token gather and MAE.mask_head's cat + gather + pos-embed add (arch.py:674-698)."""
    return _GatherTokensFn.apply(src, idx, fill, pos, complete)


def assemble_tokens(tok, prefix=None, pos=None, pos_has_prefix=True):
    """concat(prefix tokens, tok) + pos  ->  bf16 [B, P+L, D]  (VIT._pos_embed, arch.py:367-393)."""
    return _AssembleFn.apply(tok, prefix, pos, pos_has_prefix)


# ---- UNETR decoder block bodies (channels-last bf16) ------------------------------------------------------------------------
class _InstNormActFn(torch.autograd.Function):
    """y = lrelu(IN(a) [+ IN(b) | + b]): MONAI UnetResBlock / UnetBasicBlock bodies between the convolutions
    (/root/reference/src/UCF_VIT/simple/arch.py:808-940 build them; InstanceNorm affine=False, LeakyReLU 0.01)."""

    @staticmethod
    def forward(ctx, a, b, norm_b, slope, eps):
        stats_a = ops.inorm_stats(a, eps)
        stats_b = ops.inorm_stats(b, eps) if (b is not None and norm_b) else None
        y = ops.inorm_apply(a, stats_a, b, stats_b, slope)
        ctx.save_for_backward(a, b, y if slope != 1.0 else None, stats_a, stats_b)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        a, b, y, stats_a, stats_b = ctx.saved_tensors
        if dy.stride() != a.stride():
            dy = torch.empty_like(a, memory_format=torch.preserve_format).copy_(dy)
        da, db = ops.inorm_bwd(dy, y, a, stats_a, b, stats_b, ctx.slope)
        return da, db, None, None, None


def channels_last(x):
    """x [N, C, *spatial] in channels-last memory ([N, *spatial, C]); a no-op when it already is."""
    if x.movedim(1, -1).is_contiguous():
        return x
    return x.contiguous(memory_format=torch.channels_last if x.dim() == 4 else torch.channels_last_3d)


def instance_norm_act(a, residual=None, norm_residual=False, negative_slope=0.01, eps=1e-5):
    """LeakyReLU(InstanceNorm(a) + [InstanceNorm(residual) | residual]) on channels-last bf16 CUDA tensors [N, C, ...];
    negative_slope = 1 leaves the activation out."""
    if a.dtype != torch.bfloat16 or not a.is_cuda:
        raise RuntimeError("ucf_vit_b200: instance_norm_act takes bf16 CUDA tensors (no CPU fallback)")
    a = channels_last(a)
    if residual is not None:
        residual = channels_last(residual.to(torch.bfloat16))
    return _InstNormActFn.apply(a, residual, bool(norm_residual), float(negative_slope), float(eps))


class _Conv3x3x3Fn(torch.autograd.Function):
    """conv3d(x, w, stride 1, padding 1, no bias) on channels-last bf16 x with an fp32 (or bf16) master weight: cuDNN for the
    forward and the data gradient, ucf_conv3d_wgrad for the weight gradient (fp32, handed to the master weight directly)."""

    @staticmethod
    def forward(ctx, x, w):
        wb = w.detach().to(torch.bfloat16)
        ctx.save_for_backward(x, wb)
        ctx.w_dtype = w.dtype
        return torch.nn.functional.conv3d(x, wb, None, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        x, wb = ctx.saved_tensors
        dy = channels_last(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.ops.aten.convolution_backward(dy, x, wb, None, [1, 1, 1], [1, 1, 1], [1, 1, 1], False, [0, 0, 0], 1,
                                                     [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            Ci = x.shape[1]
            if Ci < 16:       # the decoder's first layer (in_chans -> 16): zero channels appended, their gradient rows dropped
                xp = x.new_zeros(x.shape[0], *x.shape[2:], 16).movedim(-1, 1)
                xp[:, :Ci] = x
                dw = ops.conv3d_wgrad(xp, dy)[:, :Ci].to(ctx.w_dtype)
            else:
                dw = ops.conv3d_wgrad(x, dy).to(ctx.w_dtype)
        return dx, dw


def conv3x3x3(x, weight):
    """3x3x3 convolution (stride 1, padding 1, no bias) of a channels-last bf16 CUDA tensor whose weight gradient comes from
    this package's kernel; callers check `ops.conv3d_wgrad_supported` first."""
    return _Conv3x3x3Fn.apply(channels_last(x), weight)


class _Conv1x1x1Fn(torch.autograd.Function):
    """nn.Conv{2,3}d(kernel_size=1, stride=1) on channels-last bf16 x with fp32 (or bf16) master weight [Co, Ci, 1, ...] and
    optional bias: forward, data gradient and weight / bias gradient through ucf_pointwise_conv*."""

    @staticmethod
    def forward(ctx, x, w, b):
        w2 = w.detach().reshape(w.shape[0], w.shape[1]).float().contiguous()
        ctx.save_for_backward(x, w2)
        ctx.w_shape, ctx.w_dtype, ctx.has_bias = w.shape, w.dtype, b is not None
        ctx.b_dtype = b.dtype if b is not None else None
        return ops.pointwise_conv(x, w2, b.detach().float().contiguous() if b is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, w2 = ctx.saved_tensors
        dy = channels_last(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.pointwise_conv(dy, w2.t().contiguous())
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.pointwise_conv_wgrad(x, dy, with_bias=ctx.has_bias)
            dw = dw.reshape(ctx.w_shape).to(ctx.w_dtype)
            db = db.to(ctx.b_dtype) if db is not None else None
        return dx, dw, db


def conv1x1x1(x, weight, bias=None):
    """1x1 convolution (stride 1) of a channels-last bf16 CUDA tensor; callers check `ops.pointwise_conv_supported` first."""
    return _Conv1x1x1Fn.apply(channels_last(x), weight, bias)
