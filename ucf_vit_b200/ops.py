"""Thin tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers,
launch on torch's current stream.  No autograd here (see `functional.py`)."""
import ctypes
import threading

import torch

from . import _lib as L


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw C accessors: torch.cuda.current_stream()
    costs ~20 us of Python per call, which was 10 % of a ViT-B step's host time)."""
    return _raw_stream(_cur_device())


def _require_cuda(*ts):
    """CUDA tensors only, and on the CURRENT device: the launchers run on the current device's stream, so a tensor
    that lives on another GPU would be addressed from the wrong context."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("ucf_vit_b200 ops run on CUDA tensors only (sm_100a); there is no CPU fallback")
        if cur is None:
            cur = _cur_device()
        if t.device.index != cur:
            raise RuntimeError(f"ucf_vit_b200: tensor on cuda:{t.device.index} but the current device is cuda:{cur}; "
                               "wrap the call in `with torch.cuda.device(tensor.device):`")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _dt(t, allow_u8=False):
    if t.dtype == torch.bfloat16:
        return L.UCF_DTYPE_BF16
    if t.dtype == torch.float32:
        return L.UCF_DTYPE_F32
    if allow_u8 and t.dtype == torch.uint8:
        return L.UCF_DTYPE_U8
    raise TypeError(f"unsupported dtype {t.dtype}")


def gemm(a, b, *, M, N, K, a_mn=False, b_mn=False, epilogue=L.EPI_BIAS, bias=None, aux=None, out=None,
         splits=1, tile_n=0, bias_grad=None):
    """C[M,N] = epi(A * B^T).  `a` is stored [M,K] (a_mn=False) or [K,M] (a_mn=True); same for b with N.
    Only the pitch of dim 0 is free; dim 1 must be contiguous."""
    _require_cuda(a, b, bias, aux, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    assert tuple(a.shape) == ((K, M) if a_mn else (M, K)), (a.shape, M, K, a_mn)
    assert tuple(b.shape) == ((K, N) if b_mn else (N, K)), (b.shape, N, K, b_mn)
    if out is None:
        assert epilogue != L.EPI_F32_ADD, "accumulating epilogue needs an explicit output"
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    assert out.stride(1) == 1 and tuple(out.shape) == (M, N)
    assert out.dtype == (torch.float32 if epilogue == L.EPI_F32_ADD else torch.bfloat16)
    if epilogue == L.EPI_BIAS_GELU_AUX and aux is None:
        aux = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and tuple(aux.shape) == (M, N) and aux.stride(1) == 1
    if bias_grad is not None:
        assert bias_grad.dtype == torch.float32 and bias_grad.numel() == M and bias_grad.is_contiguous()
    rc = L.lib().ucf_gemm_bf16(
        a.data_ptr(), b.data_ptr(), out.data_ptr(), _ptr(bias), _ptr(aux), M, N, K,
        a.stride(0), b.stride(0), out.stride(0), aux.stride(0) if aux is not None else 0,
        int(a_mn), int(b_mn), epilogue, _dt(bias) if bias is not None else 0, splits, tile_n, _ptr(bias_grad),
        _stream())
    L.check(rc, "gemm_bf16")
    return (out, aux) if epilogue == L.EPI_BIAS_GELU_AUX else out


def gemm_dgrad_delta(dy, w, o, tokens, heads):
    """dX = dy @ w (w = nn.Linear weight [out, in] read in place) and delta[b, h, n] = sum_d dX * o per head, from the GEMM's
    epilogue.  dy bf16 [M, out], w bf16 [out, in], o bf16 [M, in] -> (dX bf16 [M, in], delta fp32 [M // tokens, heads, tokens])."""
    _require_cuda(dy, w, o)
    M, K = dy.shape
    N = w.shape[1]
    assert w.shape[0] == K and tuple(o.shape) == (M, N) and M % tokens == 0
    assert dy.dtype == w.dtype == o.dtype == torch.bfloat16 and dy.stride(1) == 1 and w.stride(1) == 1 and o.stride(1) == 1
    if not L.lib().ucf_gemm_dgrad_delta_supported(M, N, K, heads):
        raise RuntimeError(f"gemm_dgrad_delta: no fused kernel for M={M} N={N} heads={heads}")
    dx = torch.empty((M, N), dtype=torch.bfloat16, device=dy.device)
    delta = torch.empty((M // tokens, heads, tokens), dtype=torch.float32, device=dy.device)
    L.check(L.lib().ucf_gemm_dgrad_delta(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), o.data_ptr(), delta.data_ptr(), M, N, K,
                                         dy.stride(0), w.stride(0), dx.stride(0), o.stride(0), tokens, heads, _stream()),
            "gemm_dgrad_delta")
    return dx, delta


def layernorm_fwd(x, gamma, beta, eps):
    _require_cuda(x, gamma, beta)
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    assert x2.is_contiguous()
    rows = x2.shape[0]
    y = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    pd = _dt(gamma) if gamma is not None else L.UCF_DTYPE_F32
    rc = L.lib().ucf_layernorm_fwd(x2.data_ptr(), _ptr(gamma), _ptr(beta), y.data_ptr(), mean.data_ptr(),
                                   rstd.data_ptr(), rows, D, float(eps), _dt(x2), pd, _stream())
    L.check(rc, "layernorm_fwd")
    return y.view(*x.shape[:-1], D), mean, rstd


def layernorm_stats(x, eps):
    """(mean, rstd) fp32 [rows] of LayerNorm over the last dim: one read of x, nothing else written."""
    _require_cuda(x)
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    assert x2.is_contiguous()
    rows = x2.shape[0]
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    L.check(L.lib().ucf_layernorm_stats(x2.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, D, float(eps), _dt(x2), _stream()),
            "layernorm_stats")
    return mean, rstd


def ln_gemm_supported(M, N, K):
    return bool(L.lib().ucf_ln_gemm_supported(M, N, K))


def ln_gemm(x, w_gamma, bias_folded, colsum, mean, rstd):
    """y = LayerNorm(x) W^T + b from the RAW rows x: y[r, n] = rstd[r] * (x[r] . Wg[n] - mean[r] * colsum[n]) + bias_folded[n]
    (Wg = W diag(gamma) in bf16, colsum = Wg.sum(1), bias_folded = W beta + b: see functional.fold_layernorm)."""
    _require_cuda(x, w_gamma, bias_folded, colsum, mean, rstd)
    M, K = x.shape
    N = w_gamma.shape[0]
    assert x.dtype == torch.bfloat16 and w_gamma.dtype == torch.bfloat16 and x.stride(1) == 1 and w_gamma.stride(1) == 1
    assert bias_folded.dtype == colsum.dtype == mean.dtype == rstd.dtype == torch.float32
    assert bias_folded.numel() == N and colsum.numel() == N and mean.numel() == M and rstd.numel() == M
    y = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().ucf_ln_gemm(x.data_ptr(), w_gamma.data_ptr(), y.data_ptr(), bias_folded.data_ptr(), colsum.data_ptr(),
                                mean.data_ptr(), rstd.data_ptr(), M, N, K, x.stride(0), w_gamma.stride(0), y.stride(0), _stream()),
            "ln_gemm")
    return y


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, dgamma=None, dbeta=None):
    """Returns dx (bf16).  dgamma/dbeta (fp32 [D]) are accumulated into when given."""
    _require_cuda(dy, x, gamma, dres)
    D = x.shape[-1]
    dy2, x2 = dy.reshape(-1, D), x.reshape(-1, D)
    assert dy2.is_contiguous() and x2.is_contiguous() and dy2.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16
    if dres is not None:
        dres = dres.reshape(-1, D)
        assert dres.is_contiguous() and dres.dtype == torch.bfloat16
    dx = torch.empty_like(x2)
    pd = _dt(gamma) if gamma is not None else L.UCF_DTYPE_F32
    rc = L.lib().ucf_layernorm_bwd(dy2.data_ptr(), x2.data_ptr(), _ptr(gamma), mean.data_ptr(), rstd.data_ptr(),
                                   _ptr(dres), dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), x2.shape[0], D, pd,
                                   _stream())
    L.check(rc, "layernorm_bwd")
    return dx.view(x.shape)


def cast_to_bf16(x, out=None):
    _require_cuda(x)
    if x.dtype == torch.bfloat16:
        return x
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().ucf_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "cast_f32_to_bf16")
    return out


def cast_to_bf16_multi(xs):
    """bf16 copies of up to 8 contiguous fp32 tensors in one launch (views of one allocation)."""
    _require_cuda(*xs)
    n = len(xs)
    assert 1 <= n <= 8 and all(x.dtype == torch.float32 and x.is_contiguous() for x in xs)
    sizes = [-(-x.numel() // 64) * 64 for x in xs]
    flat = torch.empty(sum(sizes), dtype=torch.bfloat16, device=xs[0].device)
    outs, off = [], 0
    for x, sz in zip(xs, sizes):
        outs.append(flat[off:off + x.numel()].view(x.shape))
        off += sz
    srcs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
    dsts = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    cnts = (ctypes.c_longlong * n)(*[x.numel() for x in xs])
    L.check(L.lib().ucf_cast_f32_to_bf16_multi(n, srcs, dsts, cnts, _stream()), "cast_f32_to_bf16_multi")
    return outs


def cast_to_f32(x, out=None, accumulate=False):
    _require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    if out is None:
        assert not accumulate
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    L.check(L.lib().ucf_cast_bf16_to_f32(x.data_ptr(), out.data_ptr(), x.numel(), int(accumulate), _stream()),
            "cast_bf16_to_f32")
    return out


def colsum(x, out=None, accumulate=False):
    """out[n] (+)= sum_m x[m, n]; x bf16 [M, N] (row pitch free), out fp32 [N]."""
    _require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.bfloat16
    M, N = x.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x.device)
        accumulate = False
    L.check(L.lib().ucf_colsum_bf16(x.data_ptr(), out.data_ptr(), M, N, x.stride(0), int(accumulate), _stream()),
            "colsum_bf16")
    return out


def patchify(x, p):
    """x [B,C,H,W] or [B,C,H,W,Z] (fp32 / bf16 / uint8 raw pixels, contiguous) -> bf16 [B*L, K8], K = C*p^dims ordered (c,p0,p1[,p2]) and
    K8 = K rounded up to a multiple of 8 (zero pad columns; the GEMM's K extent).  Like Conv(k = s = p), pixels past
    the last whole patch are ignored."""
    _require_cuda(x)
    assert x.is_contiguous()
    dims = x.dim() - 2
    B, C = x.shape[:2]
    S = list(x.shape[2:])
    G = [s // p for s in S]
    L_ = 1
    for g in G:
        L_ *= g
    K = C * p ** dims
    K8 = -(-K // 8) * 8
    out = torch.empty((B * L_, K8), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().ucf_patchify(x.data_ptr(), out.data_ptr(), B, C, G[0], G[1], G[2] if dims == 3 else 1, p, dims,
                                 S[0], S[1], S[2] if dims == 3 else 1, K8, _dt(x, allow_u8=True), _stream()), "patchify")
    return out


def _bnhd(t):
    assert t.dim() == 4 and t.stride(3) == 1 and t.dtype == torch.bfloat16, (t.shape, t.stride(), t.dtype)
    return t.stride(0), t.stride(1), t.stride(2)


def attention_fwd(q, k, v, scale):
    """q [B,Nq,H,hd], k/v [B,Nk,H,hd] bf16 views (last dim contiguous; e.g. slices of the packed qkv
    projection).  Returns o [B,Nq,H,hd] (contiguous) and lse [B,H,Nq] fp32."""
    _require_cuda(q, k, v)
    B, Nq, H, hd = q.shape
    Nk = k.shape[1]
    o = torch.empty((B, Nq, H, hd), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    rc = L.lib().ucf_attention_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                   B, H, Nq, Nk, hd, *_bnhd(q), *_bnhd(k), *_bnhd(v), *_bnhd(o), float(scale), _stream())
    L.check(rc, "attention_fwd")
    return o, lse


def attention_bwd(q, k, v, o, d_o, lse, scale, dq=None, dk=None, dv=None):
    """Gradients of attention_fwd.  dq/dk/dv may be views (e.g. into one packed dqkv buffer)."""
    _require_cuda(q, k, v, o, d_o)
    B, Nq, H, hd = q.shape
    Nk = k.shape[1]
    assert d_o.is_contiguous() and o.is_contiguous() and d_o.shape == o.shape
    d_o = d_o.view(B, Nq, H, hd)
    if dq is None:
        dq = torch.empty((B, Nq, H, hd), dtype=torch.bfloat16, device=q.device)
    if dk is None:
        dk = torch.empty((B, Nk, H, hd), dtype=torch.bfloat16, device=q.device)
    if dv is None:
        dv = torch.empty((B, Nk, H, hd), dtype=torch.bfloat16, device=q.device)
    # Nq <= 256: dQ is accumulated in tensor memory, no fp32 workspace
    dq_acc = torch.empty((B, Nq, H, hd), dtype=torch.float32, device=q.device) if Nq > 256 else None
    delta = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    rc = L.lib().ucf_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr(),
                                   lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                   dq_acc.data_ptr() if dq_acc is not None else None,
                                   delta.data_ptr(), B, H, Nq, Nk, hd, *_bnhd(q), *_bnhd(k), *_bnhd(v), *_bnhd(o),
                                   *_bnhd(dq), *_bnhd(dk), *_bnhd(dv), float(scale), _stream())
    L.check(rc, "attention_bwd")
    return dq, dk, dv


def var_attention_fwd(q, kv, scale):
    """q [Bq,Na,H,hd] (Bq == rows or 1), kv [rows,V,2,H,hd] -> o [rows,Na,H,hd], lse [rows,Na,H]."""
    _require_cuda(q, kv)
    assert q.is_contiguous() and kv.is_contiguous() and q.dtype == torch.bfloat16 and kv.dtype == torch.bfloat16
    rows, V, _, H, hd = kv.shape
    Bq, Na = q.shape[0], q.shape[1]
    assert Bq in (1, rows)
    shared = int(Bq == 1 and rows != 1) or int(Bq == 1)
    o = torch.empty((rows, Na, H, hd), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((rows, Na, H), dtype=torch.float32, device=q.device)
    L.check(L.lib().ucf_var_attention_fwd(q.data_ptr(), kv.data_ptr(), o.data_ptr(), lse.data_ptr(), rows, Na, V, H, hd,
                                          shared, float(scale), _stream()), "var_attention_fwd")
    return o, lse


def var_attention_bwd(q, kv, o, d_o, lse, scale):
    _require_cuda(q, kv, o, d_o)
    rows, V, _, H, hd = kv.shape
    Bq, Na = q.shape[0], q.shape[1]
    shared = int(Bq == 1)
    d_o = d_o.contiguous()
    dkv = torch.empty_like(kv)
    dq_acc = torch.empty((Bq, Na, H, hd), dtype=torch.float32, device=q.device)
    L.check(L.lib().ucf_var_attention_bwd(q.data_ptr(), kv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                          dkv.data_ptr(), dq_acc.data_ptr(), rows, Na, V, H, hd, shared, float(scale),
                                          _stream()), "var_attention_bwd")
    return dq_acc, dkv


# ---------------------------------------------------------------------------------------------
# adaptive patching (SAP)
# ---------------------------------------------------------------------------------------------
def sap_build_tree(domain, fixed_length, norm_factor=255.0):
    """Greedy quadtree (2-D numpy array) / octree (3-D cubic array) on the HOST (integer work, no GPU
    needed).  Returns (boxes int32 [n, 4|6], values int64 [n]) in the reference's list order."""
    import numpy as np
    d = np.ascontiguousarray(domain)
    if d.dtype == np.uint8:
        dt = L.UCF_DTYPE_U8
    elif d.dtype == np.float32:
        dt = L.UCF_DTYPE_F32
    else:
        d = np.ascontiguousarray(d, dtype=np.float64)
        dt = L.UCF_DTYPE_F64
    nd = d.ndim
    assert nd in (2, 3)
    nc = 4 if nd == 2 else 6
    rows = fixed_length + (3 if nd == 2 else 7) - 1          # UCF_SAP_TREE_ROWS: the last split may overshoot
    boxes = np.zeros((rows, nc), dtype=np.int32)
    values = np.zeros((rows,), dtype=np.int64)
    shp = list(d.shape) + [0] * (3 - nd)
    n = L.lib().ucf_sap_build_tree_host(d.ctypes.data, dt, nd, shp[0], shp[1], shp[2], float(norm_factor), fixed_length,
                                        boxes.ctypes.data, values.ctypes.data)
    if n < 0:
        L.check(n, "sap_build_tree_host")
    return boxes[:n].copy(), values[:n].copy()


def sap_build_trees(domains, fixed_length, norm_factor=255.0, threads=0):
    """Trees of several same-shape edge maps built on host threads (`threads` <= 0: one per core).
    Returns a list of (boxes int32 [n_i, 4|6], values int64 [n_i]) like `sap_build_tree`."""
    import numpy as np
    ds = [np.ascontiguousarray(d) for d in domains]
    if not ds:
        return []
    kind = ds[0].dtype
    if kind not in (np.uint8, np.float32):
        kind = np.dtype(np.float64)
    ds = [np.ascontiguousarray(d, dtype=kind) for d in ds]
    if any(d.shape != ds[0].shape for d in ds):
        raise ValueError("sap_build_trees: every edge map of a batch must have the same shape")
    dt = {np.dtype(np.uint8): L.UCF_DTYPE_U8, np.dtype(np.float32): L.UCF_DTYPE_F32,
          np.dtype(np.float64): L.UCF_DTYPE_F64}[np.dtype(kind)]
    nd = ds[0].ndim
    assert nd in (2, 3)
    nc = 4 if nd == 2 else 6
    n = len(ds)
    rows = fixed_length + (3 if nd == 2 else 7) - 1          # UCF_SAP_TREE_ROWS
    boxes = np.zeros((n, rows, nc), dtype=np.int32)
    values = np.zeros((n, rows), dtype=np.int64)
    counts = np.zeros((n,), dtype=np.int32)
    ptrs = (ctypes.c_void_p * n)(*[d.ctypes.data for d in ds])
    shp = list(ds[0].shape) + [0] * (3 - nd)
    rc = L.lib().ucf_sap_build_tree_batch_host(ptrs, n, dt, nd, shp[0], shp[1], shp[2], float(norm_factor), fixed_length,
                                               boxes.ctypes.data, values.ctypes.data, counts.ctypes.data, int(threads))
    L.check(rc, "sap_build_tree_batch_host")
    return [(boxes[i, :counts[i]].copy(), values[i, :counts[i]].copy()) for i in range(n)]


def sap_gather(img, boxes, fixed_length, p):
    """img: cuda uint8/float32 [H,W,C] or float32 [Z,Y,X,C]; boxes: cuda int32 [n,4|6].
    -> seq f32 [L,p,p(,p),C], seq_size int64 [L], seq_pos f64 [L,2|3] (all on the device)."""
    _require_cuda(img, boxes)
    assert img.is_contiguous() and boxes.is_contiguous() and boxes.dtype == torch.int32
    nd = img.dim() - 1
    C = img.shape[-1]
    n = boxes.shape[0]
    dt = L.UCF_DTYPE_U8 if img.dtype == torch.uint8 else L.UCF_DTYPE_F32
    assert img.dtype in (torch.uint8, torch.float32)
    seq = torch.empty((fixed_length,) + (p,) * nd + (C,), dtype=torch.float32, device=img.device)
    size = torch.empty((fixed_length,), dtype=torch.int64, device=img.device)
    pos = torch.empty((fixed_length, nd), dtype=torch.float64, device=img.device)
    shp = list(img.shape[:-1]) + [0] * (3 - nd)
    L.check(L.lib().ucf_sap_gather(img.data_ptr(), dt, nd, shp[0], shp[1], shp[2], C, boxes.data_ptr(), n, fixed_length, p,
                                   seq.data_ptr(), size.data_ptr(), pos.data_ptr(), _stream()), "sap_gather")
    return seq, size, pos


def sap_scatter(seq, boxes, out_shape, p, C, truncate_to_int=False):
    """seq f32 [>=n, p, p(, p), C] -> mask f32 [H,W,C] / [Z,Y,X,C] (zero where no leaf writes)."""
    _require_cuda(seq, boxes)
    nd = len(out_shape)
    seq = seq.contiguous().float()
    mask = torch.zeros(tuple(out_shape) + (C,), dtype=torch.float32, device=seq.device)
    shp = list(out_shape) + [0] * (3 - nd)
    L.check(L.lib().ucf_sap_scatter(seq.data_ptr(), nd, shp[0], shp[1], shp[2], C, boxes.data_ptr(), boxes.shape[0], p,
                                    int(truncate_to_int), mask.data_ptr(), _stream()), "sap_scatter")
    return mask


def assemble_tokens(tok, prefix, pos, pos_has_prefix):
    """tok bf16 [B,L,D]; prefix [P,D] or None; pos [N or L, D] (shared) / [B, N or L, D] (per sample) or None.
    -> bf16 [B, P+L, D] = concat(prefix, tok) + pos."""
    _require_cuda(tok, prefix, pos)
    B, L_, D = tok.shape
    assert tok.dtype == torch.bfloat16 and tok.is_contiguous()
    P = 0 if prefix is None else prefix.shape[0]
    pd = L.UCF_DTYPE_F32
    for t in (prefix, pos):
        if t is not None:
            assert t.is_contiguous()
            pd = _dt(t)
    if prefix is not None and pos is not None:
        assert prefix.dtype == pos.dtype
    bstride = 0
    if pos is not None and pos.dim() == 3 and pos.shape[0] != 1:
        assert pos.shape[0] == B
        bstride = pos.shape[1] * D
    out = torch.empty((B, P + L_, D), dtype=torch.bfloat16, device=tok.device)
    L.check(L.lib().ucf_assemble_tokens(tok.data_ptr(), _ptr(prefix), _ptr(pos), out.data_ptr(), B, L_, P, D, bstride,
                                        0 if pos_has_prefix else P, pd, _stream()), "assemble_tokens")
    return out


def _patch_geometry(pred, img, grid, patch):
    """Validate pred [B, L, P*C] against img [B, C, G0*p0, G1*p1, G2*p2]; returns the int arguments."""
    G0, G1, G2 = grid
    p0, p1, p2 = patch
    B, C = img.shape[0], img.shape[1]
    if img.numel() != B * C * G0 * p0 * G1 * p1 * G2 * p2:
        raise ValueError(f"image {tuple(img.shape)} is not [B, C] x grid {grid} x patch {patch}")
    if tuple(pred.shape) != (B, G0 * G1 * G2, p0 * p1 * p2 * C):
        raise ValueError(f"pred {tuple(pred.shape)} does not match {(B, G0 * G1 * G2, p0 * p1 * p2 * C)}")
    return B, C, G0, G1, G2, p0, p1, p2


def patch_mse_fwd(pred, img, grid, patch, mask=None):
    """Loss of `pred` against the patchified `img` (never materialised).  Returns a fp32 [2] tensor:
    [loss, 1 / denominator]; the second entry is what `patch_mse_bwd` needs."""
    _require_cuda(pred, img, mask)
    assert pred.is_contiguous() and img.is_contiguous()
    dims = _patch_geometry(pred, img, grid, patch)
    if mask is not None:
        assert mask.dtype == torch.float32 and mask.is_contiguous() and mask.numel() == pred.shape[0] * pred.shape[1]
    ws = torch.empty(2 * L.PATCH_MSE_MAX_BLOCKS, dtype=torch.float64, device=pred.device)
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    L.check(L.lib().ucf_patch_mse_fwd(pred.data_ptr(), _dt(pred), img.data_ptr(), _dt(img), _ptr(mask), *dims,
                                      ws.data_ptr(), out.data_ptr(), _stream()), "patch_mse_fwd")
    return out


def patch_mse_bwd(pred, img, grid, patch, mask, fwd_out, grad_out):
    """d loss / d pred scaled by the device scalar `grad_out`; same dtype / shape as pred."""
    _require_cuda(pred, img, mask, fwd_out, grad_out)
    dims = _patch_geometry(pred, img, grid, patch)
    assert fwd_out.dtype == torch.float32 and fwd_out.numel() == 2 and fwd_out.is_contiguous()
    assert grad_out.dtype == torch.float32 and grad_out.numel() == 1
    dpred = torch.empty_like(pred)
    L.check(L.lib().ucf_patch_mse_bwd(pred.data_ptr(), _dt(pred), img.data_ptr(), _dt(img), _ptr(mask),
                                      fwd_out.data_ptr(), grad_out.data_ptr(), *dims, dpred.data_ptr(), _stream()),
            "patch_mse_bwd")
    return dpred


def _is_dense(t):
    """Non-overlapping and dense in some dimension order (e.g. channels_last): an elementwise kernel may walk the storage."""
    n = 1
    for size, stride in sorted(zip(t.shape, t.stride()), key=lambda ss: ss[1]):
        if size == 1:
            continue
        if stride != n:
            return False
        n *= size
    return True


def adamw_multi(params, grads, exp_avgs, exp_avg_sqs, *, lr, beta1, beta2, eps, weight_decay, step, maximize=False):
    """One in-place AdamW update of the fp32 tensors `params` (+ both moments) that share a step count.
    `lr` and `step` are host numbers, or BOTH fp32 CUDA scalars (the CUDA-graph capturable form)."""
    n = len(params)
    if n == 0:
        return
    assert len(grads) == n and len(exp_avgs) == n and len(exp_avg_sqs) == n
    for ts in (params, grads, exp_avgs, exp_avg_sqs):
        _require_cuda(*ts)
        for t, p in zip(ts, params):
            if t.dtype != torch.float32 or not ((t.is_contiguous() and p.is_contiguous()) or
                                                (t.stride() == p.stride() and _is_dense(t))):
                raise TypeError("adamw_multi: every tensor must be a dense fp32 CUDA tensor laid out like its parameter")
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        if not (p.numel() == g.numel() == m.numel() == v.numel()):
            raise ValueError("adamw_multi: parameter, gradient and moments differ in size")
    tbl = [(ctypes.c_void_p * n)(*[t.data_ptr() for t in ts]) for ts in (params, grads, exp_avgs, exp_avg_sqs)]
    cnts = (ctypes.c_longlong * n)(*[p.numel() for p in params])
    if torch.is_tensor(lr) or (torch.is_tensor(step) and step.is_cuda):
        if not (torch.is_tensor(lr) and torch.is_tensor(step) and lr.is_cuda and step.is_cuda and
                lr.dtype == torch.float32 and step.dtype == torch.float32 and lr.numel() == 1 and step.numel() == 1):
            raise TypeError("adamw_multi: device-resident lr and step must both be fp32 CUDA scalars")
        L.check(L.lib().ucf_adamw_multi_dev(n, *tbl, cnts, lr.data_ptr(), float(beta1), float(beta2), float(eps),
                                            float(weight_decay), step.data_ptr(), int(bool(maximize)), _stream()),
                "adamw_multi_dev")
        return
    L.check(L.lib().ucf_adamw_multi(n, *tbl, cnts, float(lr), float(beta1), float(beta2), float(eps),
                                    float(weight_decay), int(step), int(bool(maximize)), _stream()), "adamw_multi")


def mask_plan(noise, len_keep):
    """(ids_shuffle, ids_restore, mask) of MAE.random_masking from the per-token noise [B, L] (fp32)."""
    _require_cuda(noise)
    assert noise.dim() == 2 and noise.dtype == torch.float32 and noise.is_contiguous()
    B, Lt = noise.shape
    ids_shuffle = torch.empty((B, Lt), dtype=torch.int64, device=noise.device)
    ids_restore = torch.empty_like(ids_shuffle)
    mask = torch.empty((B, Lt), dtype=torch.float32, device=noise.device)
    L.check(L.lib().ucf_mask_plan(noise.data_ptr(), B, Lt, int(len_keep), ids_shuffle.data_ptr(), ids_restore.data_ptr(),
                                  mask.data_ptr(), _stream()), "mask_plan")
    return ids_shuffle, ids_restore, mask


def gather_tokens(src, idx, fill=None, pos=None):
    """out[b, i] = (idx[b, i] < Ls ? src[b, idx[b, i]] : fill) + pos[b or 0, i];  src bf16 [B, Ls, D], idx int64 [B, Lo]."""
    _require_cuda(src, idx, fill, pos)
    assert src.dtype == torch.bfloat16 and src.dim() == 3 and src.is_contiguous()
    assert idx.dtype == torch.int64 and idx.dim() == 2 and idx.is_contiguous() and idx.shape[0] == src.shape[0]
    B, Ls, D = src.shape
    Lo = idx.shape[1]
    pdt, bstride = 0, 0
    if fill is not None:
        assert fill.numel() == D and fill.is_contiguous()
        pdt = _dt(fill)
    if pos is not None:
        assert pos.is_contiguous() and pos.shape[-2:] == (Lo, D) and pos.numel() in (Lo * D, B * Lo * D)
        assert fill is None or pos.dtype == fill.dtype
        pdt = _dt(pos)
        bstride = Lo * D if (pos.numel() == B * Lo * D and B > 1) else 0
    out = torch.empty((B, Lo, D), dtype=torch.bfloat16, device=src.device)
    L.check(L.lib().ucf_gather_tokens(src.data_ptr(), idx.data_ptr(), _ptr(fill), _ptr(pos), out.data_ptr(), B, Ls, Lo, D,
                                      bstride, pdt, _stream()), "gather_tokens")
    return out


def scatter_tokens(dout, idx, Ls, *, need_src=True, need_fill=False, zero_first=True):
    """Gradient of gather_tokens: (dsrc bf16 [B, Ls, D] | None, dfill fp32 [D] | None)."""
    _require_cuda(dout, idx)
    assert dout.dtype == torch.bfloat16 and dout.dim() == 3 and dout.is_contiguous()
    assert idx.dtype == torch.int64 and idx.is_contiguous() and tuple(idx.shape) == tuple(dout.shape[:2])
    B, Lo, D = dout.shape
    dsrc = torch.empty((B, Ls, D), dtype=torch.bfloat16, device=dout.device) if need_src else None
    dfill = torch.zeros(D, dtype=torch.float32, device=dout.device) if need_fill else None
    L.check(L.lib().ucf_scatter_tokens(dout.data_ptr(), idx.data_ptr(), _ptr(dsrc), _ptr(dfill), B, Ls, Lo, D,
                                       int(zero_first), _stream()), "scatter_tokens")
    return dsrc, dfill


def add_bcast(x, e):
    """x bf16 [..., D] (<= 4 dims) + e (f32 | bf16, broadcastable to x, same last dim) -> bf16, one pass."""
    _require_cuda(x, e)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and 1 <= x.dim() <= 4
    assert e.is_contiguous() and e.shape[-1] == x.shape[-1] and e.dim() <= x.dim()
    D = x.shape[-1]
    shape = [1] * (4 - x.dim()) + list(x.shape)
    ev = e.reshape([1] * (4 - e.dim()) + list(e.shape))
    st = ev.expand(shape).stride()
    assert st[3] == 1 or D == 1
    out = torch.empty_like(x)
    L.check(L.lib().ucf_add_bcast(x.data_ptr(), e.data_ptr(), out.data_ptr(), shape[0], shape[1], shape[2], D,
                                  st[0], st[1], st[2], _dt(e), _stream()), "add_bcast")
    return out


def _dice_dims(logits, targets):
    if logits.dim() < 3 or logits.shape != targets.shape or logits.shape[1] < 2:
        raise ValueError(f"dice_bce: logits and targets must share a [B, C >= 2, ...] shape, got {tuple(logits.shape)} "
                         f"and {tuple(targets.shape)}")
    B, C = logits.shape[:2]
    return B, C, logits[0, 0].numel()


def dice_bce_fwd(logits, targets, weight, smooth, act):
    """fp32 [4] = (loss, 2I + smooth, denominator, 1/n) of DiceBLoss over channels 1.. of [B, C, ...] tensors."""
    _require_cuda(logits, targets)
    assert logits.is_contiguous() and targets.is_contiguous()
    B, C, HW = _dice_dims(logits, targets)
    ws = torch.empty(4 * L.PATCH_MSE_MAX_BLOCKS, dtype=torch.float64, device=logits.device)
    out = torch.empty(4, dtype=torch.float32, device=logits.device)
    L.check(L.lib().ucf_dice_bce_fwd(logits.data_ptr(), _dt(logits), targets.data_ptr(), _dt(targets), B, C, HW,
                                     float(weight), float(smooth), int(bool(act)), ws.data_ptr(), out.data_ptr(), _stream()),
            "dice_bce_fwd")
    return out


def dice_bce_bwd(logits, targets, fwd_out, grad_out, weight, act):
    _require_cuda(logits, targets, fwd_out, grad_out)
    B, C, HW = _dice_dims(logits, targets)
    assert fwd_out.dtype == torch.float32 and fwd_out.numel() == 4 and grad_out.dtype == torch.float32 and grad_out.numel() == 1
    dlogits = torch.empty_like(logits)
    L.check(L.lib().ucf_dice_bce_bwd(logits.data_ptr(), _dt(logits), targets.data_ptr(), _dt(targets), fwd_out.data_ptr(),
                                     grad_out.data_ptr(), B, C, HW, float(weight), int(bool(act)), dlogits.data_ptr(),
                                     _stream()), "dice_bce_bwd")
    return dlogits


def _class_dtype(t):
    if t.dtype == torch.uint8:
        return L.UCF_DTYPE_U8
    if t.dtype == torch.int64:
        return L.UCF_DTYPE_I64
    if t.dtype == torch.float32:
        return L.UCF_DTYPE_F32
    raise TypeError(f"dice_ce: class indices must be uint8, int64 or float32, got {t.dtype}")


def _dice_layout(logits):
    """0: [B, C, S] contiguous; 1: channels-last memory ([B, S, C]).  Anything else is refused (callers make it contiguous)."""
    if logits.is_contiguous():
        return 0
    if logits.dim() >= 3 and logits.movedim(1, -1).is_contiguous():
        return 1
    raise TypeError("dice_ce: logits must be contiguous or channels-last")


def dice_ce_fwd(logits, target, squared_pred, smooth_nr, smooth_dr, lambda_dice, lambda_ce):
    """fp32 [2 + 2 B C]: [0] = DiceCE loss of logits [B, C, ...] against class indices target [B, ...]; the rest feeds dice_ce_bwd."""
    _require_cuda(logits, target)
    assert target.is_contiguous()
    cl = _dice_layout(logits)
    B, C = logits.shape[:2]
    S = logits[0, 0].numel()
    assert target.numel() == B * S, (tuple(logits.shape), tuple(target.shape))
    nb = L.lib().ucf_dice_ce_blocks_per_sample(B, S)
    ws = torch.empty(B * nb * 25, dtype=torch.float64, device=logits.device)
    out = torch.empty(2 + 2 * B * C, dtype=torch.float32, device=logits.device)
    L.check(L.lib().ucf_dice_ce_fwd(logits.data_ptr(), _dt(logits), target.data_ptr(), _class_dtype(target), B, C, S,
                                    int(bool(squared_pred)), float(smooth_nr), float(smooth_dr), float(lambda_dice),
                                    float(lambda_ce), ws.data_ptr(), out.data_ptr(), cl, _stream()), "dice_ce_fwd")
    return out


def dice_ce_bwd(logits, target, fwd_out, grad_out, squared_pred):
    _require_cuda(logits, target, fwd_out, grad_out)
    cl = _dice_layout(logits)
    B, C = logits.shape[:2]
    S = logits[0, 0].numel()
    assert grad_out.dtype == torch.float32 and grad_out.numel() == 1
    dlogits = torch.empty_like(logits, memory_format=torch.preserve_format)
    L.check(L.lib().ucf_dice_ce_bwd(logits.data_ptr(), _dt(logits), target.data_ptr(), _class_dtype(target), fwd_out.data_ptr(),
                                    grad_out.data_ptr(), B, C, S, int(bool(squared_pred)), dlogits.data_ptr(), cl, _stream()),
            "dice_ce_bwd")
    return dlogits


# ---- UNETR decoder: InstanceNorm (+ residual) + LeakyReLU on channels-last bf16 ---------------------------------------------
def _nsc(t):
    """(N, S, C) of a [N, C, *spatial] bf16 tensor whose memory is [N, *spatial, C] (channels_last / channels_last_3d)."""
    if t.dtype != torch.bfloat16 or t.dim() < 3:
        raise TypeError(f"instance norm kernels take bf16 [N, C, ...] tensors, got {t.dtype} {tuple(t.shape)}")
    if not t.movedim(1, -1).is_contiguous():
        raise TypeError("instance norm kernels take channels-last tensors (memory [N, ..., C]); "
                        f"got shape {tuple(t.shape)} strides {t.stride()}")
    N, C = t.shape[:2]
    return N, t[0, 0].numel(), C


def inorm_stats(x, eps=1e-5):
    """fp32 [N, 2, C]: per-(sample, channel) mean and 1/sqrt(biased var + eps) of channels-last bf16 x [N, C, ...]."""
    _require_cuda(x)
    N, S, C = _nsc(x)
    ws = torch.empty(N * L.lib().ucf_inorm_chunks(N, S, C) * 2 * C, dtype=torch.float32, device=x.device)
    stats = torch.empty(N, 2, C, dtype=torch.float32, device=x.device)
    L.check(L.lib().ucf_inorm_stats(x.data_ptr(), N, S, C, float(eps), ws.data_ptr(), stats.data_ptr(), _stream()), "inorm_stats")
    return stats


def inorm_apply(a, stats_a, b=None, stats_b=None, slope=0.01):
    """lrelu_slope(IN(a) [+ IN(b) | + b]) in a's layout; slope 1 = no activation."""
    _require_cuda(a, stats_a)
    N, S, C = _nsc(a)
    if b is not None:
        _require_cuda(b)
        assert _nsc(b) == (N, S, C), "inorm_apply: operands differ in shape"
    assert stats_b is None or b is not None
    y = torch.empty_like(a, memory_format=torch.preserve_format)
    L.check(L.lib().ucf_inorm_apply(a.data_ptr(), stats_a.data_ptr(), b.data_ptr() if b is not None else None,
                                    stats_b.data_ptr() if stats_b is not None else None, y.data_ptr(), N, S, C, float(slope),
                                    _stream()), "inorm_apply")
    return y


def inorm_bwd(dy, y, a, stats_a, b=None, stats_b=None, slope=0.01):
    """(da, db) of inorm_apply; db is None when there was no second operand."""
    _require_cuda(dy, a, stats_a)
    N, S, C = _nsc(a)
    assert _nsc(dy) == (N, S, C) and (y is None or _nsc(y) == (N, S, C))
    assert y is not None or slope == 1.0
    da = torch.empty_like(a, memory_format=torch.preserve_format)
    db = torch.empty_like(a, memory_format=torch.preserve_format) if b is not None else None
    ws = torch.empty(N * L.lib().ucf_inorm_chunks(N, S, C) * 3 * C, dtype=torch.float32, device=a.device)
    coef = torch.empty(N, 3, C, dtype=torch.float32, device=a.device)
    ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
    L.check(L.lib().ucf_inorm_bwd(dy.data_ptr(), ptr(y), a.data_ptr(), stats_a.data_ptr(), ptr(b), ptr(stats_b), da.data_ptr(),
                                  ptr(db), N, S, C, float(slope), ws.data_ptr(), coef.data_ptr(), _stream()), "inorm_bwd")
    return da, db


# ---- SAP front end: Gaussian blur + Canny of a uint8 image on the device (bit-exact with OpenCV) ---------------------------
def _hwc_u8(img):
    _require_cuda(img)
    if img.dtype != torch.uint8 or img.dim() not in (2, 3) or not img.is_contiguous():
        raise TypeError(f"expected a contiguous uint8 [H, W] or [H, W, C] CUDA image, got {img.dtype} {tuple(img.shape)}")
    H, W = img.shape[:2]
    return H, W, (img.shape[2] if img.dim() == 3 else 1)


def gaussian_blur_u8(img, ksize):
    """cv.GaussianBlur(img, (ksize, ksize), 0) of a uint8 image, ksize in {1, 3, 5}."""
    H, W, C = _hwc_u8(img)
    out = torch.empty_like(img)
    L.check(L.lib().ucf_gaussian_blur_u8(img.data_ptr(), out.data_ptr(), H, W, C, int(ksize), _stream()), "gaussian_blur_u8")
    return out


_canny_tls = threading.local()      # per host thread: {device index: (device flag, pinned host flag)}


def canny_u8(img, low, high, return_sweeps=False):
    """cv.Canny(img, low, high) (aperture 3, L1 norm) of a uint8 [H, W(, C)] image: uint8 [H, W] of 0 / 255.
    Synchronises the current stream (the hysteresis loop reads a convergence flag back)."""
    H, W, C = _hwc_u8(img)
    flags = _canny_tls.__dict__.setdefault("flags", {})
    if img.device.index not in flags:
        flags[img.device.index] = (torch.zeros(1, dtype=torch.int32, device=img.device),
                                   torch.zeros(1, dtype=torch.int32).pin_memory())
    fdev, fhost = flags[img.device.index]
    cmap = torch.empty(H, W, dtype=torch.uint8, device=img.device)
    edges = torch.empty(H, W, dtype=torch.uint8, device=img.device)
    sweeps = ctypes.c_int(0)
    L.check(L.lib().ucf_canny_u8(img.data_ptr(), H, W, C, float(low), float(high), cmap.data_ptr(), edges.data_ptr(),
                                 fdev.data_ptr(), fhost.data_ptr(), ctypes.addressof(sweeps), _stream()), "canny_u8")
    return (edges, sweeps.value) if return_sweeps else edges


# ---- UNETR decoder: weight gradient of the 3x3x3 convolutions (channels-last bf16) ------------------------------------------
def conv3d_wgrad_supported(Ci, Co, D, H, W):
    return bool(L.lib().ucf_conv3d_wgrad_supported(int(Ci), int(Co), int(D), int(H), int(W)))


def conv3d_wgrad(x, dy):
    """fp32 [Co, Ci, 3, 3, 3]: weight gradient of conv3d(x, w, stride 1, padding 1) given dy; x [N, Ci, D, H, W] and
    dy [N, Co, D, H, W] are channels-last bf16."""
    _require_cuda(x, dy)
    N, S, Ci = _nsc(x)
    N2, S2, Co = _nsc(dy)
    assert x.dim() == 5 and dy.dim() == 5 and N == N2 and x.shape[2:] == dy.shape[2:], (tuple(x.shape), tuple(dy.shape))
    D, H, W = x.shape[2:]
    ctas = L.lib().ucf_conv3d_wgrad_ctas(N, D, H, W)
    ws = torch.empty(ctas * 27 * Co * Ci, dtype=torch.float32, device=x.device)
    dw = torch.empty(Co, Ci, 3, 3, 3, dtype=torch.float32, device=x.device)
    L.check(L.lib().ucf_conv3d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, D, H, W, Ci, Co, ws.data_ptr(), _stream()),
            "conv3d_wgrad")
    return dw


# ---- UNETR decoder: 1x1x1 convolutions (channels-last bf16) -----------------------------------------------------------------
def pointwise_conv_supported(Ci, Co):
    return bool(L.lib().ucf_pointwise_conv_supported(int(Ci), int(Co)))


def pointwise_conv(x, w, bias=None):
    """y [N, Co, ...] (channels-last bf16) = 1x1 convolution of channels-last bf16 x [N, Ci, ...] with fp32 w [Co, Ci]."""
    _require_cuda(x, w)
    N, S, Ci = _nsc(x)
    Co = w.shape[0]
    assert w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape) == (Co, Ci)
    assert bias is None or (bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == Co)
    y = torch.empty((N, *x.shape[2:], Co), dtype=torch.bfloat16, device=x.device).movedim(-1, 1)
    L.check(L.lib().ucf_pointwise_conv(x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                       N * S, Ci, Co, _stream()), "pointwise_conv")
    return y


def pointwise_conv_wgrad(x, dy, with_bias=False):
    """(dw fp32 [Co, Ci], dbias fp32 [Co] | None) of pointwise_conv for channels-last bf16 x [N, Ci, ...], dy [N, Co, ...]."""
    _require_cuda(x, dy)
    N, S, Ci = _nsc(x)
    N2, S2, Co = _nsc(dy)
    assert (N, S) == (N2, S2)
    V = N * S
    ws = torch.empty(L.lib().ucf_pointwise_conv_ctas(V) * (Co * Ci + Co), dtype=torch.float32, device=x.device)
    dw = torch.empty(Co, Ci, dtype=torch.float32, device=x.device)
    db = torch.empty(Co, dtype=torch.float32, device=x.device) if with_bias else None
    L.check(L.lib().ucf_pointwise_conv_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr() if with_bias else None, V,
                                             Ci, Co, ws.data_ptr(), _stream()), "pointwise_conv_wgrad")
    return dw, db
