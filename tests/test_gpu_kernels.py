"""GPU (-m gpu): kernel-level parity through the C ABI (ctypes) against fp32 torch math.
Tolerances: bf16 outputs 1e-2 * max|ref| (one bf16 rounding of an fp32-accumulated result plus
bf16 inputs), fp32 outputs 2e-3, exact for pure data movement."""
import pytest
import torch

from ucf_vit_b200 import _lib as L
from ucf_vit_b200 import ops

pytestmark = pytest.mark.gpu
dev = "cuda"


def bf(x):
    return x.to(torch.bfloat16)


def _ok(got, ref, tol):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= tol * ref.abs().max().item() + 1e-5, (err, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 192, 192), (1000, 576, 192), (333, 1000, 768), (77, 40, 72),
                                   (4096, 3072, 768)])
@pytest.mark.parametrize("tn", [128, 256, 512])      # 512 = CTA-pair (cta_group::2) kernel
def test_gemm_family(M, N, K, tn):
    torch.manual_seed(M + N + K)
    a, b = bf(torch.randn(M, K, device=dev) * 0.5), bf(torch.randn(N, K, device=dev) * 0.5)
    bias, res = torch.randn(N, device=dev), bf(torch.randn(M, N, device=dev))
    ref = a.float() @ b.float().t()
    _ok(ops.gemm(a, b, M=M, N=N, K=K, bias=bias, tile_n=tn), ref + bias, 1e-2)
    _ok(ops.gemm(a, b, M=M, N=N, K=K, bias=bf(bias), tile_n=tn), ref + bf(bias).float(), 1e-2)
    _ok(ops.gemm(a, b, M=M, N=N, K=K, bias=bias, aux=res, epilogue=L.EPI_BIAS_RESIDUAL, tile_n=tn), ref + bias + res.float(), 1e-2)
    u, z = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_GELU_AUX, tile_n=tn)
    _ok(z, ref + bias, 1e-2)
    _ok(u, torch.nn.functional.gelu(z.float()), 1e-2)
    dy = bf(torch.randn(M, N, device=dev) * 0.5)
    dref = dy.float() @ b.float()
    _ok(ops.gemm(dy, b, M=M, N=K, K=N, b_mn=True, tile_n=tn), dref, 1e-2)
    zz = bf(torch.randn(M, K, device=dev)).float().requires_grad_(True)
    g = torch.autograd.grad(torch.nn.functional.gelu(zz).sum(), zz)[0]
    _ok(ops.gemm(dy, b, M=M, N=K, K=N, b_mn=True, aux=bf(zz.detach()), epilogue=L.EPI_DGELU, tile_n=tn), dref * g, 1e-2)
    wref = dy.float().t() @ a.float()
    for splits in (1, 4):
        dw = torch.ones(N, K, device=dev)
        dbias = torch.full((N,), 2.0, device=dev)        # fused bias gradient: += colsum(dy)
        ops.gemm(dy, a, M=N, N=K, K=M, a_mn=True, b_mn=True, epilogue=L.EPI_F32_ADD, out=dw, splits=splits, tile_n=tn,
                 bias_grad=dbias)
        _ok(dw, wref + 1.0, 2e-3)
        _ok(dbias, dy.float().sum(0) + 2.0, 2e-3)


def test_gemm_linearity_at_benchmark_size():
    """ViT-B/16 batch-256 sized GEMM (M = 50432): size-independent property instead of a CPU
    reference -- (a1 + a2) W == a1 W + a2 W for operands exactly representable in bf16."""
    M, N, K = 50432, 768, 768
    a1 = torch.randint(-4, 5, (M, K), device=dev).to(torch.bfloat16)
    a2 = torch.randint(-4, 5, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-2, 3, (N, K), device=dev).to(torch.bfloat16)
    y12 = ops.gemm(a1 + a2, w, M=M, N=N, K=K).float()
    y1, y2 = ops.gemm(a1, w, M=M, N=N, K=K).float(), ops.gemm(a2, w, M=M, N=N, K=K).float()
    # small-integer operands: every product and partial sum is exact in fp32, so the only error is
    # the final bf16 rounding of each of the three outputs (half an ulp = 2^-9 relative, each)
    assert ((y12 - (y1 + y2)).abs() <= 2.0 ** -8 * (y1.abs() + y2.abs() + y12.abs()) + 1e-3).all()
    idx = torch.randint(0, M, (64,), device=dev)
    ref = (a1[idx].float() @ w.float().t())
    assert ((y1[idx] - ref).abs() <= 0.005 * ref.abs() + 1e-3).all()


@pytest.mark.parametrize("rows,D", [(7, 64), (1000, 192), (3000, 1024), (513, 512), (100, 2048), (50432, 768), (333, 4096),
                                     (40, 3072)])
def test_layernorm_fwd_bwd(rows, D):
    torch.manual_seed(rows)
    x = torch.randn(rows, D, device=dev) * 2 + 0.5
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    for xx in (x, bf(x)):
        y, mean, rstd = ops.layernorm_fwd(xx, g, b, 1e-6)
        _ok(y, torch.nn.functional.layer_norm(xx.float(), (D,), g, b, 1e-6), 1e-2)
        _ok(mean, xx.float().mean(-1), 1e-4)
    xb = bf(x)
    y, mean, rstd = ops.layernorm_fwd(xb, g, b, 1e-6)
    dy, dres = bf(torch.randn(rows, D, device=dev)), bf(torch.randn(rows, D, device=dev))
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    dx = ops.layernorm_bwd(dy, xb, g, mean, rstd, dres=dres, dgamma=dg, dbeta=db)
    xf, gf, bfp = xb.float().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xf, (D,), gf, bfp, 1e-6).backward(dy.float())
    _ok(dx, xf.grad + dres.float(), 1e-2)
    _ok(dg, gf.grad, 2e-3)
    _ok(db, bfp.grad, 2e-3)


def _ref_attn(q, k, v, scale):
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    return torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1), v), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,N,H,hd", [(1, 128, 1, 64), (2, 197, 3, 64), (2, 50, 2, 64), (1, 512, 2, 64), (1, 1000, 1, 64),
                                      (2, 197, 2, 32), (1, 384, 2, 32), (3, 1, 2, 64),
                                      # Nq <= 128: the two warpgroups take two consecutive (b, h); odd and large item counts
                                      (3, 49, 3, 64), (64, 49, 16, 64), (5, 128, 7, 32), (1, 100, 1, 64), (37, 33, 5, 32)])
def test_attention_fwd_bwd(B, N, H, hd):
    torch.manual_seed(N)
    scale = hd ** -0.5
    qkv = bf(torch.randn(B, N, 3, H, hd, device=dev))
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, scale)
    qf, kf, vf = [t.float().detach().requires_grad_(True) for t in (q, k, v)]
    oref, lref = _ref_attn(qf, kf, vf, scale)
    _ok(o, oref, 1.5e-2)
    _ok(lse, lref, 1e-3)
    do = bf(torch.randn(B, N, H, hd, device=dev) * 0.5)
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(q, k, v, o, do, lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
    oref.backward(do.float())
    _ok(dqkv[:, :, 0], qf.grad, 2e-2)
    _ok(dqkv[:, :, 1], kf.grad, 2e-2)
    _ok(dqkv[:, :, 2], vf.grad, 2e-2)


@pytest.mark.parametrize("B,N,H,hd", [(2, 197, 2, 64), (1, 512, 2, 64), (2, 700, 2, 32), (1, 1100, 1, 64)])
def test_attention_fwd_stabiliser_jumps(B, N, H, hd):
    """Rows whose maximum moves by far more than 2^60 inside a key tile and from one tile to the next: the
    single-sweep forward kernel must take its rescale path (O and l in tensor memory / registers are
    multiplied by exp2(m_used - m_new) and the tile is swept again) and still return the exact softmax."""
    torch.manual_seed(N + hd)
    scale = hd ** -0.5
    qkv = torch.randn(B, N, 3, H, hd, device=dev)
    qkv[:, :, 0] *= 6.0
    for k0, f in ((40, 8.0), (170, 30.0), (300, 90.0), (1000, 200.0)):
        if k0 < N:
            qkv[:, k0:k0 + 3, 1] *= f
    qkv = bf(qkv)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale * 1.4426950408889634
    first = s[..., :32].max(-1).values
    assert ((s.max(-1).values - first) > 60.0).any(), "the inputs do not reach the rescale path"
    o, lse = ops.attention_fwd(q, k, v, scale)
    oref, lref = _ref_attn(q.float(), k.float(), v.float(), scale)
    _ok(o, oref, 1.5e-2)
    _ok(lse, lref, 1e-3)


def test_attention_cross_lengths():
    """Nq != Nk (decoder-style)."""
    B, Nq, Nk, H, hd = 2, 70, 300, 2, 64
    q, k, v = [bf(torch.randn(B, n, H, hd, device=dev)) for n in (Nq, Nk, Nk)]
    o, lse = ops.attention_fwd(q, k, v, hd ** -0.5)
    oref, lref = _ref_attn(q.float(), k.float(), v.float(), hd ** -0.5)
    _ok(o, oref, 1.5e-2)
    _ok(lse, lref, 1e-3)


@pytest.mark.parametrize("B,Nq,Nk,H,hd", [(2, 70, 300, 2, 64), (2, 300, 70, 2, 64), (3, 256, 257, 2, 32), (40, 197, 197, 4, 64),
                                          (20, 130, 520, 8, 64)])
def test_attention_bwd_cross_lengths_and_many_items(B, Nq, Nk, H, hd):
    """Backward with Nq != Nk on both schedules (Nq <= 256: dQ accumulated in tensor memory over all key
    tiles; longer: fp32 reduce-add), and with more work items than SMs so every CTA walks several."""
    torch.manual_seed(Nq * 1000 + Nk)
    scale = hd ** -0.5
    q, k, v = [bf(torch.randn(B, n, H, hd, device=dev)) for n in (Nq, Nk, Nk)]
    o, lse = ops.attention_fwd(q, k, v, scale)
    qf, kf, vf = [t.float().detach().requires_grad_(True) for t in (q, k, v)]
    oref, _ = _ref_attn(qf, kf, vf, scale)
    do = bf(torch.randn(B, Nq, H, hd, device=dev) * 0.5)
    dq, dk, dv = ops.attention_bwd(q, k, v, o, do, lse, scale)
    oref.backward(do.float())
    _ok(dq, qf.grad, 2e-2)
    _ok(dk, kf.grad, 2e-2)
    _ok(dv, vf.grad, 2e-2)


def test_attention_rows_sum_to_one_at_long_sequence():
    """N = 4096 (SAP config): with V = 1 every output must be exactly ~1 (softmax rows sum to 1)."""
    B, N, H, hd = 1, 4096, 2, 64
    q, k = bf(torch.randn(B, N, H, hd, device=dev)), bf(torch.randn(B, N, H, hd, device=dev))
    v = torch.ones(B, N, H, hd, device=dev, dtype=torch.bfloat16)
    o, _ = ops.attention_fwd(q, k, v, hd ** -0.5)
    assert (o.float() - 1.0).abs().max().item() <= 1e-2


def test_attention_properties_at_benchmark_size():
    """BASELINE configs[1] shape (B 256, N 197, H 12, hd 64), size-independent properties:
    forward -- with V = 1 every output is 1 (softmax rows sum to one) and lse = logsumexp bounds hold;
    backward -- when every value row is the same vector, O does not depend on Q / K, so dQ = dK = 0, and since
    the softmax rows sum to one, sum_j dV[b, j, h, :] = sum_i dO[b, i, h, :]."""
    torch.manual_seed(5)
    B, N, H, hd = 256, 197, 12, 64
    scale = hd ** -0.5
    qkv = bf(torch.randn(B, N, 3, H, hd, device=dev))
    vrow = bf(torch.randn(1, 1, H, hd, device=dev))
    qkv[:, :, 2] = vrow                                   # all keys share one value vector per head
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, scale)
    assert (o.float() - vrow.float()).abs().max().item() <= 2e-2 * vrow.float().abs().max().item() + 1e-3
    do = bf(torch.randn(B, N, H, hd, device=dev))
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(q, k, v, o, do, lse, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
    ref_scale = do.float().abs().max().item()
    # dQ, dK vanish up to the bf16 rounding of P, dP and of the saved O (delta is formed from the rounded O)
    assert dqkv[:, :, 0].float().abs().max().item() <= 5e-2 * ref_scale
    assert dqkv[:, :, 1].float().abs().max().item() <= 5e-2 * ref_scale
    dv_sum = dqkv[:, :, 2].float().sum(dim=1)            # [B, H, hd]
    do_sum = do.float().sum(dim=1)
    assert (dv_sum - do_sum).abs().max().item() <= 2e-2 * do_sum.abs().max().item() + 0.1
    assert torch.isfinite(dqkv.float()).all() and torch.isfinite(lse).all()


@pytest.mark.parametrize("rows,Na,V,H,hd,shared", [(100, 1, 4, 3, 64, True), (37, 1, 7, 2, 32, True), (50, 2, 3, 2, 64, False)])
def test_var_attention(rows, Na, V, H, hd, shared):
    torch.manual_seed(V)
    scale = hd ** -0.5
    q = bf(torch.randn(1 if shared else rows, Na, H, hd, device=dev))
    kv = bf(torch.randn(rows, V, 2, H, hd, device=dev))
    o, lse = ops.var_attention_fwd(q, kv, scale)
    qf = q.float().detach().requires_grad_(True)
    kvf = kv.float().detach().requires_grad_(True)
    s = torch.einsum("rahd,rvhd->rhav", qf.expand(rows, -1, -1, -1), kvf[:, :, 0]) * scale
    oref = torch.einsum("rhav,rvhd->rahd", s.softmax(-1), kvf[:, :, 1])
    _ok(o, oref, 1.5e-2)
    do = bf(torch.randn_like(oref) * 0.5)
    dq_acc, dkv = ops.var_attention_bwd(q, kv, o, do, lse, scale)
    oref.backward(do.float())
    _ok(dkv, kvf.grad, 2e-2)
    _ok(dq_acc, qf.grad, 2e-2)


def test_cast_multi_matches_single_casts():
    """One launch casting several fp32 tensors must equal the per-tensor casts (ragged sizes, 1..8 tensors)."""
    torch.manual_seed(3)
    for shapes in ([(2304, 768), (768, 768), (3072, 768), (768, 3072)], [(17,)], [(64, 3), (5, 5, 5), (1,), (1024,), (9, 8), (8,), (3, 3), (40,)]):
        xs = [torch.randn(*sh, device=dev) for sh in shapes]
        outs = ops.cast_to_bf16_multi(xs)
        for x, o in zip(xs, outs):
            assert o.shape == x.shape and o.dtype == torch.bfloat16
            assert torch.equal(o, x.to(torch.bfloat16))


def test_elementwise_helpers():
    x = torch.randn(1000000, device=dev)
    assert torch.equal(ops.cast_to_bf16(x), x.to(torch.bfloat16))
    xb = bf(torch.randn(5000, 776, device=dev))
    _ok(ops.colsum(xb), xb.float().sum(0), 1e-4)
    _ok(ops.colsum(xb[:, 8:520]), xb[:, 8:520].float().sum(0), 1e-4)
    acc = torch.ones(776, device=dev)
    _ok(ops.colsum(xb, out=acc, accumulate=True), xb.float().sum(0) + 1, 1e-4)
    img = torch.randn(3, 3, 32, 48, device=dev)
    ref = img.reshape(3, 3, 2, 16, 3, 16).permute(0, 2, 4, 1, 3, 5).reshape(18, 768)
    assert torch.equal(ops.patchify(img, 16), ref.to(torch.bfloat16))
    vol = torch.randn(2, 2, 16, 32, 16, device=dev)
    ref = vol.reshape(2, 2, 2, 8, 4, 8, 2, 8).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(32, 1024)
    assert torch.equal(ops.patchify(vol, 8), ref.to(torch.bfloat16))
    assert torch.equal(ops.patchify(bf(vol), 8), ref.to(torch.bfloat16))


def _conv_rows(x, p):
    """What Conv(k = s = p) sees: unfold over whole patches only, K ordered (c, p0, p1(, p2))."""
    nd = x.dim() - 2
    B, C = x.shape[:2]
    G = [s // p for s in x.shape[2:]]
    x = x[(slice(None), slice(None)) + tuple(slice(0, g * p) for g in G)]
    if nd == 2:
        return x.reshape(B, C, G[0], p, G[1], p).permute(0, 2, 4, 1, 3, 5).reshape(B * G[0] * G[1], C * p * p)
    return x.reshape(B, C, G[0], p, G[1], p, G[2], p).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(-1, C * p ** 3)


@pytest.mark.parametrize("shape,p", [((2, 1, 64, 64), 4), ((2, 3, 30, 45), 4), ((1, 1, 20, 24, 28), 4), ((2, 1, 18, 18), 2),
                                     ((2, 3, 21, 35), 7), ((1, 2, 37, 50), 16), ((2, 1, 17, 16, 19), 8)])
def test_patchify_any_patch_size_and_cropped_images(shape, p):
    """basic_ct configs use patch_size 4 (reference configs/basic_ct/*/base_config.yaml); Conv(k = s = p) floors
    a spatial size that is not a multiple of p (building_blocks.py:58-60)."""
    x = torch.randn(*shape, device=dev)
    ref = _conv_rows(x, p).to(torch.bfloat16)
    for xin in (x, bf(x)):
        got = ops.patchify(xin.contiguous(), p)
        K = ref.shape[1]
        assert got.shape == (ref.shape[0], -(-K // 8) * 8)
        assert torch.equal(got[:, :K], ref)
        assert not got[:, K:].any()
    # raw uint8 pixels (a quarter of the fp32 host->device bytes): exact, 0..255 are bf16-representable
    xu = torch.randint(0, 256, shape, device=dev, dtype=torch.uint8)
    got = ops.patchify(xu, p)
    assert torch.equal(got[:, :K], _conv_rows(xu.float(), p).to(torch.bfloat16))


def test_patch_embed_matches_conv_for_small_patches():
    from ucf_vit_b200.simple.building_blocks import PatchEmbed
    torch.manual_seed(5)
    for twoD, img, p, C in ((True, 32, 4, 1), (False, 16, 4, 1), (True, 18, 2, 1)):
        pe = PatchEmbed(img_size=img, patch_size=p, in_chans=C, embed_dim=64, twoD=twoD).to(dev)
        x = torch.randn(2, C, *([img] * (2 if twoD else 3)), device=dev)
        y = pe(x)
        ref = pe.proj(x).flatten(2).transpose(1, 2)
        _ok(y, ref, 2e-2)
        y.float().sum().backward()
        gref = torch.autograd.grad(pe.proj(x).sum(), pe.proj.weight)[0]
        _ok(pe.proj.weight.grad, gref, 2e-2)
        pe.zero_grad()


def test_assemble_tokens_fwd_bwd():
    from ucf_vit_b200 import functional as UF
    B, Lp, D = 3, 10, 64
    tok = bf(torch.randn(B, Lp, D, device=dev)).requires_grad_(True)
    cls = torch.randn(1, 1, D, device=dev, requires_grad=True)
    pos = torch.randn(1, Lp + 1, D, device=dev, requires_grad=True)
    y = UF.assemble_tokens(tok, cls.reshape(1, -1), pos, pos_has_prefix=True)
    ref = torch.cat([cls.expand(B, -1, -1), tok.float()], 1) + pos
    _ok(y, ref, 1e-2)
    g = bf(torch.randn_like(ref))
    y.backward(g)
    gt, gc, gp = tok.grad.clone(), cls.grad.clone(), pos.grad.clone()
    tok.grad = cls.grad = pos.grad = None
    ref.backward(g.float())
    _ok(gt, tok.grad, 1e-2); _ok(gc, cls.grad, 1e-2); _ok(gp, pos.grad, 1e-2)
    # per-sample (adaptive) embedding without a row for the cls token; and no prefix at all
    pe = torch.randn(B, Lp, D, device=dev)
    y2 = UF.assemble_tokens(tok.detach(), cls.detach().reshape(1, -1), pe, pos_has_prefix=False)
    ref2 = torch.cat([cls.detach().expand(B, -1, -1), tok.detach().float() + pe], 1)
    _ok(y2, ref2, 1e-2)
    y3 = UF.assemble_tokens(tok.detach(), None, pos.detach()[:, 1:], pos_has_prefix=True)
    _ok(y3, tok.detach().float() + pos.detach()[:, 1:], 1e-2)


def test_errors_are_loud():
    a = bf(torch.randn(16, 20, device=dev))
    with pytest.raises(RuntimeError, match="16-byte"):
        ops.gemm(a, a, M=16, N=16, K=20)               # K*2 bytes not a multiple of 16
    with pytest.raises(RuntimeError, match="head_dim"):
        q = bf(torch.randn(1, 8, 2, 48, device=dev))
        ops.attention_fwd(q, q, q, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.layernorm_fwd(torch.randn(4, 64), None, None, 1e-5)


@pytest.mark.parametrize("B,Lt,keep", [(4, 196, 49), (3, 64, 16), (2, 1000, 250), (1, 4096, 1024), (5, 7, 0), (2, 9, 9),
                                       (256, 196, 49)])
def test_mask_plan_matches_argsort(B, Lt, keep):
    """MAE.random_masking's index work (arch.py:663-681): bit-exact against two stable argsorts, with and
    without ties in the noise."""
    torch.manual_seed(B + Lt)
    for noise in (torch.rand(B, Lt, device=dev), torch.randint(0, 5, (B, Lt), device=dev).float()):
        sh, rs, mk = ops.mask_plan(noise, keep)
        ref_sh = torch.argsort(noise, dim=1, stable=True)
        ref_rs = torch.argsort(ref_sh, dim=1)
        ref_mk = torch.ones(B, Lt, device=dev)
        ref_mk[:, :keep] = 0
        ref_mk = torch.gather(ref_mk, 1, ref_rs)
        assert torch.equal(sh, ref_sh) and torch.equal(rs, ref_rs) and torch.equal(mk, ref_mk)


@pytest.mark.parametrize("B,Ls,Lo,D,per_sample_pos", [(3, 20, 50, 64, False), (2, 49, 196, 512, True), (4, 1, 9, 8, False)])
def test_gather_tokens_restores_the_sequence_like_mask_head(B, Ls, Lo, D, per_sample_pos):
    """cat(x, mask_token.repeat) -> gather(ids_restore) -> + pos (arch.py:687-698) and its gradients."""
    from ucf_vit_b200 import functional as UF
    torch.manual_seed(Ls + Lo)
    src = bf(torch.randn(B, Ls, D, device=dev)).requires_grad_(True)
    idx = torch.stack([torch.randperm(Lo, device=dev) for _ in range(B)])
    fill = torch.randn(1, 1, D, device=dev, requires_grad=True)
    pos = torch.randn(B if per_sample_pos else 1, Lo, D, device=dev, requires_grad=True)
    out = UF.gather_tokens(src, idx, fill=fill, pos=pos, complete=True)
    s32, f32_, p32 = (t.detach().float().requires_grad_(True) for t in (src, fill, pos))
    full = torch.cat([s32, f32_.expand(B, Lo - Ls, D)], dim=1)
    ref = torch.gather(full, 1, idx.unsqueeze(-1).expand(-1, -1, D)) + p32
    _ok(out, ref, 1e-2)
    g = bf(torch.randn(B, Lo, D, device=dev))
    out.backward(g)
    ref.backward(g.float())
    assert torch.equal(src.grad.float(), s32.grad)                     # pure row moves
    _ok(fill.grad, f32_.grad, 1e-2 if B * (Lo - Ls) > 100 else 2e-3)    # fp32 sum of bf16 rows, different order
    _ok(pos.grad, p32.grad, 2e-3)
    # without pos / fill the rows are copied bit for bit, and out-of-range indices give zero rows
    out2 = ops.gather_tokens(src.detach(), idx)
    ref2 = torch.gather(torch.cat([src.detach(), torch.zeros(B, Lo - Ls, D, device=dev, dtype=torch.bfloat16)], 1), 1,
                        idx.unsqueeze(-1).expand(-1, -1, D))
    assert torch.equal(out2, ref2)


def test_gather_tokens_keeps_tokens_like_random_masking():
    from ucf_vit_b200 import functional as UF
    torch.manual_seed(9)
    B, Lt, D, keep = 6, 196, 192, 49
    seq = bf(torch.randn(B, Lt, D, device=dev)).requires_grad_(True)
    sh, rs, mk = ops.mask_plan(torch.rand(B, Lt, device=dev), keep)
    ids_keep = sh[:, :keep].contiguous()
    kept = UF.gather_tokens(seq, ids_keep)
    s32 = seq.detach().float().requires_grad_(True)
    ref = torch.gather(s32, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    assert torch.equal(kept.float(), ref)
    g = bf(torch.randn(B, keep, D, device=dev))
    kept.backward(g)
    ref.backward(g.float())
    assert torch.equal(seq.grad.float(), s32.grad)                     # removed tokens get exact zeros
    assert (seq.grad.float().abs().sum(-1) == 0).sum().item() == B * (Lt - keep)


@pytest.mark.parametrize("xs,es,edt", [((3, 4, 10, 64), (1, 4, 1, 64), torch.float32),     # variable embedding
                                       ((5, 197, 128), (5, 1, 128), torch.float32),         # time embedding
                                       ((2, 50, 64), (1, 1, 64), torch.bfloat16),
                                       ((2, 3, 7, 8), (2, 3, 7, 8), torch.float32),         # no broadcast at all
                                       ((6, 16), (16,), torch.float32)])
def test_add_bcast_fwd_bwd(xs, es, edt):
    from ucf_vit_b200 import functional as UF
    torch.manual_seed(sum(xs))
    x = bf(torch.randn(*xs, device=dev)).requires_grad_(True)
    e = torch.randn(*es, device=dev).to(edt).requires_grad_(True)
    out = UF.add_bcast(x, e)
    x32, e32 = x.detach().float().requires_grad_(True), e.detach().float().requires_grad_(True)
    ref = x32 + e32
    _ok(out, ref, 1e-2)
    g = bf(torch.randn(*xs, device=dev))
    out.backward(g)
    ref.backward(g.float())
    assert torch.equal(x.grad.float(), x32.grad)
    assert e.grad.dtype == edt and e.grad.shape == e.shape
    _ok(e.grad, e32.grad, 1e-2 if edt == torch.bfloat16 else 2e-3)


@pytest.mark.parametrize("B,N,D,H", [(256, 197, 768, 12), (5, 300, 512, 16), (3, 1000, 1024, 16)])
def test_dgrad_with_fused_attention_delta(B, N, D, H):
    """ucf_gemm_dgrad_delta: the attention-output dgrad whose epilogue also leaves rowsum(dO o O) per head."""
    torch.manual_seed(B + N)
    M, hd = B * N, D // H
    dy = bf(torch.randn(M, D, device=dev))
    w = bf(torch.randn(D, D, device=dev) * 0.05)
    o = bf(torch.randn(M, D, device=dev))
    dx, delta = ops.gemm_dgrad_delta(dy, w, o, N, H)
    ref = dy.float() @ w.float()
    _ok(dx, ref, 1e-2)
    assert torch.equal(dx, ops.gemm(dy, w, M=M, N=D, K=D, b_mn=True))          # same product as the plain dgrad kernel
    dref = (ref * o.float()).view(B, N, H, hd).sum(-1).permute(0, 2, 1)
    _ok(delta, dref, 2e-3)


@pytest.mark.parametrize("M,D,N", [(50432, 768, 2304), (1000, 512, 1536), (4096, 1024, 3072)])
def test_layernorm_folded_into_gemm(M, D, N):
    """ucf_layernorm_stats + ucf_ln_gemm (LayerNorm folded into the projection that consumes it) against LN -> Linear."""
    from ucf_vit_b200 import functional as UF
    torch.manual_seed(M % 97)
    x = bf(torch.randn(M, D, device=dev) * 1.5 + 0.7)            # rows with a non-zero mean: the rank-1 correction matters
    ln_w = torch.randn(D, device=dev) * 0.3 + 1.0
    ln_b = torch.randn(D, device=dev) * 0.2
    w = torch.randn(N, D, device=dev) * 0.03
    b = torch.randn(N, device=dev) * 0.1
    mean, rstd = ops.layernorm_stats(x, 1e-6)
    xf = x.float()
    _ok(mean, xf.mean(-1), 1e-4)
    _ok(rstd, (xf.var(-1, unbiased=False) + 1e-6).rsqrt(), 1e-4)
    wg, colsum, bfold = UF.fold_layernorm(w, b, ln_w, ln_b)
    y = ops.ln_gemm(x, wg, bfold, colsum, mean, rstd)
    ref = torch.nn.functional.layer_norm(xf, (D,), ln_w, ln_b, 1e-6) @ w.t() + b
    _ok(y, ref, 1.5e-2)


def test_block_eval_path_matches_training_path():
    """Block in eval mode under no_grad (LayerNorm1 folded into QKV) == the training-mode forward of the same weights."""
    from functools import partial
    from ucf_vit_b200.simple.building_blocks import Block
    torch.manual_seed(1)
    blk = Block(dim=768, num_heads=12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).to(dev)
    with torch.no_grad():
        blk.norm1.weight.add_(torch.randn_like(blk.norm1.weight) * 0.2)
        blk.norm1.bias.add_(torch.randn_like(blk.norm1.bias) * 0.2)
    x = torch.randn(8, 197, 768, device=dev)
    blk.train()
    y_train = blk(x).detach()
    blk.eval()
    with torch.no_grad():
        y_eval = blk(x)
        assert blk._ln_fold is not None
        _ok(y_eval, y_train.float(), 1e-2)
        blk.norm1.weight.mul_(1.5)                       # an in-place edit bumps the version: the fold is redone
        y2 = blk(x)
    blk.train()
    assert blk._ln_fold is None
    _ok(y2, blk(x).detach().float(), 1e-2)
