"""InstanceNorm (+ residual) + LeakyReLU kernels of the UNETR decoder (csrc/instnorm.cu) through the C ABI, against
torch.nn.InstanceNorm{2,3}d + LeakyReLU evaluated in fp32 on the same bf16 inputs; then the whole fused decoder
(UNETR.use_fused_decoder) against the oracle decoder on the same features."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from ucf_vit_b200 import functional as UF, ops
dev = "cuda"


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last if t.dim() == 4 else torch.channels_last_3d)


def _ref(a, b, mode, slope, eps=1e-5):
    inorm = (nn.InstanceNorm2d if a.dim() == 4 else nn.InstanceNorm3d)(a.shape[1], eps=eps)
    z = inorm(a)
    if mode == 1:
        z = z + b
    if mode == 2:
        z = z + inorm(b)
    return torch.nn.functional.leaky_relu(z, slope) if slope != 1.0 else z


# (shape, second operand: 0 none / 1 raw / 2 normalised, slope); C = 16..256 (vector 8), 12 (vector 4), 6 (vector 2),
# 48 (6 vectors: lanes do not divide 256), 384 / 768 (more than 1024 partial columns: the finish kernel's second pass), S ragged against the rows a CTA covers, one 2-D case
CASES = [((2, 16, 8, 8, 8), 0, 0.01), ((2, 16, 8, 8, 8), 1, 0.01), ((2, 16, 8, 8, 8), 2, 0.01),
         ((3, 32, 9, 7, 5), 2, 0.01), ((1, 128, 16, 16, 16), 1, 0.01), ((2, 256, 5, 3, 2), 0, 0.2),
         ((2, 48, 6, 6, 6), 2, 0.01), ((2, 12, 7, 5, 3), 1, 0.01), ((1, 6, 11, 3, 2), 2, 0.01),
         ((2, 64, 33, 17), 2, 0.01), ((2, 16, 40, 40, 40), 0, 1.0), ((2, 768, 4, 4, 4), 2, 0.01), ((1, 384, 6, 5, 4), 1, 0.01), ((4, 16, 64, 64, 64), 2, 0.01)]


@pytest.mark.parametrize("shape,mode,slope", CASES)
def test_instance_norm_act_fwd_bwd(shape, mode, slope):
    g = torch.Generator().manual_seed(sum(shape) + mode)
    a = (torch.randn(shape, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    b = (torch.randn(shape, generator=g) * 0.7 - 0.1).to(torch.bfloat16)
    dy = torch.randn(shape, generator=g).to(torch.bfloat16)
    a1 = _cl(a.to(dev)).requires_grad_(True)
    b1 = _cl(b.to(dev)).requires_grad_(True) if mode else None
    y = UF.instance_norm_act(a1, b1, norm_residual=(mode == 2), negative_slope=slope)
    assert y.dtype == torch.bfloat16 and y.stride() == a1.stride()
    y.backward(_cl(dy.to(dev)))
    a2 = a.to(dev).float().requires_grad_(True)
    b2 = b.to(dev).float().requires_grad_(True)
    yr = _ref(a2, b2, mode, slope)
    yr.backward(dy.to(dev).float())
    # bf16 outputs: half an ulp of the result (2^-9 relative) + the fp32 reference's own error
    assert torch.allclose(y.float(), yr, rtol=8e-3, atol=8e-3), (y.float() - yr).abs().max().item()
    ga = a1.grad.float()
    scale = a2.grad.abs().max().item()
    assert (ga - a2.grad).abs().max().item() <= 1e-2 * scale + 1e-6, ((ga - a2.grad).abs().max().item(), scale)
    if mode:
        scale = b2.grad.abs().max().item()
        assert (b1.grad.float() - b2.grad).abs().max().item() <= 1e-2 * scale + 1e-6


def test_instance_norm_statistics_are_fp32_exact_enough_and_reproducible():
    """mean / rstd against float64 on a tensor with a large offset (the E[x^2] - mean^2 form is finished in double), and
    bit-identical results run to run (fixed-order two-stage reduction, no atomics)."""
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(2, 32, 24, 24, 24, generator=g) * 0.5 + 20.0).to(torch.bfloat16)
    xd = _cl(x.to(dev))
    s1 = ops.inorm_stats(xd)
    s2 = ops.inorm_stats(xd)
    assert torch.equal(s1, s2)
    x64 = x.double().flatten(2)
    mean = x64.mean(2)
    rstd = 1.0 / torch.sqrt(x64.var(2, unbiased=False) + 1e-5)
    assert torch.allclose(s1[:, 0].cpu().double(), mean, rtol=1e-6, atol=1e-6)
    assert torch.allclose(s1[:, 1].cpu().double(), rstd, rtol=1e-4)


def test_instance_norm_kernels_are_loud():
    with pytest.raises(TypeError, match="channels-last"):
        ops.inorm_stats(torch.randn(2, 16, 4, 4, 4, device=dev).to(torch.bfloat16))         # NCDHW memory
    with pytest.raises(TypeError, match="bf16"):
        ops.inorm_stats(_cl(torch.randn(2, 16, 4, 4, 4, device=dev)))
    with pytest.raises(RuntimeError, match="C=7"):
        ops.inorm_stats(_cl(torch.randn(2, 7, 4, 4, 4, device=dev).to(torch.bfloat16)))      # odd channel count
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        UF.instance_norm_act(torch.randn(2, 16, 4, 4, 4).to(torch.bfloat16))


@pytest.mark.parametrize("Ci,Co", [(16, 16), (32, 16), (32, 32), (64, 32)])
@pytest.mark.parametrize("N,D,H,W", [(1, 4, 4, 8), (2, 8, 8, 16), (3, 12, 8, 24), (2, 16, 20, 40)])
def test_conv3d_wgrad_matches_autograd(Ci, Co, N, D, H, W):
    """ucf_conv3d_wgrad against autograd's weight gradient of F.conv3d evaluated in fp32 on the same bf16 values (single tile,
    several tiles per CTA, more tiles than CTAs; every channel pair the kernel serves)."""
    g = torch.Generator().manual_seed(Ci + Co + N + D)
    x = torch.randn(N, Ci, D, H, W, generator=g).to(torch.bfloat16)
    dy = torch.randn(N, Co, D, H, W, generator=g).to(torch.bfloat16)
    assert ops.conv3d_wgrad_supported(Ci, Co, D, H, W)
    dw = ops.conv3d_wgrad(_cl(x.to(dev)), _cl(dy.to(dev)))
    dw2 = ops.conv3d_wgrad(_cl(x.to(dev)), _cl(dy.to(dev)))
    assert torch.equal(dw, dw2)                                             # fixed-order reduction
    w = torch.zeros(Co, Ci, 3, 3, 3, device=dev, requires_grad=True)
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = torch.nn.functional.conv3d(x.to(dev).float(), w, None, 1, 1)
        y.backward(dy.to(dev).float())
    finally:
        torch.backends.cudnn.allow_tf32 = True
    ref = w.grad
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_conv3x3x3_first_layer_pads_the_input_channels():
    """UF.conv3x3x3 with 4 input channels (the decoder's first layer): forward and dx from cuDNN, dW from ucf_conv3d_wgrad on a
    zero-padded copy -- against autograd in fp32."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 4, 8, 8, 16, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(16, 4, 3, 3, 3, generator=g) * 0.1).to(dev)
    dy = torch.randn(2, 16, 8, 8, 16, generator=g).to(torch.bfloat16).to(dev)
    x1 = _cl(x).requires_grad_(True)
    w1 = w.clone().requires_grad_(True)
    y = UF.conv3x3x3(x1, w1)
    y.backward(_cl(dy))
    x2 = x.float().requires_grad_(True)
    w2 = w.to(torch.bfloat16).float().requires_grad_(True)
    torch.backends.cudnn.allow_tf32 = False
    try:
        y2 = torch.nn.functional.conv3d(x2, w2, None, 1, 1)
        y2.backward(dy.float())
    finally:
        torch.backends.cudnn.allow_tf32 = True
    assert w1.grad.dtype == torch.float32 and w1.grad.shape == w.shape
    assert (y.float() - y2).abs().max().item() <= 1e-2 * y2.abs().max().item()
    assert (w1.grad - w2.grad).abs().max().item() <= 2e-3 * w2.grad.abs().max().item()
    assert (x1.grad.float() - x2.grad).abs().max().item() <= 1e-2 * x2.grad.abs().max().item()


@pytest.mark.parametrize("Ci,Co,bias", [(4, 16, False), (16, 4, True), (32, 16, False), (16, 32, False), (8, 8, True), (32, 32, True),
                                        (4, 4, False)])
@pytest.mark.parametrize("shape", [(2, 5, 7, 3), (1, 16, 16, 16), (3, 9, 8, 11)])
def test_conv1x1x1_matches_autograd(Ci, Co, bias, shape):
    """UF.conv1x1x1 (ucf_pointwise_conv / _wgrad) against nn.functional.conv3d in fp32 on the same bf16 values: output, data
    gradient, weight and bias gradients; voxel counts that leave partial tiles."""
    N, D, H, W = shape
    g = torch.Generator().manual_seed(Ci * 100 + Co + N)
    x = torch.randn(N, Ci, D, H, W, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(Co, Ci, 1, 1, 1, generator=g) * 0.3).to(dev)
    b = torch.randn(Co, generator=g).to(dev) if bias else None
    dy = torch.randn(N, Co, D, H, W, generator=g).to(torch.bfloat16).to(dev)
    x1 = _cl(x).requires_grad_(True)
    w1 = w.clone().requires_grad_(True)
    b1 = b.clone().requires_grad_(True) if bias else None
    y = UF.conv1x1x1(x1, w1, b1)
    assert y.dtype == torch.bfloat16 and y.movedim(1, -1).is_contiguous()
    y.backward(_cl(dy))
    x2 = x.float().requires_grad_(True)
    w2 = w.clone().requires_grad_(True)
    b2 = b.clone().requires_grad_(True) if bias else None
    y2 = torch.nn.functional.conv3d(x2, w2, b2)
    y2.backward(dy.float())
    assert (y.float() - y2).abs().max().item() <= 8e-3 * y2.abs().max().item() + 1e-6
    assert (x1.grad.float() - x2.grad).abs().max().item() <= 8e-3 * x2.grad.abs().max().item() + 1e-6
    assert (w1.grad - w2.grad).abs().max().item() <= 1e-4 * w2.grad.abs().max().item() + 1e-5
    if bias:
        assert (b1.grad - b2.grad).abs().max().item() <= 1e-4 * b2.grad.abs().max().item() + 1e-5
    dw_a, _ = ops.pointwise_conv_wgrad(x1.detach(), _cl(dy))
    dw_b, _ = ops.pointwise_conv_wgrad(x1.detach(), _cl(dy))
    assert torch.equal(dw_a, dw_b)


def test_conv3d_wgrad_is_loud_about_unsupported_shapes():
    assert not ops.conv3d_wgrad_supported(48, 16, 8, 8, 16)
    assert not ops.conv3d_wgrad_supported(16, 16, 6, 8, 16)
    x = _cl(torch.randn(1, 48, 8, 8, 16, device=dev).to(torch.bfloat16))
    dy = _cl(torch.randn(1, 16, 8, 8, 16, device=dev).to(torch.bfloat16))
    with pytest.raises(RuntimeError, match="not served"):
        ops.conv3d_wgrad(x, dy)


def _unetr(mode, seed=0):
    from ucf_vit_b200.simple.arch import UNETR
    torch.manual_seed(seed)
    m = UNETR(img_size=[32] * 3, patch_size=16, in_chans=2, num_classes=3, embed_dim=96, depth=4, num_heads=3, twoD=False,
              use_varemb=True, default_vars=["a", "b"], feature_size=16, skip_connection=True, linear_decoder=False,
              class_token=False).to(dev).train()
    if mode == "fused":
        m.use_fused_decoder()
    if mode == "autocast":
        m.conv_autocast_dtype = torch.bfloat16
    return m


def test_fused_decoder_matches_the_fp32_decoder_on_the_same_weights():
    """UNETR with the channels-last bf16 decoder (cuDNN convolutions + ucf_inorm_* kernels) against the same model with the
    default fp32 PyTorch decoder -- the reference's arithmetic: logits and every parameter gradient (rel-L2 per tensor).
    The bar is bf16 activations through ~20 convolutions + InstanceNorms: 3e-2 on the logits, 8e-2 on a gradient -- or, for
    the gradients that reach the encoder through all four skip paths (InstanceNorm backward cancels heavily there), no worse
    than 1.5 x what PyTorch's own bf16 decoder (the same convolutions under autocast, PyTorch InstanceNorm / LeakyReLU)
    deviates from fp32 on the same tensor."""
    m0, m1, m2 = _unetr("fp32"), _unetr("fused"), _unetr("autocast")
    m1.load_state_dict(m0.state_dict())
    m2.load_state_dict(m0.state_dict())
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 2, 32, 32, 32, generator=g).to(dev)
    w = torch.randn(2, 3, 32, 32, 32, generator=g).to(dev)
    ys = [m(x, ["a", "b"]) for m in (m0, m1, m2)]
    assert ys[1].shape == ys[0].shape
    rel = ((ys[1].float() - ys[0].float()).norm() / ys[0].float().norm()).item()
    assert rel < 3e-2, rel
    for y in ys:
        (y.float() * w).sum().backward()
    worst = (0.0, 0.0, "")
    for (n0, p0), (_, p1), (_, p2) in zip(m0.named_parameters(), m1.named_parameters(), m2.named_parameters()):
        if p0.grad is None:
            assert p1.grad is None, n0
            continue
        den = p0.grad.float().norm().item()
        if den == 0.0:
            continue
        r1 = (p1.grad.float() - p0.grad.float()).norm().item() / den
        r2 = (p2.grad.float() - p0.grad.float()).norm().item() / den
        worst = max(worst, (r1, r2, n0))
        assert r1 < max(8e-2, 1.5 * r2), (n0, r1, r2)
    print("fused decoder: logits rel-L2", rel, "worst gradient rel-L2 (fused, torch bf16, name)", worst)
