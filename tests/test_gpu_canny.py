"""Device Gaussian blur + Canny (csrc/canny.cu) through the C ABI: byte-identical to OpenCV -- against the committed cv2
fixture, the pinned oracle on fresh images, and cv2 itself (where importable) up to 4096 x 4096 x 3; then the quadtree built
from the device edge map equals the one built from OpenCV's."""
import os
import random

import numpy as np
import pytest
import torch

from oracle.canny_np import canny_u8 as canny_ref, gaussian_blur_u8 as blur_ref

pytestmark = pytest.mark.gpu
if torch.cuda.is_available():
    from ucf_vit_b200 import ops
GOLD = os.path.join(os.path.dirname(__file__), "golden", "front_end", "canny_cv2.npz")


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def test_blur_and_canny_match_the_opencv_fixture():
    g = np.load(GOLD)
    for name in sorted({k.split("/")[0] for k in g.files if "/" in k}):
        img = g[f"{name}/img"]
        for k in (1, 3, 5):
            b = ops.gaussian_blur_u8(_dev(img), k)
            assert np.array_equal(b.cpu().numpy(), g[f"{name}/blur{k}"]), (name, k)
            for lo in (50, 77, 99):
                e = ops.canny_u8(b, lo, lo + 50)
                assert np.array_equal(e.cpu().numpy(), g[f"{name}/canny{k}_{lo}"]), (name, k, lo)


@pytest.mark.parametrize("shape", [(37, 53, 3), (1, 9, 3), (9, 1), (2, 2, 3), (64, 130, 4), (50, 50, 2), (33, 1027, 3),
                                   (200, 513, 1), (129, 65, 3), (300, 300, 3)])
def test_blur_and_canny_match_the_oracle_on_ragged_sizes(shape):
    """sizes that leave partial tiles in every kernel (blur 32 x 512 bytes, nms 32 x 64, hysteresis 64 x 64), rows whose
    byte length is not a multiple of 4, 1..4 channels, degenerate 1-pixel axes"""
    rng = np.random.default_rng(sum(shape))
    noise = (rng.random(shape) * 255).astype(np.uint8)
    smooth = blur_ref(blur_ref(blur_ref(noise, 5), 5), 5)            # sparse edges with long connected chains
    for img in (noise, smooth):
        for k in (1, 3, 5):
            want = blur_ref(img, k)
            b = ops.gaussian_blur_u8(_dev(img), k)
            assert np.array_equal(b.cpu().numpy(), want), (shape, k)
            for lo, hi in ((50, 100), (10, 20), (120, 60)):
                assert np.array_equal(ops.canny_u8(b, lo, hi).cpu().numpy(), canny_ref(want, lo, hi)), (shape, k, lo, hi)


def _scene(n, rng):
    """n x n x 3 synthetic 'photograph': smooth shading, discs, bars and a little noise -- long closed contours."""
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    img = np.stack([110 + 60 * np.sin(xx / (n / 9)) * np.cos(yy / (n / 7)), 90 + 0.03 * xx * 255 / n * 30, 140 - 40 * np.cos((xx + yy) / (n / 5))], -1)
    for _ in range(40):
        cx, cy, r = rng.random() * n, rng.random() * n, (0.01 + 0.08 * rng.random()) * n
        m = (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
        img[m] = rng.random(3) * 255
    for _ in range(12):
        x0, w = int(rng.random() * n), int(2 + rng.random() * n * 0.02)
        img[:, x0:x0 + w] = rng.random(3) * 255
    img += rng.normal(0, 3, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("n", [1024, 4096])
def test_full_size_images_match_opencv(n):
    """BASELINE configs[4]'s 4096 x 4096 x 3 images: OpenCV itself is the witness (it is part of the image; the CPU oracle's
    Python hysteresis is for small cases)."""
    cv2 = pytest.importorskip("cv2")
    img = _scene(n, np.random.default_rng(n))
    x = _dev(img)
    for k, lo in ((5, 60), (3, 99), (1, 50)):
        want_b = cv2.GaussianBlur(img, (k, k), 0)
        b = ops.gaussian_blur_u8(x, k)
        assert np.array_equal(b.cpu().numpy(), want_b), (n, k)
        e, sweeps = ops.canny_u8(b, lo, lo + 50, return_sweeps=True)
        want = cv2.Canny(want_b, lo, lo + 50)
        assert np.array_equal(e.cpu().numpy(), want), (n, k, lo, int((e.cpu().numpy() != want).sum()))
        assert 0 < want.mean() < 64, want.mean()          # a sparse but non-empty edge map
        print(f"canny {n}x{n} k={k} low={lo}: {sweeps} hysteresis sweeps, {100 * (want > 0).mean():.2f} % edge pixels")


def test_canny_is_loud():
    with pytest.raises(TypeError, match="uint8"):
        ops.canny_u8(torch.zeros(8, 8, 3, device="cuda"), 50, 100)
    with pytest.raises(RuntimeError, match="ksize"):
        ops.gaussian_blur_u8(torch.zeros(8, 8, 3, dtype=torch.uint8, device="cuda"), 7)
    with pytest.raises(RuntimeError, match="C=5"):
        ops.canny_u8(torch.zeros(8, 8, 5, dtype=torch.uint8, device="cuda"), 50, 100)


def test_patchify_with_device_edges_builds_the_same_tree_as_opencv():
    """Patchify(edges='device') against Patchify(edges='host') (the reference's cv2 calls) with the same random draws:
    identical edge maps, node boxes and sequences -- 'bit-exact SAP patch indices'."""
    pytest.importorskip("cv2")
    from ucf_vit_b200.dataloaders.transform import Patchify
    img = _scene(512, np.random.default_rng(3))
    for seed in range(4):
        outs = []
        for mode in ("host", "device"):
            random.seed(seed)
            np.random.seed(seed)
            t = Patchify(sths=[1, 3, 5], fixed_length=196, patch_size=8, num_channels=3, dataset="imagenet", return_edges=True,
                         edges=mode)
            outs.append(t(img))
        (s0, z0, p0, q0, e0), (s1, z1, p1, q1, e1) = outs
        assert np.array_equal(e0, e1)
        assert np.array_equal(z0, z1) and np.array_equal(p0, p1) and np.array_equal(s0, s1)
