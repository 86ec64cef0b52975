"""The CPU restatement of OpenCV's GaussianBlur / Canny (oracle/canny_np.py) against OpenCV's own outputs: the committed
fixture (tests/golden/front_end/canny_cv2.npz, oracle/gen_golden_canny.py) and, where cv2 is importable, cv2 itself on fresh images."""
import os

import numpy as np
import pytest

from oracle.canny_np import canny_u8, gaussian_blur_u8

GOLD = os.path.join(os.path.dirname(__file__), "golden", "front_end", "canny_cv2.npz")


def test_oracle_matches_the_opencv_fixture_bit_for_bit():
    g = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in g.files if "/" in k})
    assert len(names) == 5
    n = 0
    for name in names:
        img = g[f"{name}/img"]
        for k in (1, 3, 5):
            b = gaussian_blur_u8(img, k)
            assert np.array_equal(b, g[f"{name}/blur{k}"]), (name, k)
            for lo in (50, 77, 99):
                assert np.array_equal(canny_u8(b, lo, lo + 50), g[f"{name}/canny{k}_{lo}"]), (name, k, lo)
                n += 1
    assert n == 45


def test_oracle_matches_cv2_on_fresh_images():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for shape in [(37, 53, 3), (1, 9, 3), (9, 1), (2, 2, 3), (64, 130, 4), (50, 50, 2)]:
        img = (rng.random(shape) * 255).astype(np.uint8)
        img2 = cv2.GaussianBlur(img, (9, 9), 0).reshape(img.shape)           # smoother content: sparse edges, long chains
        for im in (img, img2):
            for k in (1, 3, 5):
                ref = cv2.GaussianBlur(im, (k, k), 0).reshape(im.shape)
                assert np.array_equal(gaussian_blur_u8(im, k), ref), (shape, k)
                for lo, hi in ((50, 100), (99, 149), (10, 20), (120, 60)):
                    assert np.array_equal(canny_u8(ref, lo, hi), cv2.Canny(ref, lo, hi)), (shape, k, lo, hi)
