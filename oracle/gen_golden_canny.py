"""Writes tests/golden/front_end/canny_cv2.npz: small uint8 images with OpenCV's own GaussianBlur / Canny outputs (the two calls of
/root/reference/src/UCF_VIT/dataloaders/transform.py:33-34), so that the oracle (oracle/canny_np.py) and the CUDA kernels
stay pinned where cv2 is not importable.   python oracle/gen_golden_canny.py   (needs opencv-python; 4.13.0 here)"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def images():
    rng = np.random.default_rng(20260)
    yy, xx = np.mgrid[0:96, 0:120]
    smooth = np.stack([127 + 120 * np.sin(xx / 9.0) * np.cos(yy / 13.0), 255 * ((xx // 17 + yy // 23) % 2), 40 + xx * 1.5], -1)
    return {
        "noise_rgb": (rng.random((61, 83, 3)) * 255).astype(np.uint8),
        "noise_gray": (rng.random((70, 66)) * 255).astype(np.uint8),
        "smooth_rgb": np.clip(smooth, 0, 255).astype(np.uint8),
        "blobs_rgba": cv2.GaussianBlur((rng.random((90, 77, 4)) * 255).astype(np.uint8), (15, 15), 0),
        "tiny": (rng.random((3, 5, 3)) * 255).astype(np.uint8),
    }


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, img in images().items():
        out[f"{name}/img"] = img
        for k in (1, 3, 5):
            b = cv2.GaussianBlur(img, (k, k), 0).reshape(img.shape)
            out[f"{name}/blur{k}"] = b
            for lo in (50, 77, 99):
                out[f"{name}/canny{k}_{lo}"] = cv2.Canny(b, lo, lo + 50)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "front_end", "canny_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
