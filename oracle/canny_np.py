"""TEST INFRASTRUCTURE ONLY (never imported by the product): numpy restatement of the two OpenCV calls the reference's SAP
front end makes on natural (uint8) images,

    grey_img = cv.GaussianBlur(img, (k, k), 0)            /root/reference/src/UCF_VIT/dataloaders/transform.py:33
    edges    = cv.Canny(grey_img, c, c + 50)              /root/reference/src/UCF_VIT/dataloaders/transform.py:34

OpenCV (`opencv-python`, a dependency of the reference that is not vendored under /root/reference; 4.13.0 in this image)
is the algorithm's home, so this restates its published behaviour:

* GaussianBlur, 8-bit, sigma 0, k in {1, 3, 5}: the fixed small-kernel table ([1], [1 2 1]/4, [1 4 6 4 1]/16), separable,
  BORDER_REFLECT_101, evaluated in fixed point with one round-half-up at the end -- i.e. exactly
  (sum_ij w_i w_j p_ij + half) >> shift on integers.
* Canny, aperture 3, L1 gradient: 3x3 Sobel with BORDER_REPLICATE per channel, per pixel the channel with the largest
  |dx| + |dy| (first one on ties), non-maximum suppression with the integer tan(22.5 deg) test, double threshold
  (m > floor(high): strong, m > floor(low): candidate), hysteresis over 8-neighbours.

PINNED: tests/test_canny_oracle.py checks both functions bit-for-bit against cv2 itself (present in the image) on random
and structured images of ragged sizes, and against the committed fixture tests/golden/front_end/canny_cv2.npz (made by
oracle/gen_golden_canny.py)."""
import numpy as np

_TAPS = {1: np.array([1], np.int64), 3: np.array([1, 2, 1], np.int64), 5: np.array([1, 4, 6, 4, 1], np.int64)}


def _reflect101(n, r):
    idx = np.arange(-r, n + r)
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.abs(idx) % period
    return np.where(idx >= n, period - idx, idx)


def gaussian_blur_u8(img, k):
    """cv.GaussianBlur(img, (k, k), 0) for uint8 img [H, W] or [H, W, C], k in {1, 3, 5}."""
    assert img.dtype == np.uint8 and k in _TAPS
    if k == 1:
        return img.copy()
    w = _TAPS[k]
    r = k // 2
    x = img.astype(np.int64)
    H, W = x.shape[:2]
    rows = _reflect101(H, r)
    cols = _reflect101(W, r)
    acc = np.zeros_like(x)
    for i in range(k):
        for j in range(k):
            acc += w[i] * w[j] * x[rows[i:i + H]][:, cols[j:j + W]]
    shift = 2 * int(np.log2(w.sum()))
    return ((acc + (1 << (shift - 1))) >> shift).astype(np.uint8)


def _sobel_replicate(x):
    """3x3 Sobel dx, dy (int32) of an int32 [H, W] plane with BORDER_REPLICATE."""
    p = np.pad(x, 1, mode="edge")
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    return dx, dy


TG22 = int(0.4142135623730950488016887242097 * (1 << 15) + 0.5)


def canny_nms(img, low, high):
    """The map before hysteresis: 2 = strong edge, 0 = candidate (weak, survives suppression), 1 = no edge.  Returns
    (pmap uint8 [H, W], magnitude int32 [H, W])."""
    assert img.dtype == np.uint8
    if img.ndim == 2:
        img = img[:, :, None]
    low, high = int(np.floor(min(low, high))), int(np.floor(max(low, high)))
    H, W, C = img.shape
    dx = np.zeros((H, W), np.int32)
    dy = np.zeros((H, W), np.int32)
    mag = np.full((H, W), -1, np.int32)
    for c in range(C):
        cx, cy = _sobel_replicate(img[:, :, c].astype(np.int32))
        m = np.abs(cx) + np.abs(cy)
        take = m > mag                      # strictly larger: the first channel wins ties
        dx = np.where(take, cx, dx)
        dy = np.where(take, cy, dy)
        mag = np.where(take, m, mag)
    mp = np.pad(mag, 1)                     # magnitude is 0 outside the image
    ctr = mp[1:-1, 1:-1]
    x = np.abs(dx).astype(np.int64)
    y = np.abs(dy).astype(np.int64) << 15
    tg22x = x * TG22
    tg67x = tg22x + (x << 16)
    horiz = y < tg22x
    vert = ~horiz & (y > tg67x)
    diag = ~horiz & ~vert
    left, right = mp[1:-1, :-2], mp[1:-1, 2:]
    up, down = mp[:-2, 1:-1], mp[2:, 1:-1]
    neg = (dx ^ dy) < 0                    # gradient components of opposite sign: s = -1, else s = +1
    # previous row at column j - s, next row at column j + s
    p_diag = np.where(neg, mp[:-2, 2:], mp[:-2, :-2])
    n_diag = np.where(neg, mp[2:, :-2], mp[2:, 2:])
    keep = (horiz & (ctr > left) & (ctr >= right)) | (vert & (ctr > up) & (ctr >= down)) | \
           (diag & (ctr > p_diag) & (ctr > n_diag))
    keep &= ctr > low
    pmap = np.ones((H, W), np.uint8)
    pmap[keep] = 0
    pmap[keep & (ctr > high)] = 2
    return pmap, mag


def hysteresis(pmap):
    """255 where a candidate (0) is 8-connected to a strong pixel (2) through candidates; strong pixels included."""
    H, W = pmap.shape
    p = np.pad(pmap, 1, constant_values=1).copy()
    stack = [tuple(ix) for ix in np.argwhere(p == 2)]
    while stack:
        i, j = stack.pop()
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                if p[i + di, j + dj] == 0:
                    p[i + di, j + dj] = 2
                    stack.append((i + di, j + dj))
    return np.where(p[1:-1, 1:-1] == 2, 255, 0).astype(np.uint8)


def canny_u8(img, low, high):
    """cv.Canny(img, low, high) for uint8 img [H, W] or [H, W, C] (aperture 3, L2gradient=False)."""
    return hysteresis(canny_nms(img, low, high)[0])
