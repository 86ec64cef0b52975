"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

Numpy / pure-Python restatement of the reference's adaptive-patching (SAP) data path:
  * greedy quadtree / octree construction      dataloaders/quadtree.py:115-137, octree.py:72-102
  * serialize  (leaf crop -> p x p resample)    quadtree.py:144-174, octree.py:104-150
  * deserialize (patch -> leaf box resample)    quadtree.py:209-221 + Rect.set_area :25-36,
                                                octree.py:201-213 + Cube.set_area :28-55
(paths relative to /root/reference/src/UCF_VIT/).  Integer results (node boxes, order, sizes,
centres) must be reproduced bit-for-bit by the product; resampled pixels within tolerance.

The 2-D resampler is OpenCV's `cv.resize(..., INTER_CUBIC)` (third-party, `opencv-python`,
version unpinned by the reference's pyproject.toml:22; cv2 4.13 in this image) and the 3-D one is
scipy's `RegularGridInterpolator` (linear / nearest).  Their published algorithms are restated
below; oracle/gen_golden_sap.py pins this file against the reference classes calling the real
cv2 / scipy in the build container.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# tree construction (list semantics of the reference: first maximum wins, children replace the
# parent in place, stop when the selected leaf is 2 wide)
# ------------------------------------------------------------------------------------------------
def _contains2d(domain, x1, x2, y1, y2, norm=255):
    return int(np.sum(domain[y1:y2, x1:x2]) / norm)


def build_quadtree(domain, fixed_length):
    """-> list of (x1, x2, y1, y2, value) in the reference's node order."""
    h, w = domain.shape
    nodes = [(0, w, 0, h, _contains2d(domain, 0, w, 0, h))]
    while len(nodes) < fixed_length:
        best = 0
        for i in range(1, len(nodes)):          # first maximum (max() + list.index())
            if nodes[i][4] > nodes[best][4]:
                best = i
        x1, x2, y1, y2, _ = nodes[best]
        if x2 - x1 == 2:
            break
        xm, ym = int((x1 + x2) / 2), int((y1 + y2) / 2)
        kids = [(x1, xm, ym, y2), (xm, x2, ym, y2), (x1, xm, y1, ym), (xm, x2, y1, ym)]   # lt, rt, lb, rb
        kids = [k + (_contains2d(domain, *k),) for k in kids]
        nodes = nodes[:best] + kids + nodes[best + 1:]
    return nodes


def _contains3d(domain, c, norm):
    x1, x2, y1, y2, z1, z2 = c
    return int(np.sum(domain[z1:z2, y1:y2, x1:x2]) / norm)


def build_octree(domain, fixed_length, norm_factor=255):
    """-> list of (x1, x2, y1, y2, z1, z2, value).  NB the reference's root is Cube(0,h,0,w,0,d)
    while `contains` indexes domain[z, y, x]: consistent for cubic domains only (asserted)."""
    h, w, d = domain.shape
    assert h == w == d, "the reference's octree is only self-consistent for cubic tiles"
    root = (0, h, 0, w, 0, d)
    nodes = [root + (_contains3d(domain, root, norm_factor),)]
    while len(nodes) < fixed_length:
        best = 0
        for i in range(1, len(nodes)):
            if nodes[i][6] > nodes[best][6]:
                best = i
        x1, x2, y1, y2, z1, z2, _ = nodes[best]
        if x2 - x1 == 2:
            break
        xm, ym, zm = int((x1 + x2) / 2), int((y1 + y2) / 2), int((z1 + z2) / 2)
        kids = []
        for zz in ((z1, zm), (zm, z2)):             # x fastest, then y, then z
            for yy in ((y1, ym), (ym, y2)):
                for xx in ((x1, xm), (xm, x2)):
                    c = xx + yy + zz
                    kids.append(c + (_contains3d(domain, c, norm_factor),))
        nodes = nodes[:best] + kids + nodes[best + 1:]
    return nodes


# ------------------------------------------------------------------------------------------------
# OpenCV INTER_CUBIC (imgproc/src/resize.cpp: interpolateCubic, HResizeCubic, VResizeCubic)
# ------------------------------------------------------------------------------------------------
def _cubic_coeffs(fx):
    """Keys kernel with A = -0.75 evaluated in float32 like OpenCV."""
    A = np.float32(-0.75)
    x = np.float32(fx)
    one = np.float32(1)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def _axis_table(src, dst):
    """Per destination index: first tap (sx - 1) and the 4 float32 coefficients."""
    scale = 1.0 / (dst / src)            # double, as in cv::resize
    taps, coef = [], []
    for d in range(dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        taps.append(sx - 1)
        coef.append(_cubic_coeffs(fx))
    return taps, coef


def resize_cubic(img, dst_w, dst_h):
    """cv.resize(img, (dst_w, dst_h), interpolation=cv.INTER_CUBIC) for HxW[xC] float32 or uint8.
    float32: separable float32 filtering, border taps replicated.
    uint8  : OpenCV's fixed-point path -- coefficients rounded to 1/2048, two passes in int32,
             result (v + 2^21) >> 22 saturated to [0, 255]."""
    a = np.asarray(img)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    H, W, C = a.shape
    if (W, H) == (dst_w, dst_h):
        out = a.copy()
        return out[:, :, 0] if squeeze else out
    xt, xc = _axis_table(W, dst_w)
    yt, yc = _axis_table(H, dst_h)
    if a.dtype == np.uint8:
        xi = [np.clip(np.rint(c.astype(np.float64) * 2048), -32768, 32767).astype(np.int64) for c in xc]
        yi = [np.clip(np.rint(c.astype(np.float64) * 2048), -32768, 32767).astype(np.int64) for c in yc]
        src = a.astype(np.int64)
        tmp = np.zeros((H, dst_w, C), dtype=np.int64)
        for d in range(dst_w):
            for k in range(4):
                tmp[:, d] += src[:, min(max(xt[d] + k, 0), W - 1)] * xi[d][k]
        out = np.zeros((dst_h, dst_w, C), dtype=np.int64)
        for d in range(dst_h):
            for k in range(4):
                out[d] += tmp[min(max(yt[d] + k, 0), H - 1)] * yi[d][k]
        out = np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    else:
        src = a.astype(np.float32)
        tmp = np.zeros((H, dst_w, C), dtype=np.float32)
        for d in range(dst_w):
            acc = np.zeros((H, C), dtype=np.float32)
            for k in range(4):
                acc = acc + src[:, min(max(xt[d] + k, 0), W - 1)] * xc[d][k]
            tmp[:, d] = acc
        out = np.zeros((dst_h, dst_w, C), dtype=np.float32)
        for d in range(dst_h):
            acc = np.zeros((dst_w, C), dtype=np.float32)
            for k in range(4):
                acc = acc + tmp[min(max(yt[d] + k, 0), H - 1)] * yc[d][k]
            out[d] = acc
    return out[:, :, 0] if squeeze else out


# ------------------------------------------------------------------------------------------------
# serialize / deserialize, 2-D
# ------------------------------------------------------------------------------------------------
def serialize2d(nodes, img, p, fixed_length):
    """FixedQuadTree.serialize + the dtype handling of Patchify.forward (transform.py:40-48):
    -> seq_img float32 [L, p, p, C], seq_size int64 [L], seq_pos float64 [L, 2] (x, y centres);
    padding entries: zero patch, size 0, pos (-1, -1)."""
    C = img.shape[2]
    seq = np.zeros((fixed_length, p, p, C), dtype=np.float32)
    size = np.zeros((fixed_length,), dtype=np.int64)
    pos = np.full((fixed_length, 2), -1.0, dtype=np.float64)
    for i, (x1, x2, y1, y2, _) in enumerate(nodes):
        crop = img[y1:y2, x1:x2, :]
        r = resize_cubic(crop, p, p)
        seq[i] = np.asarray(r, dtype=np.float32).reshape(p, p, C)
        size[i] = x2 - x1
        pos[i] = ((x2 + x1) / 2, (y2 + y1) / 2)
    return seq, size, pos


def deserialize2d(nodes, seq, p, C, H, W):
    """FixedQuadTree.deserialize: patches truncated to int, resized to the leaf box, pasted."""
    seq = np.reshape(seq, (-1, p, p, C)).astype(int)
    mask = np.zeros((H, W, C), dtype=np.float64)
    for i, (x1, x2, y1, y2, _) in enumerate(nodes):
        patch = seq[i].astype('float32')
        r = resize_cubic(patch, x2 - x1, y2 - y1)
        mask[y1:y2, x1:x2, :] = np.reshape(r, (y2 - y1, x2 - x1, C))
    return mask


# ------------------------------------------------------------------------------------------------
# serialize / deserialize, 3-D (scipy RegularGridInterpolator, method='linear', grid
# linspace(0, s, s) -> query linspace(0, s, p): i.e. align-corners trilinear)
# ------------------------------------------------------------------------------------------------
def _lin_axis(src, dst):
    """index / weight pairs of 1-D align-corners linear interpolation (scipy's find-interval rule)."""
    if src == 1:
        return np.zeros(dst, dtype=np.int64), np.zeros(dst)
    grid = np.linspace(0, src, src)
    q = np.linspace(0, src, dst)
    idx = np.clip(np.searchsorted(grid, q, side='right') - 1, 0, src - 2)
    w = (q - grid[idx]) / (grid[idx + 1] - grid[idx])
    return idx.astype(np.int64), w


def resize_trilinear(vol, p):
    """vol [s, s, s, C] -> [p, p, p, C] float64."""
    s = vol.shape[0]
    v = vol.astype(np.float64)
    i0, w0 = _lin_axis(s, p)
    j = np.minimum(i0 + 1, s - 1)
    a = v[i0] * (1 - w0)[:, None, None, None] + v[j] * w0[:, None, None, None]
    a = a[:, i0] * (1 - w0)[None, :, None, None] + a[:, j] * w0[None, :, None, None]
    a = a[:, :, i0] * (1 - w0)[None, None, :, None] + a[:, :, j] * w0[None, None, :, None]
    return a


def serialize3d(nodes, vol, p, fixed_length):
    """FixedOctTree.serialize: -> float32 [L, p, p, p, C], int64 [L], float64 [L, 3]."""
    C = vol.shape[3]
    seq = np.zeros((fixed_length, p, p, p, C), dtype=np.float32)
    size = np.zeros((fixed_length,), dtype=np.int64)
    pos = np.full((fixed_length, 3), -1.0, dtype=np.float64)
    for i, (x1, x2, y1, y2, z1, z2, _) in enumerate(nodes):
        crop = vol[z1:z2, y1:y2, x1:x2, :]
        seq[i] = resize_trilinear(crop, p).astype(np.float32)
        size[i] = x2 - x1
        pos[i] = ((x2 + x1) / 2, (y2 + y1) / 2, (z2 + z1) / 2)
    return seq, size, pos


def deserialize3d(nodes, seq, p, C, shape):
    H, W, D = shape
    seq = np.reshape(seq, (-1, p, p, p, C))
    mask = np.zeros((H, W, D, C), dtype=np.float64)
    for i, (x1, x2, y1, y2, z1, z2, _) in enumerate(nodes):
        mask[z1:z2, y1:y2, x1:x2, :] = resize_trilinear(seq[i], x2 - x1)
    return mask
