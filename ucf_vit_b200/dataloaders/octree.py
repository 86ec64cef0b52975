"""Adaptive-patching octree with the reference's class API
(/root/reference/src/UCF_VIT/dataloaders/octree.py: Cube :6-66, FixedOctTree :68-213).
Host C++ tree build + GPU trilinear (align-corners) gather / scatter; see quadtree.py."""
import numpy as np
import torch

from .. import ops
from .quadtree import _as_device_image


class Cube:
    def __init__(self, x1, x2, y1, y2, z1, z2) -> None:
        assert x1 <= x2 and y1 <= y2 and z1 <= z2, 'wrong coordinate.'
        self.x1, self.x2, self.y1, self.y2, self.z1, self.z2 = x1, x2, y1, y2, z1, z2

    def contains(self, domain, norm_factor):
        return int(np.sum(domain[self.z1:self.z2, self.y1:self.y2, self.x1:self.x2]) / norm_factor)

    def get_area(self, img):
        return img[self.z1:self.z2, self.y1:self.y2, self.x1:self.x2, :]

    def get_coord(self):
        return self.x1, self.x2, self.y1, self.y2, self.z1, self.z2

    def get_size(self):
        return self.x2 - self.x1, self.y2 - self.y1, self.z2 - self.z1

    def get_center(self):
        return (self.x2 + self.x1) / 2, (self.y2 + self.y1) / 2, (self.z2 + self.z1) / 2


def _nearest_grid_index(src, dst):
    """scipy RegularGridInterpolator(method='nearest') on the reference's grids (octree.py:167-185): source points
    linspace(0, src, src), queries linspace(0, src, dst); ties (normalised distance exactly 0.5) go to the lower index."""
    if src == 1:
        return np.zeros(dst, dtype=np.int64)
    grid = np.linspace(0, src, src)
    q = np.linspace(0, src, dst)
    lo = np.clip(np.searchsorted(grid, q) - 1, 0, src - 2)
    frac = (q - grid[lo]) / (grid[lo + 1] - grid[lo])
    return np.where(frac <= 0.5, lo, lo + 1).astype(np.int64)


class FixedOctTree:
    def __init__(self, domain, fixed_length=128, norm_factor=255, device="cuda") -> None:
        self.domain, self.fixed_length, self.norm_factor, self.device = domain, fixed_length, norm_factor, device
        self._boxes_dev = None
        self._build_tree()

    def _build_tree(self):
        h, w, d = self.domain.shape
        assert h > 0 and w > 0 and d > 0, "Wrong img size."
        self.boxes, values = ops.sap_build_tree(np.asarray(self.domain), self.fixed_length, float(self.norm_factor))
        self.nodes = [[Cube(*[int(v) for v in b]), int(val)] for b, val in zip(self.boxes, values)]

    def _dev_boxes(self):
        if self._boxes_dev is None:
            self._boxes_dev = torch.from_numpy(np.ascontiguousarray(self.boxes, dtype=np.int32)).to(self.device)
        return self._boxes_dev

    def serialize_device(self, img, size=(8, 8, 8, 1)):
        h2, w2, d2, c2 = size
        assert h2 == w2 == d2
        t = _as_device_image(img, self.device).float()
        # like the reference's serialize: a fixed_length that the last split overshoots (!= 1 mod 3 / mod 7) is an error
        assert len(self.nodes) <= self.fixed_length, "Not equal fixed legnth."
        return ops.sap_gather(t, self._dev_boxes(), self.fixed_length, h2)

    def deserialize_device(self, seq, patch_size, channel):
        s = seq if torch.is_tensor(seq) else torch.from_numpy(np.asarray(seq, dtype=np.float32))
        s = s.to(self.device).float().reshape(self.fixed_length, patch_size, patch_size, patch_size, channel)
        return ops.sap_scatter(s, self._dev_boxes(), tuple(self.domain.shape), patch_size, channel)

    def serialize_labels(self, img, size=(8, 8, 8, 1)):
        """FixedOctTree.serialize_labels, octree.py:152-199: nearest-neighbour resampling of every leaf of the label
        volume (scipy RegularGridInterpolator(method='nearest') in the reference) as an integer gather on the host."""
        h2, w2, d2, c2 = size
        assert len(self.boxes) <= self.fixed_length, "Not equal fixed legnth."
        img = np.asarray(img)
        patches, sizes, pos = [], [], []
        for x1, x2, y1, y2, z1, z2 in np.asarray(self.boxes).tolist():
            h1, w1, d1 = z2 - z1, y2 - y1, x2 - x1
            assert h1 == w1 == d1, "Need squared input."
            zi = z1 + _nearest_grid_index(h1, h2)
            yi = y1 + _nearest_grid_index(w1, w2)
            xi = x1 + _nearest_grid_index(d1, d2)
            patches.append(img[zi[:, None, None], yi[None, :, None], xi[None, None, :], :].astype(np.float64))
            sizes.append(x2 - x1)
            pos.append(((x2 + x1) / 2, (y2 + y1) / 2, (z2 + z1) / 2))
        pad = self.fixed_length - len(patches)
        if pad > 0:
            patches += [np.zeros(shape=(h2, w2, d2, c2))] * pad
            sizes += [0] * pad
            pos += [(-1, -1, -1)] * pad
        return patches, sizes, pos

    def serialize(self, img, size=(8, 8, 8, 1)):
        seq, ssize, spos = self.serialize_device(img, size)
        seq = seq.cpu().numpy().astype(np.float64)
        return ([seq[i] for i in range(self.fixed_length)], [int(v) for v in ssize.cpu().tolist()],
                [tuple(p) for p in spos.cpu().tolist()])

    def deserialize(self, seq, patch_size, channel):
        return self.deserialize_device(seq, patch_size, channel).cpu().numpy().astype(np.float64)
