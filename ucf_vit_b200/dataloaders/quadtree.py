"""Adaptive-patching quadtree with the reference's class API
(/root/reference/src/UCF_VIT/dataloaders/quadtree.py: Rect :6-82, FixedQuadTree :84-242).

`nodes` keeps the reference's `[[Rect, value], ...]` list (same order, same integers).  The tree
is built by the C++ host routine `ucf_sap_build_tree_host` (8x8-cell summed-area table + priority queue,
O(L log L) instead of the reference's O(L^2) Python scans) and `serialize` / `deserialize`
resample every leaf on the GPU (`ucf_sap_gather` / `ucf_sap_scatter`) instead of one
`cv.resize` call per leaf.  Drawing helpers (matplotlib) are not part of the hot path and are
not provided."""
import numpy as np
import torch

from .. import ops


class Rect:
    def __init__(self, x1, x2, y1, y2) -> None:
        assert x1 <= x2, 'x1 > x2, wrong coordinate.'
        assert y1 <= y2, 'y1 > y2, wrong coordinate.'
        self.x1, self.x2, self.y1, self.y2 = x1, x2, y1, y2

    def contains(self, domain):
        return int(np.sum(domain[self.y1:self.y2, self.x1:self.x2]) / 255)

    def get_area(self, img):
        return img[self.y1:self.y2, self.x1:self.x2, :]

    def get_coord(self):
        return self.x1, self.x2, self.y1, self.y2

    def get_size(self):
        return self.x2 - self.x1, self.y2 - self.y1

    def get_center(self):
        return (self.x2 + self.x1) / 2, (self.y2 + self.y1) / 2

    def set_area(self, mask, patch):
        """Paste `patch` ([p, p, C]), cubic-resampled to this box, into `mask` ([H, W, C] numpy, modified in place and returned):
        Rect.set_area, quadtree.py:25-36 (cv.resize INTER_CUBIC of the float32 patch), on the device scatter kernel."""
        p_ = np.asarray(patch, dtype=np.float32)
        if p_.ndim == 2:
            p_ = p_[:, :, None]
        C = p_.shape[2]
        seq = torch.from_numpy(np.ascontiguousarray(p_)).cuda().reshape(1, p_.shape[0], p_.shape[1], C)
        box = torch.tensor([[self.x1, self.x2, self.y1, self.y2]], dtype=torch.int32, device="cuda")
        full = ops.sap_scatter(seq, box, (mask.shape[0], mask.shape[1]), p_.shape[0], C)
        mask[self.y1:self.y2, self.x1:self.x2, :] = full[self.y1:self.y2, self.x1:self.x2, :].cpu().numpy()
        return mask

    def __eq__(self, other):
        return isinstance(other, Rect) and self.get_coord() == other.get_coord()

    def __hash__(self):
        return hash(self.get_coord())


def _nearest_index(src, dst):
    """Source index of every destination sample of cv.resize(..., interpolation=cv.INTER_NEAREST): floor(x * (1 / (dst / src)))
    in double, clamped to src - 1 (OpenCV resizeNN; the same rule as torch's legacy 'nearest' mode)."""
    inv = 1.0 / (np.float64(dst) / np.float64(src))
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * inv).astype(np.int64), src - 1)


def _as_device_image(img, device):
    t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
    if t.dtype not in (torch.uint8, torch.float32):
        t = t.float()
    return t.to(device).contiguous()


class FixedQuadTree:
    def __init__(self, domain, fixed_length=128, build_from_info=False, meta_info=None, device="cuda") -> None:
        self.domain = domain
        self.fixed_length = fixed_length
        self.device = device
        self._boxes_dev = None
        if build_from_info:
            self.nodes = self.decoder_nodes(meta_info=meta_info)
            self.boxes = np.array([r.get_coord() for r, _ in self.nodes], dtype=np.int32).reshape(-1, 4)
        else:
            self._build_tree()

    def _build_tree(self):
        h, w = self.domain.shape
        assert h > 0 and w > 0, "Wrong img size."
        self._adopt(*ops.sap_build_tree(np.asarray(self.domain), self.fixed_length, 255.0))

    def _adopt(self, boxes, values):
        self.boxes = boxes
        self._values = values
        self._nodes = None            # the reference's `nodes` list ([[Rect, value], ...]) is materialised on first access:
                                      # 4096 Python objects per tree that the training path (boxes -> device gather) never reads

    @property
    def nodes(self):
        if self._nodes is None:
            self._nodes = [[Rect(int(b[0]), int(b[1]), int(b[2]), int(b[3])), int(v)] for b, v in zip(self.boxes, self._values)]
        return self._nodes

    @nodes.setter
    def nodes(self, value):
        self._nodes = value

    _pool = None

    @classmethod
    def build_many_async(cls, domains, fixed_length=128, device="cuda", threads=0):
        """`build_many` on a background host thread (the C++ builder runs without the GIL): returns a
        concurrent.futures.Future whose result() is the list of trees.  A loader submits batch k+1 here while the GPU
        works on batch k, which takes the ~5 ms per 4096^2 image of integer work off the step's critical path."""
        from concurrent.futures import ThreadPoolExecutor
        if cls._pool is None:
            cls._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="ucf_sap_tree")
        return cls._pool.submit(cls.build_many, domains, fixed_length, device, threads)

    @classmethod
    def build_many(cls, domains, fixed_length=128, device="cuda", threads=0):
        """Trees of several edge maps at once: maps of one shape and dtype are built together on host threads
        (`ucf_sap_build_tree_batch_host`).  Returns one FixedQuadTree per map, in order -- identical to
        `[FixedQuadTree(d, fixed_length, device=device) for d in domains]`."""
        doms = [np.asarray(d) for d in domains]
        groups = {}
        for i, d in enumerate(doms):
            groups.setdefault((d.shape, d.dtype.str), []).append(i)
        trees = [None] * len(doms)
        for idx in groups.values():
            built = ops.sap_build_trees([doms[i] for i in idx], fixed_length, 255.0, threads=threads)
            for i, (boxes, values) in zip(idx, built):
                t = cls.__new__(cls)
                t.domain, t.fixed_length, t.device, t._boxes_dev = domains[i], fixed_length, device, None
                h, w = doms[i].shape
                assert h > 0 and w > 0, "Wrong img size."
                t._adopt(boxes, values)
                trees[i] = t
        return trees

    # ---- bookkeeping helpers with reference semantics
    def nodes_value(self):
        return [[r.get_size()[0] / 8] for r, _ in self.nodes]

    def encode_nodes(self):
        return [[r.x1, r.x2, r.y1, r.y2] for r, _ in self.nodes]

    def decoder_nodes(self, meta_info):
        out = []
        for x1, x2, y1, y2 in meta_info:
            r = Rect(x1, x2, y1, y2)
            out.append([r, r.contains(self.domain)])
        return out

    def count_patches(self):
        return len(self.boxes)

    def _dev_boxes(self):
        if self._boxes_dev is None:
            self._boxes_dev = torch.from_numpy(np.ascontiguousarray(self.boxes, dtype=np.int32)).to(self.device)
        return self._boxes_dev

    # ---- device path
    def serialize_device(self, img, size=(8, 8, 3)):
        """img: numpy / torch [H, W, C] (uint8 or float32).  -> cuda tensors
        seq [L, p, p, C] f32, seq_size [L] i64, seq_pos [L, 2] f64."""
        h2, w2, c2 = size
        assert h2 == w2, "square target patches only"
        t = _as_device_image(img, self.device)
        assert t.dim() == 3 and t.shape[2] == c2
        # like the reference's serialize: a fixed_length that the last split overshoots (!= 1 mod 3 / mod 7) is an error
        assert len(self.boxes) <= self.fixed_length, "Not equal fixed legnth."
        return ops.sap_gather(t, self._dev_boxes(), self.fixed_length, h2)

    def deserialize_device(self, seq, patch_size, channel):
        H, W = self.domain.shape
        s = seq if torch.is_tensor(seq) else torch.from_numpy(np.asarray(seq, dtype=np.float32))
        s = s.to(self.device).float().reshape(self.fixed_length, patch_size, patch_size, channel)
        return ops.sap_scatter(s, self._dev_boxes(), (H, W), patch_size, channel, truncate_to_int=True)

    def serialize_labels(self, img, size=(8, 8, 3)):
        """FixedQuadTree.serialize_labels, quadtree.py:176-207: every leaf of the label image resampled with
        cv.INTER_NEAREST to `size` (labels must not be interpolated).  Index arithmetic on the host (integer gather, a few
        KB per image); returns the reference's three lists.  A single-channel image yields 2-D patches, like cv.resize."""
        h2, w2, c2 = size
        assert len(self.boxes) <= self.fixed_length, "Not equal fixed legnth."
        img = np.asarray(img)
        if img.ndim == 2:
            img = img[:, :, None]
        patches, sizes, pos = [], [], []
        for x1, x2, y1, y2 in np.asarray(self.boxes).tolist():
            h1, w1 = y2 - y1, x2 - x1
            assert h1 == w1, "Need squared input."
            # cv.resize(patch, (h2, w2)): dsize = (width, height) -> w2 rows, h2 columns
            ys = y1 + _nearest_index(h1, w2)
            xs = x1 + _nearest_index(w1, h2)
            pt = img[ys[:, None], xs[None, :], :]
            patches.append(pt[:, :, 0] if img.shape[2] == 1 else pt)
            sizes.append(x2 - x1)
            pos.append(((x2 + x1) / 2, (y2 + y1) / 2))
        pad = self.fixed_length - len(patches)
        if pad > 0:
            patches += [np.zeros(shape=(h2, w2, c2)) if c2 > 1 else np.zeros(shape=(h2, w2))] * pad
            sizes += [0] * pad
            pos += [(-1, -1)] * pad
        return patches, sizes, pos

    # ---- reference-typed API (lists / numpy back on the host)
    def serialize(self, img, size=(8, 8, 3)):
        seq, ssize, spos = self.serialize_device(img, size)
        c2 = size[2]
        seq = seq.cpu().numpy()
        patches = [seq[i] if c2 > 1 else seq[i, :, :, 0] for i in range(self.fixed_length)]
        return patches, [int(v) for v in ssize.cpu().tolist()], [tuple(p) for p in spos.cpu().tolist()]

    def deserialize(self, seq, patch_size, channel):
        return self.deserialize_device(seq, patch_size, channel).cpu().numpy().astype(np.float64)
