"""SAP adaptive patching.
CPU part : oracle (numpy restatement) and the product's C++ host tree builder against golden
           vectors produced by the reference FixedQuadTree / FixedOctTree -- node boxes, order,
           values, sizes and centres are BIT-EXACT.
GPU part : gather / scatter kernels against the same goldens (pixels: float32 within 2e-5 of
           OpenCV INTER_CUBIC / scipy linear; uint8 within 1 LSB of OpenCV's fixed-point path)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import quadtree_np as Q
from ucf_vit_b200 import ops

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "sap_tree_*.npz")))


def _load(name):
    cfg, _, a = fx.load_case(os.path.join(GOLDEN, name + ".npz"))
    n = cfg["size"]
    shape = (n, n) if cfg["kind"] == "quadtree" else (n, n, n)
    dom = (np.unpackbits(a["domain"])[:int(np.prod(shape))].reshape(shape) * 255).astype(np.uint8)
    return cfg, a, dom


def test_appendix_e_known_answer():
    dom = np.zeros((16, 16), dtype=np.uint8)
    for r, c in [(1, 1), (2, 3), (3, 12), (9, 9), (10, 10), (9, 10), (14, 2)]:
        dom[r, c] = 255
    want = [(0, 8, 8, 16, 1), (8, 12, 12, 16, 0), (12, 16, 12, 16, 0), (8, 10, 10, 12, 0), (10, 12, 10, 12, 1),
            (8, 10, 8, 10, 1), (10, 12, 8, 10, 1), (12, 16, 8, 12, 0), (0, 8, 0, 8, 2), (8, 16, 0, 8, 1)]
    assert Q.build_quadtree(dom, 10) == want
    boxes, values = ops.sap_build_tree(dom, 10)
    assert [tuple(b) + (v,) for b, v in zip(boxes.tolist(), values.tolist())] == want
    img = np.arange(256, dtype=np.float32).reshape(16, 16, 1)
    seq, size, pos = Q.serialize2d(want, img, 2, 10)
    assert np.allclose(seq[0, :, :, 0], [[153.5, 157.5], [217.5, 221.5]])
    assert size.tolist() == [8, 4, 4, 2, 2, 2, 2, 4, 8, 8]
    assert pos[:3].tolist() == [[4, 12], [10, 14], [14, 14]]


@pytest.mark.parametrize("name", CASES)
def test_tree_oracle_and_host_builder_bit_exact(name):
    cfg, a, dom = _load(name)
    want = [tuple(int(v) for v in row) for row in a["nodes"]]
    if cfg["kind"] == "quadtree":
        assert Q.build_quadtree(dom, cfg["L"]) == want
    else:
        assert Q.build_octree(dom, cfg["L"]) == want
    boxes, values = ops.sap_build_tree(dom, cfg["L"])
    got = [tuple(b) + (v,) for b, v in zip(boxes.tolist(), values.tolist())]
    assert got == want
    for dt in (np.float32, np.float64):        # float edge maps take the double-precision SAT path
        boxes_f, values_f = ops.sap_build_tree(dom.astype(dt), cfg["L"])
        assert [tuple(b) + (v,) for b, v in zip(boxes_f.tolist(), values_f.tolist())] == want


def test_tree_edge_cases():
    # fixed_length 1: root only; empty edge map: ties resolved in list order; stop rule at 2 px
    z = np.zeros((8, 8), dtype=np.uint8)
    b, v = ops.sap_build_tree(z, 1)
    assert b.tolist() == [[0, 8, 0, 8]] and v.tolist() == [0]
    assert [tuple(x) for x in ops.sap_build_tree(z, 7)[0].tolist()] == [n[:4] for n in Q.build_quadtree(z, 7)]
    full = np.full((4, 4), 255, dtype=np.uint8)
    b, v = ops.sap_build_tree(full, 64)       # splits 4x4 -> four 2x2, then stops at the 2-px rule
    assert len(b) == 4 and [tuple(x) + (y,) for x, y in zip(b.tolist(), v.tolist())] == Q.build_quadtree(full, 64)
    with pytest.raises(RuntimeError):
        ops.sap_build_tree(np.zeros((4, 8, 8), dtype=np.uint8), 8)      # non-cubic octree domain


@pytest.mark.parametrize("shape,L", [((227, 203), 150), ((100, 260), 64), ((64, 64), 256), ((8, 8), 16), ((5, 9), 4),
                                     ((513, 1027), 1000), ((16, 16), 1000)])
def test_uint8_cell_table_gives_the_per_pixel_sums(shape, L):
    """The uint8 path (8 x 8 cell table + border strips, SIMD byte sums) against the oracle's per-box numpy sums and
    against the float64 path (per-pixel summed-area table, exact for integer data): same boxes, order and values
    on sizes whose box edges are not multiples of 8."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    dom = ((rng.random(shape) < 0.2) * 255).astype(np.uint8)
    dom[: shape[0] // 3, : shape[1] // 2] = 255                      # a dense corner: deep, unbalanced tree
    b8, v8 = ops.sap_build_tree(dom, L)
    b64, v64 = ops.sap_build_tree(dom.astype(np.float64), L)
    assert np.array_equal(b8, b64) and np.array_equal(v8, v64)
    nodes = Q.build_quadtree(dom, L)
    assert np.array_equal(b8, np.array([n[:4] for n in nodes], np.int32))
    assert np.array_equal(v8, np.array([n[4] for n in nodes], np.int64))
    for (x1, x2, y1, y2), v in zip(b8[:: max(1, len(b8) // 25)], v8[:: max(1, len(b8) // 25)]):
        assert v == int(dom[y1:y2, x1:x2].astype(np.int64).sum() / 255)


def test_batch_tree_builder_matches_single_calls_on_any_thread_count():
    rng = np.random.default_rng(5)
    maps = [((rng.random((96, 96)) < d) * 255).astype(np.uint8) for d in (0.01, 0.1, 0.3, 0.6, 0.9)]
    single = [ops.sap_build_tree(m, 40) for m in maps]
    for threads in (0, 1, 3, 16):
        got = ops.sap_build_trees(maps, 40, threads=threads)
        assert len(got) == len(single)
        for (b, v), (b0, v0) in zip(got, single):
            assert np.array_equal(b, b0) and np.array_equal(v, v0)
    vols = [rng.random((16, 16, 16)).astype(np.float32) * 255 for _ in range(3)]
    for (b, v), vol in zip(ops.sap_build_trees(vols, 22), vols):
        b0, v0 = ops.sap_build_tree(vol, 22)
        assert np.array_equal(b, b0) and np.array_equal(v, v0)
    assert ops.sap_build_trees([], 8) == []
    with pytest.raises(ValueError):
        ops.sap_build_trees([maps[0], maps[1][:50]], 8)
    with pytest.raises(RuntimeError, match="image 0 was rejected"):
        ops.sap_build_trees([np.zeros((8, 8), np.uint8)] * 2, 8, norm_factor=0.0)


@pytest.mark.parametrize("name", [c for c in CASES if "oct" not in c])
def test_oracle_serialize_matches_reference_2d(name):
    cfg, a, dom = _load(name)
    nodes = [tuple(int(v) for v in row) for row in a["nodes"]]
    seq, size, pos = Q.serialize2d(nodes, a["img"], cfg["p"], cfg["L"])
    assert np.array_equal(size, a["seq_size"]) and np.array_equal(pos, a["seq_pos"])
    tol = 1.0 if cfg["dtype"] == "uint8" else 2e-6
    assert np.abs(seq - a["seq_img"]).max() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gather_scatter_kernels(name):
    cfg, a, dom = _load(name)
    L, p, C = cfg["L"], cfg["p"], cfg["C"]
    boxes, _ = ops.sap_build_tree(dom, L)
    bdev = torch.from_numpy(boxes).cuda()
    if cfg["kind"] == "quadtree":
        img = torch.from_numpy(a["img"]).cuda()
        seq, size, pos = ops.sap_gather(img, bdev, L, p)
        assert np.array_equal(size.cpu().numpy(), a["seq_size"]), "sizes must be bit-exact"
        assert np.array_equal(pos.cpu().numpy(), a["seq_pos"]), "centres must be bit-exact"
        tol = 1.0 if cfg["dtype"] == "uint8" else 2e-5
        err = np.abs(seq.cpu().numpy() - a["seq_img"]).max()
        assert err <= tol, err
        if cfg["dtype"] == "uint8":
            assert (seq.cpu().numpy() == a["seq_img"]).mean() > 0.999
        scale = 1.0 if cfg["dtype"] == "uint8" else 255.0
        back = torch.from_numpy(a["seq_img"] * scale).cuda()
        mask = ops.sap_scatter(back, bdev, (cfg["size"], cfg["size"]), p, C, truncate_to_int=True)
        assert np.abs(mask.cpu().numpy() - a["mask"]).max() <= 2e-3
        # round trip property at leaf resolution == patch resolution: identity
        same = [i for i, b in enumerate(boxes) if b[1] - b[0] == p]
        if same and cfg["dtype"] == "float32":
            i = same[0]
            x1, x2, y1, y2 = boxes[i]
            assert np.array_equal(seq[i].cpu().numpy(), a["img"][y1:y2, x1:x2].astype(np.float32))
    else:
        vol = torch.from_numpy(a["vol"]).cuda()
        seq, size, pos = ops.sap_gather(vol, bdev, L, p)
        assert np.array_equal(size.cpu().numpy(), a["seq_size"]) and np.array_equal(pos.cpu().numpy(), a["seq_pos"])
        assert np.abs(seq.cpu().numpy() - a["seq_img"].astype(np.float32)).max() <= 1e-3   # fixture stored as fp16
        nodes = [tuple(int(v) for v in row) for row in a["nodes"]]
        o_seq, _, _ = Q.serialize3d(nodes, a["vol"], p, L)
        assert np.abs(seq.cpu().numpy() - o_seq).max() <= 2e-6
        mask = ops.sap_scatter(seq, bdev, (cfg["size"],) * 3, p, C)
        o_mask = Q.deserialize3d(nodes, o_seq, p, C, (cfg["size"],) * 3)
        assert np.abs(mask.cpu().numpy() - o_mask).max() <= 2e-5


@pytest.mark.gpu
def test_quadtree_class_api_and_large_image_partition():
    """Reference-typed API + size-independent property at the SAP benchmark scale: the leaves of a
    4096^2 tree tile the image exactly (areas sum to H*W, no overlaps)."""
    from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree
    rng = np.random.RandomState(0)
    dom = (rng.rand(4096, 4096) < 0.002).astype(np.uint8) * 255
    qt = FixedQuadTree(dom, fixed_length=1024)
    assert qt.count_patches() == 1024
    cover = np.zeros((4096, 4096), dtype=np.int32)
    for r, _ in qt.nodes:
        cover[r.y1:r.y2, r.x1:r.x2] += 1
    assert cover.min() == 1 and cover.max() == 1
    img = rng.randint(0, 256, (4096, 4096, 3)).astype(np.uint8)
    patches, sizes, centres = qt.serialize(img, size=(16, 16, 3))
    assert len(patches) == 1024 and patches[0].shape == (16, 16, 3) and sizes[0] == qt.nodes[0][0].get_size()[0]
    assert centres[0] == qt.nodes[0][0].get_center()


def _fake_cv2():
    """OpenCV is not installed in the build image: a small stand-in with the three calls the edge recipes use
    (not OpenCV's arithmetic -- enough to drive the host logic around them)."""
    import types
    from scipy import ndimage
    cv = types.ModuleType("cv2")
    cv.CV_64F = 6
    cv.Sobel = lambda a, ddepth, dx, dy, ksize=3: ndimage.sobel(a.astype(np.float64), axis=1 if dx else 0)
    cv.GaussianBlur = lambda a, k, s: ndimage.gaussian_filter(a, sigma=max(k[0], 1) / 3)

    def canny(a, lo, hi):                                    # multi-channel input -> one 2-D map, like OpenCV
        a = a.astype(np.float64)
        g = np.hypot(ndimage.sobel(a, axis=0), ndimage.sobel(a, axis=1))
        g = g.max(axis=2) if g.ndim == 3 else g
        return ((g > hi) * 255).astype(np.uint8)
    cv.Canny = canny
    return cv


def test_patchify_3d_edge_map_feeds_the_octree(monkeypatch):
    import sys
    from ucf_vit_b200.dataloaders.transform import Patchify_3D
    monkeypatch.setitem(sys.modules, "cv2", _fake_cv2())
    rng = np.random.default_rng(3)
    vol = rng.random((16, 16, 16, 2)).astype(np.float32)
    vol[4:12, 4:12, 4:12, :] += 2.0                                   # a block with sharp faces
    t = Patchify_3D(fixed_length=22, patch_size=4, num_channels=2)
    t.smooth_factor, t.canny = 1, [60, 110]
    edges = t._edges(vol)
    assert edges.dtype == np.uint8 and edges.shape == (16, 16, 16)
    assert t._norm_factor == 127 and set(np.unique(edges)) <= {0, 127, 254} and edges.any()
    boxes, values = ops.sap_build_tree(edges, 22, float(t._norm_factor))
    nodes = Q.build_octree(edges, 22, t._norm_factor)
    assert np.array_equal(boxes, np.array([n[:6] for n in nodes], np.int32))
    assert np.array_equal(values, np.array([n[6] for n in nodes], np.int64))


def test_patchify_2d_edge_branches(monkeypatch):
    import sys
    from ucf_vit_b200.dataloaders.transform import Patchify
    monkeypatch.setitem(sys.modules, "cv2", _fake_cv2())
    rng = np.random.default_rng(4)
    img = (rng.random((32, 32, 3)) * 255).astype(np.uint8)
    t = Patchify(fixed_length=16, patch_size=4, num_channels=3, dataset="imagenet")
    t.smooth_factor, t.canny = 3, [50, 100]
    e = t._edges(img)
    assert e.dtype == np.uint8 and e.shape == (32, 32)
    t.smooth_factor = 0                                   # the reference's "no smoothing" branch: a random float map
    np.random.seed(0)
    e0 = t._edges(img)
    assert e0.dtype == np.float64 and e0.shape == (32, 32) and 0.0 <= e0.min() and e0.max() <= 1.0
    boxes, values = ops.sap_build_tree(e0, 16)             # float64 path of the builder
    nodes = Q.build_quadtree(e0, 16)
    assert np.array_equal(boxes, np.array([n[:4] for n in nodes], np.int32))


def test_build_many_and_forward_batch_equal_the_per_image_path(monkeypatch):
    import random
    import sys
    from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree
    from ucf_vit_b200.dataloaders.transform import Patchify
    rng = np.random.default_rng(8)
    maps = [((rng.random((64, 64)) < 0.2) * 255).astype(np.uint8), rng.random((64, 64)),            # mixed dtypes and
            ((rng.random((32, 32)) < 0.5) * 255).astype(np.uint8), ((rng.random((64, 64)) < 0.7) * 255).astype(np.uint8)]
    many = FixedQuadTree.build_many(maps, 31, device="cpu", threads=3)
    for m, t in zip(maps, many):
        one = FixedQuadTree(m, 31, device="cpu")
        assert np.array_equal(t.boxes, one.boxes) and [v for _, v in t.nodes] == [v for _, v in one.nodes]
        assert t.encode_nodes() == one.encode_nodes() and t.count_patches() == one.count_patches()
    # the transform: same random draws, same trees, whether images go one by one or as a batch
    monkeypatch.setitem(sys.modules, "cv2", _fake_cv2())
    monkeypatch.setattr(FixedQuadTree, "serialize_device",
                        lambda self, img, size: (torch.zeros(self.fixed_length, size[0], size[1], size[2]),
                                                 torch.zeros(self.fixed_length, dtype=torch.int64),
                                                 torch.zeros(self.fixed_length, 2, dtype=torch.float64)))
    imgs = [(rng.random((32, 32, 3)) * 255).astype(np.uint8) for _ in range(5)]
    t = Patchify(fixed_length=16, patch_size=4, num_channels=3, dataset="imagenet", device="cpu")
    random.seed(11)
    np.random.seed(11)
    single = [t(img) for img in imgs]
    random.seed(11)
    np.random.seed(11)
    batch = t.forward_batch(imgs, threads=2)
    assert len(batch) == len(single)
    for a, b in zip(single, batch):
        assert np.array_equal(a[3].boxes, b[3].boxes) and a[0].shape == b[0].shape == (3, 16, 16)


def test_serialize_labels_nearest_2d_and_3d():
    """serialize_labels (reference quadtree.py:176-207, octree.py:152-199): label patches are resampled with nearest
    neighbour.  2-D witness: torch's legacy 'nearest' mode (the documented twin of cv.INTER_NEAREST: floor(dst * scale));
    3-D witness: scipy's RegularGridInterpolator(method='nearest') driven exactly as the reference drives it."""
    import torch
    from scipy.interpolate import RegularGridInterpolator
    from ucf_vit_b200.dataloaders.octree import FixedOctTree
    from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree, Rect
    rng = np.random.default_rng(5)
    edge = (rng.random((128, 128)) < 0.05).astype(np.uint8) * 255
    edge[:32, :32] = 255
    qdt = FixedQuadTree(edge, fixed_length=46, device="cpu")
    lab = rng.integers(0, 5, size=(128, 128, 3)).astype(np.uint8)
    patches, sizes, pos = qdt.serialize_labels(lab, size=(8, 8, 3))
    assert len(patches) == len(sizes) == len(pos) == 46
    for (r, _), pt, sz, ps in zip(qdt.nodes, patches, sizes, pos):
        src = torch.from_numpy(r.get_area(lab).astype(np.float32)).permute(2, 0, 1)[None]
        ref = torch.nn.functional.interpolate(src, size=(8, 8), mode="nearest")[0].permute(1, 2, 0).numpy()
        assert np.array_equal(pt.astype(np.float32), ref) and pt.dtype == lab.dtype
        assert sz == r.get_size()[0] and ps == r.get_center()
    one = qdt.serialize_labels(lab[:, :, :1], size=(4, 4, 1))[0]
    assert one[0].shape == (4, 4)                       # cv.resize drops a single channel axis
    short = FixedQuadTree(np.zeros((4, 4), np.uint8), fixed_length=7, device="cpu")      # four 2x2 leaves cannot split: 3 pad slots
    p2, s2, o2 = short.serialize_labels(np.ones((4, 4, 2), np.uint8), size=(2, 2, 2))
    assert len(p2) == 7 and s2 == [2, 2, 2, 2, 0, 0, 0] and o2[-1] == (-1, -1) and not p2[-1].any() and p2[0].all()
    assert hash(Rect(0, 4, 0, 4)) == hash(Rect(0, 4, 0, 4)) and len({Rect(0, 4, 0, 4), Rect(0, 4, 0, 4)}) == 1

    vol_edge = (rng.random((32, 32, 32)) < 0.03).astype(np.float32)
    oct_ = FixedOctTree(vol_edge, fixed_length=22, norm_factor=1, device="cpu")
    labv = rng.integers(0, 4, size=(32, 32, 32, 1)).astype(np.float64)
    patches, sizes, pos = oct_.serialize_labels(labv, size=(4, 4, 4, 1))
    assert len(patches) == 22
    for (c, _), pt in zip(oct_.nodes, patches):
        area = c.get_area(labv)
        h1 = area.shape[0]
        g1 = np.linspace(0, h1, h1)
        f = RegularGridInterpolator(points=[g1, g1, g1], values=area[:, :, :, 0], method="nearest")
        q = np.linspace(0, h1, 4)
        H, W, D = np.meshgrid(q, q, q, indexing="ij")
        ref = f(np.vstack([H.ravel(), W.ravel(), D.ravel()]).T).reshape(4, 4, 4)
        assert np.array_equal(pt[:, :, :, 0], ref)
