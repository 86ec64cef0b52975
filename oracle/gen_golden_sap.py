"""ORACLE -- TEST INFRASTRUCTURE ONLY.  SAP golden-vector generator (BUILD container only).

    python oracle/gen_golden_sap.py        # writes tests/golden/sap_tree_*.npz

Runs the UNMODIFIED reference FixedQuadTree / FixedOctTree (which call the real cv2 / scipy) on
deterministic edge maps and images, asserts that oracle/quadtree_np.py reproduces node boxes,
order, sizes and centres BIT-EXACTLY and the resampled pixels within tolerance, and stores
compact fixtures (edge map seeds, node tables, patch samples)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, "/root/reference/src")
sys.path.insert(2, ROOT)

import contextlib  # noqa: E402
import io  # noqa: E402

import numpy as np  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from oracle import quadtree_np as Q  # noqa: E402
from UCF_VIT.dataloaders.quadtree import FixedQuadTree  # noqa: E402
from UCF_VIT.dataloaders.octree import FixedOctTree  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def edge_map(h, w, seed, density):
    """Canny-like 0/255 uint8 map: a few random line segments plus salt noise."""
    rng = np.random.RandomState(seed)
    e = np.zeros((h, w), dtype=np.uint8)
    for _ in range(6):
        x0, y0, x1, y1 = rng.randint(0, w), rng.randint(0, h), rng.randint(0, w), rng.randint(0, h)
        n = max(abs(x1 - x0), abs(y1 - y0)) + 1
        xs = np.linspace(x0, x1, n).round().astype(int)
        ys = np.linspace(y0, y1, n).round().astype(int)
        e[ys, xs] = 255
    e[rng.rand(h, w) < density] = 255
    return e


def nodes_of(tree):
    return [tuple(int(v) for v in r.get_coord()) + (int(val),) for r, val in tree.nodes]


def case2d(name, size, L, p, C, seed, density, dtype):
    dom = edge_map(size, size, seed, density)
    rng = np.random.RandomState(seed + 1)
    if dtype == "uint8":
        img = rng.randint(0, 256, (size, size, C)).astype(np.uint8)
        # smooth content so bicubic taps matter but values stay structured
        img = ((img.astype(np.float32) + np.roll(img, 1, 0) + np.roll(img, 1, 1)) / 3).astype(np.uint8)
    else:
        img = rng.rand(size, size, C).astype(np.float32)
    ref = FixedQuadTree(domain=dom, fixed_length=L)
    rn = nodes_of(ref)
    on = Q.build_quadtree(dom, L)
    assert rn == on, f"{name}: oracle node list differs from the reference"
    sp, ss, spos = ref.serialize(img, size=(p, p, C))
    sp = np.asarray(sp, dtype=np.float32).reshape(L, p, p, C)
    o_seq, o_size, o_pos = Q.serialize2d(on, img, p, L)
    assert np.array_equal(np.asarray(ss), o_size) and np.array_equal(np.asarray(spos, dtype=np.float64), o_pos)
    err = np.abs(sp - o_seq).max()
    tol = 1.0 if dtype == "uint8" else 2e-6
    assert err <= tol, f"{name}: pixel deviation {err}"
    exact = float((sp == o_seq).mean())
    # deserialize (reference prints a debug line)
    seq_back = (sp * (1.0 if dtype == "uint8" else 255.0))
    if C > 1:
        with contextlib.redirect_stdout(io.StringIO()):
            rm = ref.deserialize(seq_back.copy(), p, C)
    else:
        # reference defect: Rect.set_area loses the channel axis for C == 1 (cv.resize returns HxW)
        # and raises ValueError; the oracle defines the obvious single-channel behaviour.
        rm = Q.deserialize2d(on, seq_back.copy(), p, C, size, size)
    om = Q.deserialize2d(on, seq_back.copy(), p, C, size, size)
    derr = np.abs(rm - om).max()
    assert derr <= 1e-3, f"{name}: deserialize deviation {derr}"
    fx.save_case(os.path.join(OUT, name + ".npz"),
                 dict(kind="quadtree", size=size, L=L, p=p, C=C, seed=seed, density=density, dtype=dtype), {},
                 dict(domain=np.packbits(dom > 0), nodes=np.array(rn, dtype=np.int64), seq_size=np.asarray(ss, dtype=np.int64),
                      seq_pos=np.asarray(spos, dtype=np.float64), seq_img=sp.astype(np.float32),
                      img=img, mask=rm.astype(np.float32)))
    print(f"[golden] {name}: {len(rn)} leaves, nodes bit-exact, serialize max dev {err:.2e} ({100*exact:.1f}% identical), "
          f"deserialize max dev {derr:.2e}")


def kat_appendix_e():
    """The 16x16 known-answer vector of SURVEY.md Appendix E."""
    dom = np.zeros((16, 16), dtype=np.uint8)
    for r, c in [(1, 1), (2, 3), (3, 12), (9, 9), (10, 10), (9, 10), (14, 2)]:
        dom[r, c] = 255
    ref = FixedQuadTree(dom, fixed_length=10)
    rn = nodes_of(ref)
    assert rn == Q.build_quadtree(dom, 10)
    assert rn == [(0, 8, 8, 16, 1), (8, 12, 12, 16, 0), (12, 16, 12, 16, 0), (8, 10, 10, 12, 0), (10, 12, 10, 12, 1),
                  (8, 10, 8, 10, 1), (10, 12, 8, 10, 1), (12, 16, 8, 12, 0), (0, 8, 0, 8, 2), (8, 16, 0, 8, 1)]
    img = np.arange(256, dtype=np.float32).reshape(16, 16, 1)
    sp, ss, spos = ref.serialize(img, size=(2, 2, 1))
    assert np.allclose(sp[0], [[153.5, 157.5], [217.5, 221.5]])
    print("[golden] Appendix-E KAT reproduced by reference and oracle")


def case3d(name, size, L, p, C, seed):
    rng = np.random.RandomState(seed)
    dom = (rng.rand(size, size, size) < 0.02).astype(np.uint8) * 255
    dom[size // 4: size // 2, size // 3, :] = 255
    vol = rng.rand(size, size, size, C).astype(np.float32)
    ref = FixedOctTree(domain=dom, fixed_length=L)
    rn = [tuple(int(v) for v in c.get_coord()) + (int(val),) for c, val in ref.nodes]
    on = Q.build_octree(dom, L)
    assert rn == on, f"{name}: oracle octree differs from the reference"
    sp, ss, spos = ref.serialize(vol, size=(p, p, p, C))
    sp = np.asarray(sp, dtype=np.float32).reshape(L, p, p, p, C)
    o_seq, o_size, o_pos = Q.serialize3d(on, vol, p, L)
    assert np.array_equal(np.asarray(ss), o_size) and np.array_equal(np.asarray(spos, dtype=np.float64), o_pos)
    err = np.abs(sp - o_seq).max()
    assert err <= 2e-6, f"{name}: voxel deviation {err}"
    rm = ref.deserialize(sp.copy(), p, C)
    om = Q.deserialize3d(on, sp.copy(), p, C, (size, size, size))
    derr = np.abs(rm - om).max()
    assert derr <= 2e-6, f"{name}: deserialize deviation {derr}"
    fx.save_case(os.path.join(OUT, name + ".npz"),
                 dict(kind="octree", size=size, L=L, p=p, C=C, seed=seed), {},
                 dict(domain=np.packbits(dom > 0), nodes=np.array(rn, dtype=np.int64), seq_size=np.asarray(ss, dtype=np.int64),
                      seq_pos=np.asarray(spos, dtype=np.float64), seq_img=sp.astype(np.float16),
                      vol=vol.astype(np.float32), mask_checksum=np.float64(np.abs(rm).sum())))
    print(f"[golden] {name}: {len(rn)} leaves, nodes bit-exact, serialize max dev {err:.2e}, deserialize max dev {derr:.2e}")


if __name__ == "__main__":
    kat_appendix_e()
    case2d("sap_tree_u8_64_L31", 64, 31, 8, 3, 5, 0.01, "uint8")
    case2d("sap_tree_f32_128_L64", 128, 64, 16, 3, 7, 0.004, "float32")
    case2d("sap_tree_f32_64_L16_c1", 64, 16, 8, 1, 8, 0.01, "float32")
    case2d("sap_tree_u8_128_L100_stop", 32, 400, 4, 3, 9, 0.3, "uint8")      # hits the 2-px stop rule -> padding
    case3d("sap_tree_oct_32_L22", 32, 22, 4, 1, 11)
