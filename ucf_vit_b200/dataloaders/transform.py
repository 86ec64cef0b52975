"""`Patchify` / `Patchify_3D` data transforms with the reference's constructor and return values
(/root/reference/src/UCF_VIT/dataloaders/transform.py:9-54, :56-132).

Edge detection (bit-exactness of the tree depends on it): for natural uint8 images --
the imagenet / catsdogs branch, `cv.GaussianBlur` + `cv.Canny` -- `Patchify(edges="device")` runs both on the GPU
(`ucf_gaussian_blur_u8`, `ucf_canny_u8`: integer arithmetic, byte-identical to OpenCV, tests/test_gpu_canny.py) on the copy
of the image the gather needs anyway and hands the host tree builder the uint8 map; `edges="host"` (default, the reference's
own calls) and every other branch (float images, the 3-D recipe) use OpenCV / scipy on the host.  The tree build is the C++
host routine and the per-leaf resampling gather runs on the GPU.  With `device_output=True` the sequence stays on the device as torch
tensors (no D2H copy) for direct consumption by the model."""
import random

import numpy as np
import torch

from .octree import FixedOctTree
from .quadtree import FixedQuadTree


class Patchify(torch.nn.Module):
    def __init__(self, sths=[0, 1, 3, 5], fixed_length=196, cannys=[50, 100], patch_size=16, num_channels=3,
                 dataset="imagenet", return_edges=False, device="cuda", device_output=False, edges="host") -> None:
        super().__init__()
        if edges not in ("host", "device"):
            raise ValueError("edges must be 'host' (OpenCV) or 'device' (ucf_gaussian_blur_u8 + ucf_canny_u8)")
        self.edges = edges
        self.sths = sths
        self.fixed_length = fixed_length
        self.cannys = [x for x in range(cannys[0], cannys[1], 1)]
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.dataset = dataset
        self.return_edges = return_edges
        self.device, self.device_output = device, device_output

    def _edges_device(self, img):
        """The natural-image branch on the GPU: same bytes as cv.Canny(cv.GaussianBlur(img, (k, k), 0), c, c + 50)."""
        from .. import ops
        dev = torch.device(self.device)
        with torch.cuda.device(dev):
            x = torch.as_tensor(np.ascontiguousarray(img)).to(dev, non_blocking=True)
            e = ops.canny_u8(ops.gaussian_blur_u8(x, self.smooth_factor), self.canny[0], self.canny[1])
            return e.cpu().numpy()

    def _edges(self, img):
        natural = self.dataset in ("imagenet", "catsdogs")
        if (self.edges == "device" and natural and self.smooth_factor != 0 and isinstance(img, np.ndarray)
                and img.dtype == np.uint8 and self.smooth_factor in (1, 3, 5)):
            return self._edges_device(img)
        import cv2 as cv
        if self.smooth_factor == 0:
            lo, hi = (0, 1) if natural else (np.min(img), np.max(img))
            return np.random.uniform(low=lo, high=hi, size=(img.shape[0], img.shape[1]))
        blurred = cv.GaussianBlur(img, (self.smooth_factor, self.smooth_factor), 0)
        if not natural:
            blurred = (blurred * 255).astype(np.uint8)
        return cv.Canny(blurred, self.canny[0], self.canny[1])

    def forward(self, img):
        self.smooth_factor = random.choice(self.sths)
        c = random.choice(self.cannys)
        self.canny = [c, c + 50]
        edges = self._edges(img)
        qdt = FixedQuadTree(domain=edges, fixed_length=self.fixed_length, device=self.device)
        p, C = self.patch_size, self.num_channels
        seq, seq_size, seq_pos = qdt.serialize_device(img, size=(p, p, C))
        # the reference RAW-reshapes (L,p,p,C) -> (C,L,p*p) (transform.py:45-48; not a transpose)
        seq = seq.reshape(C, -1, p * p) if C > 1 else seq.reshape(-1, p * p)
        if not self.device_output:
            seq, seq_size, seq_pos = seq.cpu().numpy(), seq_size.cpu().numpy(), seq_pos.cpu().numpy()
        return (seq, seq_size, seq_pos, qdt, edges) if self.return_edges else (seq, seq_size, seq_pos, qdt)

    def forward_batch(self, imgs, threads=0):
        """`[self(img) for img in imgs]` with the trees of the whole batch built together on host threads (the
        random smoothing / Canny draws are consumed in the same order as by repeated `forward` calls)."""
        edges = []
        for img in imgs:
            self.smooth_factor = random.choice(self.sths)
            c = random.choice(self.cannys)
            self.canny = [c, c + 50]
            edges.append(self._edges(img))
        trees = FixedQuadTree.build_many(edges, self.fixed_length, device=self.device, threads=threads)
        p, C = self.patch_size, self.num_channels
        out = []
        for img, e, qdt in zip(imgs, edges, trees):
            seq, seq_size, seq_pos = qdt.serialize_device(img, size=(p, p, C))
            seq = seq.reshape(C, -1, p * p) if C > 1 else seq.reshape(-1, p * p)
            if not self.device_output:
                seq, seq_size, seq_pos = seq.cpu().numpy(), seq_size.cpu().numpy(), seq_pos.cpu().numpy()
            out.append((seq, seq_size, seq_pos, qdt, e) if self.return_edges else (seq, seq_size, seq_pos, qdt))
        return out


class Patchify_3D(torch.nn.Module):
    """Volume [Z, Y, X, C] -> adaptive octree sequence (reference transform.py:56-132).

    Edge map on the host, slice by slice, with the reference's recipe: Gaussian smoothing of the volume, the
    gradient direction from 5x5 Sobel derivatives, one Canny pass per channel; a voxel votes once per channel that
    marks it, and it counts when the min-max-normalised direction of an edge voxel exceeds 0.5.  The uint8 map
    (votes * int(255 / C)) drives the C++ octree build; the per-leaf trilinear gather runs on the GPU."""

    def __init__(self, sths=[0, 1, 3, 5], fixed_length=196, cannys=[50, 100], patch_size=16, num_channels=3,
                 dataset="basic_ct", return_edges=False, device="cuda", device_output=False) -> None:
        super().__init__()
        self.sths = sths
        self.fixed_length = fixed_length
        self.cannys = [x for x in range(cannys[0], cannys[1], 1)]
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.dataset = dataset
        self.return_edges = return_edges
        self.device, self.device_output = device, device_output

    def _slice_direction(self, planes):
        """arctan2(dy, dx) of one z-slice.  Channel 0 supplies both derivatives; a later channel takes over dx when its
        mean gradient magnitude beats channel 0's, and dy when its mean dy beats the current one (:75-90)."""
        import cv2 as cv
        dx = cv.Sobel(planes[:, :, 0], cv.CV_64F, 1, 0, ksize=5)
        dy = cv.Sobel(planes[:, :, 0], cv.CV_64F, 0, 1, ksize=5)
        strength0 = np.mean(np.sqrt(dx ** 2 + dy ** 2))
        for c in range(1, self.num_channels):
            cx = cv.Sobel(planes[:, :, c], cv.CV_64F, 1, 0, ksize=5)
            cy = cv.Sobel(planes[:, :, c], cv.CV_64F, 0, 1, ksize=5)
            if np.mean(np.sqrt(cx ** 2 + cy ** 2)) > strength0:
                dx = cx
            if np.mean(cy) > np.mean(dy):
                dy = cy
        return np.arctan2(dy, dx)

    def _edges(self, img):
        import cv2 as cv
        from scipy.ndimage import gaussian_filter
        s = self.smooth_factor
        smooth = gaussian_filter(img, sigma=(s, s, s, 0))
        direction = np.zeros_like(smooth[:, :, :, 0])
        votes = np.zeros(smooth.shape[:3], dtype=np.uint8)
        for z in range(smooth.shape[0]):
            direction[z] = self._slice_direction(smooth[z])
            for c in range(self.num_channels):
                marked = cv.Canny((smooth[z, :, :, c] * 255).astype(np.uint8), self.canny[0], self.canny[1]) > 0
                votes[z] += marked.astype(np.uint8)
        on_edge = votes > 0                       # the reference ORs the per-channel maps through a wrapping uint8 sum
        masked = np.zeros_like(direction)
        masked[on_edge] = direction[on_edge]
        with np.errstate(invalid="ignore", divide="ignore"):
            unit = (masked - masked.min()) / (masked.max() - masked.min())
        self._norm_factor = int(255 / self.num_channels)
        return (unit > 0.5).astype(np.uint8) * (votes * self._norm_factor)

    def forward(self, img):
        self.smooth_factor = random.choice(self.sths)
        c = random.choice(self.cannys)
        self.canny = [c, c + 50]
        edges = self._edges(img)
        tree = FixedOctTree(domain=edges, fixed_length=self.fixed_length, norm_factor=self._norm_factor, device=self.device)
        p, C = self.patch_size, self.num_channels
        seq, seq_size, seq_pos = tree.serialize_device(img, size=(p, p, p, C))
        # raw reshape like the 2-D transform (:121-124): (L,p,p,p,C) -> (C,L,p^3), not a transpose
        seq = seq.reshape(C, -1, p * p * p) if C > 1 else seq.reshape(-1, p * p * p)
        if not self.device_output:
            seq, seq_size, seq_pos = seq.cpu().numpy(), seq_size.cpu().numpy(), seq_pos.cpu().numpy()
        return (seq, seq_size, seq_pos, tree, edges) if self.return_edges else (seq, seq_size, seq_pos, tree)
