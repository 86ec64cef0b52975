"""`Patchify` / `Patchify_3D` data transforms with the reference's constructor and return values
(/root/reference/src/UCF_VIT/dataloaders/transform.py:9-54, :56-132).

Edge detection stays on OpenCV / scipy on the host (bit-exactness of the tree depends on it,
SURVEY.md §0.9, §8f rank 4); the tree build is the C++ host routine and the per-leaf resampling
gather runs on the GPU.  With `device_output=True` the sequence stays on the device as torch
tensors (no D2H copy) for direct consumption by the model."""
import random

import numpy as np
import torch

from .octree import FixedOctTree
from .quadtree import FixedQuadTree


class Patchify(torch.nn.Module):
    def __init__(self, sths=[0, 1, 3, 5], fixed_length=196, cannys=[50, 100], patch_size=16, num_channels=3,
                 dataset="imagenet", return_edges=False, device="cuda", device_output=False) -> None:
        super().__init__()
        self.sths = sths
        self.fixed_length = fixed_length
        self.cannys = [x for x in range(cannys[0], cannys[1], 1)]
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.dataset = dataset
        self.return_edges = return_edges
        self.device, self.device_output = device, device_output

    def _edges(self, img):
        import cv2 as cv
        natural = self.dataset in ("imagenet", "catsdogs")
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
