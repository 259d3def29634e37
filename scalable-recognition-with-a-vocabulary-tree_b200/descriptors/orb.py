"""``Orb`` (cbir/descriptors/orb.py:12-33): OpenCV ORB, up to 1,500 keypoints over 32 pyramid levels,
32-byte descriptors returned as float32 ``[n, 32]``; an image without keypoints yields one all-zero
descriptor.  CPU front-end, outside the accelerated path."""
from __future__ import annotations

import numpy as np

from .descriptor_base import DescriptorBase


class Orb(DescriptorBase):
    def __init__(self, patch_size=65, store_root=None):
        super(Orb, self).__init__(store_root)
        import cv2
        self.patch_size = (int(patch_size), int(patch_size))
        self.orb = cv2.ORB_create(1500, nlevels=32)

    def describe(self, image):
        _, desc = self.orb.detectAndCompute(image, None)
        if desc is None or np.size(desc) <= 1:
            return np.zeros((1, 32), dtype=np.float32)
        return np.asarray(desc, dtype=np.float32)
