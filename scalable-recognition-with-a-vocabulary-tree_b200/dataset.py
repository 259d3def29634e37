"""Dataset boundary (cbir/dataset.py).  ``Database`` iterates ``dataset.image_paths`` in order --
that order is the top-k tie-break (database.py:43,75) -- and calls ``dataset.read_image(path)``.
Image decoding is outside the accelerated path; ``ArrayDataset`` feeds descriptor arrays
directly (the synthetic "image is its own descriptor set" setup of SURVEY.md Appendix A)."""
from __future__ import annotations

import os

import numpy as np


class Dataset(object):
    """Folder of images, sorted by file name (dataset.py:9-40)."""

    def __init__(self, folder="data/jpg"):
        self.path = folder
        self.image_paths = sorted(
            os.path.join(folder, f) for f in os.listdir(folder)
            if f.lower().endswith((".jpg", ".jpeg", ".png", ".bmp", ".pgm"))) if os.path.isdir(folder) else []

    def __len__(self):
        return len(self.image_paths)

    def __iter__(self):
        return iter(self.image_paths)

    def __getitem__(self, i):
        return self.image_paths[i]

    def read_image(self, image_path, scale=1.):
        """RGB uint8 array; adds the .jpg suffix / folder like dataset.py:41-65 and raises
        FileNotFoundError (dataset.py:60-61) when the file is missing."""
        import cv2
        if not image_path.endswith((".jpg", ".jpeg", ".png", ".bmp", ".pgm")):
            image_path = image_path + ".jpg"
        if not os.path.exists(image_path):
            image_path = os.path.join(self.path, os.path.basename(image_path))
        if not os.path.exists(image_path):
            raise FileNotFoundError("Cannot find image %s" % image_path)
        image = cv2.cvtColor(cv2.imread(image_path), cv2.COLOR_BGR2RGB)
        if scale != 1.:
            image = cv2.resize(image, None, fx=scale, fy=scale)
        return image


class ArrayDataset(Dataset):
    """In-memory dataset whose "images" are [n_i, D] descriptor arrays (use with the identity
    descriptor).  Names sort like the reference's sorted file list."""

    def __init__(self, arrays, names=None):
        names = list(names) if names is not None else ["%06d.jpg" % (100000 + i) for i in range(len(arrays))]
        order = sorted(range(len(names)), key=lambda i: names[i])
        self.path = "memory"
        self.image_paths = [names[i] for i in order]
        self._arrays = {names[i]: np.ascontiguousarray(arrays[i]) for i in order}

    def read_image(self, image_path, scale=1.):
        if image_path not in self._arrays:
            raise FileNotFoundError("Cannot find image %s" % image_path)
        return self._arrays[image_path]
