"""Dataset boundary (cbir/dataset.py).  ``Database`` iterates ``dataset.image_paths`` in order --
that order is the top-k tie-break (database.py:43,75) -- and calls ``dataset.read_image(path)``.
Image decoding is outside the accelerated path; ``ArrayDataset`` feeds descriptor arrays
directly (the synthetic "image is its own descriptor set" setup of SURVEY.md Appendix A)."""
from __future__ import annotations

import os

import numpy as np


IMAGE_SUFFIXES = (".jpeg", ".jpg", ".jp2", ".png", ".bmp", ".tif", ".tiff", "pbm", ".pgm", "ppm")   # dataset.py:90-98


class Dataset(object):
    """Folder of images (dataset.py:9-98): ``image_paths`` holds the bare names of the regular files of
    ``folder`` sorted by name -- the order ``Database`` indexes in and breaks score ties by."""

    def __init__(self, folder="data/jpg"):
        self.path = folder
        self.image_paths = [name for name in sorted(os.listdir(folder)) if os.path.isfile(os.path.join(folder, name))]
        self.subset = Subset(self)

    def __str__(self):
        return str(list(self.image_paths[:6]))              # the first six names (dataset.py:21-27)

    __repr__ = __str__

    def __len__(self):
        return len(self.image_paths)

    def __iter__(self):
        return iter(self.image_paths)

    def __getitem__(self, idx, read=False):
        items = self.image_paths[idx]
        if read:
            return [self.read_image(item) for item in items]
        return items

    def is_image(self, path):
        return os.path.splitext(path)[-1] in IMAGE_SUFFIXES

    def read_image(self, image_path, scale=1.):
        """RGB uint8 array.  Like dataset.py:41-65: names without an image suffix get ``.jpg``, names that
        are not files as given are looked up in the dataset folder, a missing file raises
        FileNotFoundError(path)."""
        import cv2
        if not self.is_image(image_path):
            image_path = image_path + ".jpg"
        if not os.path.isfile(image_path):
            image_path = os.path.abspath(os.path.join(self.path, image_path))
        if not os.path.isfile(image_path):
            raise FileNotFoundError(image_path)
        image = cv2.imread(image_path)
        image = cv2.resize(image, (0, 0), fx=scale, fy=scale)
        return cv2.cvtColor(image, cv2.COLOR_BGR2RGB)

    def get_random_image(self):
        import random
        return self.read_image(random.choice(self.image_paths))

    def show_image(self, image, gray=False, **kwargs):
        """Displays an image or the image behind a name (matplotlib is an optional, plotting-only dependency)."""
        import matplotlib.pyplot as plt
        if isinstance(image, str):
            image = self.read_image(image)
        if gray:
            kwargs.setdefault("cmap", "gray")
        plt.imshow(image, aspect="equal", **kwargs)


class Subset(object):
    """``dataset.subset[a:b]`` narrows the dataset IN PLACE and returns it (dataset.py:100-107)."""

    def __init__(self, dataset):
        self.dataset = dataset

    def __getitem__(self, idx):
        self.dataset.image_paths = self.dataset.image_paths[idx]
        return self.dataset


class ArrayDataset(Dataset):
    """In-memory dataset whose "images" are [n_i, D] descriptor arrays (use with the identity
    descriptor).  Names sort like the reference's sorted file list."""

    def __init__(self, arrays, names=None):
        names = list(names) if names is not None else ["%06d.jpg" % (100000 + i) for i in range(len(arrays))]
        order = sorted(range(len(names)), key=lambda i: names[i])
        self.path = "memory"
        self.image_paths = [names[i] for i in order]
        self._arrays = {names[i]: np.ascontiguousarray(arrays[i]) for i in order}
        self.subset = Subset(self)

    def read_image(self, image_path, scale=1.):
        if image_path not in self._arrays:
            raise FileNotFoundError("Cannot find image %s" % image_path)
        return self._arrays[image_path]
