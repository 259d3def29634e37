"""``DescriptorBase`` (cbir/descriptors/descriptor_base.py:7-46): a descriptor is called with an image,
keys it by the sha1 of its pixels (utils.get_image_id) and memoises the features.  The reference keeps
the memo in ``<store_root>/features_<Class>.hdf5``; h5py is optional here: without it (or without a
store root) the memo lives in memory, with it the same HDF5 layout (one dataset per image id) is used."""
from __future__ import annotations

import os

import numpy as np

from .. import utils


class DescriptorBase(object):
    def __init__(self, store_root=None):
        self._memo = {}
        self._storage = None
        if store_root is not None:
            try:
                import h5py  # noqa: F401  (optional)
                self._storage = os.path.join(store_root, "features_%s.hdf5" % self.__class__.__name__)
            except ImportError:
                self._storage = None

    def __call__(self, image):
        image_id = utils.get_image_id(image)
        if not self.is_stored(image_id):
            self.store(image_id, self.describe(image))
        return self.load(image_id)

    def describe(self, image):
        raise NotImplementedError("subclasses return the [n, D] descriptor array of an image")

    def is_stored(self, image_id, store_root="data"):
        if image_id in self._memo:
            return True
        if self._storage is None or not os.path.isfile(self._storage):
            return False
        import h5py
        with h5py.File(self._storage, "r") as file:
            return image_id in file

    def load(self, image_id):
        if image_id not in self._memo:
            import h5py
            with h5py.File(self._storage, "r") as file:
                self._memo[image_id] = np.array(file[image_id])
        return self._memo[image_id]

    def store(self, image_id, features, force=False, store_root="data"):
        if self.is_stored(image_id) and not force:
            return
        features = np.array(features)
        self._memo[image_id] = features
        if self._storage is not None:
            import h5py
            os.makedirs(os.path.dirname(self._storage) or ".", exist_ok=True)
            with h5py.File(self._storage, "a") as file:
                if image_id in file:
                    del file[image_id]
                file.create_dataset(image_id, features.shape, data=features)
        return
