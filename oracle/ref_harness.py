"""oracle/ref_harness.py -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference package (``/root/reference/cbir``) in this container so that
its own code can be executed on seeded synthetic inputs.  Used only by
``oracle/make_golden.py`` (fixture generation) -- ``/root/reference`` does not exist on the GPU
box, so nothing under tests/ or bench.py may import this at run time.

What is swapped, and why:
* matplotlib / h5py / termcolor / IPython are absent here; the reference imports them eagerly
  (cbir/__init__.py:1-6) but never touches them on the hot path -> empty stub modules.
* ``cbir.encoders.vocabulary.MiniBatchKMeans`` (unseeded, stochastic, vocabulary.py:62) is
  replaced by ``vt_oracle.SeededLloyd``; the recursion, stable partition, leaf rule, DFS
  numbering and centre hand-down of ``VocabularyTree.fit`` stay the reference's own code.
"""
from __future__ import annotations

import contextlib
import importlib.machinery
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def import_reference(root: str = REFERENCE_ROOT):
    if not os.path.isdir(os.path.join(root, "cbir")):
        raise RuntimeError("reference checkout not available at %s" % root)
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("matplotlib.patches", Circle=object)
    _stub("matplotlib.gridspec", GridSpec=object)
    _stub("h5py")
    _stub("termcolor", colored=lambda s, *a, **k: s)
    _stub("IPython")
    _stub("IPython.display", HTML=object, display=lambda *a, **k: None)
    if root not in sys.path:
        sys.path.insert(0, root)
    import cbir  # noqa: the reference package, unmodified
    return cbir


def make_dataset_class(cbir):
    class SynthDataset(cbir.Dataset):
        """Passes ``isinstance(dataset, Dataset)`` (database.py:12); an "image" is its own
        C-contiguous [n, D] descriptor array and the descriptor is the identity."""

        def __init__(self, arrays, names=None):
            self.path = "synthetic"
            names = names or ["%06d.jpg" % (100000 + i) for i in range(len(arrays))]
            order = sorted(range(len(names)), key=lambda i: names[i])
            self.image_paths = [names[i] for i in order]
            self._arrays = {names[i]: np.ascontiguousarray(arrays[i]) for i in order}
            self.subset = None

        def read_image(self, image_path, scale=1.):
            return self._arrays[image_path]

        def __len__(self):
            return len(self.image_paths)

    return SynthDataset


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def reference_build(cbir, features: np.ndarray, k: int, depth: int, n_iter: int):
    """``VocabularyTree(k, depth, identity).fit(features)`` with the seeded Lloyd patched in."""
    from . import vt_oracle
    import cbir.encoders.vocabulary as voc_mod

    class _Lloyd(vt_oracle.SeededLloyd):
        pass
    _Lloyd.n_iter = n_iter
    voc_mod.MiniBatchKMeans = _Lloyd
    voc = cbir.encoders.VocabularyTree(n_branches=k, depth=depth, descriptor=lambda im: im)
    with quiet():
        voc.fit(features)
    return voc
