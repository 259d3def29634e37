"""Descriptor front-ends (cbir/descriptors/).  The descriptor protocol of the reference is a callable
``image -> [n, D] ndarray`` (descriptor_base.py:14-21, new_descriptor_example.py:3-6); anything with that
shape plugs into ``VocabularyTree(n_branches, depth, descriptor)``.  Feature extraction itself is NOT
part of the accelerated path (it runs on the CPU through OpenCV exactly like the reference's) -- these
classes exist so that the notebook flow runs unchanged."""
from .descriptor_base import DescriptorBase  # noqa: F401
from .orb import Orb  # noqa: F401
