"""Encoders of the retrieval path: the vocabulary tree (cbir/encoders/vocabulary.py).  The reference's
AlexNet encoder is outside the accelerated path."""
from .vocabulary import VocabularyTree  # noqa: F401

__all__ = ["VocabularyTree"]
