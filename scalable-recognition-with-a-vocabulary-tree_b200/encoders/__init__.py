from .vocabulary import VocabularyTree  # noqa: F401
