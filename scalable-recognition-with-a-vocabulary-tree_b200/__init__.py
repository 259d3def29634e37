"""B200-native vocabulary-tree retrieval hot path behind the ``cbir`` API of
epignatelli/scalable-recognition-with-a-vocabulary-tree.

    import vtree_b200 as cbir          # alias module at the repo root
    voc = cbir.encoders.VocabularyTree(n_branches=10, depth=3, descriptor=my_descriptor)
    voc.fit(features)
    db = cbir.Database(dataset, voc); db.index(); db.retrieve(path)

Only the path named in BASELINE.json is here: tree build, descent, bag-of-words encoding,
inverted-file scoring + top-k.  ``descriptors`` holds thin CPU front-ends (OpenCV
ORB) so that the notebook flow runs; GPU feature extraction, plotting, downloads and the teaching
helpers of the reference are out of scope.
"""
from . import utils  # noqa: F401
from . import synth  # noqa: F401
from .dataset import Dataset, ArrayDataset  # noqa: F401
from .tree import FlatTree  # noqa: F401
from . import encoders  # noqa: F401
from . import descriptors  # noqa: F401
from .database import Database  # noqa: F401
from . import engine  # noqa: F401
from . import dist  # noqa: F401
from . import _native  # noqa: F401

__all__ = ["Database", "Dataset", "ArrayDataset", "FlatTree", "encoders", "descriptors", "utils", "engine", "synth"]
