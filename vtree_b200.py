"""Importable alias of the package directory ``scalable-recognition-with-a-vocabulary-tree_b200``
(whose name is not a Python identifier):  ``import vtree_b200 as cbir``."""
import importlib
import sys

_REAL = "scalable-recognition-with-a-vocabulary-tree_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
