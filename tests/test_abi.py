"""CPU: the C-ABI shared library builds, loads and exports every symbol include/vtree_b200.h
declares (no compute calls here -- those need a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vtree_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(vt):
    path = vt._native.ensure_built()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "%s declared in the header but not exported" % n


def test_binding_table_matches_header(vt):
    assert sorted(vt._native.SIGNATURES) == _declared()


def test_version_and_error_string(vt):
    lib = vt._native.load()
    assert lib.vt_abi_version() == 1
    assert isinstance(lib.vt_last_error(), bytes)


def test_argument_validation_without_gpu(vt):
    """Argument checks run before any CUDA call, so they are testable on the CPU box."""
    lib = vt._native.load()
    rc = lib.vt_quantize(None, 0, 10, 128, None, None, None, 10, 3, 0, None, None, None, None)
    assert rc == -1 and b"null tree" in lib.vt_last_error()
    flip = ctypes.c_int(0)
    rc = lib.vt_sort_pairs(None, None, None, None, 1 << 40, 8, None, ctypes.byref(flip), None)
    assert rc == -1
    assert lib.vt_encode_max_words() == 8192 and lib.vt_score_shard_size() == 8192
