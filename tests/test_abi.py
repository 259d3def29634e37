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


def test_host_pack_is_pure_host_code(vt):
    """vt_host_pack (the native multi-threaded packing of a batch's descriptor arrays into one staging buffer) touches no
    device: many arrays of ragged sizes, empty ones included, land back to back, with one thread and with eight."""
    import numpy as np
    lib = vt._native.load()
    rng = np.random.default_rng(0)
    arrs = [rng.integers(0, 256, (int(n), 128), dtype=np.uint8) for n in rng.integers(0, 900, 400)] + [np.zeros((0, 128), np.uint8)]
    want = np.concatenate(arrs)
    assert want.nbytes > (8 << 20)                                  # large enough for the threaded path
    n = len(arrs)
    src = (ctypes.c_void_p * n)(*[a.__array_interface__["data"][0] for a in arrs])
    nbytes = (ctypes.c_int64 * n)(*[a.nbytes for a in arrs])
    for threads in (1, 8):
        dst = np.full(want.shape, 7, np.uint8)
        assert lib.vt_host_pack(src, nbytes, n, ctypes.c_void_p(dst.ctypes.data), threads) == 0
        assert np.array_equal(dst, want)
    assert lib.vt_host_pack(src, nbytes, n, None, 4) == -1 and b"bad arguments" in lib.vt_last_error()
