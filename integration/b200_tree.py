# cbir/encoders/b200_tree.py  (new file in the reference tree; kept here as integration/b200_tree.py and executed by
# tests/test_integration_plugin.py)
"""Encoder plugin for the UNMODIFIED reference package: binds the C ABI of libvtree_b200.so with ctypes and
satisfies the reference's encoder protocol (`embedding(image) -> 1-D ndarray`, consumed at cbir/database.py:56).
It replaces the Python descent loop of `VocabularyTree.propagate_feature` (cbir/encoders/vocabulary.py:104-124)."""
import ctypes
import os

import numpy as np

LIB_ENV = "VTREE_B200_LIB"          # path of libvtree_b200.so (default: next to the vtree_b200 package)

# vt_quantize(d_desc, dtype, n, dp, centroids, internal, ref_id, k, depth, start, leaf, word, stats, stream)
VT_QUANTIZE_ARGTYPES = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                        ctypes.c_int, ctypes.c_int, ctypes.c_int32,
                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
VT_F32 = 1


def load_library(path=None):
    path = path or os.environ.get(LIB_ENV) or "libvtree_b200.so"
    lib = ctypes.CDLL(path)
    lib.vt_quantize.restype = ctypes.c_int
    lib.vt_quantize.argtypes = VT_QUANTIZE_ARGTYPES
    lib.vt_last_error.restype = ctypes.c_char_p
    return lib


def heap_arrays(voc, dp=128):
    """The reference's `nodes` / `tree` dicts (vocabulary.py:16-17) as the flat heap the ABI takes: children of
    heap slot h are h*k+1 .. h*k+k; `ref[h]` is the reference's node id of the slot (-1 = unused)."""
    k, depth = voc.n_branches, voc.depth
    n_heap = sum(k ** l for l in range(depth + 1))
    cent = np.zeros((n_heap, dp), np.float32)
    internal = np.zeros(n_heap, np.uint8)
    ref = np.full(n_heap, -1, np.int32)
    stack = [(0, 0)]                                       # (reference id, heap slot)
    while stack:
        rid, h = stack.pop()
        cent[h, :len(voc.nodes[rid])] = voc.nodes[rid]
        ref[h] = rid
        for j, c in enumerate(voc.tree.get(rid, [])):
            internal[h] = 1
            stack.append((c, h * k + 1 + j))
    return cent, internal, ref


class B200Tree:
    """Same visit-count embedding as VocabularyTree.embedding (vocabulary.py:126-137), descent on the GPU.
    `voc` is a fitted reference VocabularyTree (anything with n_branches, depth, nodes, tree, descriptor)."""

    def __init__(self, voc, dp=128, lib=None):
        self.k, self.L, self.dp, self.voc = voc.n_branches, voc.depth, dp, voc
        self.cent_h, self.internal_h, self.ref_h = heap_arrays(voc, dp)
        self.parent = {c: p for p, kids in voc.tree.items() for c in kids}
        self._lib = lib
        self._dev = None

    def _device(self):
        import torch
        if self._dev is None:
            self._lib = self._lib or load_library()
            self._dev = tuple(torch.from_numpy(a).cuda() for a in (self.cent_h, self.internal_h, self.ref_h))
        return self._dev

    def quantize(self, x):
        """reference node id of the stop node of every descriptor row (int32 [n])"""
        import torch
        cent, internal, ref = self._device()
        x = np.ascontiguousarray(x, np.float32)
        rows = torch.zeros((len(x), self.dp), dtype=torch.float32, device="cuda")
        rows[:, :x.shape[1]] = torch.from_numpy(x).cuda()
        word = torch.empty(len(x), dtype=torch.int32, device="cuda")
        rc = self._lib.vt_quantize(rows.data_ptr(), VT_F32, len(x), self.dp, cent.data_ptr(), internal.data_ptr(),
                                   ref.data_ptr(), self.k, self.L, 0, None, word.data_ptr(), None,
                                   torch.cuda.current_stream().cuda_stream)
        if rc:
            raise RuntimeError(self._lib.vt_last_error().decode())
        return word.cpu().numpy()

    def embedding(self, image):
        e = np.zeros(len(self.voc.nodes))
        for w in self.quantize(self.voc.descriptor(image)):
            while True:                                    # counts on every node of the path (vocabulary.py:92-100)
                e[w] += 1
                if w == 0:
                    break
                w = self.parent[w]
        with np.errstate(invalid="ignore", divide="ignore"):
            return e / np.linalg.norm(e, ord=2)

# notebook:  db = cbir.Database(dataset, B200Tree(voc));  db.index();  db.retrieve(path)
