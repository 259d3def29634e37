"""oracle/vt_oracle.py -- TEST INFRASTRUCTURE ONLY (not product code).

CPU restatement of the reference's vocabulary-tree hot path.  Heavy inner loops are the
plain-C functions of ``oracle/vt_oracle.c`` (loaded with ctypes); the control flow around
them follows the reference line by line:

* ``build_tree``       -> ``VocabularyTree.fit``             cbir/encoders/vocabulary.py:36-76
* ``SeededLloyd``      -> replaces sklearn ``MiniBatchKMeans`` at vocabulary.py:62-63
* ``descend``          -> ``VocabularyTree.propagate_feature`` vocabulary.py:104-124
* ``path_counts``      -> ``VocabularyTree.propagate``        vocabulary.py:78-102
* ``embedding_dense``  -> ``VocabularyTree.embedding``        vocabulary.py:126-137
* ``score_dense``      -> ``Database.score``                  cbir/database.py:59-70
* ``retrieve_order``   -> ``Database.retrieve``               cbir/database.py:72-81

Parity status: pinned.  ``oracle/make_golden.py`` runs the unmodified reference package on
seeded inputs and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this module against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/vt_oracle.c with the committed Makefile (gcc)."""
    src = os.path.join(_HERE, "vt_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.vt_or_dist2.restype = ctypes.c_double
        L.vt_or_dist2.argtypes = [f32p, f32p, ctypes.c_int]
        L.vt_or_assign.restype = None
        L.vt_or_assign.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int, i32p, f64p, f64p]
        L.vt_or_lloyd.restype = ctypes.c_int
        L.vt_or_lloyd.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, i32p]
        L.vt_or_mean.restype = None
        L.vt_or_mean.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p]
        L.vt_or_descend.restype = None
        L.vt_or_descend.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, u8p, ctypes.c_int, ctypes.c_int,
                                    i32p, i32p, f64p]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


# ----------------------------------------------------------------------------- layout
def pad_dim(d: int) -> int:
    """Descriptor rows are zero padded to 32, 64, 128, ... (the canonical chunking cuts the
    padded row into 32 equal chunks; the CUDA path pads the same way)."""
    dp = 32
    while dp < int(d):
        dp *= 2
    return dp


def padded(x) -> np.ndarray:
    """[n, D] anything -> C-contiguous float32 [n, pad_dim(D)] (exact for byte data)."""
    x = np.asarray(x)
    if x.ndim == 1:
        x = x[None, :]
    n, d = x.shape
    dp = pad_dim(d)
    out = np.zeros((n, dp), dtype=np.float32)
    out[:, :d] = x
    return out


def heap_size(k: int, depth: int) -> int:
    return sum(k ** l for l in range(depth + 1))


def level_offset(k: int, level: int) -> int:
    return sum(k ** l for l in range(level))


# ----------------------------------------------------------------------------- k-means
def assign(xp: np.ndarray, cp: np.ndarray, want_gap: bool = False):
    """Canonical first-wins argmin (vt_or_assign)."""
    xp = np.ascontiguousarray(xp, dtype=np.float32)
    cp = np.ascontiguousarray(cp, dtype=np.float32)
    m, dp = xp.shape
    k = cp.shape[0]
    labels = np.empty(m, dtype=np.int32)
    best = np.empty(m, dtype=np.float64) if want_gap else None
    second = np.empty(m, dtype=np.float64) if want_gap else None
    lib().vt_or_assign(_p(xp, ctypes.c_float), m, dp, _p(cp, ctypes.c_float), k,
                       _p(labels, ctypes.c_int32), _p(best, ctypes.c_double), _p(second, ctypes.c_double))
    return (labels, best, second) if want_gap else labels


def lloyd(xp: np.ndarray, k: int, n_iter: int):
    """Seeded exact Lloyd on one node -> (centres [k, dp] float32, labels [m] int32, updates)."""
    xp = np.ascontiguousarray(xp, dtype=np.float32)
    m, dp = xp.shape
    assert m >= k
    c = np.zeros((k, dp), dtype=np.float32)
    labels = np.empty(m, dtype=np.int32)
    done = lib().vt_or_lloyd(_p(xp, ctypes.c_float), m, dp, k, n_iter, _p(c, ctypes.c_float),
                             _p(labels, ctypes.c_int32))
    return c, labels, done


def exact_mean(xp: np.ndarray) -> np.ndarray:
    xp = np.ascontiguousarray(xp, dtype=np.float32)
    out = np.zeros(xp.shape[1], dtype=np.float32)
    lib().vt_or_mean(_p(xp, ctypes.c_float), xp.shape[0], xp.shape[1], _p(out, ctypes.c_float))
    return out


class SeededLloyd:
    """Drop-in for ``sklearn.cluster.MiniBatchKMeans`` as used at vocabulary.py:62-66,75
    (only ``fit``, ``labels_`` and ``cluster_centers_`` are touched by the reference)."""

    n_iter = 10

    def __init__(self, n_clusters, n_iter=None):
        self.n_clusters = int(n_clusters)
        if n_iter is not None:
            self.n_iter = int(n_iter)

    def fit(self, features):
        x = np.asarray(features)
        d = x.shape[1]
        c, labels, _ = lloyd(padded(x), self.n_clusters, self.n_iter)
        self.cluster_centers_ = np.ascontiguousarray(c[:, :d])
        self.labels_ = labels
        return self


# ----------------------------------------------------------------------------- tree build
def build_tree(x, k: int, depth: int, n_iter: int = 10) -> dict:
    """Restatement of ``VocabularyTree.fit`` (vocabulary.py:36-76) producing flat arrays.

    Heap layout: children of heap node n are n*k+1 .. n*k+k.  ``ref_id`` is the id the
    reference gives the node (DFS pre-order counter, vocabulary.py:70-75).
    """
    x = np.asarray(x)
    d = x.shape[1]
    xp = padded(x)
    dp = xp.shape[1]
    nh = heap_size(k, depth)
    t = dict(k=k, depth=depth, D=d, DP=dp, n_iter=n_iter,
             centroids=np.zeros((nh, dp), np.float32),
             exists=np.zeros(nh, np.uint8), internal=np.zeros(nh, np.uint8),
             ref_id=np.full(nh, -1, np.int32), count=np.zeros(nh, np.int64))
    counter = [0]

    def fit(rows, heap, level):
        t["exists"][heap] = 1
        t["ref_id"][heap] = counter[0]
        counter[0] += 1
        t["count"][heap] = len(rows)
        if level >= depth or len(rows) < k:          # vocabulary.py:55
            return
        c, labels, _ = lloyd(rows, k, n_iter)         # vocabulary.py:62-63
        t["internal"][heap] = 1
        for j in range(k):                            # vocabulary.py:70-75
            child = heap * k + 1 + j
            t["centroids"][child] = c[j]
            fit(rows[labels == j], child, level + 1)  # stable partition, vocabulary.py:64-66

    t["centroids"][0] = exact_mean(xp)                # vocabulary.py:48-49
    fit(xp, 0, 0)
    n_ref = counter[0]
    t["n_ref"] = n_ref
    hor = np.full(n_ref, -1, np.int64)
    ex = np.nonzero(t["exists"])[0]
    hor[t["ref_id"][ex]] = ex
    t["heap_of_ref"] = hor
    return t


def tree_from_reference(voc, d: int) -> dict:
    """Flatten a *reference* ``VocabularyTree`` (``voc.tree`` id->child ids, ``voc.nodes``
    id->centre) into the same arrays ``build_tree`` produces."""
    k, depth = voc.n_branches, voc.depth
    dp = pad_dim(d)
    nh = heap_size(k, depth)
    t = dict(k=k, depth=depth, D=d, DP=dp,
             centroids=np.zeros((nh, dp), np.float32),
             exists=np.zeros(nh, np.uint8), internal=np.zeros(nh, np.uint8),
             ref_id=np.full(nh, -1, np.int32))
    stack = [(0, 0)]
    while stack:
        rid, heap = stack.pop()
        t["exists"][heap] = 1
        t["ref_id"][heap] = rid
        t["centroids"][heap, :d] = np.asarray(voc.nodes[rid], dtype=np.float32)
        kids = voc.tree.get(rid, [])
        if kids:
            t["internal"][heap] = 1
            for j, cid in enumerate(kids):
                stack.append((cid, heap * k + 1 + j))
    n_ref = len(voc.nodes)
    t["n_ref"] = n_ref
    hor = np.full(n_ref, -1, np.int64)
    ex = np.nonzero(t["exists"])[0]
    hor[t["ref_id"][ex]] = ex
    t["heap_of_ref"] = hor
    return t


# ----------------------------------------------------------------------------- descent
def descend(tree: dict, x, want_path: bool = False, want_gap: bool = False):
    """Greedy descent -> heap index of the stop node per descriptor (+ optional heap path,
    + optional smallest relative runner-up gap met on the way)."""
    xp = padded(x)
    n = xp.shape[0]
    depth = tree["depth"]
    leaf = np.empty(n, np.int32)
    path = np.empty((n, depth + 1), np.int32) if want_path else None
    gap = np.empty(n, np.float64) if want_gap else None
    cent = np.ascontiguousarray(tree["centroids"], np.float32)
    internal = np.ascontiguousarray(tree["internal"], np.uint8)
    lib().vt_or_descend(_p(xp, ctypes.c_float), n, xp.shape[1], _p(cent, ctypes.c_float),
                        _p(internal, ctypes.c_uint8), tree["k"], depth,
                        _p(leaf, ctypes.c_int32), _p(path, ctypes.c_int32), _p(gap, ctypes.c_double))
    out = [leaf]
    if want_path:
        out.append(path)
    if want_gap:
        out.append(gap)
    return out[0] if len(out) == 1 else tuple(out)


def descend_literal(tree: dict, x) -> np.ndarray:
    """Line-by-line restatement of ``VocabularyTree.propagate_feature`` (vocabulary.py:104-124) with the
    reference's own arithmetic: float32 ``np.linalg.norm([centre - feature])`` per child, strict ``<``.
    Pure-Python loops -- only for small samples (it is what the reference costs per descriptor) and for
    pinning: on the machine that made the goldens it reproduces ``ref_leaf`` exactly."""
    x = np.asarray(x, dtype=np.float32)
    d = x.shape[1]
    k = tree["k"]
    cent = tree["centroids"][:, :d]
    internal = tree["internal"]
    out = np.empty(len(x), np.int32)
    for i, feature in enumerate(x):
        node = 0
        while node * k + k < len(internal) and internal[node]:      # graph.out_degree(node) > 0
            min_dist = float("inf")
            closest = None
            for child in range(node * k + 1, node * k + 1 + k):     # insertion order == child index order
                distance = np.linalg.norm([cent[child] - feature])
                if distance < min_dist:
                    min_dist = distance
                    closest = child
            node = closest
        out[i] = node
    return out


def heap_path_to_ref(tree: dict, path: np.ndarray) -> list:
    """heap paths [n, depth+1] (-1 padded) -> list of reference-id lists, like
    ``propagate_feature`` returns (vocabulary.py:111-124)."""
    rid = tree["ref_id"]
    return [[int(rid[h]) for h in row if h >= 0] for row in path]


# ----------------------------------------------------------------------------- encoding
def path_counts(tree: dict, x_img, all_nodes: bool = True) -> np.ndarray:
    """Visit counts per reference node id for one image (vocabulary.py:92-100): dense int64
    [n_ref].  ``all_nodes=False`` counts the stop node only ("leaf words")."""
    counts = np.zeros(tree["n_ref"], np.int64)
    x_img = np.asarray(x_img)
    if x_img.shape[0] == 0:
        return counts
    leaf, path = descend(tree, x_img, want_path=True)
    if all_nodes:
        h = path[path >= 0]
    else:
        h = leaf
    np.add.at(counts, tree["ref_id"][h], 1)
    return counts


def embedding_dense(counts: np.ndarray) -> np.ndarray:
    """vocabulary.py:132-137: counts / ||counts||_2 in float64 (0/0 -> NaN like the reference)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return counts / np.linalg.norm(counts, ord=2)


def idf_weights(count_matrix: np.ndarray) -> np.ndarray:
    """Nister & Stewenius w_i = ln(N / N_i) (the weighting vocabulary.py:131,137 leaves
    commented out).  Nodes no image visits get weight 0."""
    n = count_matrix.shape[0]
    df = (count_matrix > 0).sum(axis=0)
    w = np.zeros(count_matrix.shape[1], np.float64)
    nz = df > 0
    w[nz] = np.log(n / df[nz])
    return w


def weighted_embedding(counts, weights=None, norm: str = "l2") -> np.ndarray:
    v = counts.astype(np.float64)
    if weights is not None:
        v = v * weights
    with np.errstate(invalid="ignore", divide="ignore"):
        if norm == "l2":
            return v / np.sqrt(np.sum(v * v))
        return v / np.sum(np.abs(v))


# ----------------------------------------------------------------------------- scoring
def score_dense(d: np.ndarray, q: np.ndarray) -> float:
    """database.py:59-70 verbatim arithmetic."""
    with np.errstate(invalid="ignore", divide="ignore"):
        d = d / np.linalg.norm(d, ord=2)
        q = q / np.linalg.norm(q, ord=2)
        s = np.linalg.norm(d - q, ord=2)
    return float(s) if not np.isnan(s) else 1e6


def score_lp(d: np.ndarray, q: np.ndarray, norm: str = "l2") -> float:
    """Distance between two already normalised vectors (opt-in tf-idf / L1 modes)."""
    with np.errstate(invalid="ignore"):
        s = np.sqrt(np.sum((d - q) ** 2)) if norm == "l2" else np.sum(np.abs(d - q))
    return float(s) if not np.isnan(s) else 1e6


def retrieve_order(scores: np.ndarray) -> np.ndarray:
    """database.py:79-80: ascending, Python's sort is stable => ties keep image order."""
    return np.argsort(scores, kind="stable")


def postings(count_matrix: np.ndarray) -> dict:
    """node ref id -> {image index: count}: the dict-per-node inverted file that
    ``propagate`` grows on ``graph.nodes[node]`` (vocabulary.py:96-100)."""
    out = {}
    img, node = np.nonzero(count_matrix)
    for i, n_ in zip(img.tolist(), node.tolist()):
        out.setdefault(n_, {})[i] = int(count_matrix[i, n_])
    return out
