"""Shared test helpers: golden loading and reference-shaped conversions (tests only)."""
import json
import os

import numpy as np

from oracle import vt_oracle as vo
from oracle.make_golden import case_images, inputs_digest  # noqa: F401  (case definitions are shared)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, "%s.npz" % name))
    meta = json.loads(str(g["meta"]))
    return g, meta


def golden_inputs(name):
    """(images uint8 list, k, depth, n_iter) regenerated from seeds and checked against the digest
    stored with the golden file."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = case_images(name)
    assert inputs_digest(images) == meta["inputs_sha256"], "synthetic generator drifted from the goldens"
    return images, k, depth, n_iter


def golden_tree_dicts(g, k):
    """reference ``nodes`` / ``tree`` dicts from the stored arrays."""
    nodes = {i: row for i, row in enumerate(g["ref_nodes"])}
    tree = {}
    for i, kids in enumerate(g["ref_children"]):
        if kids[0] >= 0:
            tree[i] = [int(c) for c in kids[:k]]
    return nodes, tree


def oracle_tree_from_flat(ft):
    """dict the oracle functions accept, from a product FlatTree (or anything with its fields)."""
    return dict(k=ft.k, depth=ft.depth, D=ft.D, DP=ft.DP, centroids=ft.centroids, exists=ft.exists,
                internal=ft.internal, ref_id=ft.ref_id, n_ref=ft.n_ref, heap_of_ref=ft.heap_of_ref)


def near_tie_rows(tree, x, rel=1e-5):
    """rows whose descent met a runner-up within ``rel`` (relative, binary64) of the winner."""
    _, gap = vo.descend(tree, x, want_gap=True)
    return gap < rel


def csr_to_dense(row_off, node, count, n_ref):
    n_img = len(row_off) - 1
    out = np.zeros((n_img, n_ref), np.int64)
    for i in range(n_img):
        out[i, node[row_off[i]:row_off[i + 1]]] = count[row_off[i]:row_off[i + 1]]
    return out
