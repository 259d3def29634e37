"""CPU: the oracle (oracle/vt_oracle.*) against outputs of the reference itself (tests/golden/,
made by oracle/make_golden.py from the unmodified reference package)."""
import hashlib

import numpy as np
import pytest

from oracle import vt_oracle as vo
from helpers import golden_inputs, golden_tree_dicts, load_golden

SMALL = ["toy", "small", "orb", "adversarial"]


def _ref_tables(t, d, k):
    ex = np.nonzero(t["exists"])[0]
    nodes = np.zeros((t["n_ref"], d), np.float32)
    nodes[t["ref_id"][ex]] = t["centroids"][ex, :d]
    ch = np.full((t["n_ref"], k), -1, np.int32)
    for h in ex:
        if t["internal"][h]:
            ch[t["ref_id"][h]] = t["ref_id"][h * k + 1:h * k + 1 + k]
    return nodes, ch


@pytest.mark.parametrize("name", SMALL)
def test_build_matches_reference_fit(name):
    """vocabulary.py:36-76 with the seeded Lloyd: same node ids, children and centres."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    t = vo.build_tree(np.concatenate(images), k, depth, n_iter)
    assert t["n_ref"] == meta["n_ref"]
    nodes, ch = _ref_tables(t, meta["D"], k)
    assert np.array_equal(ch, g["ref_children"])
    assert np.array_equal(nodes[1:], g["ref_nodes"][1:])                       # bit exact
    # root: the reference takes np.mean in float32, the oracle the exact mean
    np.testing.assert_allclose(nodes[0], g["ref_nodes"][0], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", SMALL)
def test_descent_matches_propagate_feature(name):
    """vocabulary.py:104-124 (float32 literal) vs canonical binary64: equal except near-ties."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    x = np.concatenate(images)
    t = vo.build_tree(x, k, depth, n_iter)
    leaf, path, gap = vo.descend(t, x, want_path=True, want_gap=True)
    mine = t["ref_id"][leaf]
    mism = np.nonzero(mine != g["ref_leaf"])[0]
    assert np.all(gap[mism] < 1e-5), "a non-near-tie descent differs from the reference"
    assert len(mism) <= max(2, int(2e-5 * len(x)))
    depth_of = (path >= 0).sum(1) - 1
    ok = np.setdiff1d(np.arange(len(x)), mism)
    assert np.array_equal(depth_of[ok], g["ref_depth"][ok])


@pytest.mark.parametrize("name", SMALL)
def test_encoding_and_scoring_match_reference(name):
    """propagate/embedding (vocabulary.py:78-137), score/retrieve (database.py:59-81)."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    t = vo.build_tree(np.concatenate(images), k, depth, n_iter)
    counts = np.stack([vo.path_counts(t, im) for im in images])
    assert np.array_equal(counts, g["counts"])                                 # posting lists
    emb = np.stack([vo.embedding_dense(c) for c in counts])
    assert np.array_equal(emb, g["emb"], equal_nan=True)                        # bit exact float64
    for qi, q in enumerate(g["q_idx"]):
        sc = np.array([vo.score_dense(emb[i], emb[q]) for i in range(len(images))])
        order = vo.retrieve_order(sc)
        assert np.array_equal(order, g["order"][qi])
        assert np.array_equal(sc[order], g["scores"][qi])


@pytest.mark.parametrize("name", ["small", "orb"])
def test_literal_restatement_reproduces_reference_paths(name):
    """The Python-loop restatement with the reference's float32 arithmetic gives the reference's leaves
    (checked on a slice; exact on the machine/BLAS that generated the goldens, near-ties aside elsewhere)."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    x = np.concatenate(images)[:1500]
    t = vo.build_tree(np.concatenate(images), k, depth, n_iter)
    got = t["ref_id"][vo.descend_literal(t, x)]
    _, gap = vo.descend(t, x, want_gap=True)
    bad = np.nonzero(got != g["ref_leaf"][:len(x)])[0]
    assert np.all(gap[bad] < 1e-5) and len(bad) <= 2


def test_adversarial_case_has_the_edge_cases():
    g, meta = load_golden("adversarial")
    counts = g["counts"]
    assert counts[7].sum() == 0                                                # zero-descriptor image
    assert np.array_equal(counts[5], counts[2])                                # identical images
    q = list(g["q_idx"]).index(2)
    order, scores = g["order"][q], g["scores"][q]
    assert scores[0] == 0.0 and scores[1] == 0.0 and set(order[:2]) == {2, 5} and order[0] == 2
    assert scores[-1] == 1e6 and order[-1] == 7                                # NaN -> 1e6, sorted last


def test_toy_example_notebook_cell_54():
    g, meta = load_golden("toy")
    nodes, tree = golden_tree_dicts(g, 2)
    assert tree == {0: [1, 4], 1: [2, 3], 4: [5, 6]}
    assert nodes[0][0] == 6.5
    assert sorted(float(nodes[i][0]) for i in (2, 3, 5, 6)) == [1.0, 2.0, 11.0, 12.0]
    assert sorted(float(nodes[i][0]) for i in (1, 4)) == [1.5, 11.5]           # exact Lloyd centres


def test_exact_ties_first_child_wins():
    """Duplicate centroids: strict '<' keeps the first child (vocabulary.py:119)."""
    c = np.zeros((4, 32), np.float32)
    c[0, :3] = [5, 5, 5]
    c[1, :3] = [1, 2, 3]
    c[2, :3] = [1, 2, 3]
    c[3, :3] = [9, 9, 9]
    x = vo.padded(np.array([[1, 2, 3], [2, 2, 2], [9, 9, 8]], np.float32))
    assert vo.assign(x, c).tolist() == [1, 1, 3]


def test_c1_descent_and_topk_against_reference():
    """BASELINE configs[0] (1,491 x 1,000 SIFT-128, k=10, L=3): descent of all 1.49 M descriptors
    through the reference's tree, near-tie accounting, and top-10 of 64 queries."""
    g, meta = load_golden("c1")
    images, k, depth, n_iter = golden_inputs("c1")
    nodes, tree = golden_tree_dicts(g, k)
    import vtree_b200 as vt
    ft = vt.FlatTree.from_reference_dicts(nodes, tree, k, depth)
    t = dict(k=k, depth=depth, D=128, DP=128, centroids=ft.centroids, exists=ft.exists, internal=ft.internal,
             ref_id=ft.ref_id, n_ref=ft.n_ref)
    x = np.concatenate(images)
    leaf, gap = vo.descend(t, x, want_gap=True)
    mine = t["ref_id"][leaf]
    ref = g["ref_leaf"].astype(np.int32)
    assert hashlib.sha256(ref.tobytes()).hexdigest() == str(g["ref_leaf_sha256"])
    mism = np.nonzero(mine != ref)[0]
    assert np.all(gap[mism] < 1e-5)
    assert len(mism) <= 30, "expected ~5e-7 per level (SURVEY 7.1-4), got %d" % len(mism)
    # encoding of the first 16 images + retrieval order of the stored queries
    counts = np.stack([vo.path_counts(t, im) for im in images[:16]])
    touched = np.unique(np.concatenate([np.arange(s, s + len(im)) for s, im in
                                        zip(np.cumsum([0] + [len(i) for i in images[:15]]), images[:16])]))
    if not np.intersect1d(touched, mism).size:
        assert np.array_equal(counts, g["counts_head"])
        assert np.array_equal(np.stack([vo.embedding_dense(c) for c in counts]), g["emb_head"])


@pytest.mark.slow
def test_c1_build_matches_reference_fit():
    g, meta = load_golden("c1")
    images, k, depth, n_iter = golden_inputs("c1")
    t = vo.build_tree(np.concatenate(images), k, depth, n_iter)
    nodes, ch = _ref_tables(t, 128, k)
    assert np.array_equal(ch, g["ref_children"])
    assert np.array_equal(nodes[1:], g["ref_nodes"][1:])
