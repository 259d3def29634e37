"""CPU: host-side logic of the package (tree numbering, reference-shaped views, boundary types)."""
import numpy as np
import pytest

from oracle import vt_oracle as vo
from helpers import golden_inputs, golden_tree_dicts, load_golden


@pytest.mark.parametrize("name", ["toy", "small", "orb", "adversarial"])
def test_flat_tree_numbering_matches_reference(vt, name):
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    nodes, tree = golden_tree_dicts(g, k)
    ft = vt.FlatTree.from_reference_dicts(nodes, tree, k, depth)
    assert ft.n_ref == meta["n_ref"]
    assert ft.tree_dict() == tree
    back = ft.nodes_dict()
    assert all(np.array_equal(back[i], nodes[i]) for i in nodes)
    # parent / level tables
    for node, kids in tree.items():
        for c in kids:
            assert ft.parent_ref[c] == node and ft.level_ref[c] == ft.level_ref[node] + 1
    assert ft.parent_ref[0] == -1 and ft.level_ref[0] == 0
    # path reconstruction == what propagate_feature returns
    t = vo.build_tree(np.concatenate(images), k, depth, n_iter)
    x = np.concatenate(images)[:50]
    leaf, path = vo.descend(t, x, want_path=True)
    for h, p in zip(leaf, vo.heap_path_to_ref(t, path)):
        assert ft.path_of(h) == p


def test_ragged_numbering_random(vt):
    rng = np.random.default_rng(5)
    for k, depth in [(2, 5), (3, 4), (10, 3)]:
        nh = vt.tree.heap_size(k, depth)
        exists = np.zeros(nh, np.uint8)
        internal = np.zeros(nh, np.uint8)
        exists[0] = 1
        for h in range(nh):
            if exists[h] and h * k + k < nh and rng.random() < 0.7:
                internal[h] = 1
                exists[h * k + 1:h * k + 1 + k] = 1
        ref, n_ref = vt.tree.reference_ids(exists, internal, k, depth)
        # literal DFS pre-order
        want = np.full(nh, -1, np.int64)
        counter = [0]

        def dfs(h):
            want[h] = counter[0]
            counter[0] += 1
            if internal[h]:
                for c in range(k):
                    dfs(h * k + 1 + c)
        dfs(0)
        assert n_ref == counter[0] and np.array_equal(ref, want)


def test_database_type_errors(vt):
    with pytest.raises(TypeError):
        vt.Database(42, encoder=None)                       # database.py:17
    ds = vt.ArrayDataset([np.zeros((2, 4), np.uint8)])
    with pytest.raises(FileNotFoundError):
        ds.read_image("nope.jpg")                           # dataset.py:60-61
    with pytest.raises(ValueError):
        vt.Database(ds, encoder=None, weighting="tf", norm="l1")


def test_utils_semantics(vt):
    out = vt.utils.show_progress(lambda v: [v, v], [1, 2, 3], quiet=True)
    assert out == [1, 1, 2, 2, 3, 3]                        # extend, not append (utils.py:23)
    a = np.arange(12, dtype=np.uint8).reshape(3, 4)
    import hashlib
    assert vt.utils.get_image_id(a) == hashlib.sha1(a).hexdigest()


def test_array_dataset_is_sorted(vt):
    ds = vt.ArrayDataset([np.zeros((1, 4)), np.ones((1, 4))], names=["b.jpg", "a.jpg"])
    assert ds.image_paths == ["a.jpg", "b.jpg"] and ds.read_image("a.jpg")[0, 0] == 1


def test_no_cpu_fallback(vt):
    """Without a device the product path must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    voc = vt.encoders.VocabularyTree(2, 2, descriptor=lambda im: im)
    with pytest.raises(vt._native.NativeError):
        voc.fit(np.array([[1.], [2.], [11.], [12.]]))


def test_product_does_not_import_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "scalable-recognition-with-a-vocabulary-tree_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
