"""The reference-side binding INTEGRATION.md shows (integration/b200_tree.py) is executed, not just printed.

CPU box: constructed on a tree built by the UNMODIFIED reference (/root/reference, when present), heap arrays
checked against the product's own import of that tree, ctypes signature checked against the header binding table.
GPU box: its embeddings must equal the reference's golden embeddings (the reference itself is not there; the
plugin only needs the `nodes` / `tree` dicts, which the goldens hold)."""
import importlib.util
import os
import types

import numpy as np
import pytest

from helpers import golden_inputs, golden_tree_dicts, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plugin():
    spec = importlib.util.spec_from_file_location("b200_tree", os.path.join(ROOT, "integration", "b200_tree.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _golden_voc(name):
    g, meta = load_golden(name)
    images, k, depth, _ = golden_inputs(name)
    nodes, tree = golden_tree_dicts(g, k)
    voc = types.SimpleNamespace(n_branches=k, depth=depth, nodes=nodes, tree=tree, descriptor=lambda im: im)
    return voc, g, images


def test_integration_md_shows_the_executed_file():
    code = open(os.path.join(ROOT, "integration", "b200_tree.py")).read()
    assert code in open(os.path.join(ROOT, "INTEGRATION.md")).read()


def test_plugin_signature_matches_binding_table(vt):
    mod = _plugin()
    res, args = vt._native.SIGNATURES["vt_quantize"]
    assert [a.__name__ for a in mod.VT_QUANTIZE_ARGTYPES] == [a.__name__ for a in args]
    lib = mod.load_library(vt._native.ensure_built())          # loads and binds vt_quantize / vt_last_error
    assert lib.vt_quantize.argtypes == mod.VT_QUANTIZE_ARGTYPES and mod.VT_F32 == vt._native.VT_F32


@pytest.mark.parametrize("name", ["toy", "small", "adversarial"])
def test_plugin_heap_arrays_equal_product_import(vt, name):
    mod = _plugin()
    voc, g, _ = _golden_voc(name)
    cent, internal, ref = mod.heap_arrays(voc, dp=vt.tree.pad_dim(len(voc.nodes[0])))
    ft = vt.tree.FlatTree.from_reference_dicts(voc.nodes, voc.tree, voc.n_branches, voc.depth)
    assert np.array_equal(cent, ft.centroids) and np.array_equal(internal, ft.internal)
    assert np.array_equal(ref, ft.ref_id)


@pytest.mark.skipif(not os.path.isdir("/root/reference/cbir"), reason="reference checkout not on this box")
def test_plugin_constructs_on_a_tree_built_by_the_unmodified_reference(vt):
    from oracle import ref_harness
    cbir = ref_harness.import_reference()
    images, k, depth, n_iter = golden_inputs("small")
    feats = np.concatenate(images).astype(np.float32)
    voc = ref_harness.reference_build(cbir, feats, k, depth, n_iter)
    voc.descriptor = lambda im: im
    mod = _plugin()
    plug = mod.B200Tree(voc, dp=vt.tree.pad_dim(feats.shape[1]))
    g, _ = load_golden("small")
    assert int((plug.ref_h >= 0).sum()) == len(voc.nodes) == len(g["ref_nodes"])
    assert np.array_equal(plug.cent_h[plug.ref_h >= 0][np.argsort(plug.ref_h[plug.ref_h >= 0])][:, :feats.shape[1]],
                          g["ref_nodes"])
    # the reference's Database accepts it as an encoder (duck typing, database.py:9-23)
    ds = ref_harness.make_dataset_class(cbir)(images)
    db = cbir.Database(ds, plug)
    assert db.encoder is plug and hasattr(plug, "embedding")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "adversarial"])
def test_plugin_embeddings_match_reference_goldens(vt, cuda, name):
    mod = _plugin()
    voc, g, images = _golden_voc(name)
    plug = mod.B200Tree(voc, dp=vt.tree.pad_dim(len(voc.nodes[0])), lib=mod.load_library(vt._native.LIB_PATH))
    counts = g["counts"].astype(np.float64)
    for i in range(min(8, len(images))):
        if len(images[i]) == 0:
            continue
        e = plug.embedding(images[i])
        with np.errstate(invalid="ignore", divide="ignore"):
            want = counts[i] / np.linalg.norm(counts[i])
        # float32 near-ties of the literal reference may move a single descriptor
        assert np.abs(e - want).max() <= 2.5 / max(np.linalg.norm(counts[i]), 1.0)
