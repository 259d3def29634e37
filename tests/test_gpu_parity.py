"""GPU parity tests: every CUDA entry point (through the C ABI) against the CPU oracle and the
reference's golden vectors.  Integer / index results are compared bit-exactly; float64 scores
within 1e-5 relative (BASELINE.json north_star).  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from oracle import vt_oracle as vo
from helpers import (csr_to_dense, golden_inputs, golden_tree_dicts, load_golden, near_tie_rows,
                     oracle_tree_from_flat)

pytestmark = pytest.mark.gpu
CASES = ["toy", "small", "orb", "adversarial"]


def _torch():
    import torch
    return torch


def _flat_from_golden(vt, name):
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    nodes, tree = golden_tree_dicts(g, k)
    return vt.FlatTree.from_reference_dicts(nodes, tree, k, depth), g, meta, images


# ------------------------------------------------------------------------------------------ sort
@pytest.mark.parametrize("n,bits", [(1, 8), (31, 3), (1000, 8), (4097, 13), (100003, 20), (3_000_001, 32)])
def test_sort_pairs_is_a_stable_sort(vt, cuda, n, bits):
    torch = _torch()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k_d = torch.from_numpy(keys.view(np.int32)).to(cuda)
    v_d = torch.from_numpy(vals.view(np.int32)).to(cuda)
    sk, sv = vt.engine.sort_pairs(k_d, v_d, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(sk.cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(sv.cpu().numpy().view(np.uint32), vals[order])


# ------------------------------------------------------------------------------------------ descent
@pytest.mark.parametrize("name", CASES + ["c1"])
@pytest.mark.parametrize("as_float", [False, True])
def test_quantize_matches_oracle_and_reference(vt, cuda, name, as_float):
    torch = _torch()
    ft, g, meta, images = _flat_from_golden(vt, name)
    x = np.concatenate(images)
    if name == "c1":
        x = x[:400_000] if as_float else x
    ot = oracle_tree_from_flat(ft)
    want_heap, gap = vo.descend(ot, x, want_gap=True)
    desc, d = vt.engine.stage(x.astype(np.float32) if as_float else x, cuda, prefer_bytes=not as_float)
    assert desc.dtype == (torch.float32 if as_float else torch.uint8)
    stats = torch.zeros(4, dtype=torch.int64, device=cuda)
    word, heap = vt.engine.quantize(ft.device(cuda), ft.k, ft.depth, desc, want_heap=True, stats=stats, method="fused")
    heap, word = heap.cpu().numpy(), word.cpu().numpy()
    assert np.array_equal(heap, want_heap), "%d rows differ from the binary64 oracle" % (heap != want_heap).sum()
    assert np.array_equal(word, ft.ref_id[want_heap])
    if not as_float:                                             # level-synchronous sorted path: same ids
        st2 = torch.zeros(4, dtype=torch.int64, device=cuda)
        w2, h2 = vt.engine.quantize(ft.device(cuda), ft.k, ft.depth, desc, want_heap=True, stats=st2, method="sorted")
        assert np.array_equal(h2.cpu().numpy(), want_heap) and np.array_equal(w2.cpu().numpy(), word)
    # against the reference's own float32 propagate_feature: only near-ties may differ
    ref_leaf = g["ref_leaf"].astype(np.int32)[:len(x)]
    mism = np.nonzero(word != ref_leaf)[0]
    assert np.all(gap[mism] < 1e-5)
    assert len(mism) <= max(2, int(3e-5 * len(x)))
    print("%s: %d/%d near-tie rows differ from the float32 reference; %d binary64 rechecks"
          % (name, len(mism), len(x), int(stats[0].item())))


def _random_tree(vt, rng, k, depth, d, ragged, tie_heavy):
    dp = vt.tree.pad_dim(d)
    nh = vt.tree.heap_size(k, depth)
    cent = np.zeros((nh, dp), np.float32)
    if tie_heavy:      # small integer centres + duplicated siblings: exact and near ties everywhere
        cent[:, :d] = rng.integers(0, 4, (nh, d)).astype(np.float32)
        dup = rng.random(nh) < 0.3
        cent[1:][dup[1:]] = cent[:-1][dup[1:]]
    else:
        cent[:, :d] = rng.random((nh, d), dtype=np.float32) * 255.0
    exists = np.zeros(nh, np.uint8)
    internal = np.zeros(nh, np.uint8)
    exists[0] = 1
    for h in range(nh):
        if exists[h] and h * k + k < nh and (not ragged or h == 0 or rng.random() < 0.8):
            internal[h] = 1
            exists[h * k + 1:h * k + 1 + k] = 1
    return vt.FlatTree(k, depth, d, cent, exists, internal)


@pytest.mark.parametrize("k,depth,d", [(2, 6, 1), (3, 5, 20), (4, 4, 32), (10, 3, 64), (10, 4, 128), (16, 3, 100),
                                       (31, 2, 128), (1, 3, 32)])
@pytest.mark.parametrize("tie_heavy", [False, True])
def test_quantize_random_trees(vt, cuda, k, depth, d, tie_heavy):
    torch = _torch()
    rng = np.random.default_rng(1000 * k + 10 * depth + d + tie_heavy)
    ft = _random_tree(vt, rng, k, depth, d, ragged=True, tie_heavy=tie_heavy)
    n = 20011
    x = rng.integers(0, 4 if tie_heavy else 256, (n, d)).astype(np.uint8)
    want, gap = vo.descend(oracle_tree_from_flat(ft), x, want_gap=True)
    for arr in (x, x.astype(np.float32) + (0.0 if tie_heavy else 0.25)):
        exp = want if arr.dtype == np.uint8 or tie_heavy else vo.descend(oracle_tree_from_flat(ft), arr)
        desc, _ = vt.engine.stage(arr, cuda)
        stats = torch.zeros(4, dtype=torch.int64, device=cuda)
        _, heap = vt.engine.quantize(ft.device(cuda), k, depth, desc, want_heap=True, stats=stats, method="fused")
        assert np.array_equal(heap.cpu().numpy(), exp)
        if tie_heavy and k > 1:
            assert int(stats[0].item()) > 0            # the binary64 path really ran
        if desc.dtype == torch.uint8:
            st2 = torch.zeros(4, dtype=torch.int64, device=cuda)
            w2, h2 = vt.engine.quantize(ft.device(cuda), k, depth, desc, want_heap=True, stats=st2, method="sorted")
            assert np.array_equal(h2.cpu().numpy(), exp)
            assert np.array_equal(w2.cpu().numpy(), ft.ref_id[exp])
    if tie_heavy and k > 1:
        assert (gap == 0).any()                        # exact ties were present (first child must win)


@pytest.mark.parametrize("case", ["small", "c1", "rand-k10", "rand-k16-d100", "ties-k4", "ragged-k3"])
def test_quantize_tensor_core_path_matches_oracle(vt, cuda, case):
    """vt_quantize_mma (tcgen05 kind::i8, exact integer distances + provable rounding margin) must
    return the same ids as the binary64 oracle."""
    torch = _torch()
    rng = np.random.default_rng(len(case))
    if case in ("small", "c1"):
        ft, g, meta, images = _flat_from_golden(vt, case)
        x = np.concatenate(images)
    else:
        k, depth, d, tie = {"rand-k10": (10, 4, 128, False), "rand-k16-d100": (16, 3, 100, False),
                            "ties-k4": (4, 5, 128, True), "ragged-k3": (3, 6, 128, False)}[case]
        ft = _random_tree(vt, rng, k, depth, d, ragged=True, tie_heavy=tie)
        x = rng.integers(0, 4 if tie else 256, (50_003, d)).astype(np.uint8)
    want = vo.descend(oracle_tree_from_flat(ft), x)
    desc, _ = vt.engine.stage(x, cuda)
    stats = torch.zeros(4, dtype=torch.int64, device=cuda)
    word, heap = vt.engine.quantize(ft.device(cuda), ft.k, ft.depth, desc, want_heap=True, stats=stats, method="mma")
    torch.cuda.synchronize()
    assert int(stats[1].item()) == 0                       # no centroid outside the fixed-point range
    heap = heap.cpu().numpy()
    assert np.array_equal(heap, want), "%d of %d rows differ" % ((heap != want).sum(), len(want))
    assert np.array_equal(word.cpu().numpy(), ft.ref_id[want])
    print(case, "binary64 rechecks:", int(stats[0].item()), "of", len(x) * ft.depth)


@pytest.mark.parametrize("scale", [2.0 ** -6, 2.0 ** -10, 2.0 ** -14, 2.0 ** -18])
def test_tensor_core_margin_under_near_ties(vt, cuda, scale):
    """The tensor-core kernel compares scores of ROUNDED centroids (8.16 fixed point) in units of 2^-7;
    whatever that loses must be inside its margin.  Siblings that differ by perturbations of the order of
    those granularities (and descriptors drawn close to the centroids, so squared distances are small and
    gaps tiny) make the margin decide a large share of the rows; ids must still equal the binary64 oracle."""
    torch = _torch()
    rng = np.random.default_rng(int(-np.log2(scale)))
    k, depth, d = 10, 3, 128
    ft = _random_tree(vt, rng, k, depth, d, ragged=False, tie_heavy=False)
    cent = ft.centroids.copy()
    for h in range(vt.tree.heap_size(k, depth - 1)):            # every internal node: children = one base + noise
        kids = slice(h * k + 1, h * k + 1 + k)
        base = cent[h * k + 1].copy()
        noise = rng.standard_normal((k, cent.shape[1])).astype(np.float32) * np.float32(scale * 40.0)
        cent[kids] = np.clip(base[None, :] + noise, 0.0, 255.0)
        cent[h * k + 1 + 3] = cent[h * k + 1 + 2]               # and one exact duplicate (first child must win)
    cent[:, d:] = 0
    ft = vt.FlatTree(k, depth, d, cent, ft.exists, ft.internal)
    leaves = cent[vt.tree.heap_size(k, depth - 1):]
    pick = rng.integers(0, len(leaves), 40_003)
    x = np.clip(np.rint(leaves[pick][:, :d] + rng.standard_normal((len(pick), d)) * 2.0), 0, 255).astype(np.uint8)
    want = vo.descend(oracle_tree_from_flat(ft), x)
    desc, _ = vt.engine.stage(x, cuda)
    stats = torch.zeros(4, dtype=torch.int64, device=cuda)
    word, heap = vt.engine.quantize(ft.device(cuda), k, depth, desc, want_heap=True, stats=stats, method="mma")
    torch.cuda.synchronize()
    assert int(stats[1].item()) == 0
    heap = heap.cpu().numpy()
    assert np.array_equal(heap, want), "%d of %d rows differ" % ((heap != want).sum(), len(want))
    rechecks = int(stats[0].item())
    assert rechecks > 0                                          # the margin was exercised
    print("scale 2^%d: %d binary64 rechecks for %d descriptor-levels" % (int(np.log2(scale)), rechecks, len(x) * depth))


def test_tensor_core_deferred_output_on_small_batches(cuda):
    """The last level of batches >= 2^22 rows leaves through the deferred-output path (row-range buckets).
    A child process lowers that threshold to 1 row (the library reads VT_MMA_DEFER_MIN once) and repeats
    the tensor-core parity cases -- ragged trees, exact ties, k = 16 -- on small batches."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, VT_MMA_DEFER_MIN="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-x",
                          "-m", "gpu", "-k", "test_quantize_tensor_core_path_matches_oracle", "-p", "no:cacheprovider"],
                         env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "6 passed" in out.stdout


@pytest.mark.parametrize("defer", ["0", "1"])
def test_two_levels_per_launch_on_small_batches(cuda, defer):
    """The tcgen05 descent can take TWO tree levels per launch (vt_quantize_pair.cu: persistent warp-specialised CTAs, TMA
    gathers, the second level re-gathers from L2; opt-in).  A child process sets the threshold to 1 row (VT_MMA_PAIR_MIN,
    read once by the library) and repeats the tensor-core parity cases -- ragged trees, exact ties, k = 16, odd and even depths, near-tie
    margins -- with the ids written directly (defer=0) and through the deferred (row, child) pairs (defer=1)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, VT_MMA_PAIR_MIN="1", VT_MMA_DEFER_MIN="1" if defer == "1" else str(1 << 40))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-x",
                          "-m", "gpu", "-k", "test_quantize_tensor_core_path_matches_oracle or test_tensor_core_margin_under_near_ties",
                          "-p", "no:cacheprovider"],
                         env=env, cwd=root, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "10 passed" in out.stdout


def test_quantize_start_node_and_empty_batch(vt, cuda):
    torch = _torch()
    ft, g, meta, images = _flat_from_golden(vt, "small")
    x = np.concatenate(images)[:100]
    desc, _ = vt.engine.stage(x, cuda)
    kids = ft.tree_dict()[0]
    node = kids[3]
    start = int(ft.heap_of_ref[node])
    w = vt.engine.quantize(ft.device(cuda), ft.k, ft.depth, desc, start=start).cpu().numpy()
    ot = oracle_tree_from_flat(ft)
    # oracle: descend a subtree by re-rooting through explicit loops
    for i in range(len(x)):
        h = start
        while ft.internal[h]:
            c = vo.assign(vo.padded(x[i:i + 1]), ft.centroids[h * ft.k + 1:h * ft.k + 1 + ft.k])[0]
            h = h * ft.k + 1 + c
        assert w[i] == ft.ref_id[h]
    empty = torch.empty((0, ft.DP), dtype=torch.uint8, device=cuda)
    assert vt.engine.quantize(ft.device(cuda), ft.k, ft.depth, empty).numel() == 0
    del ot


def test_quantize_host_pipeline_equals_device_path(vt, cuda):
    torch = _torch()
    ft, g, meta, images = _flat_from_golden(vt, "c1")
    x = np.concatenate(images)[:300_001]
    desc, _ = vt.engine.stage(x, cuda)
    want = vt.engine.quantize(ft.device(cuda), ft.k, ft.depth, desc).cpu().numpy()
    h_desc = torch.from_numpy(x).pin_memory()
    h_word = torch.empty(len(x), dtype=torch.int32).pin_memory()
    vt.engine.quantize_host(ft.device(cuda), ft.k, ft.depth, h_desc, h_word, chunk_rows=70_000)
    assert np.array_equal(h_word.numpy(), want)


# ------------------------------------------------------------------------------------------ build
@pytest.mark.parametrize("name", CASES)
def test_build_matches_reference_fit(vt, cuda, name):
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    x = np.concatenate(images)
    desc, d = vt.engine.stage(x, cuda)
    ft = vt.engine.build_tree(desc, d, k, depth, n_iter=n_iter)
    t = vo.build_tree(x, k, depth, n_iter)
    assert np.array_equal(ft.exists, t["exists"]) and np.array_equal(ft.internal, t["internal"])
    assert np.array_equal(ft.ref_id, t["ref_id"]) and ft.n_ref == meta["n_ref"]
    assert np.array_equal(ft.centroids, t["centroids"]), "centroids are not bit-identical to the oracle"
    assert np.array_equal(ft.count, t["count"])
    nodes = ft.nodes_dict()
    assert all(np.array_equal(nodes[i], g["ref_nodes"][i]) for i in range(1, ft.n_ref))
    assert ft.tree_dict() == golden_tree_dicts(g, k)[1]


def test_build_medium_tree_bit_exact(vt, cuda):
    """200 k descriptors, k=10, L=4 (deep levels use the global-atomic accumulation path)."""
    x = vt.synth.sift_like(200_000, 128, n_centres=3000, seed=77)
    desc, d = vt.engine.stage(x, cuda)
    log = []
    ft = vt.engine.build_tree(desc, d, 10, 4, n_iter=6, log=lambda **kw: log.append(kw))
    t = vo.build_tree(x, 10, 4, 6)
    assert np.array_equal(ft.internal, t["internal"]) and np.array_equal(ft.count, t["count"])
    assert np.array_equal(ft.centroids, t["centroids"])
    assert len(log) == 4


def test_build_simt_and_tensor_paths_agree(vt, cuda):
    """the FP32/atomic assign kernel and the tcgen05 assign+accumulate kernel build the same tree"""
    x = vt.synth.sift_like(60_000, 128, n_centres=500, seed=9)
    x[1000:1400] = x[1000]                                      # a big block of duplicates -> empty clusters
    desc, d = vt.engine.stage(x, cuda)
    a = vt.engine.build_tree(desc, d, 10, 3, n_iter=5, ops=vt.engine.CudaLevelOps(use_tensor_cores=True))
    b = vt.engine.build_tree(desc, d, 10, 3, n_iter=5, ops=vt.engine.CudaLevelOps(use_tensor_cores=False))
    t = vo.build_tree(x, 10, 3, 5)
    for ft in (a, b):
        assert np.array_equal(ft.centroids, t["centroids"]) and np.array_equal(ft.count, t["count"])
        assert np.array_equal(ft.internal, t["internal"])


def test_build_degenerate_inputs(vt, cuda):
    # fewer points than branches: the root is a leaf (vocabulary.py:55)
    x = np.array([[3, 4], [5, 6]], np.uint8)
    desc, d = vt.engine.stage(x, cuda)
    ft = vt.engine.build_tree(desc, d, 4, 3)
    assert ft.n_ref == 1 and not ft.internal.any() and ft.centroids[0, 0] == 4.0
    # all points identical: one full cluster, k-1 empty leaves that keep the initial centre
    x = np.full((50, 8), 7, np.uint8)
    desc, d = vt.engine.stage(x, cuda)
    ft = vt.engine.build_tree(desc, d, 3, 2)
    t = vo.build_tree(x, 3, 2, 10)
    assert np.array_equal(ft.centroids, t["centroids"]) and np.array_equal(ft.count, t["count"])
    assert np.array_equal(ft.ref_id, t["ref_id"])


# ------------------------------------------------------------------------------------------ encode
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("all_nodes", [True, False])
def test_encode_matches_propagate(vt, cuda, name, all_nodes):
    torch = _torch()
    ft, g, meta, images = _flat_from_golden(vt, name)
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    csr = voc.encode_batch(images, want_df=True, all_nodes=all_nodes)
    ro, node, count = (csr[k_].cpu().numpy() for k_ in ("row_off", "node", "count"))
    dense = csr_to_dense(ro, node, count, ft.n_ref)
    ot = oracle_tree_from_flat(ft)
    want = np.stack([vo.path_counts(ot, im, all_nodes=all_nodes) for im in images])
    assert np.array_equal(dense, want)
    if all_nodes:
        assert np.array_equal(dense, g["counts"])                       # the reference's postings
    assert np.array_equal(csr["sumsq"].cpu().numpy(), (want.astype(np.int64) ** 2).sum(1))
    assert np.array_equal(csr["df"].cpu().numpy(), (want > 0).sum(0))
    assert (ro[1:] - ro[:-1] == (want > 0).sum(1)).all()
    assert (count > 0).all()
    del torch


@pytest.mark.parametrize("name", ["small", "orb", "adversarial"])
def test_encode_level_mask(vt, cuda, name):
    """min_level = m: the reference's visit counts (golden) with every node above level m zeroed --
    the level mask of the paper (weights the reference leaves commented out, vocabulary.py:131,137).
    Early leaves above the mask vanish; retrieval still ranks a database image first for itself."""
    ft, g, meta, images = _flat_from_golden(vt, name)
    for m in range(0, ft.depth + 1):
        voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im, min_level=m).set_tree(ft)
        csr = voc.encode_batch(images, want_df=True)
        dense = csr_to_dense(*(csr[k_].cpu().numpy() for k_ in ("row_off", "node", "count")), ft.n_ref)
        want = g["counts"].astype(np.int64) * (ft.level_ref >= m)[None, :]
        assert np.array_equal(dense, want), m
        assert np.array_equal(csr["sumsq"].cpu().numpy(), (want ** 2).sum(1))
        assert np.array_equal(csr["df"].cpu().numpy(), (want > 0).sum(0))
    with pytest.raises(ValueError):
        vt.engine.encode(csr["node"], csr["row_off"], ft.device(cuda), ft.depth, ft.n_ref, min_level=ft.depth + 1)


def test_encode_large_images(vt, cuda):
    """images near the shared-memory capacity (8192 descriptors) and a mix of sizes"""
    ft, g, meta, images = _flat_from_golden(vt, "small")
    rng = np.random.default_rng(3)
    pool = np.concatenate(images)
    imgs = [pool[rng.integers(0, len(pool), n)] for n in (8192, 1, 2049, 0, 1025, 5000, 33)]
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    ot = oracle_tree_from_flat(ft)
    for all_nodes in (True, False):
        csr = voc.encode_batch(imgs, all_nodes=all_nodes)
        dense = csr_to_dense(*(csr[k_].cpu().numpy() for k_ in ("row_off", "node", "count")), ft.n_ref)
        want = np.stack([vo.path_counts(ot, im, all_nodes=all_nodes) for im in imgs])
        assert np.array_equal(dense, want)
    # beyond the per-CTA capacity: encoded in pieces and merged (the reference has no limit, vocabulary.py:78-102)
    big = [pool[rng.integers(0, len(pool), n)] for n in (8193, 7, 30000)]
    for all_nodes in (True, False):
        csr = voc.encode_batch(big, all_nodes=all_nodes)
        dense = csr_to_dense(*(csr[k_].cpu().numpy() for k_ in ("row_off", "node", "count")), ft.n_ref)
        assert np.array_equal(dense, np.stack([vo.path_counts(ot, im, all_nodes=all_nodes) for im in big]))


# ------------------------------------------------------------------------------------------ index + score
def _expected_scores(counts, q_rows):
    emb = np.stack([vo.embedding_dense(c) for c in counts])
    return np.array([[vo.score_dense(emb[i], emb[q]) for i in range(len(counts))] for q in q_rows])


def _assert_topk(ids, scores, want_scores_full, k, rtol=1e-5):
    """ids must be the reference's order; swaps are only tolerated inside groups whose reference
    scores agree to 1e-9 (rounding-level ties, documented in DESIGN.md)."""
    for qi in range(len(ids)):
        ws = want_scores_full[qi]
        order = vo.retrieve_order(ws)[:k]
        got = ids[qi][:len(order)]
        np.testing.assert_allclose(scores[qi][:len(order)], ws[order], rtol=rtol, atol=1e-7)
        bad = np.nonzero(got != order)[0]
        for b in bad:
            assert abs(ws[got[b]] - ws[order[b]]) <= 1e-9 * max(1.0, abs(ws[order[b]])), (qi, b, got, order)


@pytest.mark.parametrize("name", CASES)
def test_postings_and_topk_match_reference(vt, cuda, name):
    ft, g, meta, images = _flat_from_golden(vt, name)
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    csr = voc.encode_batch(images, want_df=True)
    index = vt.engine.build_index(csr, ft.n_ref)
    # posting lists: per node the (image, count) pairs, in image order
    off = index["post_off"].cpu().numpy().astype(np.int64)
    packed = index["postings"].cpu().numpy().view(np.uint32)          # (tf << 13) | local image index
    post = np.stack([packed & 8191, packed >> 13], axis=1).astype(np.int64)
    counts = g["counts"]
    assert off[-1] == (counts > 0).sum() and index["n_shards"] == 1
    for node in range(ft.n_ref):
        seg = post[off[node]:off[node + 1]]
        imgs = np.nonzero(counts[:, node])[0]
        assert np.array_equal(seg[:, 0], imgs) and np.array_equal(seg[:, 1], counts[imgs, node])
    # top-k for every image as query
    k = min(10, len(images))
    out = vt.engine.score(index, csr, k, want_all=True)
    ids, sc = out["id"].cpu().numpy(), out["score"].cpu().numpy()
    ref_sc = np.empty_like(g["scores"])
    for qi in range(len(g["q_idx"])):
        ref_sc[qi, g["order"][qi]] = g["scores"][qi]
    assert np.array_equal(g["q_idx"], np.arange(len(images)))
    _assert_topk(ids, sc, ref_sc, k)
    # exact reference order (no tolerance) wherever the reference scores are distinct
    for qi in range(len(images)):
        top = g["order"][qi][:k]
        if len(np.unique(g["scores"][qi][:k + 1])) == min(k + 1, len(images)):
            assert np.array_equal(ids[qi], top)


def test_score_multi_shard_random_csr(vt, cuda):
    """20 k images (3 shards), synthetic sparse count vectors, leaf-words shaped."""
    torch = _torch()
    import scipy.sparse as sp
    rng = np.random.default_rng(9)
    n_img, n_ref, nnz_per = 20_000, 5000, 40
    rows = np.repeat(np.arange(n_img), nnz_per)
    cols = rng.integers(0, n_ref, n_img * nnz_per)
    m = sp.coo_matrix((np.ones(len(rows), np.int64), (rows, cols)), shape=(n_img, n_ref)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    m[17] = 0
    m.eliminate_zeros()                                         # an image without words
    m = m.tocsr()
    csr = dict(row_off=torch.from_numpy(m.indptr.astype(np.int64)).to(cuda),
               node=torch.from_numpy(m.indices.astype(np.int32)).to(cuda),
               count=torch.from_numpy(m.data.astype(np.int32)).to(cuda),
               sumsq=torch.from_numpy(np.asarray(m.multiply(m).sum(1)).ravel().astype(np.int64)).to(cuda),
               n_img=n_img, nnz=m.nnz)
    index = vt.engine.build_index(csr, n_ref)
    assert index["n_shards"] == 3
    q_rows = np.array([0, 5, 8191, 8192, 16384, 19999, 17, 12345])
    mq = m[q_rows]
    q = dict(row_off=torch.from_numpy(mq.indptr.astype(np.int64)).to(cuda),
             node=torch.from_numpy(mq.indices.astype(np.int32)).to(cuda),
             count=torch.from_numpy(mq.data.astype(np.int32)).to(cuda),
             sumsq=torch.from_numpy(np.asarray(mq.multiply(mq).sum(1)).ravel().astype(np.int64)).to(cuda),
             n_img=len(q_rows), nnz=mq.nnz)
    k = 10
    out = vt.engine.score(index, q, k, want_all=True)
    ids, sc, allk = out["id"].cpu().numpy(), out["score"].cpu().numpy(), out["all_key"].cpu().numpy()
    norm = np.sqrt(np.asarray(m.multiply(m).sum(1)).ravel().astype(np.float64))
    dots = np.asarray((mq @ m.T).todense()).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        key = np.where(norm > 0, dots / norm, -1.0)
    np.testing.assert_allclose(allk, key, rtol=1e-12)
    for qi, qr in enumerate(q_rows):
        if norm[qr] == 0:
            assert (sc[qi] == 1e6).all()
            continue
        order = np.lexsort((np.arange(n_img), -key[qi]))[:k]
        assert np.array_equal(ids[qi], order), (qi, ids[qi], order)
        want = np.sqrt(np.maximum(0, 2 - 2 * key[qi][order] / norm[qr]))
        want[order == qr] = 0.0
        np.testing.assert_allclose(sc[qi], want, rtol=1e-5, atol=1e-7)
        assert ids[qi][0] == qr and sc[qi][0] == 0.0            # self match first, exactly zero


@pytest.mark.parametrize("n_ref", [3000, 400_000])
@pytest.mark.parametrize("shards_per_step", [1, 2, 4])
@pytest.mark.parametrize("groups", [1, 2, 5])
def test_carried_topk_equals_exhaustive_selection(vt, cuda, groups, shards_per_step, n_ref):
    """vt_score_query.cu: a CTA carries the k best of a query from shard to shard and prunes with their worst key.  40 k
    images (5 shards) with planted duplicates (exact key ties across shards, broken by the smaller image index like
    database.py:79-80), images without words, queries with no match at all, long and short posting runs; every grouping
    of the shards must return the order the full key dump gives -- with 32-bit accumulators (one shard per step) and with
    16-bit accumulators (2 or 4 shards per step, groups that end inside a step), on an index with runs of ~80 postings and
    on a sparse one (~0.6 postings per run)."""
    torch = _torch()
    import scipy.sparse as sp
    rng = np.random.default_rng(21 + groups)
    n_img, nnz_per = 40_000, 30                                 # n_ref 3000: runs of ~80 postings per (word, shard); 400 k: ~0.6
    rows = np.repeat(np.arange(n_img), nnz_per)
    cols = rng.integers(0, n_ref, n_img * nnz_per)
    cols[rng.random(len(cols)) < 0.2] = 7                       # one very frequent word: runs of thousands of postings
    m = sp.coo_matrix((rng.integers(1, 4, len(rows)).astype(np.int64), (rows, cols)), shape=(n_img, n_ref)).tolil()
    dup_src = [11, 9000, 17000, 30000]
    for j, src in enumerate(dup_src):                           # copies of an image in other shards: equal keys
        for dst in ((src + 8192 * (1 + j % 2) + 3) % n_img, (src + 20011) % n_img):
            m[dst] = m[src]
    for e in (17, 8192, 39999):
        m[e] = 0
    m = m.tocsr()
    m.eliminate_zeros()
    m.sort_indices()

    def to_csr(mat):
        return dict(row_off=torch.from_numpy(mat.indptr.astype(np.int64)).to(cuda),
                    node=torch.from_numpy(mat.indices.astype(np.int32)).to(cuda),
                    count=torch.from_numpy(mat.data.astype(np.int32)).to(cuda),
                    sumsq=torch.from_numpy(np.asarray(mat.multiply(mat).sum(1)).ravel().astype(np.int64)).to(cuda),
                    n_img=mat.shape[0], nnz=mat.nnz)
    index = vt.engine.build_index(to_csr(m), n_ref)
    assert index["n_shards"] == 5
    q_rows = np.concatenate([dup_src, [0, 17, 8191, 8193, 25000, 39998], rng.integers(0, n_img, 54)])
    mq = m[q_rows].tolil()
    mq[5] = 0
    mq[5, n_ref - 1] = 2                                        # a query whose only word is (almost surely) rare
    mq = mq.tocsr()
    q = to_csr(mq)
    k = 12
    full = vt.engine.score(index, q, k, want_all=True)
    allk = full["all_key"].cpu().numpy()
    got = vt.engine.score(index, q, k, groups=groups, shards_per_step=shards_per_step)
    auto = vt.engine.score(index, q, k)
    assert torch.equal(auto["id"], got["id"]) and torch.equal(auto["key"], got["key"])
    ids, key, sc = got["id"].cpu().numpy(), got["key"].cpu().numpy(), got["score"].cpu().numpy()
    for qi in range(len(q_rows)):
        order = np.lexsort((np.arange(n_img), -allk[qi]))[:k]
        assert np.array_equal(ids[qi], order), (groups, qi, ids[qi], order)
        assert np.array_equal(key[qi], allk[qi][order])
    assert np.array_equal(ids, full["id"].cpu().numpy()) and np.array_equal(sc, full["score"].cpu().numpy())
    for j, src in enumerate(dup_src):                           # the copies tie with the original, smaller index first
        assert sc[j][0] == 0.0 and sc[j][1] == 0.0 and sc[j][2] == 0.0 and list(ids[j][:3]) == sorted(ids[j][:3])
    if shards_per_step != 1:
        return
    # tf-idf / L2 through the same carried kernel with binary64 accumulators: the keys of the full dump (per-shard kernel),
    # up to the order of the binary64 additions
    csr = to_csr(m)
    df = torch.bincount(csr["node"].long(), minlength=n_ref).to(torch.int32)
    idf = vt.engine.idf_from_df(df, n_img)
    weights = dict(idf=idf, norm="l2", img_rnorm=vt.engine.weighted_rnorm(csr, idf, "l2"))
    mode = vt._native.SCORE_W_L2
    fullw = vt.engine.score(index, q, k, mode=mode, weights=weights, want_all=True)
    gotw = vt.engine.score(index, q, k, mode=mode, weights=weights, groups=groups)
    wk, wid, wkey = fullw["all_key"].cpu().numpy(), gotw["id"].cpu().numpy(), gotw["key"].cpu().numpy()
    for qi in range(len(q_rows)):
        best = np.sort(wk[qi])[::-1][:k]
        np.testing.assert_allclose(wkey[qi], best, rtol=1e-12, atol=1e-15)
        assert len(set(wid[qi].tolist())) == k
        np.testing.assert_allclose(wk[qi][wid[qi]], wkey[qi], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(gotw["score"].cpu().numpy(), fullw["score"].cpu().numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("norm", ["l2", "l1"])
def test_tfidf_modes_against_dense_oracle(vt, cuda, norm):
    ft, g, meta, images = _flat_from_golden(vt, "orb")
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    ds = vt.ArrayDataset(images)
    db = vt.Database(ds, voc, weighting="tfidf", norm=norm)
    db.index()
    counts = g["counts"].astype(np.int64)
    w = vo.idf_weights(counts)
    emb = np.stack([vo.weighted_embedding(c, w, norm) for c in counts])
    q_rows = [0, 3, 40, 79]
    ids, sc = db.retrieve_batch([ds.image_paths[i] for i in q_rows], k=8)
    for qi, q in enumerate(q_rows):
        ws = np.array([vo.score_lp(emb[i], emb[q], norm) for i in range(len(images))])
        order = vo.retrieve_order(ws)[:8]
        # binary64 weights, norms and accumulation (csrc/vt_weight.cu): north_star's 1e-5 relative on scores
        np.testing.assert_allclose(sc[qi], ws[order], rtol=1e-5, atol=1e-12)
        assert sc[qi][0] == 0.0
        # top-k ids in the reference's order (stable ascending sort); only rounding-level score ties may swap
        for a, b in zip(ids[qi], order):
            assert a == b or abs(ws[a] - ws[b]) <= 1e-9, (qi, ids[qi], order)
    # the weights themselves: idf and the per-image norms against the dense binary64 restatement
    np.testing.assert_allclose(db._weights["idf"].cpu().numpy(), w, rtol=1e-12, atol=0)
    v = counts * w
    ref_norm = np.sqrt((v * v).sum(1)) if norm == "l2" else np.abs(v).sum(1)
    np.testing.assert_allclose(1.0 / db._weights["img_rnorm"].cpu().numpy(), ref_norm, rtol=1e-12)
    # the full ranking of one query through retrieve() equals the oracle's
    full = db.retrieve(ds.image_paths[3])
    ws3 = np.array([vo.score_lp(emb[i], emb[3], norm) for i in range(len(images))])
    got = [ds.image_paths.index(p) for p in full]
    for a, b in zip(got, vo.retrieve_order(ws3)):
        assert a == b or abs(ws3[a] - ws3[b]) <= 1e-9
    np.testing.assert_allclose(list(full.values()), np.sort(ws3, kind="stable"), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------ API flow
@pytest.mark.parametrize("name", CASES)
def test_notebook_flow_matches_reference(vt, cuda, name):
    """NB[78-91]: VocabularyTree(...).fit -> Database(...).index -> retrieve, through the public
    API, against the reference's outputs on the same inputs."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    ds = vt.ArrayDataset([im.astype(np.float32) for im in images])
    voc = vt.encoders.VocabularyTree(n_branches=k, depth=depth, descriptor=lambda im: im, n_iter=n_iter)
    features = np.concatenate([ds.read_image(p) for p in ds.image_paths])
    voc.fit(features)
    nodes, tree = golden_tree_dicts(g, k)
    assert voc.tree == tree
    assert all(np.array_equal(voc.nodes[i], nodes[i]) for i in range(1, len(nodes)))
    np.testing.assert_allclose(voc.nodes[0], nodes[0], rtol=1e-5, atol=1e-5)
    # propagate_feature paths end at the reference's leaf
    for i in range(0, len(features), max(1, len(features) // 25)):
        p = voc.propagate_feature(features[i])
        assert p[0] == 0 and p[-1] == g["ref_leaf"][i] and len(p) - 1 == g["ref_depth"][i]
    db = vt.Database(ds, voc)
    db.index()
    for i in (0, len(images) // 2, len(images) - 1):
        assert np.array_equal(db.embedding(ds.image_paths[i]), g["emb"][i], equal_nan=True)
        assert np.array_equal(voc.embedding(ds.read_image(ds.image_paths[i])), g["emb"][i], equal_nan=True)
    for qi in (0, len(images) // 3, len(images) - 1):
        res = db.retrieve(ds.image_paths[qi])
        got_order = [ds.image_paths.index(p) for p in res]
        ws = np.empty(len(images))
        ws[g["order"][qi]] = g["scores"][qi]
        np.testing.assert_allclose(list(res.values()), g["scores"][qi], rtol=1e-5, atol=1e-7)
        for r, (a, b) in enumerate(zip(got_order, g["order"][qi])):
            assert a == b or abs(ws[a] - ws[b]) <= 1e-9, (qi, r)
        other = min(1, len(images) - 1)
        s = db.score(ds.image_paths[other], ds.image_paths[qi])
        assert s == ws[other]
    # reference-shaped postings
    post = voc.postings()
    sha = vt.utils.get_image_id(ds.read_image(ds.image_paths[0]))
    got0 = {n_: d_[sha] for n_, d_ in post.items() if sha in d_}
    assert got0 == {int(n_): int(c) for n_, c in enumerate(g["counts"][0]) if c > 0}


def test_streaming_index_equals_one_batch(vt, cuda):
    """Database.index(batch_images=b): images encoded b at a time give the same inverted file
    (streaming index) -- CSR, document frequencies and retrieval are identical."""
    import torch
    g, meta = load_golden("adversarial")
    images, k, depth, n_iter = golden_inputs("adversarial")
    ds = vt.ArrayDataset(images)
    voc = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, n_iter=n_iter)
    voc.fit(np.concatenate(images))
    dbs = []
    for b in (4096, 7, 1):
        for weighting in ("tf", "tfidf"):
            db = vt.Database(ds, voc, weighting=weighting)
            db.index(batch_images=b)
            dbs.append((weighting, db))
    for weighting in ("tf", "tfidf"):
        same = [db for w, db in dbs if w == weighting]
        for db in same[1:]:
            for a, b_ in zip(same[0]._host_csr(), db._host_csr()):
                assert np.array_equal(a, b_)
            assert torch.equal(same[0]._csr["df"], db._csr["df"])
            ids0, sc0 = same[0].retrieve_batch(ds.image_paths, k=5)
            ids1, sc1 = db.retrieve_batch(ds.image_paths, k=5)
            assert np.array_equal(ids0, ids1)
            if weighting == "tf":                                  # integer arithmetic: identical
                assert np.array_equal(sc0, sc1)
            else:                                                  # float32 weights, atomically summed norms
                np.testing.assert_allclose(sc0, sc1, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["small", "orb", "adversarial"])
def test_dense_encoder_protocol(vt, cuda, name):
    """ANY object with ``embedding(image)`` is an encoder for Database
    (encoders/new_encoder_example.py:1-4, e.g. the reference's AlexNet).  Here the plugin returns the
    reference's own embeddings (golden), so Database.retrieve through the dense GPU path must reproduce
    the reference's scores and order -- including the NaN -> 1e6 image and the exact 0 of identical images."""
    g, meta = load_golden(name)
    images, k, depth, n_iter = golden_inputs(name)
    ds = vt.ArrayDataset(images)
    row_of = {}
    for i, im in enumerate(images):
        row_of.setdefault(vt.utils.get_image_id(np.ascontiguousarray(im)), i)

    class Plugin(object):
        calls = 0

        def embedding(self, cv_image):
            Plugin.calls += 1
            return g["emb"][row_of[vt.utils.get_image_id(np.ascontiguousarray(cv_image))]]

    db = vt.Database(ds, Plugin())
    db.index()
    n_img = len(images)
    assert Plugin.calls == n_img and all(db.is_indexed(p) for p in ds.image_paths)
    assert np.array_equal(db.embedding(ds.image_paths[0]), g["emb"][0], equal_nan=True) and Plugin.calls == n_img
    for qi in range(0, n_img, max(1, n_img // 12)):
        res = db.retrieve(ds.image_paths[qi])
        ws = np.empty(n_img)
        ws[g["order"][qi]] = g["scores"][qi]
        np.testing.assert_allclose(list(res.values()), g["scores"][qi], rtol=1e-9, atol=1e-12)
        for a, b in zip([ds.image_paths.index(p) for p in res], g["order"][qi]):
            assert a == b or abs(ws[a] - ws[b]) <= 1e-9
        assert db.score(ds.image_paths[min(1, n_img - 1)], ds.image_paths[qi]) == pytest.approx(ws[min(1, n_img - 1)], abs=1e-12)
    # batched top-k (matrix product + exact re-score of the best candidates when the set is large)
    ids, sc = db.retrieve_batch(ds.image_paths, k=5)
    for qi in range(n_img):
        ws = np.empty(n_img)
        ws[g["order"][qi]] = g["scores"][qi]
        np.testing.assert_allclose(sc[qi], g["scores"][qi][:5], rtol=1e-9, atol=1e-12)
        for a, b in zip(ids[qi], g["order"][qi][:5]):
            assert a == b or abs(ws[a] - ws[b]) <= 1e-9
    with pytest.raises(TypeError):
        vt.Database(42, Plugin())


def test_database_of_images_without_descriptors(vt, cuda):
    """Every image empty: the reference's embeddings are NaN, every score is the 1e6 sentinel (database.py:70) and
    the stable sort leaves the images in dataset order -- the inverted file is empty but must still work."""
    ft, g, meta, images = _flat_from_golden(vt, "adversarial")
    d = images[0].shape[1]
    empties = [np.zeros((0, d), np.uint8) for _ in range(5)]
    ds = vt.ArrayDataset(empties)
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    db = vt.Database(ds, voc)
    db.index()
    res = db.retrieve(ds.image_paths[2])
    assert list(res.keys()) == ds.image_paths and all(v == 1e6 for v in res.values())
    ids, sc = db.retrieve_batch(ds.image_paths[:2], k=3)
    assert np.array_equal(ids, [[0, 1, 2], [0, 1, 2]]) and np.all(sc == 1e6)


# ------------------------------------------------------------------------------------------ full size
def test_full_size_quantize_properties(vt, cuda):
    """BASELINE configs[2] shape at a size the oracle cannot check row by row (20 M x 128 through a
    1,111,111-node tree): size-independent properties + an oracle spot check.
    * all three engines (fused / sorted SIMT / tcgen05) agree on every row;
    * the result does not depend on how the batch is chunked or ordered (permutation equivariance);
    * a checksum of checksums over chunks equals the whole-batch checksum;
    * every word is a deepest-level node; a 50 k sample equals the binary64 oracle."""
    torch = _torch()
    k, depth, d, n = 10, 6, 128, 20_000_000
    cent = vt.synth.hierarchical_tree(k, depth, d, seed=7, device=cuda)
    nh = vt.tree.heap_size(k, depth)
    exists = np.ones(nh, np.uint8)
    internal = np.zeros(nh, np.uint8)
    internal[:vt.tree.level_offset(k, depth)] = 1
    ref_id, n_ref = vt.tree.reference_ids(exists, internal, k, depth)
    tree_dev = dict(centroids=cent, internal=torch.from_numpy(internal).to(cuda), ref_id=torch.from_numpy(ref_id).to(cuda))
    desc = vt.synth.descriptors_from_tree(cent, k, depth, n, seed=3)
    w_mma = vt.engine.quantize(tree_dev, k, depth, desc, method="mma")
    w_sorted = vt.engine.quantize(tree_dev, k, depth, desc, method="sorted")
    assert torch.equal(w_mma, w_sorted)
    w_fused = vt.engine.quantize(tree_dev, k, depth, desc[:4_000_000], method="fused")
    assert torch.equal(w_fused, w_mma[:4_000_000])
    # chunk invariance + checksum of checksums
    total = int(w_mma.to(torch.int64).sum().item())
    parts = 0
    for lo in range(0, n, 7_000_003):
        hi = min(n, lo + 7_000_003)
        wc = vt.engine.quantize(tree_dev, k, depth, desc[lo:hi], method="mma")
        assert torch.equal(wc, w_mma[lo:hi])
        parts += int(wc.to(torch.int64).sum().item())
    assert parts == total
    # permutation equivariance
    perm = torch.randperm(2_000_000, device=cuda)
    wp = vt.engine.quantize(tree_dev, k, depth, desc[:2_000_000][perm].contiguous(), method="mma")
    assert torch.equal(wp, w_mma[:2_000_000][perm])
    # words are leaves of the full tree
    leaf_ids = torch.from_numpy(ref_id[vt.tree.level_offset(k, depth):]).to(cuda)
    assert bool(torch.isin(w_mma[:1_000_000], leaf_ids).all().item())
    # oracle spot check
    sample = torch.randint(0, n, (50_000,), device=cuda)
    ot = dict(k=k, depth=depth, D=d, DP=d, centroids=cent.cpu().numpy(), exists=exists, internal=internal,
              ref_id=ref_id, n_ref=n_ref)
    want = ref_id[vo.descend(ot, desc[sample].cpu().numpy())]
    assert np.array_equal(w_mma[sample].cpu().numpy(), want)


def test_full_size_retrieval_properties(vt, cuda):
    """200 k-image inverted file (25 shards): every database image queried against the index ranks
    itself first with score exactly 0, scores ascend, and top-k is a prefix of top-(k+5)."""
    torch = _torch()
    rng = np.random.default_rng(4)
    n_img, n_ref, wpi = 200_000, 1_111_111, 300
    words = torch.from_numpy(rng.integers(1, n_ref, n_img * wpi).astype(np.int32)).to(cuda)
    off = torch.arange(n_img + 1, dtype=torch.int64, device=cuda) * wpi
    dummy = dict(parent_ref=torch.zeros(1, dtype=torch.int32, device=cuda), level_ref=torch.zeros(1, dtype=torch.uint8, device=cuda))
    csr = vt.engine.encode(words, off, dummy, 6, n_ref, all_nodes=False)
    index = vt.engine.build_index(csr, n_ref)
    assert index["n_shards"] == 25
    q_rows = torch.from_numpy(rng.integers(0, n_img, 512)).to(cuda)
    ro = csr["row_off"]
    sizes = (ro[q_rows + 1] - ro[q_rows])
    qoff = torch.zeros(len(q_rows) + 1, dtype=torch.int64, device=cuda)
    qoff[1:] = torch.cumsum(sizes, 0)
    idx = torch.cat([torch.arange(int(ro[r]), int(ro[r + 1]), device=cuda) for r in q_rows.tolist()])
    q = dict(row_off=qoff, node=csr["node"][idx].contiguous(), count=csr["count"][idx].contiguous(),
             sumsq=csr["sumsq"][q_rows].contiguous(), n_img=len(q_rows), nnz=int(qoff[-1]))
    o10 = vt.engine.score(index, q, 10)
    o15 = vt.engine.score(index, q, 15)
    assert torch.equal(o10["id"][:, 0].long(), q_rows) and bool((o10["score"][:, 0] == 0).all())
    assert bool((o10["score"][:, 1:] >= o10["score"][:, :-1]).all())
    assert torch.equal(o10["id"], o15["id"][:, :10])


# ------------------------------------------------------------------------------------------ persistence
def test_save_load_round_trip_and_reference_import(vt, cuda, tmp_path):
    """VocabularyTree.save/load (vocabulary.py:148-159 has no load; ours adds it), Database.save/load
    (database.py:83-101), and import of a tree the REFERENCE built (its nodes/tree dicts)."""
    import pickle
    g, meta = load_golden("small")
    images, k, depth, n_iter = golden_inputs("small")
    ds = vt.ArrayDataset(images)
    voc = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, n_iter=n_iter)
    voc.fit(np.concatenate(images))
    db = vt.Database(ds, voc)
    db.index()
    want_ids, want_sc = db.retrieve_batch(ds.image_paths[:10], k=5)
    assert voc.save(str(tmp_path)) and db.save(str(tmp_path))
    # reference-format nodes.pickle: dict id -> centre
    nodes = pickle.load(open(tmp_path / "nodes.pickle", "rb"))
    assert all(np.array_equal(nodes[i], g["ref_nodes"][i]) for i in range(1, len(nodes)))
    voc2 = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im)
    assert voc2.load(str(tmp_path))
    db2 = vt.Database(ds, voc2)
    assert db2.load(str(tmp_path)) and db2.is_indexed(ds.image_paths[3])
    ids, sc = db2.retrieve_batch(ds.image_paths[:10], k=5)
    assert np.array_equal(ids, want_ids) and np.array_equal(sc, want_sc)
    # a tree handed over by the reference (dicts as in voc.nodes / voc.tree)
    ref_nodes, ref_tree = golden_tree_dicts(g, k)
    voc3 = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im)
    voc3.set_tree(vt.FlatTree.from_reference_dicts(ref_nodes, ref_tree, k, depth))
    x = np.concatenate(images)
    assert np.array_equal(voc3.quantize(x), voc.quantize(x))
    # loading a missing index prints and carries on, like the reference's bare except (database.py:99-100)
    assert vt.Database(ds, voc2).load(str(tmp_path / "nowhere"))


def test_reference_written_files_drive_the_gpu_engine(vt, cuda, tmp_path):
    """the files the reference's own save() wrote (tests/golden/persist, generated by
    oracle/make_golden.py persist) are imported -- tree from nodes.pickle + graph.pickle, index from
    index.pickle -- and the GPU engine then reproduces the reference's descent and retrieval without
    ever seeing the descriptors of the database images; export in the reference's format round-trips."""
    import json
    import os
    import pickle
    from helpers import GOLDEN
    pdir = os.path.join(GOLDEN, "persist")
    g, meta = load_golden("adversarial")
    images, k, depth, n_iter = golden_inputs("adversarial")
    ds = vt.ArrayDataset([im.astype(np.float32) for im in images])
    keys = json.load(open(os.path.join(pdir, "keys.json")))["keys"]
    voc = vt.encoders.VocabularyTree(1, 1, descriptor=lambda im: im)
    assert voc.load(pdir)
    x = np.concatenate(images)
    mism = np.nonzero(voc.quantize(x) != g["ref_leaf"])[0]      # float32 near-ties only
    assert len(mism) <= 2 and np.all(near_tie_rows(oracle_tree_from_flat(voc.flat), x[mism].astype(np.float32)))
    db = vt.Database(ds, voc)
    assert db.load(pdir, key=lambda p: keys[p])
    assert all(db.is_indexed(p) for p in ds.image_paths)
    n_img = len(images)
    for qi in range(n_img):
        res = db.retrieve(ds.image_paths[qi])
        ws = np.empty(n_img)
        ws[g["order"][qi]] = g["scores"][qi]
        np.testing.assert_allclose(list(res.values()), g["scores"][qi], rtol=1e-5, atol=1e-7)
        for a, b in zip([ds.image_paths.index(p) for p in res], g["order"][qi]):
            assert a == b or abs(ws[a] - ws[b]) <= 1e-9
    # export in the reference's formats and compare with what the reference wrote
    assert voc.save(str(tmp_path)) and db.save(str(tmp_path), reference_format=True, key=lambda p: keys[p])
    ref_index = pickle.load(open(os.path.join(pdir, "index.pickle"), "rb"))
    our_index = pickle.load(open(tmp_path / "index.pickle", "rb"))
    assert set(ref_index) == set(our_index)
    for key_, e in ref_index.items():
        np.testing.assert_allclose(our_index[key_], e, rtol=1e-12, atol=0, equal_nan=True)
    ref_graph = pickle.load(open(os.path.join(pdir, "graph.pickle"), "rb"))
    our_graph = pickle.load(open(tmp_path / "graph.pickle", "rb"))
    assert list(ref_graph.edges) == list(our_graph.edges) or sorted(ref_graph.edges) == sorted(our_graph.edges)
    assert {n_: dict(a) for n_, a in ref_graph.nodes(data=True)} == {n_: dict(a) for n_, a in our_graph.nodes(data=True)}
    # a wrong key function fails like the reference's load: message, database unchanged
    db_bad = vt.Database(ds, voc)
    assert db_bad.load(pdir, key=lambda p: 12345) and not db_bad.is_indexed(ds.image_paths[0])


def test_c1_full_pipeline_matches_reference(vt, cuda):
    """BASELINE configs[0] at full size through the public API: 1,491 images x 1,000 SIFT-128, k=10, L=3.
    fit -> the reference's tree (bit exact); index -> the reference's postings / embeddings; retrieve ->
    the reference's top-32 order for the 64 stored queries (float32 near-tie rows aside)."""
    import hashlib
    g, meta = load_golden("c1")
    images, k, depth, n_iter = golden_inputs("c1")
    ds = vt.ArrayDataset(images)
    voc = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, n_iter=n_iter)
    x = np.concatenate(images)
    voc.fit(x)
    nodes, tree = golden_tree_dicts(g, k)
    assert voc.tree == tree
    mine = voc.nodes
    assert all(np.array_equal(mine[i], nodes[i]) for i in range(1, len(nodes)))
    words = voc.quantize(x)
    ref_leaf = g["ref_leaf"].astype(np.int32)
    diff = np.nonzero(words != ref_leaf)[0]
    assert len(diff) <= 30                                      # the float32 near-ties of the literal reference
    db = vt.Database(ds, voc)
    db.index()
    touched = set((np.searchsorted(np.cumsum([len(i) for i in images]), diff, side="right")).tolist())
    ro, node, count = db._host_csr()
    dense = csr_to_dense(ro[:17], node, count, voc.flat.n_ref)
    for i in range(16):
        if i not in touched:
            assert np.array_equal(dense[i], g["counts_head"][i])
            assert np.array_equal(db.embedding(ds.image_paths[i]), g["emb_head"][i])
    if not touched:
        full = csr_to_dense(ro, node, count, voc.flat.n_ref).astype(np.int32)
        assert hashlib.sha256(full.tobytes()).hexdigest() == str(g["counts_sha256"])
    q_idx = g["q_idx"]
    ids, sc = db.retrieve_batch([ds.image_paths[i] for i in q_idx], k=32)
    ref_order, ref_sc = g["order"].astype(np.int64), g["scores"]
    for qi in range(len(q_idx)):
        if touched and (int(q_idx[qi]) in touched or touched & set(ref_order[qi].tolist())):
            continue                                            # a near-tie descriptor moved one count: skip that row
        np.testing.assert_allclose(sc[qi], ref_sc[qi], rtol=1e-5, atol=1e-7)
        bad = np.nonzero(ids[qi] != ref_order[qi])[0]
        for b in bad:                                           # only rounding-level score ties may swap
            assert abs(ref_sc[qi][b] - sc[qi][b]) <= 1e-9 and ids[qi][b] in ref_order[qi]
    assert ids[0][0] == q_idx[0] and sc[0][0] == 0.0


# ------------------------------------------------------------------------------------------ literal notebook
def test_notebook_cells_78_to_91_with_real_images(vt, cuda, tmp_path):
    """The literal notebook flow (NB[78-91]) on image FILES: Dataset(folder) -> Orb descriptor ->
    VocabularyTree(4, 4, orb).extract_features / fit -> Database(dataset, voc).index / save / retrieve.
    Images are synthetic textures; every image must retrieve itself first with score exactly 0, a
    near-duplicate (same texture, small shift) must come second, and reloading the saved tree/index
    gives the same ranking."""
    import cv2
    rng = np.random.default_rng(78)
    base = []
    for i in range(6):
        small = rng.integers(0, 256, (24, 32, 3), dtype=np.uint8)
        img = cv2.resize(small, (320, 240), interpolation=cv2.INTER_NEAREST)
        img = cv2.GaussianBlur(img, (3, 3), 0)
        base.append(img)
        assert cv2.imwrite(str(tmp_path / ("%d00.png" % (i + 1))), img)
        assert cv2.imwrite(str(tmp_path / ("%d01.png" % (i + 1))), np.roll(img, 3, axis=1))     # near-duplicate
    dataset = vt.Dataset(str(tmp_path))
    orb = vt.descriptors.Orb()
    voc = vt.encoders.VocabularyTree(n_branches=4, depth=4, descriptor=orb)
    features = voc.extract_features(dataset)               # every descriptor of every image, stacked (utils.py:8-27)
    assert features.ndim == 2 and features.shape[1] == 32 and features.shape[0] > 100 * len(dataset)
    voc.fit(features)
    db = vt.Database(dataset, voc)
    db.index()
    order = {}
    for name in dataset.image_paths:
        res = db.retrieve(name)
        ranked = list(res.items())
        assert ranked[0][0] == name and ranked[0][1] == 0.0
        twin = name[0] + ("01.png" if name.endswith("00.png") else "00.png")
        assert ranked[1][0] == twin, (name, ranked[:3])
        order[name] = [p for p, _ in ranked]
    store = tmp_path / "store"
    store.mkdir()
    assert voc.save(str(store)) and db.save(str(store))
    voc2 = vt.encoders.VocabularyTree(n_branches=4, depth=4, descriptor=orb)
    assert voc2.load(str(store))
    db2 = vt.Database(dataset, voc2)
    assert db2.load(str(store))
    for name in dataset.image_paths[:4]:
        assert [p for p in db2.retrieve(name)] == order[name]


def test_fit_recursion_arguments(vt, cuda):
    """vocabulary.py:36: fit(features, node, root, current_depth).  root replaces the stored value of node 0, current_depth
    = c builds the depth - c levels the recursion would still add; node != 0 is refused with an explanation."""
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, (6000, 128), dtype=np.uint8)
    full = vt.encoders.VocabularyTree(3, 2, descriptor=lambda im: im).fit(x)
    root = np.full(128, 7.5, np.float32)
    part = vt.encoders.VocabularyTree(3, 3, descriptor=lambda im: im).fit(x, node=0, root=root, current_depth=1)
    assert part.flat.depth == 2 and part.flat.n_ref == full.flat.n_ref
    a, b = full.flat.nodes_dict(), part.flat.nodes_dict()
    assert np.array_equal(b[0], root) and not np.array_equal(a[0], root)
    assert all(np.array_equal(a[i], b[i]) for i in range(1, full.flat.n_ref))
    img = x[:500]
    assert np.array_equal(full.embedding(img), part.embedding(img))
    with pytest.raises(NotImplementedError):
        vt.encoders.VocabularyTree(3, 2, descriptor=lambda im: im).fit(x, node=4)
    with pytest.raises(NotImplementedError):
        vt.encoders.VocabularyTree(3, 2, descriptor=lambda im: im).fit(x, current_depth=2)


def test_pipelined_retrieve_batch_equals_one_shot(vt, cuda):
    """Large query batches are staged chunk by chunk on a loader thread while the device works on the previous chunk
    (Database._retrieve_batch_pipelined): same ids and scores as the one-shot path, for indexed images, images that are
    not in the database, repeated queries and an image without descriptors."""
    ft, g, meta, images = _flat_from_golden(vt, "orb")
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    rng = np.random.default_rng(2)
    pool = np.concatenate(images)
    extra = [pool[rng.integers(0, len(pool), int(n))] for n in rng.integers(1, 300, 25)] + [pool[:0]]
    ds = vt.ArrayDataset(list(images) + extra)
    n_db = len(images)
    db = vt.Database(vt.ArrayDataset(list(images)), voc)
    db.index()
    db.dataset = ds                                              # queries may name images beyond the indexed ones
    order = rng.permutation(len(ds.image_paths)).tolist() + [3, 3, n_db + 1, n_db + 1]
    paths = [ds.image_paths[i] for i in order]
    one_ids, one_sc = db.retrieve_batch(paths, k=5)
    db._extra_csr.clear()
    db.query_chunk = 7
    ids, sc = db.retrieve_batch(paths, k=5)
    assert np.array_equal(ids, one_ids) and np.array_equal(sc, one_sc)
    assert all(sc[j][0] == 0.0 for j, i in enumerate(order) if i < n_db and len(images[i]))


def test_record_postings_reproduces_the_reference_side_effect(vt, cuda):
    """vocabulary.py:126-127: the reference's embedding() propagates every image it sees -- database images and queries --
    into the tree's per-node {image sha1: count} postings (vocabulary.py:96-100).  Opt-in here (Database(record_postings=
    True)); the postings then equal the golden visit counts image by image, a query that is not in the database appears
    after it was retrieved, and without the flag nothing is recorded.  Retrieval results are the same either way."""
    ft, g, meta, images = _flat_from_golden(vt, "small")
    counts = g["counts"].astype(np.int64)
    extra = np.concatenate([images[0][:7], images[1][:5]])
    ds = vt.ArrayDataset(list(images) + [extra])
    plain_voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    plain = vt.Database(vt.ArrayDataset(list(images)), plain_voc)
    plain.index()
    plain.dataset = ds
    want_ids, want_sc = plain.retrieve_batch(ds.image_paths, k=4)
    assert plain_voc.postings() == {}
    voc = vt.encoders.VocabularyTree(ft.k, ft.depth, descriptor=lambda im: im).set_tree(ft)
    db = vt.Database(vt.ArrayDataset(list(images)), voc, record_postings=True)
    db.index()
    post = voc.postings()
    for i, im in enumerate(images):
        key = vt.utils.get_image_id(im)
        got = np.zeros(ft.n_ref, np.int64)
        for node, by_image in post.items():
            if key in by_image:
                got[node] = by_image[key]
        assert np.array_equal(got, counts[i]), i
    q_key = vt.utils.get_image_id(extra)
    assert not any(q_key in by_image for by_image in post.values())
    db.dataset = ds
    ids, sc = db.retrieve_batch(ds.image_paths, k=4)
    assert np.array_equal(ids, want_ids) and np.array_equal(sc, want_sc)
    ot = oracle_tree_from_flat(ft)
    got = np.zeros(ft.n_ref, np.int64)
    for node, by_image in voc.postings().items():
        if q_key in by_image:
            got[node] = by_image[q_key]
    assert np.array_equal(got, vo.path_counts(ot, extra, all_nodes=True))
