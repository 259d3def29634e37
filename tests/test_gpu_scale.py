"""GPU: full-size certificates that do not need the CPU oracle, multi-GPU runs over NCCL, persistence corners.

* build certificate at BASELINE configs[1] (10 M x 128, k=10, L=6): independent per-level checks with plain torch
  integer arithmetic -- every level's member counts equal what the canonical descent of the finished tree gives
  (final labels are the canonical argmin), and on every converged level each centre is (float)(sum / count) of
  its members (exact integer sums);
* NCCL: 2 ranks (when >= 2 GPUs are visible) drive the public classes -- sharded build, sharded tf and tf-idf
  index, merged top-k -- and must reproduce the single-GPU results bit for bit;
* VT_SLOW=1: the C oracle builds a 2 M-row, L=6 tree on the host and the GPU tree must equal it."""
import os
import socket

import numpy as np
import pytest

from oracle import vt_oracle as vo

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _level_of_heap(h, k, depth):
    """level of heap indices (int64 tensor)"""
    torch = _torch()
    lvl = torch.zeros_like(h)
    off = 0
    for level in range(1, depth + 1):
        off += k ** (level - 1)
        lvl += (h >= off).to(h.dtype)
    return lvl


def _certificate(vt, cuda, n, k, depth, n_iter, seed):
    torch = _torch()
    cent0 = vt.synth.hierarchical_tree(k, depth, 128, seed=7, device=cuda)
    x = vt.synth.descriptors_from_tree(cent0, k, depth, n, seed=seed)
    del cent0
    log = []
    ft = vt.engine.build_tree(x, 128, k, depth, n_iter=n_iter, log=lambda **kw: log.append(kw))
    td = ft.device(cuda)
    _, heap = vt.engine.quantize(td, k, depth, x, want_heap=True)          # canonical descent of the finished tree
    heap = heap.long()
    count = torch.from_numpy(ft.count).to(cuda)
    cent = td["centroids"]
    stop_level = _level_of_heap(heap, k, depth)
    checked_levels, converged_levels = 0, 0
    for level in range(depth, 0, -1):
        # ancestor at `level` of every row that reaches it
        a = heap.clone()
        steps = stop_level - level
        for _ in range(depth):
            up = steps > 0
            a = torch.where(up, (a - 1) // k, a)
            steps = steps - up.to(steps.dtype)
        reach = stop_level >= level
        o = vt.tree.level_offset(k, level)
        n_l = k ** level
        idx = (a[reach] - o)
        got = torch.bincount(idx, minlength=n_l)
        # (1) member counts of the build == counts under the canonical descent of the final centres
        assert torch.equal(got, count[o:o + n_l]), "level %d: partition differs from the canonical argmin" % level
        checked_levels += 1
        conv = [l for l in log if l["level"] == level - 1]
        if conv and conv[0]["converged"]:
            # (2) converged level: every centre is (float)((double)sum / (double)count) of its members
            sums = torch.zeros((n_l, 128), dtype=torch.int64, device=cuda)
            rows = torch.nonzero(reach).flatten()
            for c0 in range(0, rows.numel(), 1 << 20):
                r = rows[c0:c0 + (1 << 20)]
                sums.index_add_(0, a[r] - o, x[r].to(torch.int64))
            has = got > 0
            want = (sums[has].to(torch.float64) / got[has].to(torch.float64)[:, None]).to(torch.float32)
            assert torch.equal(cent[o:o + n_l][has], want), "level %d: a centre is not the mean of its members" % level
            converged_levels += 1
    # (3) the partition is stable: internal nodes are exactly those with >= k members (vocabulary.py:55)
    inner = vt.tree.level_offset(k, depth)
    assert torch.equal(torch.from_numpy(ft.internal[:inner].astype(bool)).to(cuda),
                       (count[:inner] >= k) & torch.from_numpy(ft.exists[:inner].astype(bool)).to(cuda))
    return checked_levels, converged_levels, log


def test_build_certificate_at_c2_size(vt, cuda):
    """BASELINE configs[1]: 10 M descriptors, k=10, L=6, 10 Lloyd iterations."""
    checked, converged, log = _certificate(vt, cuda, 10_000_000, 10, 6, 10, seed=4242)
    assert checked == 6
    assert converged >= 1, log           # the deep levels (a handful of members per cluster) reach their fixed point


def test_build_certificate_converged_run(vt, cuda):
    """Enough iterations for every level to reach its fixed point: the centroid check then covers the whole tree."""
    checked, converged, log = _certificate(vt, cuda, 400_000, 10, 4, 200, seed=99)
    assert checked == 4 and converged == 4, log


@pytest.mark.slow
def test_build_equals_oracle_at_2m_rows(vt, cuda):
    """VT_SLOW=1: the C oracle (seeded exact Lloyd, binary64) builds k=10, L=6 on 2 M rows on the host cores."""
    torch = _torch()
    cent0 = vt.synth.hierarchical_tree(10, 6, 128, seed=7, device=cuda)
    x = vt.synth.descriptors_from_tree(cent0, 10, 6, 2_000_000, seed=5)
    ft = vt.engine.build_tree(x, 128, 10, 6, n_iter=10)
    t = vo.build_tree(x.cpu().numpy(), 10, 6, 10)
    assert np.array_equal(ft.exists, t["exists"]) and np.array_equal(ft.internal, t["internal"])
    assert np.array_equal(ft.count, t["count"]) and np.array_equal(ft.centroids, t["centroids"])


# ------------------------------------------------------------------------------------------ NCCL, 2 ranks
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import vtree_b200 as vt
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        comm = vt.dist.Comm()
        rng = np.random.default_rng(3)
        images = [np.clip(np.rint(rng.normal(90, 40, (int(rng.integers(40, 90)), 128))), 0, 255).astype(np.uint8)
                  for _ in range(300)]
        images[7] = images[5].copy()                                  # duplicate image: exact score ties
        images[11] = np.zeros((0, 128), np.uint8)                     # image without descriptors
        feats = np.concatenate(images)
        lo, hi = vt.dist.shard_range(len(feats), rank, world)
        res = {}
        # sharded build through the public class == single-GPU build
        voc_d = vt.encoders.VocabularyTree(6, 3, descriptor=lambda im: im, device=dev, comm=comm).fit(feats[lo:hi])
        voc_1 = vt.encoders.VocabularyTree(6, 3, descriptor=lambda im: im, device=dev).fit(feats)
        res["tree"] = bool(np.array_equal(voc_d.flat.centroids, voc_1.flat.centroids)
                           and np.array_equal(voc_d.flat.internal, voc_1.flat.internal)
                           and np.array_equal(voc_d.flat.count, voc_1.flat.count))
        ds = vt.ArrayDataset(images)
        queries = ds.image_paths[:40]
        for weighting, norm in (("tf", "l2"), ("tfidf", "l2"), ("tfidf", "l1")):
            db_d = vt.Database(ds, voc_d, weighting=weighting, norm=norm, comm=comm)
            db_d.index(batch_images=64)
            db_1 = vt.Database(ds, voc_1, weighting=weighting, norm=norm)
            db_1.index()
            ids_d, sc_d = db_d.retrieve_batch(queries, k=10)
            ids_1, sc_1 = db_1.retrieve_batch(queries, k=10)
            same_ids = bool(np.array_equal(ids_d, ids_1))
            tol = 0.0 if weighting == "tf" else 1e-12
            same_sc = bool(np.allclose(sc_d, sc_1, rtol=tol, atol=tol))
            full_d = db_d.retrieve(queries[3])
            full_1 = db_1.retrieve(queries[3])
            same_full = list(full_d) == list(full_1)
            if weighting == "tfidf":
                same_full = same_full and bool(torch.equal(db_d._weights["idf"], db_1._weights["idf"]))
            res["%s/%s" % (weighting, norm)] = same_ids and same_sc and same_full
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_nccl_equal_single_gpu(vt, cuda):
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        assert out.get(r) and all(out[r].values()), dict(out)


# ------------------------------------------------------------------------------------------ persistence corner
def test_save_load_with_trailing_images_without_descriptors(vt, cuda, tmp_path):
    """np.add.reduceat raised when the LAST images were empty while earlier ones were not."""
    rng = np.random.default_rng(1)
    images = [rng.integers(0, 256, (30, 32), dtype=np.uint8) for _ in range(6)]
    images += [np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8)]
    voc = vt.encoders.VocabularyTree(3, 2, descriptor=lambda im: im).fit(np.concatenate(images))
    ds = vt.ArrayDataset(images)
    db = vt.Database(ds, voc)
    db.index()
    ids0, sc0 = db.retrieve_batch(ds.image_paths, k=4)
    db.save(str(tmp_path))
    db2 = vt.Database(ds, voc)
    db2.load(str(tmp_path))
    ids1, sc1 = db2.retrieve_batch(ds.image_paths, k=4)
    assert np.array_equal(ids0, ids1) and np.array_equal(sc0, sc1)
    assert (sc1[-1] == 1e6).all()                                    # NaN -> 1e6 (database.py:70)
    db.save(str(tmp_path), reference_format=True, key=lambda p: p)
    db3 = vt.Database(ds, voc)
    db3.load(str(tmp_path), key=lambda p: p)
    assert np.array_equal(db3.retrieve_batch(ds.image_paths, k=4)[0], ids0)


def test_native_rank_merge_equals_test_double(vt, cuda):
    """dist.merge_ranked (vt_topk_merge_scored) against the torch-sort restatement used in the gloo tests."""
    torch = _torch()
    from oracle_ops import merge_ranked_cpu
    rng = np.random.default_rng(5)
    w, q, k = 8, 50, 10
    keys = np.round(rng.random((w, q, k)), 2)                        # coarse -> exact ties across ranks
    ids = rng.permutation(w * q * k).reshape(w, q, k).astype(np.int32) % 100000
    ids = np.stack([np.stack([rng.permutation(5000)[:k] + 5000 * r for _ in range(q)]) for r in range(w)]).astype(np.int32)
    ids[3, :, 6:] = -1                                               # short lists
    sc = np.sqrt(2 - 2 * keys)
    a = merge_ranked_cpu(torch.from_numpy(keys), torch.from_numpy(ids), torch.from_numpy(sc), k)
    b = vt.dist.merge_ranked(torch.from_numpy(keys).to(cuda), torch.from_numpy(ids).to(cuda),
                             torch.from_numpy(sc).to(cuda), k)
    assert torch.equal(a[0], b[0].cpu()) and torch.equal(a[1], b[1].cpu())


def test_tensors_on_a_non_current_device_are_served(vt, cuda):
    """kernels launched on the current device against another device's pointers.  Needs 2 GPUs."""
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, (5000, 128), dtype=np.uint8)
    torch.cuda.set_device(0)
    v0 = vt.encoders.VocabularyTree(4, 3, descriptor=lambda im: im, device="cuda:0").fit(x)
    v1 = vt.encoders.VocabularyTree(4, 3, descriptor=lambda im: im, device="cuda:1").fit(x)
    assert torch.cuda.current_device() == 0
    assert np.array_equal(v0.flat.centroids, v1.flat.centroids)
    assert np.array_equal(v0.quantize(x), v1.quantize(x))


# ------------------------------------------------------------------------------------------ dense top of all-nodes indexes
def _full_tree(vt, cuda, k, depth, d, seed):
    torch = _torch()
    cent = vt.synth.hierarchical_tree(k, depth, d, seed=seed, device=cuda)
    nh = vt.tree.heap_size(k, depth)
    internal = torch.zeros(nh, dtype=torch.uint8, device=cuda)
    internal[:vt.tree.level_offset(k, depth)] = 1
    pad = torch.zeros((nh, vt.tree.pad_dim(d)), dtype=torch.float32, device=cuda)
    pad[:, :d] = cent
    return vt.FlatTree.from_device(k, depth, d, pad, torch.ones(nh, dtype=torch.uint8, device=cuda), internal, byte_built=True), cent


@pytest.mark.parametrize("k,depth", [(4, 6), (10, 3), (3, 2), (16, 3)])
def test_dense_top_equals_plain_postings(vt, cuda, k, depth):
    """All-nodes tf retrieval with the dense top (levels 0..1 int32, levels 2..4 uint8 block scored by the tcgen05 GEMM,
    deeper levels postings) must give the ids, scores and rank keys of the plain all-postings index, bit for bit --
    including images without descriptors, database sizes that are no multiple of the tile, several shards, and a
    query with more than 255 descriptors under one dense node."""
    torch = _torch()
    rng = np.random.default_rng(k * 10 + depth)
    ft, cent = _full_tree(vt, cuda, k, depth, 32, seed=3)
    voc = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, all_nodes=True, device=cuda).set_tree(ft)
    n_img = 17000 if k == 4 else 700                                 # 17000 images: three shards of 8192
    sizes = rng.integers(5, 60, n_img)
    sizes[13] = 0
    sizes[n_img - 1] = 0
    x = vt.synth.descriptors_from_tree(cent, k, depth, int(sizes.sum()), seed=8).cpu().numpy()[:, :32]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    images = [x[offs[i]:offs[i + 1]] for i in range(n_img)]
    csr = voc.encode_batch(images)
    tdev = ft.device(cuda)
    plain = vt.engine.build_index(csr, ft.n_ref)
    dense = vt.engine.DenseTopBuilder(tdev, depth, ft.n_ref, n_img).add(csr, 0).finish()
    assert dense is not None and dense["layout"]["n_top"] == 1 + k
    heavy = np.repeat(x[:3], 300, axis=0)                            # 300 copies of 3 descriptors: counts of 300 on their paths
    q_images = [images[0], images[13], images[4321 % n_img], heavy, x[100:140], images[n_img - 2]]
    q = voc.encode_batch(q_images)
    a = vt.engine.score(plain, q, 10, want_all=True)
    b = vt.engine.score_dense(dense, q, 10, want_all=True)
    assert torch.equal(a["id"], b["id"])
    assert torch.equal(a["score"], b["score"]) and torch.equal(a["key"], b["key"])
    assert torch.equal(a["all_key"], b["all_key"])
    assert int(b["id"][0, 0]) == 0 and float(b["score"][0, 0]) == 0.0
    assert (b["score"][1] == 1e6).all()                              # query without descriptors (database.py:70)
    # small query budget: several GEMM chunks
    c = vt.engine.score_dense(dense, q, 10, want_all=True, d_budget_bytes=1)
    assert torch.equal(a["id"], c["id"]) and torch.equal(a["all_key"], c["all_key"])
    # the top-k carried from shard to shard (vt_score_query.cu), seeded with the GEMM dots or over the plain postings:
    # one CTA per query over all shards, one list per shard, and the default grouping
    for groups in (1, None, plain["n_shards"]):
        for got in (vt.engine.score_dense(dense, q, 10, groups=groups), vt.engine.score(plain, q, 10, groups=groups)):
            assert torch.equal(a["id"], got["id"]) and torch.equal(a["key"], got["key"]) and torch.equal(a["score"], got["score"])


def test_dense_top_falls_back_when_a_database_count_overflows(vt, cuda):
    rng = np.random.default_rng(0)
    ft, cent = _full_tree(vt, cuda, 4, 5, 32, seed=3)
    voc = vt.encoders.VocabularyTree(4, 5, descriptor=lambda im: im, all_nodes=True, device=cuda).set_tree(ft)
    x = vt.synth.descriptors_from_tree(cent, 4, 5, 4000, seed=8).cpu().numpy()[:, :32]
    images = [x[i * 40:(i + 1) * 40] for i in range(99)] + [np.repeat(x[:1], 400, axis=0)]
    ds = vt.ArrayDataset(images)
    db = vt.Database(ds, voc)
    db.index()
    assert db._dense_index is None                                   # a count of 400 does not fit the byte block
    ids, sc = db.retrieve_batch(ds.image_paths, k=3)
    assert (ids[:, 0] == np.arange(100)).all() and (sc[:, 0] == 0).all()
    ok = vt.Database(vt.ArrayDataset(images[:99]), voc)
    ok.index()
    assert ok._dense_index is not None


def test_two_levels_per_launch_at_scale_equals_one_level_per_launch(vt, cuda):
    """16 M descriptors through the k=10, L=6 tree: the two-levels-per-launch kernel (3 DRAM reads of every row) and the
    default one-level kernel (6 reads) must return the same ids -- regular chunks, chunk pipelining over ~100 chunks per CTA,
    node boundaries inside chunks, the deferred output."""
    torch = _torch()
    lib = vt._native.load()
    ft, cent = _full_tree(vt, cuda, 10, 6, 128, seed=7)
    x = vt.synth.descriptors_from_tree(cent, 10, 6, 16_000_000, seed=21)
    td = ft.device(cuda)
    a = vt.engine.quantize(td, 10, 6, x, method="mma")
    prev = lib.vt_quantize_mma_pair_min(1 << 20)
    try:
        b = vt.engine.quantize(td, 10, 6, x, method="mma")
        c = vt.engine.quantize(td, 10, 6, x[:3_000_001], method="mma")      # below the deferred-output threshold, ragged last chunk
    finally:
        lib.vt_quantize_mma_pair_min(prev if prev < (1 << 62) else -1)
    assert torch.equal(a, b) and torch.equal(a[:3_000_001], c)


@pytest.mark.parametrize("all_nodes", [False, True])
def test_images_with_more_descriptors_than_one_cta_sorts(vt, cuda, all_nodes):
    """The reference puts no limit on the descriptors of an image (vocabulary.py:78-102); vt_encode sorts at most 8192
    words per CTA, larger images are encoded in pieces and merged.  20,000- and 8,193-descriptor images between ordinary
    and empty ones: counts against a numpy restatement (every node on the path counts once per descriptor), rows in
    vt_encode's order, and the big image retrieves itself with score 0."""
    torch = _torch()
    k, depth = 4, 5
    ft, cent = _full_tree(vt, cuda, k, depth, 32, seed=5)
    voc = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, all_nodes=all_nodes, device=cuda).set_tree(ft)
    sizes = [30, 20000, 0, 8193, 55, 8192]
    x = vt.synth.descriptors_from_tree(cent, k, depth, sum(sizes), seed=11).cpu().numpy()[:, :32]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    images = [x[offs[i]:offs[i + 1]] for i in range(len(sizes))]
    csr = voc.encode_batch(images, want_df=True)
    tdev = ft.device(cuda)
    words = vt.engine.quantize(tdev, k, depth, torch.from_numpy(x).to(cuda)).cpu().numpy()
    parent, level = tdev["parent_ref"].cpu().numpy(), tdev["level_ref"].cpu().numpy()
    ro, node, cnt = csr["row_off"].cpu().numpy(), csr["node"].cpu().numpy(), csr["count"].cpu().numpy()
    df = np.zeros(ft.n_ref, np.int64)
    for i in range(len(sizes)):
        want = np.zeros(ft.n_ref, np.int64)
        w = words[offs[i]:offs[i + 1]].copy()
        np.add.at(want, w, 1)
        while all_nodes and len(w) and (level[w] > 0).any():
            w = parent[w[level[w] > 0]]
            np.add.at(want, w, 1)
        got = np.zeros(ft.n_ref, np.int64)
        got[node[ro[i]:ro[i + 1]]] = cnt[ro[i]:ro[i + 1]]
        assert np.array_equal(got, want), i
        assert len(np.unique(node[ro[i]:ro[i + 1]])) == ro[i + 1] - ro[i]                  # one entry per node
        assert int(csr["sumsq"][i]) == int((want * want).sum())
        df += want > 0
        small = voc.encode_batch([images[i]])                                              # the row order vt_encode emits
        if sizes[i] <= 8192:
            assert np.array_equal(small["node"].cpu().numpy(), node[ro[i]:ro[i + 1]])
    assert np.array_equal(csr["df"].cpu().numpy(), df)
    ds = vt.ArrayDataset(images)
    db = vt.Database(ds, voc)
    db.index()
    ids, sc = db.retrieve_batch([ds.image_paths[1], ds.image_paths[3]], k=2)
    assert ids[0][0] == 1 and sc[0][0] == 0.0 and ids[1][0] == 3 and sc[1][0] == 0.0
