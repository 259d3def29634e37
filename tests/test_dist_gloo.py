"""CPU, world_size 2 over gloo: the multi-rank protocol of the build (all-reduce of exact integer
sums, ownership of the seeded initial centres) and of retrieval (all-gather + merge of per-shard
top-k).  The per-rank compute is the oracle test double; the product uses the CUDA ops."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vt_oracle as vo


def vt_split_level(k, depth, world):
    import vtree_b200 as vt
    return vt.engine.subtree_split_level(k, depth, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _build_worker(rank, world, port, out):
    import vtree_b200 as vt
    from oracle_ops import OracleLevelOps
    _init(rank, world, port)
    try:
        x = vt.synth.sift_like(3000, 32, n_centres=60, seed=42, scale=40.0, noise=6.0)
        x[100:140] = x[100]                                     # duplicates -> empty clusters somewhere
        lo, hi = vt.dist.shard_range(len(x), rank, world)
        comm = vt.dist.Comm()
        ft = vt.engine.build_tree(torch.from_numpy(x[lo:hi].copy()), 32, 4, 4, n_iter=5, comm=comm,
                                  ops=OracleLevelOps())
        t = vo.build_tree(x, 4, 4, 5)
        ok = (np.array_equal(ft.centroids, t["centroids"]) and np.array_equal(ft.internal, t["internal"])
              and np.array_equal(ft.ref_id, t["ref_id"]) and np.array_equal(ft.count, t["count"]))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_build_is_bit_identical_to_single_process(world):
    """k=4, depth=4: levels 0-1 all-reduce integer sums, level 2 (16 nodes >= 4 per rank) is the subtree
    exchange (all-to-all of member rows), levels 2-3 run without collectives, one publish at the end."""
    assert vt_split_level(4, 4, world) == 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_build_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world)), dict(out)


def test_single_process_build_with_oracle_ops_matches_oracle():
    import vtree_b200 as vt
    from oracle_ops import OracleLevelOps
    x = vt.synth.sift_like(1500, 32, n_centres=30, seed=7, scale=40.0, noise=6.0)
    ft = vt.engine.build_tree(torch.from_numpy(x), 32, 3, 4, n_iter=4, ops=OracleLevelOps())
    t = vo.build_tree(x, 3, 4, 4)
    assert np.array_equal(ft.centroids, t["centroids"]) and np.array_equal(ft.count, t["count"])


def _retrieve_worker(rank, world, port, out):
    import vtree_b200 as vt
    from oracle_ops import merge_ranked_cpu
    _init(rank, world, port)
    try:
        rng = np.random.default_rng(0)
        n_img, q, k = 101, 7, 5
        key = np.round(rng.random((q, n_img)), 2)               # coarse -> many exact ties
        lo, hi = vt.dist.shard_range(n_img, rank, world)

        def score_local():
            kk = key[:, lo:hi]
            order = np.lexsort((np.broadcast_to(np.arange(hi - lo), kk.shape), -kk), axis=1)[:, :k]
            ks = np.take_along_axis(kk, order, 1)
            return dict(key=torch.from_numpy(ks), id=torch.from_numpy(order.astype(np.int32)),
                        score=torch.from_numpy(np.sqrt(2 - 2 * ks)))
        ids, sc, ks = vt.dist.retrieve_sharded(score_local, lo, k, vt.dist.Comm(), merge=merge_ranked_cpu)
        want = np.lexsort((np.broadcast_to(np.arange(n_img), key.shape), -key), axis=1)[:, :k]
        out[rank] = bool(np.array_equal(ids.numpy(), want) and
                         np.allclose(sc.numpy(), np.sqrt(2 - 2 * np.take_along_axis(key, want, 1))))
    finally:
        dist.destroy_process_group()


def test_sharded_retrieval_merge_matches_global_order():
    world = 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_retrieve_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world)), dict(out)


def test_merge_ranked_handles_empty_slots():
    from oracle_ops import merge_ranked_cpu
    keys = torch.tensor([[[0.5, 0.2, -2.0]], [[0.5, 0.4, 0.1]]], dtype=torch.float64)     # [world=2, Q=1, k=3]
    ids = torch.tensor([[[3, 9, -1]], [[1, 7, 8]]], dtype=torch.int32)
    sc = torch.zeros_like(keys)
    got, _, _ = merge_ranked_cpu(keys, ids, sc, 4)
    assert got.tolist() == [[1, 3, 7, 9]]


def test_shard_range_covers_everything():
    import vtree_b200 as vt
    for n in (0, 1, 7, 100):
        for w in (1, 2, 3, 8):
            spans = [vt.dist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
