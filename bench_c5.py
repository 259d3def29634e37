#!/usr/bin/env python
"""bench_c5.py -- BASELINE.json configs[4]: end-to-end build + index + retrieve at scale.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench_c5.py            # 10 M images, k=10 L=7
    python bench_c5.py --images 200000 --sample 2000000 --depth 5                           # single-GPU smoke

Every rank owns a contiguous slice of the image ids:
  build     k-means tree on a descriptor sample sharded over the ranks; integer sums all-reduced (NCCL)
  index     images streamed in batches (descriptors are generated on the device per batch and never all
            resident: 10 M x 1,000 x 128 B = 1.28 TB), quantised, encoded (leaf words), inverted per rank
  retrieve  queries = database images (regenerated from their batch seeds on every rank), per-rank top-k,
            all-gather + merge; every query must find itself first with score 0
Rank 0 prints one JSON line with the stage timings.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=10_000_000)
    ap.add_argument("--words", type=int, default=1000, help="descriptors per image")
    ap.add_argument("--depth", type=int, default=7)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--sample", type=int, default=100_000_000, help="descriptors used to build the tree (whole job)")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch-images", type=int, default=8192)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--topk", type=int, default=10)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import vtree_b200 as vt
    from vtree_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator is created; stdout carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    comm = vt.dist.Comm() if world > 1 else None
    k, depth, d, wpi = args.k, args.depth, 128, args.words

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # synthetic generator: a random k-ary centroid hierarchy; descriptors = deepest-level centroid + noise
    gen_cent = vt.synth.hierarchical_tree(k, depth, d, seed=7, device=dev)

    def batch_descriptors(batch_id, n_images):
        return vt.synth.descriptors_from_tree(gen_cent, k, depth, n_images * wpi, seed=100_000 + batch_id)

    # ---- build --------------------------------------------------------------------------------------
    lo_s, hi_s = vt.dist.shard_range(args.sample, rank, world)
    sample = vt.synth.descriptors_from_tree(gen_cent, k, depth, args.sample, seed=99)[lo_s:hi_s].contiguous() \
        if args.sample <= 20_000_000 else vt.synth.descriptors_from_tree(gen_cent, k, depth, hi_s - lo_s, seed=99 + rank)
    # warm-up like bench.py's build leg: module load, allocator pools, and the NCCL point-to-point connections the first
    # all-to-all of the subtree exchange sets up (seconds on 8 ranks) -- none of it is the build
    warm = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, n_iter=1, all_nodes=False, device=dev, comm=comm)
    warm.fit(sample[:max(1, min(len(sample), 2_000_000))].contiguous())
    del warm
    sync()
    t0 = time.perf_counter()
    voc = vt.encoders.VocabularyTree(k, depth, descriptor=lambda im: im, n_iter=args.iters, all_nodes=False, device=dev,
                                     comm=comm)
    voc.fit(sample)                                        # top levels: all-reduce of integer sums; below: subtree ownership
    sync()
    t_build = tmax(time.perf_counter() - t0)
    del sample
    tree = voc.flat
    tree_dev = tree.device(dev)
    checksum = int(tree_dev["centroids"].view(torch.int32).to(torch.int64).sum().item())

    # ---- index: stream this rank's images -----------------------------------------------------------
    lo, hi = vt.dist.shard_range(args.images, rank, world)
    n_mine = hi - lo
    assert n_mine * wpi < (1 << 31), "per-rank postings must stay below 2^31 (use more GPUs or fewer images)"
    words = torch.empty(n_mine * wpi, dtype=torch.int32, device=dev)
    bi = args.batch_images
    sync()
    t0 = time.perf_counter()
    n_desc = 0
    for b0 in range(lo, hi, bi):                           # batches are aligned to the global image grid of this rank
        nb = min(bi, hi - b0)
        x = batch_descriptors(b0 // bi + rank * 1_000_000, nb)
        w = engine.quantize(tree_dev, k, depth, x)
        words[(b0 - lo) * wpi:(b0 - lo + nb) * wpi] = w
        n_desc += nb * wpi
    sync()
    t_quant = tmax(time.perf_counter() - t0)
    t0 = time.perf_counter()
    off = torch.arange(n_mine + 1, dtype=torch.int64, device=dev) * wpi
    csr = engine.encode(words, off, tree_dev, depth, tree.n_ref, all_nodes=False)
    del words
    import numpy as np
    db = vt.Database(vt.ArrayDataset([]), voc, comm=comm)
    db._adopt(csr, {}, args.images, lo)                    # this rank's shard of the inverted file (Database.index minus the image loop)
    sync()
    t_index = tmax(time.perf_counter() - t0)

    # ---- queries: the first images of every rank's first batch, regenerated everywhere, as HOST descriptor arrays ----
    per_rank = max(1, min(args.queries // world, bi))
    q_desc, q_ids = [], []
    for r in range(world):
        lo_r, hi_r = vt.dist.shard_range(args.images, r, world)
        nb = min(bi, hi_r - lo_r)
        x = batch_descriptors(lo_r // bi + r * 1_000_000, nb)[:per_rank * wpi]
        q_desc.append(x.cpu().numpy().reshape(per_rank, wpi, d))
        q_ids.append(np.arange(lo_r, lo_r + per_rank))
    q_desc, q_ids = np.concatenate(q_desc), np.concatenate(q_ids)
    nq = len(q_ids)
    names = ["q%07d" % i for i in range(nq)]
    db.dataset = vt.ArrayDataset([q_desc[i] for i in range(nq)], names)

    def one_pass():
        db._extra_csr.clear()
        return db.retrieve_batch(names, k=args.topk)         # host descriptors -> quantise -> encode -> score -> all-gather + merge

    ids, sc = one_pass()
    sync()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ids, sc = one_pass()
    sync()
    t_ret = tmax((time.perf_counter() - t0) / reps)
    ok = bool((ids[:, 0] == q_ids).all() and (sc[:, 0] == 0).all())
    # the same queries with their bag-of-words CSR resident in HBM: per-rank scoring + all-gather + merge only
    qcsr, back = db._query_csr(names)
    for _ in range(2):
        db._score_global(qcsr, args.topk)
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r_ids, r_sc, _ = db._score_global(qcsr, args.topk)
    sync()
    t_res = tmax((time.perf_counter() - t0) / reps)
    ok_res = bool(np.array_equal(r_ids.cpu().numpy()[back], ids) and np.array_equal(r_sc.cpu().numpy()[back], sc))

    if rank == 0:
        print(json.dumps({
            "config": "BASELINE configs[4]: build + index + retrieve, %d images x %d descriptors, k=%d L=%d, %d GPU(s)"
                      % (args.images, wpi, k, depth, world),
            "n_gpus": world, "data": "synthetic",
            "build": {"seconds": t_build, "sample_descriptors": args.sample, "n_iter": args.iters, "nodes": int(tree.n_ref),
                      "tree_checksum": checksum},
            "quantize_stream": {"seconds": t_quant, "descriptors": n_desc * world,
                                "descriptors_per_s": n_desc * world / t_quant,
                                "note": "includes on-device generation of every batch"},
            "encode_and_invert": {"seconds": t_index, "postings": int(csr["nnz"]) * world,
                                  "images_per_s": args.images / t_index},
            "retrieve": {"queries": nq, "topk": args.topk, "seconds_per_batch": t_ret, "queries_per_s": nq / t_ret,
                         "api": "Database.retrieve_batch (host query descriptors in, host ids out)",
                         "every_query_finds_itself_first_with_score_0": ok,
                         "device_resident": {"seconds_per_batch": t_res, "queries_per_s": nq / t_res,
                                             "equals_host_path": ok_res}},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
