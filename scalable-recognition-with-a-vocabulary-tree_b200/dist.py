"""Multi-GPU layer: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for
the two real exchange steps of the path:

* build     -- descriptors sharded contiguously.  Top levels: per level-iteration ONE all-reduce of the
               [rows, D] integer sums + counts (+ changed counter) the assign kernel wrote.  From the
               first level with >= 4 nodes per rank on: ONE all-to-all hands every rank the members of
               the subtrees it owns (``engine.build_tree``), the levels below need no collective, and one
               all-reduce with a single contributor per row publishes the finished subtrees.  Integer
               sums and order-preserving exchange make the tree independent of the number of GPUs, bit
               for bit.
* index     -- images sharded by id; tf-idf only: ONE all-reduce of the per-node document
               frequencies (``Database.index``).
* quantise  -- descriptors sharded, tree replicated: no data-path collective.
* retrieve  -- inverted file sharded by image id; per-shard top-k, then ONE all-gather of the
               [Q, k] (rank key, global id, score) triples and a merge with the reference's
               comparator (score ascending == key descending, image index ascending).

The same code runs on CPU tensors over gloo, which is how the protocol is tested without GPUs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class Comm:
    """Thin wrapper over a process group (default: WORLD)."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_counts(self, t: torch.Tensor) -> torch.Tensor:
        flat = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(flat, t.contiguous().view(-1), group=self.group)
        return flat.view((self.world,) + tuple(t.shape))

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        return self.all_gather_counts(t)

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_to_all_rows(self, out: torch.Tensor, inp: torch.Tensor, out_rows, in_rows) -> torch.Tensor:
        """Row exchange: ``inp`` [sum(in_rows), ...] is cut into per-destination row blocks, ``out`` receives the
        blocks of every source in rank order."""
        dist.all_to_all_single(out, inp, [int(x) for x in out_rows], [int(x) for x in in_rows], group=self.group)
        return out

    def barrier(self):
        dist.barrier(group=self.group)


def shard_range(n: int, rank: int, world: int):
    """Contiguous split that keeps the original order: rank r owns [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_ranked(keys: torch.Tensor, ids: torch.Tensor, scores: torch.Tensor, topk: int):
    """Merge per-rank top-k lists.  keys/ids/scores: [world, Q, k] (ids global, -1 = empty slot).
    Order: key descending, ties by ascending image index -- the reference's stable ascending sort
    on scores (database.py:79-80).  Returns ([Q, topk] ids, scores, keys).  One launch of
    ``vt_topk_merge_scored`` (the comparator of the single-GPU merge); no CPU path."""
    from . import _native as nat
    lib = nat.load()
    if not keys.is_cuda:
        raise nat.NativeError("merge_ranked runs on the device (vt_topk_merge_scored); got CPU tensors")
    w, q, k = keys.shape
    topk = int(topk)
    with torch.cuda.device(keys.device):
        ck = keys.permute(1, 0, 2).reshape(q, w * k).contiguous()
        ci = ids.permute(1, 0, 2).reshape(q, w * k).to(torch.int32).contiguous()
        cs = scores.permute(1, 0, 2).reshape(q, w * k).contiguous()
        ok = torch.empty((max(q, 1), topk), dtype=torch.float64, device=keys.device)
        oi = torch.empty((max(q, 1), topk), dtype=torch.int32, device=keys.device)
        osc = torch.empty((max(q, 1), topk), dtype=torch.float64, device=keys.device)
        nat.check(lib.vt_topk_merge_scored(nat.ptr(ck), nat.ptr(ci), nat.ptr(cs), q, w * k, topk, nat.ptr(ok), nat.ptr(oi),
                                           nat.ptr(osc), nat.stream_ptr()), "vt_topk_merge_scored")
    return oi[:q], osc[:q], ok[:q]


def retrieve_sharded(score_local, img_base: int, topk: int, comm: Comm, merge=None):
    """``score_local()`` -> dict(key, id, score) [Q, topk] over this rank's image shard (ids local).
    ONE all-gather of the packed (key, global id, score) triples, then the merge; every rank returns
    the same global result.  ``merge`` replaces ``merge_ranked`` (CPU protocol tests pass a test double)."""
    out = score_local()
    ids = torch.where(out["id"] >= 0, out["id"] + int(img_base), out["id"])
    packed = torch.stack([out["key"].to(torch.float64), ids.to(torch.float64), out["score"].to(torch.float64)])
    allp = comm.all_gather(packed.contiguous())                  # [world, 3, Q, topk]; ids < 2^31 are exact in binary64
    keys, gids, scores = allp[:, 0], allp[:, 1].to(torch.int32), allp[:, 2]
    return (merge or merge_ranked)(keys, gids, scores, topk)
