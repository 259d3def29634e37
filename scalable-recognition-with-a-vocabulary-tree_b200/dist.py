"""Multi-GPU layer: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for
the two real exchange steps of the path:

* build     -- descriptors sharded contiguously; per level-iteration ONE all-reduce of the
               [rows, D] integer sums + counts (+ changed counter) the assign kernel wrote.
               Integer sums make the tree independent of the number of GPUs, bit for bit.
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
    on scores (database.py:79-80).  Returns ([Q, topk] ids, scores, keys)."""
    w, q, k = keys.shape
    keys = keys.permute(1, 0, 2).reshape(q, w * k)
    ids = ids.permute(1, 0, 2).reshape(q, w * k)
    scores = scores.permute(1, 0, 2).reshape(q, w * k)
    big = torch.iinfo(torch.int32).max
    ids_key = torch.where(ids >= 0, ids, torch.full_like(ids, big))
    keys = torch.where(ids >= 0, keys, torch.full_like(keys, float("-inf")))
    o1 = torch.sort(ids_key, dim=1, stable=True).indices                    # secondary: image index
    k1 = torch.gather(keys, 1, o1)
    o2 = torch.sort(k1, dim=1, descending=True, stable=True).indices        # primary: key
    order = torch.gather(o1, 1, o2)[:, :topk]
    return torch.gather(ids, 1, order), torch.gather(scores, 1, order), torch.gather(keys, 1, order)


def retrieve_sharded(score_local, img_base: int, topk: int, comm: Comm):
    """``score_local()`` -> dict(key, id, score) [Q, topk] over this rank's image shard (ids local).
    All-gathers the lists and merges them; every rank returns the same global result."""
    out = score_local()
    ids = torch.where(out["id"] >= 0, out["id"] + int(img_base), out["id"])
    keys = comm.all_gather(out["key"].contiguous())
    gids = comm.all_gather(ids.contiguous())
    scores = comm.all_gather(out["score"].contiguous())
    return merge_ranked(keys, gids, scores, topk)
