"""Device-side orchestration of the hot path: thin torch-tensor wrappers over the C ABI
(``include/vtree_b200.h``).  torch only owns memory and streams here; every compute step is one
of the hand-written sm_100a kernels.  Nothing in this module runs on the CPU -- without the
native library and a B200-class device every entry point raises ``NativeError``.

Reference functions covered (cbir/encoders/vocabulary.py, cbir/database.py):
``stage``      input side of ``extract_features`` (:29-34)        -> byte / float rows on device
``build_tree`` ``VocabularyTree.fit`` (:36-76), level synchronous  -> FlatTree
``quantize``   ``propagate_feature`` (:104-124), batched           -> reference node ids
               (three engines with identical ids: fused SIMT, level-synchronous SIMT, tcgen05)
``encode``     ``propagate`` + ``embedding`` (:78-137), batched    -> CSR (node, tf), norms
``build_index``the per-node posting dicts (:96-100)                -> sharded inverted CSR
``score``      ``Database.score`` / ``retrieve`` (database.py:59-81) -> top-k ids and scores
"""
from __future__ import annotations

import ctypes as C
import functools
import os

import numpy as np
import torch

from . import _native as nat
from .tree import FlatTree, heap_size, level_offset, pad_dim

VT_U8, VT_F32 = nat.VT_U8, nat.VT_F32


def _lib():
    return nat.load()


def _cuda_device_of(obj):
    """device of the first CUDA tensor found in ``obj`` (tensor, dict / list / tuple of tensors), else None."""
    if isinstance(obj, torch.Tensor):
        return obj.device if obj.is_cuda else None
    if isinstance(obj, dict):
        obj = obj.values()
    if isinstance(obj, (list, tuple)) or type(obj).__name__ == "dict_values":
        for v in obj:
            if isinstance(v, (torch.Tensor, dict)):
                d = _cuda_device_of(v)
                if d is not None:
                    return d
    return None


def _device_scoped(fn):
    """Run ``fn`` with the device of its tensor arguments current: the C ABI launches on the CURRENT device and
    ``nat.stream_ptr()`` is that device's current stream, so tensors on cuda:1 must not be driven from device 0
    (that would be an illegal address on the wrong context)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = None
        for a in list(args) + list(kw.values()):
            dev = _cuda_device_of(a)
            if dev is not None:
                break
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapper


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.uint8:
        return VT_U8
    if t.dtype == torch.float32:
        return VT_F32
    raise TypeError("descriptors must be uint8 or float32 on the device, got %s" % t.dtype)


# --------------------------------------------------------------------------------------- staging
def stage(x, device="cuda", prefer_bytes: bool = True):
    """[n, D] descriptors (numpy or torch; uint8, any int, float32/64) -> (rows [n, DP] on the
    device, D).  Integer-valued data in 0..255 is kept as bytes (lossless, 4x less HBM traffic);
    anything else stays float32."""
    nat.require_device()
    lib = _lib()
    target = torch.device(device)
    if target.type == "cuda" and target.index is not None and target.index != torch.cuda.current_device():
        with torch.cuda.device(target):
            return stage(x, target, prefer_bytes)
    if isinstance(x, np.ndarray):
        if x.dtype not in (np.uint8, np.float32):
            x = x.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(x))
    else:
        t = x
        if t.dtype not in (torch.uint8, torch.float32):
            t = t.to(torch.float32)
    if t.dim() == 1:
        t = t[None, :]
    t = t.to(device, non_blocking=True).contiguous()
    n, d = t.shape
    dp = pad_dim(d)
    s = nat.stream_ptr()
    if t.dtype == torch.uint8:
        if d == dp:
            return t, d
        out = torch.empty((n, dp), dtype=torch.uint8, device=t.device)
        nat.check(lib.vt_stage_pad(nat.ptr(t), n, d, nat.ptr(out), dp, 1, s), "vt_stage_pad")
        return out, d
    if prefer_bytes:
        out = torch.empty((n, dp), dtype=torch.uint8, device=t.device)
        flag = torch.zeros(1, dtype=torch.int32, device=t.device)
        nat.check(lib.vt_stage_f32_to_u8(nat.ptr(t), n, d, nat.ptr(out), dp, nat.ptr(flag), s), "vt_stage_f32_to_u8")
        if int(flag.item()) == 0:
            return out, d
    if d == dp:
        return t, d
    out = torch.empty((n, dp), dtype=torch.float32, device=t.device)
    nat.check(lib.vt_stage_pad(nat.ptr(t), n, d, nat.ptr(out), dp, 4, s), "vt_stage_pad")
    return out, d


# --------------------------------------------------------------------------------------- descent
SORTED_MIN_ROWS = 1 << 18      # below this the fused single-launch descent wins (launch count, short segments)


@_device_scoped
def quantize(tree_dev: dict, k: int, depth: int, desc: torch.Tensor, want_heap: bool = False,
             stats: torch.Tensor = None, start: int = 0, method: str = "auto", work: torch.Tensor = None):
    """Reference node id of the stop node of every descriptor (int32 [n]); optionally also the
    heap index.  ``stats`` (int64 [4], zeroed by the caller) receives the binary64 recheck count.

    ``method``: "fused" = one launch, warp-per-descriptor descent (any dtype, any start node);
    "sorted" = level-synchronous descent over the batch sorted by node (bytes, large batches);
    "mma" = the same with the distance step on the INT8 tensor pipe (tcgen05, 128-byte rows, k <= 16,
    centroids in [0, 255.99]); "auto" picks by batch size.  All return identical ids."""
    lib = _lib()
    n, dp = desc.shape
    word = torch.empty(n, dtype=torch.int32, device=desc.device)
    leaf = torch.empty(n, dtype=torch.int32, device=desc.device) if want_heap else None
    if method == "auto":
        # the level-synchronous engines pack the descent path into 32 bits and index the heap with int32 -- the
        # same predicate vt_quantize_host applies before it leaves the fused kernel
        bits = max(1, int(k).bit_length())
        packed_ok = (depth - 1) * bits <= 32 and heap_size(k, depth) < (1 << 31)
        big = desc.dtype == torch.uint8 and start == 0 and n >= SORTED_MIN_ROWS and depth >= 2 and packed_ok
        method = ("mma" if tree_dev.get("mma_ok", False) else "sorted") if big else "fused"
    if method == "mma":
        # the byte planes depend only on the tree: prepared once and kept with the tree's device tensors
        planes = tree_dev.get("_mma_planes")
        if planes is None or tree_dev.get("_mma_planes_for") != (tree_dev["centroids"].data_ptr(), k, depth):
            planes = torch.empty(int(lib.vt_quantize_mma_planes_bytes(k, depth)), dtype=torch.uint8, device=desc.device)
            nat.check(lib.vt_quantize_mma_prepare(nat.ptr(tree_dev["centroids"]), k, depth, nat.ptr(planes), nat.ptr(stats),
                                                  nat.stream_ptr()), "vt_quantize_mma_prepare")
            tree_dev["_mma_planes"] = planes
            tree_dev["_mma_planes_for"] = (tree_dev["centroids"].data_ptr(), k, depth)
        need = 16 * max(n, 1) + int(lib.vt_sort_scratch_bytes(n))
        if work is None or work.numel() * work.element_size() < need:
            work = torch.empty(need, dtype=torch.uint8, device=desc.device)
        nat.check(lib.vt_quantize_mma_prepared(nat.ptr(desc), _dtype_code(desc), n, dp, nat.ptr(tree_dev["centroids"]),
                                               nat.ptr(tree_dev["internal"]), nat.ptr(tree_dev["ref_id"]), k, depth,
                                               nat.ptr(leaf), nat.ptr(word), nat.ptr(stats), nat.ptr(planes),
                                               nat.ptr(work), work.numel() * work.element_size(), nat.stream_ptr()),
                  "vt_quantize_mma_prepared")
    elif method == "sorted":
        need = int(lib.vt_quantize_sorted_work_bytes(n))
        if work is None or work.numel() * work.element_size() < need:
            work = torch.empty(need, dtype=torch.uint8, device=desc.device)
        nat.check(lib.vt_quantize_sorted(nat.ptr(desc), _dtype_code(desc), n, dp, nat.ptr(tree_dev["centroids"]),
                                         nat.ptr(tree_dev["internal"]), nat.ptr(tree_dev["ref_id"]), k, depth,
                                         nat.ptr(leaf), nat.ptr(word), nat.ptr(stats), nat.ptr(work),
                                         work.numel() * work.element_size(), nat.stream_ptr()), "vt_quantize_sorted")
    else:
        nat.check(lib.vt_quantize(nat.ptr(desc), _dtype_code(desc), n, dp, nat.ptr(tree_dev["centroids"]),
                                  nat.ptr(tree_dev["internal"]), nat.ptr(tree_dev["ref_id"]), k, depth, start,
                                  nat.ptr(leaf), nat.ptr(word), nat.ptr(stats), nat.stream_ptr()), "vt_quantize")
    return (word, leaf) if want_heap else word


@_device_scoped
def quantize_host(tree_dev: dict, k: int, depth: int, h_desc, h_word, chunk_rows: int = 0):
    """End-to-end variant: host (pinned) descriptor rows in, host word ids out, chunked and
    overlapped inside the native library (``vt_quantize_host``)."""
    lib = _lib()
    n, dp = h_desc.shape
    code = VT_U8 if h_desc.dtype in (torch.uint8, np.uint8) else VT_F32
    nat.check(lib.vt_quantize_host(nat.ptr(h_desc), code, n, dp, nat.ptr(tree_dev["centroids"]),
                                   nat.ptr(tree_dev["internal"]), nat.ptr(tree_dev["ref_id"]), k, depth,
                                   nat.ptr(h_word), chunk_rows), "vt_quantize_host")
    return h_word


# --------------------------------------------------------------------------------------- sort
@_device_scoped
def sort_pairs(keys: torch.Tensor, vals: torch.Tensor, n_bits: int):
    """Stable radix sort of (uint32 key, uint32 value) pairs held in int32 tensors."""
    lib = _lib()
    n = keys.numel()
    if n == 0 or n_bits == 0:
        return keys, vals
    tk, tv = torch.empty_like(keys), torch.empty_like(vals)
    hist = torch.empty(int(lib.vt_sort_scratch_bytes(n)), dtype=torch.uint8, device=keys.device)
    flip = C.c_int(0)
    nat.check(lib.vt_sort_pairs(nat.ptr(keys), nat.ptr(vals), nat.ptr(tk), nat.ptr(tv), n, n_bits, nat.ptr(hist),
                                C.byref(flip), nat.stream_ptr()), "vt_sort_pairs")
    return (tk, tv) if flip.value else (keys, vals)


# --------------------------------------------------------------------------------------- build
class _NoComm:
    """Single-GPU stand-in for the collective hooks ``build_tree`` uses."""
    world = 1
    rank = 0

    def all_reduce(self, t):
        return t

    def all_gather_counts(self, t):
        return t[None]


def subtree_split_level(k: int, depth: int, world: int):
    """First level whose nodes give every rank a few whole subtrees (>= 4 nodes per rank); None when the tree
    is too small for that and the build stays with one all-reduce per Lloyd pass."""
    for level in range(1, depth):
        if k ** level >= 4 * world:
            return level
    return None


def subtree_owners(seg_global: torch.Tensor, world: int) -> torch.Tensor:
    """Owner rank of every node of the split level: contiguous node ranges balanced by member count (midpoint of
    the node's member range against equal cuts).  Deterministic from the all-reduced counts, so every rank
    computes the same map."""
    c = seg_global.to(torch.float64)
    total = float(c.sum().item())
    if total <= 0:
        return torch.zeros_like(seg_global)
    mid = torch.cumsum(c, 0) - 0.5 * c
    return torch.clamp((mid * (world / total)).floor().to(torch.int64), 0, world - 1)


class CudaLevelOps:
    """The five per-level device steps of the build, each one C-ABI call.  ``build_tree`` only
    talks to this interface so that the multi-rank protocol around it can be exercised on CPU
    (gloo) with a test double; the product always uses this class."""

    def __init__(self, use_tensor_cores: bool = True):
        nat.require_device()
        self.lib = _lib()
        self.use_tensor_cores = use_tensor_cores
        self._planes = None

    def column_sums(self, desc):
        n, dp = desc.shape
        col = torch.zeros(dp, dtype=torch.int64, device=desc.device)
        nat.check(self.lib.vt_column_sums(nat.ptr(desc), VT_U8, n, dp, nat.ptr(col), nat.stream_ptr()), "vt_column_sums")
        return col

    def kmeans_init(self, desc, perm, seg_begin, seg_local, seg_global, rank0, n_nodes, k, child_cent):
        nat.check(self.lib.vt_kmeans_init(nat.ptr(desc), VT_U8, desc.shape[1], nat.ptr(perm), nat.ptr(seg_begin),
                                          nat.ptr(seg_local), nat.ptr(seg_global), nat.ptr(rank0), n_nodes, k,
                                          nat.ptr(child_cent), nat.stream_ptr()), "vt_kmeans_init")

    def kmeans_assign(self, desc, perm, parent, n_active, node_internal, n_nodes, k, child_cent, label, sums,
                      counts, changed, stats):
        if self.use_tensor_cores and desc.shape[1] == 128 and k <= 16 and n_active > 0:
            # byte descriptors => centres lie in [0, 255]: inside the 8.16 fixed-point range of the planes
            need = int(self.lib.vt_kmeans_planes_bytes(n_nodes, k))
            if self._planes is None or self._planes.numel() < need:
                self._planes = torch.empty(need, dtype=torch.uint8, device=desc.device)
            s = nat.stream_ptr()
            nat.check(self.lib.vt_kmeans_planes(nat.ptr(child_cent), n_nodes, k, 128, nat.ptr(self._planes), None, s),
                      "vt_kmeans_planes")
            nat.check(self.lib.vt_kmeans_assign_mma(nat.ptr(desc), VT_U8, 128, nat.ptr(perm), nat.ptr(parent), n_active,
                                                    nat.ptr(node_internal), n_nodes, k, nat.ptr(child_cent),
                                                    nat.ptr(self._planes), nat.ptr(label), nat.ptr(sums), nat.ptr(counts),
                                                    nat.ptr(changed), nat.ptr(stats), s), "vt_kmeans_assign_mma")
            return
        nat.check(self.lib.vt_kmeans_assign(nat.ptr(desc), VT_U8, desc.shape[1], nat.ptr(perm), nat.ptr(parent),
                                            n_active, nat.ptr(node_internal), n_nodes, k, nat.ptr(child_cent),
                                            nat.ptr(label), nat.ptr(sums), nat.ptr(counts), nat.ptr(changed),
                                            nat.ptr(stats), nat.stream_ptr()), "vt_kmeans_assign")

    def kmeans_update(self, sums, counts, n_rows, dp, child_cent):
        nat.check(self.lib.vt_kmeans_update(nat.ptr(sums), nat.ptr(counts), n_rows, dp, nat.ptr(child_cent),
                                            nat.stream_ptr()), "vt_kmeans_update")

    def stable_sort(self, keys, vals, n_bits):
        """stable sort of (int32 key, int32 value) pairs by key (subtree exchange: regroup received rows by node)."""
        return sort_pairs(keys, vals, n_bits)

    def partition(self, parent, label, n_active, node_internal, k, n_rows, perm):
        """stable partition of every node's members into its children (vocabulary.py:64-66):
        returns (child index per member, member rows), both sorted by child index; members of leaf
        nodes carry the sentinel key ``n_rows`` and end up last."""
        keys = torch.empty(max(n_active, 1), dtype=torch.int32, device=parent.device)
        nat.check(self.lib.vt_partition_keys(nat.ptr(parent), nat.ptr(label), n_active, nat.ptr(node_internal), k,
                                             n_rows, nat.ptr(keys), nat.stream_ptr()), "vt_partition_keys")
        return sort_pairs(keys[:n_active], perm, int(n_rows).bit_length())


class _ChangedProbes:
    """Asynchronous reads of the device-side `changed` counter of the Lloyd passes: ``add`` copies the counter into a
    pinned slot and records an event right behind the copy; ``get(i)`` waits for pass i only.  One pinned buffer per
    build.  CPU tensors (gloo protocol tests) are read directly."""

    def __init__(self, slots: int, cuda: bool):
        self.host = torch.empty(slots, dtype=torch.int64)
        if cuda:
            self.host = self.host.pin_memory()
        self.cuda, self.events, self.base = cuda, [], 0

    def reset(self):
        self.events = []

    def add(self, changed: torch.Tensor):
        i = len(self.events)
        if not self.cuda:
            self.host[i] = int(changed.item())
            self.events.append(None)
            return
        self.host[i:i + 1].copy_(changed, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events.append(ev)

    def get(self, i: int) -> int:
        if self.events[i] is not None:
            self.events[i].synchronize()
        return int(self.host[i])


@_device_scoped
def build_tree(desc: torch.Tensor, d: int, k: int, depth: int, n_iter: int = 10, comm=None,
               stats: torch.Tensor = None, log=None, ops=None) -> FlatTree:
    """Level-synchronous hierarchical k-means over byte descriptors [N, DP] (this rank's shard).

    Per level: seeded initial centres (member floor(j*m/k) of every node), ``n_iter`` exact Lloyd
    steps over ALL nodes of the level at once (early exit when no label changes), stable
    partition of the members into the children.  With ``comm`` (see ``dist.Comm``) the integer
    sums/counts are all-reduced, so every rank ends with the same tree, bit for bit."""
    if desc.dtype != torch.uint8:
        raise TypeError("build_tree needs byte descriptors (integer valued 0..255); stage() keeps float rows only "
                        "when the data is not byte-exact, which the exact-sum k-means does not support")
    ops = ops or CudaLevelOps()
    comm = comm or _NoComm()
    dev = desc.device
    n, dp = desc.shape
    n_heap = heap_size(k, depth)
    cent = torch.zeros((n_heap, dp), dtype=torch.float32, device=dev)
    internal = torch.zeros(n_heap, dtype=torch.uint8, device=dev)
    exists = torch.zeros(n_heap, dtype=torch.uint8, device=dev)
    count_heap = torch.zeros(n_heap, dtype=torch.int64, device=dev)
    exists[0] = 1

    # root = mean of all features (vocabulary.py:48-49), exact
    col = torch.cat([ops.column_sums(desc), torch.tensor([n], dtype=torch.int64, device=dev)])
    col = comm.all_reduce(col)
    n_total = int(col[dp].item())
    if n_total > 0:
        cent[0] = (col[:dp].to(torch.float64) / float(n_total)).to(torch.float32)
    count_heap[0] = n_total

    perm = torch.arange(n, dtype=torch.int32, device=dev)
    parent = torch.zeros(n, dtype=torch.int32, device=dev)
    seg_local = torch.tensor([n], dtype=torch.int64, device=dev)
    seg_global = torch.tensor([n_total], dtype=torch.int64, device=dev)

    probes = _ChangedProbes(n_iter + 1, dev.type == "cuda")
    split = subtree_split_level(k, depth, comm.world) if comm.world > 1 else None
    publish, published_from = None, None
    for level in range(depth):
        n_nodes = k ** level
        n_rows = n_nodes * k
        if split is not None and level == split:
            # ---- subtree ownership: ONE all-to-all moves every member of a node to the rank that owns the node's
            # subtree; the levels from here on run without any collective.  Ranks hold contiguous ranges of the
            # global row order and send their members node by node, so concatenating the blocks in rank order and
            # regrouping them by node with a STABLE sort keeps every node's members in global order -- the seeded
            # initial centres (member floor(j*m/k)) and hence the tree are those of the single-process build.
            allc = comm.all_gather_counts(seg_local)                      # [world, n_nodes]
            owner = subtree_owners(seg_global, comm.world)
            mine = owner == comm.rank
            send_rows = desc.index_select(0, perm.long())                 # members in node order (perm is node-grouped)
            send_counts = torch.zeros(comm.world, dtype=torch.int64, device=dev).index_add_(0, owner, seg_local)
            recv_per_src = (allc * mine[None, :].to(allc.dtype))          # [world, n_nodes] rows I receive
            recv_counts = recv_per_src.sum(1)
            sc_h, rc_h = send_counts.cpu().tolist(), recv_counts.cpu().tolist()
            recv = torch.empty((int(sum(rc_h)), dp), dtype=desc.dtype, device=dev)
            comm.all_to_all_rows(recv, send_rows, rc_h, sc_h)
            node_ids = torch.arange(n_nodes, dtype=torch.int32, device=dev)
            keys = torch.repeat_interleave(node_ids.repeat(comm.world), recv_per_src.reshape(-1))
            order = torch.arange(keys.numel(), dtype=torch.int32, device=dev)
            parent, perm = ops.stable_sort(keys.contiguous(), order, max(1, int(n_nodes).bit_length()))
            desc, n = recv, int(recv.shape[0])
            seg_global = torch.where(mine, seg_global, torch.zeros_like(seg_global))
            seg_local = seg_global.clone()
            publish, published_from = comm, level
            comm = _NoComm()
            del send_rows
        n_active = perm.numel()
        node_internal = (seg_global >= k).to(torch.uint8)                 # vocabulary.py:55
        if int(node_internal.sum().item()) == 0:
            break
        o, co = level_offset(k, level), level_offset(k, level + 1)
        internal[o:o + n_nodes] = node_internal
        child_cent = cent[co:co + n_rows]
        seg_begin = torch.cumsum(seg_local, 0) - seg_local
        if comm.world > 1:
            allc = comm.all_gather_counts(seg_local)                      # [world, n_nodes]
            rank0 = (allc[:comm.rank].sum(0) if comm.rank > 0 else torch.zeros_like(seg_local)).contiguous()
            ops.kmeans_init(desc, perm, seg_begin, seg_local, seg_global, rank0, n_nodes, k, child_cent)
            comm.all_reduce(child_cent)                                   # one owner per row: exact
        else:
            ops.kmeans_init(desc, perm, seg_begin, seg_local, None, None, n_nodes, k, child_cent)

        label = torch.full((max(n_active, 1),), 255, dtype=torch.uint8, device=dev)
        acc = torch.zeros(n_rows * dp + n_rows + 1, dtype=torch.int64, device=dev)   # sums | counts | changed
        sums, counts, changed = acc[:n_rows * dp], acc[n_rows * dp:n_rows * dp + n_rows], acc[n_rows * dp + n_rows:]
        local_counts = None

        def assign():
            nonlocal local_counts
            acc.zero_()
            ops.kmeans_assign(desc, perm, parent, n_active, node_internal, n_nodes, k, child_cent, label, sums,
                              counts, changed, stats)
            if comm.world > 1:
                local_counts = counts.clone()
                comm.all_reduce(acc)                                      # sums, counts and changed in one go
            else:
                local_counts = counts

        # Lloyd passes.  "labels repeat" (changed == 0) is a fixed point: one more update + assign reproduces the same
        # sums, centres and labels bit for bit.  So the counter of pass `it` is read one pass LATE, from a pinned copy
        # whose event was recorded right behind the pass: the device never idles for the host round trip, and
        # stopping one pass after the fixed point leaves exactly the state of stopping at it.
        converged, iters = False, 0
        probes.reset()
        for it in range(n_iter):
            assign()
            probes.add(changed)
            if it > 1 and probes.get(it - 1) == 0:
                converged, iters = True, it - 1                           # fixed point reached at pass it - 1
                break
            ops.kmeans_update(sums, counts, n_rows, dp, child_cent)
            iters = it + 1
        if not converged and n_iter > 1 and probes.get(n_iter - 1) == 0:
            converged, iters = True, n_iter - 1                           # ... at the last pass: the labels are current
        if not converged:
            assign()                                                      # labels_ against the final centres
        if log is not None:
            log(level=level, nodes=n_nodes, members=n_active, iters=iters, converged=converged)

        kid_exists = node_internal.repeat_interleave(k)
        exists[co:co + n_rows] = kid_exists
        count_heap[co:co + n_rows] = counts * kid_exists
        seg_global = (counts * kid_exists).clone()
        seg_local = (local_counts * kid_exists).clone()
        if level + 1 >= depth:
            break
        skeys, svals = ops.partition(parent, label, n_active, node_internal, k, n_rows, perm)
        n_next = int(seg_local.sum().item())
        perm = svals[:n_next].contiguous()
        parent = skeys[:n_next].contiguous()

    if publish is not None:
        # every heap row below the split level was written by exactly one owner (zeros elsewhere): the sum is exact
        o, co = level_offset(k, published_from), level_offset(k, published_from + 1)
        publish.all_reduce(cent[co:])
        publish.all_reduce(count_heap[co:])
        flags = torch.cat([internal[o:], exists[co:]]).to(torch.int32)
        publish.all_reduce(flags)
        internal[o:] = flags[:n_heap - o].to(torch.uint8)
        exists[co:] = flags[n_heap - o:].to(torch.uint8)
    if dev.type == "cuda":       # everything stays on the device: numbering computed there, host mirrors on demand
        tree = FlatTree.from_device(k, depth, d, cent, exists, internal, count_heap, byte_built=True)
    else:
        tree = FlatTree(k, depth, d, cent.numpy(), exists.numpy(), internal.numpy(), count_heap.numpy(), byte_built=True)
    return tree


# --------------------------------------------------------------------------------------- encode
_SINGLE_PASS_WORDS = 1 << 31        # the single-pass encoder stages 8 B per descriptor


@_device_scoped
def encode(words: torch.Tensor, img_off: torch.Tensor, tree_dev: dict, depth: int, n_ref: int,
           all_nodes: bool = True, want_df: bool = False, min_level: int = 0) -> dict:
    """Bag-of-words CSR of a batch of images.  ``words`` int32 [n_desc] (reference ids of the
    stop nodes) grouped by image, ``img_off`` int64 [n_img+1] on the device.  ``min_level`` > 0
    (all-nodes mode) drops the nodes above that level: the level mask of Nister & Stewenius."""
    lib = _lib()
    dev = words.device
    n_img = img_off.numel() - 1
    s = nat.stream_ptr()
    sizes = img_off[1:] - img_off[:-1]
    max_words = int(sizes.max().item()) if n_img > 0 else 0
    if max_words > lib.vt_encode_max_words():
        # the reference has no limit (vocabulary.py:78-102); one CTA sorts at most vt_encode_max_words() words in shared
        # memory, so larger images are encoded in pieces whose rows are then merged (counts are additive)
        return _encode_in_pieces(words, img_off, tree_dev, depth, n_ref, all_nodes, want_df, min_level,
                                 int(lib.vt_encode_max_words()))
    nnz = torch.zeros(max(n_img, 1), dtype=torch.int32, device=dev)
    if not 0 <= min_level <= depth:
        raise ValueError("min_level must lie in [0, depth]")
    args = (nat.ptr(words), nat.ptr(img_off), n_img, max_words, 1 + int(min_level) if all_nodes else 0,
            nat.ptr(tree_dev["parent_ref"]), nat.ptr(tree_dev["level_ref"]), depth)
    sumsq = torch.zeros(max(n_img, 1), dtype=torch.int64, device=dev)
    df = torch.zeros(n_ref, dtype=torch.int32, device=dev) if want_df else None
    row_off = torch.zeros(n_img + 1, dtype=torch.int64, device=dev)
    n_words = int(words.numel())
    if not all_nodes and 0 < n_words <= _SINGLE_PASS_WORDS:
        # stop nodes only: an image has at most as many entries as descriptors, so ONE pass writes its row at the
        # image's word offset (and reports the row length); a copy kernel then packs the rows
        node_tmp = torch.empty(n_words, dtype=torch.int32, device=dev)
        count_tmp = torch.empty(n_words, dtype=torch.int32, device=dev)
        nat.check(lib.vt_encode(*args, nat.ptr(nnz), nat.ptr(img_off), nat.ptr(node_tmp), nat.ptr(count_tmp),
                                nat.ptr(sumsq), nat.ptr(df), s), "vt_encode(single pass)")
        row_off[1:] = torch.cumsum(nnz[:n_img].to(torch.int64), 0)
        total = int(row_off[-1].item())
        node = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        count = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        nat.check(lib.vt_csr_compact(nat.ptr(img_off), nat.ptr(nnz), nat.ptr(row_off), n_img, nat.ptr(node_tmp),
                                     nat.ptr(count_tmp), nat.ptr(node), nat.ptr(count), s), "vt_csr_compact")
        del node_tmp, count_tmp
    else:
        nat.check(lib.vt_encode(*args, nat.ptr(nnz), None, None, None, None, None, s), "vt_encode(count)")
        row_off[1:] = torch.cumsum(nnz[:n_img].to(torch.int64), 0)
        total = int(row_off[-1].item())
        node = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        count = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        nat.check(lib.vt_encode(*args, None, nat.ptr(row_off), nat.ptr(node), nat.ptr(count), nat.ptr(sumsq),
                                nat.ptr(df), s), "vt_encode(fill)")
    return dict(row_off=row_off, node=node[:total], count=count[:total], sumsq=sumsq[:n_img], df=df,
                n_img=n_img, nnz=total)


def _encode_in_pieces(words, img_off, tree_dev, depth, n_ref, all_nodes, want_df, min_level, limit) -> dict:
    """``encode`` for batches holding images with more than ``limit`` descriptors: every image is cut into pieces of at most
    ``limit`` words, the pieces are encoded as rows of their own and the rows of an image are merged on the device (sort by
    (image, emission order, node), add the counts of equal nodes).  Rows come out in the order ``vt_encode`` emits them."""
    dev = words.device
    n_img = img_off.numel() - 1
    off = img_off.cpu().numpy()
    piece_off, piece_img = [0], []
    for i in range(n_img):
        b, e = int(off[i]), int(off[i + 1])
        for c in range(b, e, limit):
            piece_off.append(min(c + limit, e))
            piece_img.append(i)
        if e == b:
            piece_off.append(e)
            piece_img.append(i)
    sub = encode(words, torch.tensor(piece_off, dtype=torch.int64, device=dev), tree_dev, depth, n_ref, all_nodes=all_nodes,
                 want_df=False, min_level=min_level)
    p_img = torch.tensor(piece_img, dtype=torch.int64, device=dev)
    lens = sub["row_off"][1:] - sub["row_off"][:-1]
    img = torch.repeat_interleave(p_img, lens)
    node = sub["node"].to(torch.int64)
    if all_nodes:       # vt_encode emits the deepest level first, ids ascending inside a level
        rank = depth - tree_dev["level_ref"].to(torch.int64)[node]
        key = (img * (depth + 1) + rank) * n_ref + node
    else:
        key = img * n_ref + node
    key, order = torch.sort(key, stable=True)
    uniq, inverse = torch.unique_consecutive(key, return_inverse=True)
    count = torch.zeros(uniq.numel(), dtype=torch.int64, device=dev).index_add_(0, inverse, sub["count"].to(torch.int64)[order])
    out_node = (uniq % n_ref).to(torch.int32)
    out_img = uniq // (n_ref * (depth + 1)) if all_nodes else uniq // n_ref
    row_off = torch.zeros(n_img + 1, dtype=torch.int64, device=dev)
    row_off[1:] = torch.cumsum(torch.bincount(out_img, minlength=n_img)[:n_img], 0)
    sumsq = torch.zeros(max(n_img, 1), dtype=torch.int64, device=dev).index_add_(0, out_img, count * count)
    df = torch.bincount(out_node.to(torch.int64), minlength=n_ref)[:n_ref].to(torch.int32) if want_df else None
    total = int(uniq.numel())
    return dict(row_off=row_off, node=out_node, count=count.to(torch.int32), sumsq=sumsq[:n_img], df=df, n_img=n_img, nnz=total)


def concat_csr(parts: list) -> dict:
    """Concatenates forward CSRs of consecutive image batches (streaming ``Database.index``)."""
    if len(parts) == 1:
        return parts[0]
    offs, base = [parts[0]["row_off"][:1]], 0
    for p in parts:
        offs.append(p["row_off"][1:] + base)
        base += p["nnz"]
    df = None
    if all(p["df"] is not None for p in parts):
        df = torch.stack([p["df"] for p in parts]).sum(0, dtype=torch.int32)
    return dict(row_off=torch.cat(offs), node=torch.cat([p["node"] for p in parts]),
                count=torch.cat([p["count"] for p in parts]), sumsq=torch.cat([p["sumsq"] for p in parts]),
                df=df, n_img=sum(p["n_img"] for p in parts), nnz=base)


def csr_rows(csr: dict, rows: torch.Tensor) -> dict:
    """Sub-CSR of the given rows (int64 device tensor, any order, repeats allowed), gathered on the device --
    how ``Database.retrieve_batch`` reuses the rows of indexed images as queries without a host mirror."""
    ro = csr["row_off"]
    starts, lens = ro[rows], ro[rows + 1] - ro[rows]
    off = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=ro.device)
    off[1:] = torch.cumsum(lens, 0)
    total = int(off[-1].item())
    idx = torch.repeat_interleave(starts - off[:-1], lens) + torch.arange(total, dtype=torch.int64, device=ro.device)
    return dict(row_off=off, node=csr["node"][idx], count=csr["count"][idx], sumsq=csr["sumsq"][rows], df=None,
                n_img=int(rows.numel()), nnz=total)


def entry_rows(csr: dict) -> torch.Tensor:
    """image index of every CSR entry (int64 [nnz])."""
    n_img = csr["n_img"]
    sizes = csr["row_off"][1:] - csr["row_off"][:-1]
    return torch.repeat_interleave(torch.arange(n_img, device=sizes.device), sizes)


@_device_scoped
def idf_from_df(df: torch.Tensor, n_images: int) -> torch.Tensor:
    """idf_n = ln(N / df_n) in binary64 (0 where no image visits the node) -- the Nister & Stewenius weight
    that vocabulary.py:131,137 leaves commented out.  ``df`` int32 [n_ref] (all-reduced over the ranks first
    when the images are sharded)."""
    lib = _lib()
    idf = torch.empty(df.numel(), dtype=torch.float64, device=df.device)
    nat.check(lib.vt_idf(nat.ptr(df.contiguous()), df.numel(), int(n_images), nat.ptr(idf), nat.stream_ptr()), "vt_idf")
    return idf


@_device_scoped
def weighted_rnorm(csr: dict, idf: torch.Tensor, norm: str = "l2") -> torch.Tensor:
    """1 / |idf * tf|_2 (or _1) per CSR row, binary64 [n_img]; -1 for all-zero rows."""
    lib = _lib()
    rn = torch.empty(max(csr["n_img"], 1), dtype=torch.float64, device=idf.device)
    nat.check(lib.vt_weighted_rnorm(nat.ptr(csr["row_off"]), nat.ptr(csr["node"]), nat.ptr(csr["count"]), csr["n_img"],
                                    nat.ptr(idf), int(norm == "l1"), nat.ptr(rn), nat.stream_ptr()), "vt_weighted_rnorm")
    return rn[:csr["n_img"]]


@_device_scoped
def query_weights(q: dict, idf: torch.Tensor, norm: str = "l2") -> dict:
    """Query-side factors of the weighted rank keys (csrc/vt_weight.cu): dict(qv, qi, rnorm), binary64."""
    lib = _lib()
    rn = weighted_rnorm(q, idf, norm)
    qv = torch.empty(max(q["nnz"], 1), dtype=torch.float64, device=idf.device)
    qi = torch.empty(max(q["nnz"], 1), dtype=torch.float64, device=idf.device)
    rn_full = rn if rn.numel() else torch.full((1,), -1.0, dtype=torch.float64, device=idf.device)
    nat.check(lib.vt_query_weights(nat.ptr(q["row_off"]), nat.ptr(q["node"]), nat.ptr(q["count"]), q["n_img"],
                                   nat.ptr(idf), nat.ptr(rn_full), int(norm == "l1"), nat.ptr(qv), nat.ptr(qi),
                                   nat.stream_ptr()), "vt_query_weights")
    return dict(qv=qv, qi=qi, rnorm=rn_full)


# --------------------------------------------------------------------------------------- index
@_device_scoped
def build_index(csr: dict, n_ref: int) -> dict:
    """Forward CSR -> inverted file sharded by image range (postings in image order, 4 bytes each:
    (tf << 13) | local image index).  The same index serves tf, tf-idf/L2 and tf-idf/L1 scoring."""
    lib = _lib()
    dev = csr["sumsq"].device
    s = nat.stream_ptr()
    n_img, nnz = csr["n_img"], csr["nnz"]
    shard = lib.vt_score_shard_size()
    n_shards = max(1, (n_img + shard - 1) // shard)
    n_keys = n_shards * n_ref
    n_bits = int(n_keys).bit_length()
    post_off = torch.empty(n_keys + 1, dtype=torch.int32, device=dev)
    if nnz == 0:                 # no image has a descriptor (or there is no image): every posting list is empty
        postings = torch.empty(1, dtype=torch.int32, device=dev)
        nat.check(lib.vt_posting_offsets(None, 0, n_keys, nat.ptr(post_off), s), "vt_posting_offsets")
    else:
        keys = torch.empty(nnz, dtype=torch.int32, device=dev)
        packed = torch.empty(nnz, dtype=torch.int32, device=dev)
        nat.check(lib.vt_posting_pack(nat.ptr(csr["row_off"]), nat.ptr(csr["node"]), nat.ptr(csr["count"]), n_img, shard,
                                      n_ref, nat.ptr(keys), nat.ptr(packed), s), "vt_posting_pack")
        skeys, postings = sort_pairs(keys, packed, n_bits)        # the sorted values ARE the postings
        nat.check(lib.vt_posting_offsets(nat.ptr(skeys), nnz, n_keys, nat.ptr(post_off), s), "vt_posting_offsets")
        del keys, packed, skeys
    rnorm = torch.empty(max(n_img, 1), dtype=torch.float64, device=dev)
    nat.check(lib.vt_rnorm(nat.ptr(csr["sumsq"]), n_img, nat.ptr(rnorm), s), "vt_rnorm")
    shard_empty = torch.zeros(n_shards, dtype=torch.uint8, device=dev)
    empty = torch.nonzero(csr["sumsq"][:n_img] == 0).flatten()
    if empty.numel():
        shard_empty[(empty // shard).unique()] = 1
    # per-shard spread of the norms (tf-mode fast selection): rho = min(1/|d|) / max(1/|d|), -1 with empties
    pad = n_shards * shard - n_img
    rn = rnorm[:n_img]
    lo = torch.cat([rn, rn.new_full((pad,), float("inf"))]).view(n_shards, shard).min(1).values
    hi = torch.cat([rn, rn.new_full((pad,), 0.0)]).view(n_shards, shard).max(1).values
    pos = torch.where(rn > 0, rn, rn.new_full((), float("inf")))             # described images only
    rmin = torch.cat([pos, pos.new_full((pad,), float("inf"))]).view(n_shards, shard).min(1).values
    rmin = torch.where(torch.isinf(rmin), torch.zeros_like(rmin), rmin * (1.0 - 1e-12)).contiguous()
    shard_rho = torch.where((shard_empty == 0) & (hi > 0), lo / hi.clamp(min=1e-300) * (1.0 - 1e-9),
                            torch.full_like(lo, -1.0)).contiguous()
    # size-biased mean length of a (node, shard) posting run = the run a random posting sits in: a few postings
    # for leaf-word indexes, thousands once the top tree levels (which list every image) are indexed
    lens = (post_off[1:] - post_off[:-1]).to(torch.float64)
    long_lists = bool(nnz > 0 and float((lens * lens).sum().item()) / float(nnz) > 64.0)
    run_mean = float(nnz) / max(int((lens > 0).sum().item()), 1)               # postings per non-empty (word, shard) run
    return dict(post_off=post_off, postings=postings[:nnz], rnorm=rnorm[:n_img], sumsq=csr["sumsq"], n_img=n_img,
                shard_empty=shard_empty, shard_rho=shard_rho, shard_rmax=hi.clamp(min=0.0).contiguous(), shard_rmin=rmin,
                max_sumsq=int(csr["sumsq"][:n_img].max().item()) if n_img > 0 else 0,
                long_lists=long_lists, run_mean=run_mean,
                n_ref=n_ref, shard=shard, n_shards=n_shards, nnz=nnz)


# --------------------------------------------------------------------------------------- dense top (all-nodes indexes)
DENSE_MAX_LEVEL = 4           # levels 2..4 of a k=10 tree: 11,100 nodes, every one visited by >= ~10 % of the images
DENSE_MAX_TOP = 17            # levels 0..1: 1 + k nodes, k <= 16


def dense_layout(tree_dev: dict, depth: int, n_ref: int, max_level: int = DENSE_MAX_LEVEL):
    """Which nodes of an all-nodes index leave the posting lists (csrc/vt_dense.cu): levels 0..1 -> int32 "top"
    columns, levels 2..max_level -> uint8 columns of the dense block, deeper levels stay postings.  None when the
    tree's top does not fit (k > 16)."""
    level = tree_dev["level_ref"].to(torch.int64)
    dev = level.device
    top = level <= 1
    n_top = int(top.sum().item())
    if n_top > DENSE_MAX_TOP or n_top < 1:
        return None
    dense = (level >= 2) & (level <= min(max_level, depth))
    n_dense = int(dense.sum().item())
    col = torch.full((n_ref,), -1, dtype=torch.int32, device=dev)
    col[dense] = torch.arange(n_dense, dtype=torch.int32, device=dev)
    col[top] = -2 - torch.arange(n_top, dtype=torch.int32, device=dev)
    kd = max(128, (n_dense + 127) // 128 * 128)
    return dict(col_of_ref=col, n_top=n_top, n_dense=n_dense, kd=kd, max_level=min(max_level, depth))


class DenseTopBuilder:
    """Accumulates the dense top of an all-nodes index batch by batch (``add``), then builds the posting lists of
    the deep levels (``finish``).  ``finish`` returns None when a dense count exceeded 255 (caller falls back to
    plain postings)."""

    def __init__(self, tree_dev: dict, depth: int, n_ref: int, n_img: int, layout: dict = None):
        self.layout = layout or dense_layout(tree_dev, depth, n_ref)
        self.n_ref, self.n_img = n_ref, int(n_img)
        dev = tree_dev["level_ref"].device
        self.np = max(256, (self.n_img + 255) // 256 * 256)
        self.block = torch.zeros((self.np, self.layout["kd"]), dtype=torch.uint8, device=dev)
        self.top = torch.zeros((self.np, self.layout["n_top"]), dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sparse, self.sumsq = [], []

    @staticmethod
    def scatter(csr: dict, layout: dict, block, top, overflow, img_base: int = 0):
        lib = _lib()
        nat.check(lib.vt_dense_scatter(nat.ptr(csr["row_off"]), nat.ptr(csr["node"]), nat.ptr(csr["count"]), csr["n_img"],
                                       int(img_base), nat.ptr(layout["col_of_ref"]), nat.ptr(block), layout["kd"], nat.ptr(top),
                                       layout["n_top"], nat.ptr(overflow), nat.stream_ptr()), "vt_dense_scatter")

    @staticmethod
    def sparse_part(csr: dict, layout: dict) -> dict:
        """the entries of the deep levels (posting nodes) as their own CSR; norms stay those of the full rows"""
        keep = layout["col_of_ref"][csr["node"].long()] == -1
        prefix = torch.zeros(csr["node"].numel() + 1, dtype=torch.int64, device=keep.device)
        prefix[1:] = torch.cumsum(keep, 0)
        row_off = prefix[csr["row_off"]]
        return dict(row_off=row_off, node=csr["node"][keep], count=csr["count"][keep], sumsq=csr["sumsq"], df=None,
                    n_img=csr["n_img"], nnz=int(row_off[-1].item()))

    @_device_scoped
    def add(self, csr: dict, img_base: int):
        self.scatter(csr, self.layout, self.block, self.top, self.overflow, img_base)
        self.sparse.append(self.sparse_part(csr, self.layout))
        return self

    def finish(self):
        if int(self.overflow.item()) != 0:
            return None
        sp = concat_csr(self.sparse) if self.sparse else None
        self.sparse = []
        index = build_index(sp, self.n_ref)
        return dict(index=index, block=self.block, top=self.top, layout=self.layout, np=self.np, n_img=self.n_img)


@_device_scoped
def score_dense(dense: dict, q: dict, topk: int, want_all: bool = False, d_budget_bytes: int = 4 << 30,
                groups: int = None) -> dict:
    """All-nodes tf scoring with the dense top: per chunk of queries ONE tcgen05 INT8 GEMM gives the int32 dots over
    levels 0..Ld (``vt_dense_dots``), the posting scorer starts its accumulators from them and adds the deep levels
    (``vt_score_topk_seeded``).  Same outputs as ``score`` (exact integer dots, binary64 keys)."""
    lib = _lib()
    index, layout = dense["index"], dense["layout"]
    dev = dense["block"].device
    s = nat.stream_ptr()
    nq, n_img, np_, kd, n_top = q["n_img"], dense["n_img"], dense["np"], layout["kd"], layout["n_top"]
    if topk > lib.vt_score_max_topk():
        raise ValueError("top-%d requested; at most %d per query" % (topk, lib.vt_score_max_topk()))
    n_shards = index["n_shards"]
    key = torch.empty((max(nq, 1), topk), dtype=torch.float64, device=dev)
    ids = torch.empty((max(nq, 1), topk), dtype=torch.int32, device=dev)
    sc = torch.empty((max(nq, 1), topk), dtype=torch.float64, device=dev)
    all_key = torch.empty((max(nq, 1), n_img), dtype=torch.float64, device=dev) if want_all else None
    if nq == 0:
        return dict(key=key[:0], id=ids[:0], score=sc[:0], all_key=all_key, topk=topk)
    # query side of the GEMM; a query count above 255 in the dense levels is spread over extra rows (the dot is linear)
    qp = (nq + 255) // 256 * 256
    a_rows = torch.zeros((qp, kd), dtype=torch.uint8, device=dev)
    qt = torch.zeros((qp, n_top), dtype=torch.int32, device=dev)
    over = torch.zeros(1, dtype=torch.int32, device=dev)
    DenseTopBuilder.scatter(q, layout, a_rows, qt, over, 0)
    extra = None
    if int(over.item()) != 0:
        extra = _dense_query_overflow(q, layout, dense["block"])
    qs = DenseTopBuilder.sparse_part(q, layout)
    q_val = qs["count"].to(torch.float32)
    chunk = int(max(256, min(qp, (d_budget_bytes // (4 * np_)) // 256 * 256)))
    seed = torch.empty((chunk, np_), dtype=torch.int32, device=dev)
    for c0 in range(0, nq, chunk):
        c1 = min(nq, c0 + chunk)
        rows = (c1 - c0 + 255) // 256 * 256
        nat.check(lib.vt_dense_dots(nat.ptr(a_rows[c0:]), nat.ptr(qt[c0:]), nat.ptr(dense["block"]), nat.ptr(dense["top"]),
                                    rows, np_, kd, n_top, nat.ptr(seed), np_, s), "vt_dense_dots")
        if extra is not None:
            extra(seed, c0, c1)
        off = (qs["row_off"][c0:c1 + 1] - qs["row_off"][c0]).contiguous()
        b = int(qs["row_off"][c0].item())
        if not want_all:
            # the top-k carried from shard to shard (vt_score_query.cu): a few lists per query instead of one per shard
            n_groups = int(groups or lib.vt_score_group_count(c1 - c0, n_img))
            shard_key = torch.empty((c1 - c0, n_groups, topk), dtype=torch.float64, device=dev)
            shard_id = torch.empty((c1 - c0, n_groups, topk), dtype=torch.int32, device=dev)
            nat.check(lib.vt_score_topk_carried(nat.ptr(index["post_off"]), nat.ptr(index["postings"]), nat.ptr(index["rnorm"]),
                                                n_img, index["n_ref"], nat.ptr(off), nat.ptr(qs["node"][b:]), nat.ptr(q_val[b:]),
                                                c1 - c0, topk, None, None, nat.ptr(seed), np_, n_groups, 1, nat.ptr(shard_key),
                                                nat.ptr(shard_id), s), "vt_score_topk_carried")
            m = merge_topk(shard_key.view(c1 - c0, -1), shard_id.view(c1 - c0, -1), c1 - c0, topk, nat.SCORE_TF_L2,
                           q["sumsq"][c0:c1].contiguous(), index["sumsq"])
            key[c0:c1], ids[c0:c1], sc[c0:c1] = m["key"], m["id"], m["score"]
            continue
        shard_key = torch.empty((c1 - c0, n_shards, topk), dtype=torch.float64, device=dev)
        shard_id = torch.empty((c1 - c0, n_shards, topk), dtype=torch.int32, device=dev)
        nat.check(lib.vt_score_topk_seeded(nat.ptr(index["post_off"]), nat.ptr(index["postings"]), nat.ptr(index["rnorm"]),
                                           n_img, index["n_ref"], nat.ptr(off), nat.ptr(qs["node"][b:]), nat.ptr(q_val[b:]),
                                           c1 - c0, topk, nat.ptr(seed), np_, nat.ptr(shard_key), nat.ptr(shard_id),
                                           nat.ptr(all_key[c0:]) if want_all else None, s), "vt_score_topk_seeded")
        m = merge_topk(shard_key.view(c1 - c0, -1), shard_id.view(c1 - c0, -1), c1 - c0, topk, nat.SCORE_TF_L2,
                       q["sumsq"][c0:c1].contiguous(), index["sumsq"])
        key[c0:c1], ids[c0:c1], sc[c0:c1] = m["key"], m["id"], m["score"]
    return dict(key=key[:nq], id=ids[:nq], score=sc[:nq], all_key=all_key, topk=topk)


def _dense_query_overflow(q: dict, layout: dict, block: torch.Tensor):
    """Rare: a query has more than 255 descriptors under one dense node.  The scattered row holds min(count, 255);
    the remainder is added to the seed column by column (dot products are linear in the query counts)."""
    col = layout["col_of_ref"][q["node"].long()]
    big = torch.nonzero((col >= 0) & (q["count"] > 255)).flatten()
    rows = (torch.searchsorted(q["row_off"], big, right=True) - 1).tolist()
    cols, rest = col[big].tolist(), (q["count"][big] - 255).tolist()

    def add(seed, c0, c1):
        for r, c, v in zip(rows, cols, rest):
            if c0 <= r < c1:
                seed[r - c0] += int(v) * block[:, c].to(torch.int32)
    return add


# --------------------------------------------------------------------------------------- score
@_device_scoped
def score(index: dict, q: dict, topk: int, mode: int = nat.SCORE_TF_L2, q_val: torch.Tensor = None,
          want_all: bool = False, merge: bool = True, weights: dict = None, groups: int = None,
          shards_per_step: int = None) -> dict:
    """Score a CSR batch of queries against the inverted file and select the top-k per query.

    Returns dict(key [Q, topk] f64, id [Q, topk] i32, score [Q, topk] f64, all_key [Q, n_img] | None,
    shard_key/shard_id when ``merge`` is False).  ``groups``: lists per query of the carried-top-k scorer (default: what
    ``vt_score_group_count`` proposes for the batch; 1 = one CTA takes a query through every shard).
    ``shards_per_step``: 1 (32-bit accumulators), 2 or 4 (16-bit accumulators, several shards per step; only valid when
    every dot stays below 65535); default: 2 when the Cauchy-Schwarz bound of the batch allows it."""
    lib = _lib()
    dev = index["post_off"].device
    s = nat.stream_ptr()
    nq = q["n_img"]
    n_img, n_shards = index["n_img"], index["n_shards"]
    if topk > lib.vt_score_max_topk():
        raise ValueError("top-%d requested; the block-wide selection keeps at most %d per query (use retrieve() for "
                         "the full order)" % (topk, lib.vt_score_max_topk()))
    topk = int(topk)
    shard_key = torch.empty((max(nq, 1), n_shards, topk), dtype=torch.float64, device=dev)
    shard_id = torch.empty((max(nq, 1), n_shards, topk), dtype=torch.int32, device=dev)
    all_key = torch.empty((max(nq, 1), n_img), dtype=torch.float64, device=dev) if want_all else None
    if mode != nat.SCORE_TF_L2:
        # tf-idf modes over the integer tf index: binary64 weights on the query side, per-image weighted norms
        if weights is None or "img_rnorm" not in weights:
            raise ValueError("weighted scoring needs weights=dict(idf=..., img_rnorm=..., norm=...)")
        qw = query_weights(q, weights["idf"], weights["norm"])
        if mode == nat.SCORE_W_L2 and not want_all and nq > 0:
            # tf-idf / L2 with the top-k carried from shard to shard (binary64 accumulators, vt_score_query.cu)
            n_groups = int(groups or lib.vt_score_group_count(nq, n_img))
            shard_key = torch.empty((nq, n_groups, topk), dtype=torch.float64, device=dev)
            shard_id = torch.empty((nq, n_groups, topk), dtype=torch.int32, device=dev)
            nat.check(lib.vt_score_topk_weighted_carried(nat.ptr(index["post_off"]), nat.ptr(index["postings"]),
                                                         nat.ptr(weights["img_rnorm"]), n_img, index["n_ref"],
                                                         nat.ptr(q["row_off"]), nat.ptr(q["node"]), nat.ptr(qw["qv"]),
                                                         nat.ptr(qw["rnorm"]), nq, topk, n_groups, nat.ptr(shard_key),
                                                         nat.ptr(shard_id), s), "vt_score_topk_weighted_carried")
            out = dict(all_key=None, shard_key=shard_key, shard_id=shard_id, topk=topk)
            if merge:
                out.update(merge_topk(shard_key.view(nq, -1), shard_id.view(nq, -1), nq, topk, mode, q["sumsq"], index["sumsq"]))
            return out
        nat.check(lib.vt_score_topk_weighted(nat.ptr(index["post_off"]), nat.ptr(index["postings"]),
                                             nat.ptr(weights["img_rnorm"]), n_img, index["n_ref"], nat.ptr(q["row_off"]),
                                             nat.ptr(q["node"]), nat.ptr(qw["qv"]), nat.ptr(qw["qi"]), nat.ptr(qw["rnorm"]),
                                             nq, mode, topk, nat.ptr(shard_key), nat.ptr(shard_id), nat.ptr(all_key), s),
                  "vt_score_topk_weighted")
        out = dict(all_key=all_key, shard_key=shard_key, shard_id=shard_id, topk=topk)
        if merge:
            out.update(merge_topk(shard_key.view(max(nq, 1), -1), shard_id.view(max(nq, 1), -1), nq, topk, mode,
                                  q["sumsq"], index["sumsq"]))
        return out
    if q_val is None:
        q_val = q["count"].to(torch.float32)
    if not want_all and nq > 0 and index.get("shard_rmax") is not None:
        # the top-k carried from shard to shard (vt_score_query.cu): ``shard_key`` / ``shard_id`` are then per GROUP of shards
        n_groups = int(groups or lib.vt_score_group_count(nq, n_img))
        # 16-bit accumulators (two shards per step) when no dot of the batch can reach 65535: |q . d| <= |q| |d|
        if shards_per_step is None:
            bound = float(q["sumsq"].max().item()) * float(index.get("max_sumsq", 1 << 62))
            # ... and when the runs are long enough for the longer two-shard runs to pay (measured: runs of ~8 postings 207 k
            # -> 296 k queries/s; runs of ~1 posting 1.36 M -> 1.24 M, where the one-shard kernel's 4 CTAs per SM win)
            shards_per_step = 2 if bound < 65535.0 ** 2 and n_shards > 1 and index.get("run_mean", 0.0) >= 3.0 else 1
            if shards_per_step > 1 and os.environ.get("VT_SCORE_SHARDS_PER_STEP"):     # (experiments: 1, 2 or 4)
                shards_per_step = int(os.environ["VT_SCORE_SHARDS_PER_STEP"])
        shard_key = torch.empty((nq, n_groups, topk), dtype=torch.float64, device=dev)
        shard_id = torch.empty((nq, n_groups, topk), dtype=torch.int32, device=dev)
        nat.check(lib.vt_score_topk_carried(nat.ptr(index["post_off"]), nat.ptr(index["postings"]), nat.ptr(index["rnorm"]),
                                            n_img, index["n_ref"], nat.ptr(q["row_off"]), nat.ptr(q["node"]), nat.ptr(q_val),
                                            nq, topk, nat.ptr(index["shard_rmax"]), nat.ptr(index.get("shard_rmin")), None, 0,
                                            n_groups, int(shards_per_step),
                                            nat.ptr(shard_key), nat.ptr(shard_id), s), "vt_score_topk_carried")
        out = dict(all_key=None, shard_key=shard_key, shard_id=shard_id, topk=topk)
        if merge:
            out.update(merge_topk(shard_key.view(nq, -1), shard_id.view(nq, -1), nq, topk, mode, q["sumsq"], index["sumsq"]))
        return out
    nat.check(lib.vt_score_topk(nat.ptr(index["post_off"]), nat.ptr(index["postings"]), nat.ptr(index["rnorm"]),
                                n_img, index["n_ref"], nat.ptr(q["row_off"]), nat.ptr(q["node"]), nat.ptr(q_val),
                                nq, mode, topk, nat.ptr(shard_key), nat.ptr(shard_id), nat.ptr(all_key),
                                nat.ptr(index.get("shard_empty")), nat.ptr(index.get("shard_rho")),
                                int(index.get("long_lists", 0)), s),
              "vt_score_topk")
    out = dict(all_key=all_key, shard_key=shard_key, shard_id=shard_id, topk=topk)
    if merge:
        out.update(merge_topk(shard_key.view(max(nq, 1), -1), shard_id.view(max(nq, 1), -1), nq, topk, mode,
                              q["sumsq"], index["sumsq"]))
    return out


@_device_scoped
def merge_topk(cand_key: torch.Tensor, cand_id: torch.Tensor, nq: int, topk: int, mode: int,
               q_sumsq: torch.Tensor, img_sumsq: torch.Tensor) -> dict:
    """Merge ``n_cand`` (key, id) candidates per query into the final top-k + reference scores."""
    lib = _lib()
    dev = cand_key.device
    s = nat.stream_ptr()
    n_cand = cand_key.shape[1]
    q_rnorm = torch.empty(max(nq, 1), dtype=torch.float64, device=dev)
    nat.check(lib.vt_rnorm(nat.ptr(q_sumsq), nq, nat.ptr(q_rnorm), s), "vt_rnorm")
    key = torch.empty((max(nq, 1), topk), dtype=torch.float64, device=dev)
    ids = torch.empty((max(nq, 1), topk), dtype=torch.int32, device=dev)
    sc = torch.empty((max(nq, 1), topk), dtype=torch.float64, device=dev)
    nat.check(lib.vt_topk_merge(nat.ptr(cand_key.contiguous()), nat.ptr(cand_id.contiguous()), nq, n_cand, topk, mode,
                                nat.ptr(q_rnorm), nat.ptr(q_sumsq), nat.ptr(img_sumsq), nat.ptr(key), nat.ptr(ids),
                                nat.ptr(sc), s), "vt_topk_merge")
    return dict(key=key[:nq], id=ids[:nq], score=sc[:nq])
