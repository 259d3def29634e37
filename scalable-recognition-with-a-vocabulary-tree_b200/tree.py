"""Flat vocabulary tree: the array form of the reference's ``VocabularyTree.tree`` /
``.nodes`` / ``.graph`` (cbir/encoders/vocabulary.py:16-18).

Layout.  The tree is an implicit complete k-ary heap: node ``h`` has children
``h*k+1 .. h*k+k``; level ``l`` starts at ``(k**l - 1)/(k - 1)``.  Per heap slot we keep the
centroid (float32, zero padded to ``DP``), ``exists`` (the reference created this node),
``internal`` (it has children; the reference's ``graph.out_degree(node) > 0``,
vocabulary.py:113) and ``ref_id`` -- the id the reference assigns, i.e. the DFS pre-order
counter of ``fit`` (vocabulary.py:70-75).  Ragged trees (a node with fewer than ``k`` members is
a leaf at any depth, vocabulary.py:55) simply leave heap slots unused.

Only host-side index arithmetic lives here (numpy); device copies are plain torch tensors.
"""
from __future__ import annotations

import numpy as np


def pad_dim(d: int) -> int:
    dp = 32
    while dp < d:
        dp *= 2
    if dp > 128:
        raise ValueError("descriptor dimension %d is not supported (max 128)" % d)
    return dp


def heap_size(k: int, depth: int) -> int:
    return sum(k ** l for l in range(depth + 1))


def level_offset(k: int, level: int) -> int:
    return sum(k ** l for l in range(level))


def reference_ids(exists: np.ndarray, internal: np.ndarray, k: int, depth: int):
    """DFS pre-order numbering of the existing nodes (vocabulary.py:70-75), vectorised per level.

    id(child c of n) = id(n) + 1 + sum of the subtree sizes of the children before c.
    Returns (ref_id [n_heap] int32 with -1 for unused slots, n_ref).  32-bit arithmetic in place (ids are
    int32 across the ABI): this runs on the host at the end of every build."""
    n_heap = exists.shape[0]
    size = exists.astype(np.int32)                            # subtree sizes, bottom-up
    inter = internal != 0
    for level in range(depth - 1, -1, -1):
        o, n_l = level_offset(k, level), k ** level
        co = level_offset(k, level + 1)
        kids = size[co: co + n_l * k].reshape(n_l, k).sum(axis=1, dtype=np.int32)
        np.add(size[o: o + n_l], kids, out=size[o: o + n_l], where=inter[o: o + n_l])
    ref = np.full(n_heap, -1, np.int32)
    if n_heap and exists[0]:
        ref[0] = 0
    for level in range(depth):
        o, n_l = level_offset(k, level), k ** level
        co = level_offset(k, level + 1)
        child = size[co: co + n_l * k].reshape(n_l, k)
        cid = np.cumsum(child, axis=1, dtype=np.int32)
        cid -= child                                          # exclusive prefix over siblings
        cid += (ref[o: o + n_l] + 1)[:, None]
        ok = inter[o: o + n_l, None] & (exists[co: co + n_l * k].reshape(n_l, k) != 0)
        cid[~ok] = -1
        ref[co: co + n_l * k] = cid.reshape(-1)
    return ref, int(size[0]) if n_heap else 0


def reference_tables_torch(exists, internal, k: int, depth: int):
    """``reference_ids`` + the per-id tables (heap slot, parent id, level) with torch ops, on whatever device the
    inputs live -- the build keeps them on the GPU, so the numbering of a 1.1 M-node tree costs microseconds there
    instead of tens of milliseconds of numpy on the host.  Same arithmetic as the numpy versions (tested equal)."""
    import torch
    dev = exists.device
    n_heap = exists.shape[0]
    size = exists.to(torch.int32).clone()
    inter = internal != 0
    exb = exists != 0
    zero = torch.zeros((), dtype=torch.int32, device=dev)
    for level in range(depth - 1, -1, -1):
        o, n_l, co = level_offset(k, level), k ** level, level_offset(k, level + 1)
        kids = size[co: co + n_l * k].view(n_l, k).sum(1, dtype=torch.int32)
        size[o: o + n_l] += torch.where(inter[o: o + n_l], kids, zero)
    ref = torch.full((n_heap,), -1, dtype=torch.int32, device=dev)
    ref[0] = torch.where(exb[0], 0, -1).to(torch.int32)
    parent_heap = torch.full((n_heap,), -1, dtype=torch.int32, device=dev)
    level_heap = torch.zeros(n_heap, dtype=torch.uint8, device=dev)
    minus1 = torch.full((), -1, dtype=torch.int32, device=dev)
    for level in range(depth):
        o, n_l, co = level_offset(k, level), k ** level, level_offset(k, level + 1)
        child = size[co: co + n_l * k].view(n_l, k)
        cid = torch.cumsum(child, 1, dtype=torch.int32) - child + (ref[o: o + n_l] + 1)[:, None]
        ok = inter[o: o + n_l, None] & exb[co: co + n_l * k].view(n_l, k)
        ref[co: co + n_l * k] = torch.where(ok, cid, minus1).reshape(-1)
        parent_heap[co: co + n_l * k] = ref[o: o + n_l].repeat_interleave(k)
        level_heap[co: co + n_l * k] = level + 1
    n_ref = int(size[0].item()) if n_heap else 0
    ex = torch.nonzero(exb).flatten()
    ids = ref[ex].long()
    heap_of_ref = torch.full((n_ref,), -1, dtype=torch.int64, device=dev)
    heap_of_ref[ids] = ex
    parent_ref = torch.empty(n_ref, dtype=torch.int32, device=dev)
    parent_ref[ids] = parent_heap[ex]
    level_ref = torch.empty(n_ref, dtype=torch.uint8, device=dev)
    level_ref[ids] = level_heap[ex]
    return ref, n_ref, heap_of_ref, parent_ref, level_ref


class FlatTree:
    """Host mirror of a built tree plus lazily created device tensors."""

    def __init__(self, k: int, depth: int, d: int, centroids, exists: np.ndarray,
                 internal: np.ndarray, count=None, byte_built: bool = False):
        """``centroids``: float32 [n_heap, DP] as a numpy array, or a torch tensor that stays on its
        device (the host copy is then made lazily, on first access of ``.centroids``)."""
        self.k, self.depth, self.D = int(k), int(depth), int(d)
        self.DP = int(centroids.shape[1])
        self.n_heap = heap_size(k, depth)
        assert centroids.shape[0] == self.n_heap
        self._byte_built = bool(byte_built)       # built from byte descriptors => every centre in [0, 255]
        if isinstance(centroids, np.ndarray):
            self._centroids = np.ascontiguousarray(centroids, np.float32)
            self._centroids_dev = None
        else:
            self._centroids = None
            self._centroids_dev = centroids
        self.exists = np.ascontiguousarray(exists, np.uint8)
        self.internal = np.ascontiguousarray(internal, np.uint8)
        self.count = None if count is None else np.ascontiguousarray(count, np.int64)
        self.ref_id, self.n_ref = reference_ids(self.exists, self.internal, self.k, self.depth)
        # per reference id: heap slot, parent id and depth (the bag-of-words encoder lifts counts with them).
        # Built level by level in heap space -- the parent of every child of a level is the parent's id repeated
        # k times -- and scattered through ref_id once.
        parent_heap = np.full(self.n_heap, -1, np.int32)
        level_heap = np.zeros(self.n_heap, np.uint8)
        for level in range(self.depth):
            o, n_l = level_offset(self.k, level), self.k ** level
            co = level_offset(self.k, level + 1)
            parent_heap[co: co + n_l * self.k] = np.repeat(self.ref_id[o: o + n_l], self.k)
            level_heap[co: co + n_l * self.k] = level + 1
        ex = np.flatnonzero(self.exists)
        ids = self.ref_id[ex]
        self.heap_of_ref = np.full(self.n_ref, -1, np.int64)
        self.heap_of_ref[ids] = ex
        self.parent_ref = np.empty(self.n_ref, np.int32)
        self.parent_ref[ids] = parent_heap[ex]
        self.level_ref = np.empty(self.n_ref, np.uint8)
        self.level_ref[ids] = level_heap[ex]
        self._dev = {}

    # host arrays that a device-built tree materialises on first use (FlatTree.from_device)
    _LAZY = ("exists", "internal", "count", "ref_id", "heap_of_ref", "parent_ref", "level_ref")

    def __getattr__(self, name):                      # only reached when the attribute is not set yet
        lazy = self.__dict__.get("_lazy")
        if lazy is not None and name in lazy:
            value = lazy[name].detach().cpu().numpy()
            self.__dict__[name] = value
            return value
        raise AttributeError(name)

    @classmethod
    def from_device(cls, k: int, depth: int, d: int, centroids, exists, internal, count=None,
                    byte_built: bool = False) -> "FlatTree":
        """Tree whose arrays are torch tensors on one device (the end of ``engine.build_tree``): the reference
        numbering and the per-id tables are computed there, the device dict is ready at once, and the numpy
        mirrors (``exists``, ``ref_id``, ``heap_of_ref`` ...) are copied to the host only when first read."""
        self = cls.__new__(cls)
        self.k, self.depth, self.D = int(k), int(depth), int(d)
        self.DP = int(centroids.shape[1])
        self.n_heap = heap_size(k, depth)
        assert centroids.shape[0] == self.n_heap
        self._byte_built = bool(byte_built)
        self._centroids, self._centroids_dev = None, centroids
        ref, n_ref, heap_of_ref, parent_ref, level_ref = reference_tables_torch(exists, internal, self.k, self.depth)
        self.n_ref = n_ref
        self._lazy = dict(exists=exists, internal=internal, ref_id=ref, heap_of_ref=heap_of_ref,
                          parent_ref=parent_ref, level_ref=level_ref)
        if count is None:
            self.count = None
        else:
            self._lazy["count"] = count
        self._dev = {str(centroids.device): dict(mma_ok=self.mma_ok(), centroids=centroids, internal=internal,
                                                 ref_id=ref, parent_ref=parent_ref, level_ref=level_ref)}
        return self

    @property
    def centroids(self) -> np.ndarray:
        if self._centroids is None:
            self._centroids = self._centroids_dev.detach().cpu().numpy()
        return self._centroids

    # ------------------------------------------------------------------ device copies
    def device(self, device=None):
        """dict of torch tensors on ``device`` (cached): centroids, internal, ref_id,
        parent_ref, level_ref."""
        import torch
        device = torch.device(device if device is not None else "cuda")
        if device.type == "cuda" and device.index is None:     # "cuda" and "cuda:<current>" are the same cache entry
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        if key not in self._dev:
            self._dev[key] = dict(
                mma_ok=self.mma_ok(),
                centroids=torch.from_numpy(self.centroids).to(device),
                internal=torch.from_numpy(self.internal).to(device),
                ref_id=torch.from_numpy(self.ref_id).to(device),
                parent_ref=torch.from_numpy(self.parent_ref).to(device),
                level_ref=torch.from_numpy(self.level_ref).to(device),
            )
        return self._dev[key]

    def adopt_device(self, device, centroids, internal):
        """Reuse tensors the build already holds on the device instead of re-uploading."""
        import torch
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        self._dev[key] = dict(
            mma_ok=self.mma_ok(),
            centroids=centroids, internal=internal,
            ref_id=torch.from_numpy(self.ref_id).to(device),
            parent_ref=torch.from_numpy(self.parent_ref).to(device),
            level_ref=torch.from_numpy(self.level_ref).to(device),
        )

    def mma_ok(self) -> bool:
        """The tensor-core descent (vt_quantize_mma) needs 128-byte rows, k <= 16 and every centroid
        inside the 8.16 fixed-point range -- true for any tree built from byte descriptors."""
        if self.DP != 128 or self.k > 16 or self.depth < 1:
            return False
        if self._byte_built:
            return True
        c = self.centroids
        return bool(c.size == 0 or (float(c.min()) >= 0.0 and float(c.max()) <= 255.99))

    # ------------------------------------------------------------------ reference views
    def nodes_dict(self) -> dict:
        """``VocabularyTree.nodes``: reference id -> centre (float32 [D]) (vocabulary.py:51)."""
        return {int(r): self.centroids[h, :self.D].copy() for r, h in enumerate(self.heap_of_ref)}

    def tree_dict(self) -> dict:
        """``VocabularyTree.tree``: reference id -> list of child ids (vocabulary.py:69-72)."""
        out = {}
        for h in np.nonzero(self.internal)[0]:
            out[int(self.ref_id[h])] = [int(x) for x in self.ref_id[h * self.k + 1: h * self.k + 1 + self.k]]
        return dict(sorted(out.items()))

    def to_networkx(self):
        """``VocabularyTree.graph`` (vocabulary.py:18,52,73) for small trees / visualisation."""
        import networkx as nx
        g = nx.DiGraph()
        g.add_nodes_from(range(self.n_ref))
        for node, kids in self.tree_dict().items():
            for c in kids:
                g.add_edge(node, c)
        return g

    def path_of(self, heap_leaf: int) -> list:
        """reference-id path root..stop node, what ``propagate_feature`` returns (vocabulary.py:111-124)."""
        path = []
        h = int(heap_leaf)
        while True:
            path.append(int(self.ref_id[h]))
            if h == 0:
                break
            h = (h - 1) // self.k
        return path[::-1]

    # ------------------------------------------------------------------ persistence
    def state(self) -> dict:
        return dict(k=self.k, depth=self.depth, D=self.D, centroids=self.centroids, exists=self.exists,
                    internal=self.internal, count=self.count if self.count is not None else np.zeros(0, np.int64))

    @classmethod
    def from_state(cls, st) -> "FlatTree":
        count = st["count"] if len(st["count"]) else None
        return cls(int(st["k"]), int(st["depth"]), int(st["D"]), st["centroids"], st["exists"], st["internal"], count)

    @classmethod
    def from_reference_dicts(cls, nodes: dict, tree: dict, k: int, depth: int) -> "FlatTree":
        """Import a tree built by the reference (its ``nodes`` / ``tree`` dicts, e.g. from
        ``nodes.pickle``, vocabulary.py:148-159)."""
        d = int(np.asarray(nodes[0]).shape[0])
        dp = pad_dim(d)
        nh = heap_size(k, depth)
        cent = np.zeros((nh, dp), np.float32)
        exists = np.zeros(nh, np.uint8)
        internal = np.zeros(nh, np.uint8)
        stack = [(0, 0)]
        while stack:
            rid, h = stack.pop()
            exists[h] = 1
            cent[h, :d] = np.asarray(nodes[rid], np.float32)
            kids = tree.get(rid, [])
            if kids:
                internal[h] = 1
                for j, c in enumerate(kids):
                    stack.append((c, h * k + 1 + j))
        return cls(k, depth, d, cent, exists, internal)
