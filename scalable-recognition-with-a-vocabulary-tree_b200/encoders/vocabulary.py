"""``VocabularyTree`` with the reference's API (cbir/encoders/vocabulary.py) on top of the
sm_100a kernels.

Same constructor and method names as the reference class (vocabulary.py:10-159):
``fit / learn / extract_features / propagate / propagate_feature / embedding / save`` and the
attributes ``n_branches, depth, descriptor, tree, nodes, graph``.  Additions are batched entry
points (``quantize``, ``encode_batch``) and ``load``; new options are keyword-only and default to
the reference's behaviour (all-node visit counts, no idf, L2 normalisation).

Differences that are deliberate and documented in DESIGN.md:
* the per-node k-means is the seeded exact Lloyd (the reference's MiniBatchKMeans is unseeded);
* distances are decided in binary64 on near-ties (the reference's float32 result depends on
  OpenBLAS' summation order there);
* querying does not add the query's counts to the postings (vocabulary.py:126-127 does).
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from .. import utils
from .. import engine
from .. import _native as nat
from ..tree import FlatTree


def _pack_threads() -> int:
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(16, n))


class VocabularyTree(object):
    def __init__(self, n_branches, depth, descriptor, *, n_iter=10, all_nodes=True, min_level=0, device="cuda",
                 comm=None):
        self.n_branches = n_branches
        self.depth = depth
        self.descriptor = descriptor
        # keyword-only additions
        self.n_iter = n_iter
        self.all_nodes = all_nodes          # reference semantics: count every node on the path
        self.min_level = min_level          # level mask: nodes above this level carry weight 0
        self.device = device
        # multi-GPU build (one process per GPU): ``fit`` then takes THIS RANK'S shard of the descriptors (contiguous
        # ranges of the global row order, rank by rank) and every rank ends with the same tree, bit for bit
        self.comm = comm if comm is not None and getattr(comm, "world", 1) > 1 else None
        # private
        self._pinned = {}                   # staging slot -> pinned host buffer (Database.index double buffering)
        self._pinned_free = {}              # staging slot -> CUDA event: the slot's last upload has completed
        self._flat = None                   # FlatTree
        self._propagated = {}               # image sha1 -> (node ids int32, counts int32)
        self._views = {}
        self.build_log = []

    # ------------------------------------------------------------------ reference-shaped views
    @property
    def flat(self) -> FlatTree:
        if self._flat is None:
            raise RuntimeError("the vocabulary tree has not been fitted")
        return self._flat

    @property
    def tree(self) -> dict:
        """node id -> child ids (vocabulary.py:16,69-72)."""
        if self._flat is None:
            return {}
        if "tree" not in self._views:
            self._views["tree"] = self._flat.tree_dict()
        return self._views["tree"]

    @property
    def nodes(self) -> dict:
        """node id -> centre (vocabulary.py:17,51)."""
        if self._flat is None:
            return {}
        if "nodes" not in self._views:
            self._views["nodes"] = self._flat.nodes_dict()
        return self._views["nodes"]

    @property
    def graph(self):
        """networkx view (vocabulary.py:18) including the per-node {image_id: count} postings
        the reference keeps as node attributes (vocabulary.py:96-100).  Small trees only."""
        g = self._flat.to_networkx() if self._flat is not None else __import__("networkx").DiGraph()
        for image_id, (nodes, counts) in self._propagated.items():
            for n_, c in zip(nodes.tolist(), counts.tolist()):
                g.nodes[n_][image_id] = c
        return g

    def postings(self) -> dict:
        """node id -> {image sha1: count}: the inverted file as the reference stores it."""
        out = {}
        for image_id, (nodes, counts) in self._propagated.items():
            for n_, c in zip(nodes.tolist(), counts.tolist()):
                out.setdefault(n_, {})[image_id] = c
        return out

    # ------------------------------------------------------------------ build
    def learn(self, dataset):
        features = self.extract_features(dataset)
        self.fit(features)
        return self

    def extract_features(self, dataset):
        print("Extracting features...")
        func = lambda path: self.descriptor(dataset.read_image(path))  # noqa: E731
        features = utils.show_progress(func, dataset)
        print("\n%d features extracted" % len(features))
        return np.array(features)

    def fit(self, features, node=0, root=None, current_depth=0):
        """Hierarchical k-means over ``features`` [N, D] (vocabulary.py:36-76), all nodes of a
        level per launch.  The reference's recursion arguments are honoured where they make sense for a call from
        outside: ``root`` replaces the stored value of node 0 (vocabulary.py:48-51; the descent never compares with
        it) and ``current_depth`` = c builds the ``depth - c`` levels the reference's recursion would still add below a
        node at that depth (vocabulary.py:55).  ``node != 0`` would graft a subtree into the reference's running
        ``_current_index`` numbering of a half-built tree -- state the level-synchronous build does not have."""
        if node != 0:
            raise NotImplementedError("fit(node != 0): grafting below an existing node needs the reference's recursive "
                                      "numbering state; build the whole tree with node=0")
        levels = self.depth - int(current_depth)
        if current_depth < 0 or levels < 1:
            raise NotImplementedError("fit(current_depth >= depth) leaves a one-node tree; nothing to build")
        features = np.asarray(features) if not hasattr(features, "device") else features
        desc, d = engine.stage(features, self.device, prefer_bytes=True)
        self.build_log = []
        self._flat = engine.build_tree(desc, d, self.n_branches, levels, n_iter=self.n_iter, comm=self.comm,
                                       log=lambda **kw: self.build_log.append(kw))
        if root is not None:
            import torch
            r = torch.as_tensor(np.asarray(root, dtype=np.float32).reshape(-1)[:d])
            cent = self._flat.device(self.device)["centroids"]
            cent[0, :d] = r.to(cent.device)
            if getattr(self._flat, "_centroids", None) is not None:
                self._flat._centroids[0, :d] = r.numpy()
        self._views = {}
        self._propagated = {}
        return self

    def set_tree(self, flat: FlatTree):
        self._flat = flat
        self._views = {}
        self._propagated = {}
        return self

    # ------------------------------------------------------------------ descent
    def quantize(self, features, return_heap=False):
        """Batched ``propagate_feature``: reference node id of the stop node per descriptor."""
        desc, _ = engine.stage(features, self.device)
        t = self.flat
        out = engine.quantize(t.device(self.device), t.k, t.depth, desc, want_heap=return_heap)
        if return_heap:
            return out[0].cpu().numpy(), out[1].cpu().numpy()
        return out.cpu().numpy()

    def propagate_feature(self, feature, node=0):
        """Path of node ids from ``node`` down to a leaf (vocabulary.py:104-124)."""
        t = self.flat
        desc, _ = engine.stage(np.asarray(feature)[None, :], self.device)
        start = int(t.heap_of_ref[node])
        _, heap = engine.quantize(t.device(self.device), t.k, t.depth, desc, want_heap=True, start=start)
        path = t.path_of(int(heap[0].item()))
        return path[path.index(node):]

    # ------------------------------------------------------------------ encoding
    def stage_host(self, images, slot=0):
        """Pack a list of descriptor arrays into one pinned host buffer (``slot`` selects one of the alternating
        staging buffers) -> dict(rows, sizes).  Runs on a loader thread in ``Database.index`` while the device
        works on the previous batch; the upload from pinned memory is then a true asynchronous copy."""
        import torch
        d = self.flat.D
        arrs = [np.asarray(a) for a in images]
        sizes = [int(a.shape[0]) for a in arrs]
        total = int(sum(sizes))
        as_bytes = all(a.dtype == np.uint8 for a in arrs if a.shape[0] > 0)
        dtype = torch.uint8 if as_bytes else torch.float32
        ev = self._pinned_free.get(slot)
        if ev is not None:
            ev.synchronize()                                   # the previous upload from this slot has finished
        buf = self._pinned.get(slot)
        if buf is None or buf.dtype != dtype or buf.shape[0] < total or buf.shape[1] != d:
            cap = max(total, 1) if buf is None else max(total, int(buf.shape[0] * 1.5))
            buf = torch.empty((cap, d), dtype=dtype)
            if torch.cuda.is_available():
                buf = buf.pin_memory()
            self._pinned[slot] = buf
        view = buf.numpy()
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

        def pack(lo, hi):
            for i in range(lo, hi):
                if sizes[i]:
                    np.copyto(view[offs[i]:offs[i + 1]], arrs[i].reshape(sizes[i], d), casting="unsafe")

        want = np.uint8 if as_bytes else np.float32
        plain = all(a.dtype == want and a.flags.c_contiguous and (a.ndim == 2 and a.shape[1] == d or a.shape[0] == 0)
                    for a in arrs)
        if plain and total * d * buf.element_size() >= (8 << 20):
            # one memcpy thread moves ~10 GB/s, the upload that follows 55 GB/s: vt_host_pack spreads the copy over native
            # threads (the call releases the GIL, so a loader thread packing here never stalls the thread driving the GPU)
            import ctypes as C
            lib = nat.load()
            n_arr = len(arrs)
            src = (C.c_void_p * n_arr)(*[a.__array_interface__["data"][0] for a in arrs])
            nbytes = (C.c_int64 * n_arr)(*[a.nbytes for a in arrs])
            nat.check(lib.vt_host_pack(src, nbytes, n_arr, C.c_void_p(buf.data_ptr()), _pack_threads()), "vt_host_pack")
        else:
            pack(0, len(arrs))
        return dict(rows=buf[:total], sizes=sizes, slot=slot)

    def encode_batch(self, images, want_df=False, all_nodes=None, staged=None):
        """Bag-of-words CSR (device tensors) for a list of descriptor arrays (or a ``stage_host`` result)."""
        import torch
        t = self.flat
        if staged is None:
            staged = self.stage_host(images, slot="sync")
        sizes = staged["sizes"]
        off = torch.zeros(len(sizes) + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0)
        dev = torch.device(self.device)
        if staged["rows"].shape[0] > 0:
            desc, _ = engine.stage(staged["rows"], dev)
            if dev.type == "cuda":
                with torch.cuda.device(dev):
                    ev = torch.cuda.Event()
                    ev.record()
                self._pinned_free[staged["slot"]] = ev
            words = engine.quantize(t.device(dev), t.k, t.depth, desc)
        else:
            words = torch.zeros(1, dtype=torch.int32, device=dev)
        return engine.encode(words, off.to(dev), t.device(dev), t.depth, t.n_ref,
                             all_nodes=self.all_nodes if all_nodes is None else all_nodes, want_df=want_df,
                             min_level=min(self.min_level, t.depth))

    def propagate(self, image):
        """Send an image's features down the tree and record its visit counts
        (vocabulary.py:78-102).  Idempotent per image content (sha1 key)."""
        image_id = utils.get_image_id(image)
        if image_id in self._propagated:
            return
        features = self.descriptor(image)
        csr = self.encode_batch([features])
        self._propagated[image_id] = (csr["node"].cpu().numpy(), csr["count"].cpu().numpy())
        return

    def embedding(self, image):
        """Dense visit-count vector over ALL nodes in node-id order, L2 normalised
        (vocabulary.py:126-137).  float64 [n_nodes]; 0/0 -> NaN like the reference."""
        self.propagate(image)
        nodes, counts = self._propagated[utils.get_image_id(image)]
        return self.dense_embedding(nodes, counts)

    encode = embedding      # name used by the notebook skeleton / BASELINE north_star

    def dense_embedding(self, nodes, counts):
        e = np.zeros(self.flat.n_ref, dtype=np.int64)
        e[nodes] = counts
        with np.errstate(invalid="ignore", divide="ignore"):
            return e / np.linalg.norm(e, ord=2)

    # ------------------------------------------------------------------ persistence
    GRAPH_PICKLE_MAX_NODES = 2_000_000

    def save(self, path=None):
        """Flat-array checkpoint ``tree.npz`` plus the reference's two files (vocabulary.py:148-159):
        ``nodes.pickle`` (node id -> centre) and ``graph.pickle`` (the networkx DiGraph with the
        {image sha1: count} postings as node attributes; ``nx.write_gpickle`` was a plain
        ``pickle.dump`` of the graph at the highest protocol, which is what is written here --
        the function itself is gone from networkx 3).  The graph is skipped for huge trees."""
        if path is None:
            path = "data"
        os.makedirs(path, exist_ok=True)
        np.savez_compressed(os.path.join(path, "tree.npz"), **self.flat.state())
        with open(os.path.join(path, "nodes.pickle"), "wb") as f:
            pickle.dump(self.nodes, f)
        if self.flat.n_ref <= self.GRAPH_PICKLE_MAX_NODES:
            with open(os.path.join(path, "graph.pickle"), "wb") as f:
                pickle.dump(self.graph, f, pickle.HIGHEST_PROTOCOL)
        return True

    def load(self, path="data"):
        """Restores a tree: the flat checkpoint if present, else a tree the REFERENCE saved
        (``nodes.pickle`` + ``graph.pickle``): children in edge-insertion order, postings taken
        from the node attributes (vocabulary.py:96-100).  The reference has no ``load``."""
        npz = os.path.join(path, "tree.npz")
        if os.path.exists(npz):
            self.set_tree(FlatTree.from_state(np.load(npz)))
            self.n_branches, self.depth = self.flat.k, self.flat.depth
            return True
        with open(os.path.join(path, "nodes.pickle"), "rb") as f:
            nodes = pickle.load(f)
        with open(os.path.join(path, "graph.pickle"), "rb") as f:
            graph = pickle.load(f)
        self.import_reference(nodes, graph)
        return True

    def import_reference(self, nodes, graph):
        """Adopt a reference-built tree: ``nodes`` dict and networkx ``graph`` (or the ``tree``
        dict node -> children).  Branching factor and depth are read off the structure."""
        if hasattr(graph, "successors"):
            tree = {n_: list(graph.successors(n_)) for n_ in graph.nodes if graph.out_degree(n_) > 0}
        else:
            tree = {n_: list(c) for n_, c in graph.items() if len(c) > 0}
        k = max((len(c) for c in tree.values()), default=self.n_branches)
        depth, frontier = 0, [0]
        while True:
            frontier = [c for n_ in frontier for c in tree.get(n_, [])]
            if not frontier:
                break
            depth += 1
        self.set_tree(FlatTree.from_reference_dicts(nodes, tree, k, max(depth, 1)))
        if self._flat.tree_dict() != tree:
            raise ValueError("node ids are not in the reference's creation order (vocabulary.py:69-72)")
        self.n_branches, self.depth = k, max(depth, 1)
        if hasattr(graph, "nodes"):
            per_image = {}
            for n_, attrs in graph.nodes(data=True):
                for image_id, c in attrs.items():
                    per_image.setdefault(image_id, []).append((n_, c))
            for image_id, pairs in per_image.items():
                pairs.sort()
                self._propagated[image_id] = (np.array([p_[0] for p_ in pairs], np.int32),
                                              np.array([p_[1] for p_ in pairs], np.int32))
        return self

    def subgraph(self, image_id, draw=True):
        """The nodes an image visited, as a networkx subgraph of the tree (vocabulary.py:139-146).
        ``image_id`` is the sha1 key of a propagated image (``utils.get_image_id``).  Like the reference
        the tree is drawn with those nodes highlighted -- when matplotlib is importable and ``draw``."""
        graph = self.graph
        sub = graph.subgraph([n_ for n_, v in graph.nodes(data=image_id, default=None) if v is not None])
        if draw:
            colours = ["C0"] * graph.number_of_nodes()
            for n_ in sub.nodes:
                colours[n_] = "C3"
            try:
                self.draw(node_color=colours)
            except ImportError:
                pass
        return sub

    def draw(self, figsize=None, node_color=None, layout="tree", labels=None):
        """Plots the tree with networkx (vocabulary.py:162-178; plotting-only optional dependencies:
        matplotlib, and pygraphviz for the "tree" / "radial" layouts -- without it networkx's default layout)."""
        import matplotlib.pyplot as plt
        import networkx as nx
        fig = plt.figure(figsize=(30, 10) if figsize is None else figsize)
        graph, pos = self.graph, None
        prog = "dot" if "tree" in layout.lower() else ("twopi" if "radial" in layout.lower() else None)
        if prog is not None:
            try:
                pos = nx.drawing.nx_agraph.graphviz_layout(graph, prog=prog)
            except ImportError:
                pos = None
        nx.draw(graph, pos=pos, with_labels=labels is None, labels=labels, node_color=node_color)
        return fig
