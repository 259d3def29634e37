"""``Database`` with the reference's API (cbir/database.py) on top of the sm_100a inverted-file
scorer.

Same constructor / method names (database.py:9-115): ``index / embedding / score / retrieve /
is_indexed / save / load / show_results``.  ``retrieve`` returns the reference's dict
``image_path -> score`` ordered best first over ALL images (its ``n`` argument is ignored there
too, database.py:72-81).  ``retrieve_batch`` is the batched top-k entry point.

Any object with ``embedding(image)`` works as encoder for the reference (its plugin protocol,
encoders/new_encoder_example.py:1-4).  Encoders that also offer ``encode_batch`` (``VocabularyTree``)
take the inverted-file path; any other encoder (e.g. the reference's AlexNet, encoders/alexnet.py) takes
the dense path: its embeddings are stacked into one float64 matrix on the GPU and scored by brute
force (normalised L2 distance = one matrix product + top-k, then the k best re-scored exactly like
database.py:66-69).  Image identity is the position in ``dataset.image_paths`` (the reference keys by
a salted ``hash(abspath)``, database.py:25-34, which cannot cross processes).
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from . import _native as nat
from . import engine
from . import utils
from .dataset import Dataset


class Database(object):
    def __init__(self, dataset, encoder, *, weighting="tf", norm="l2", comm=None, dense_top=True, record_postings=False):
        # public:
        if isinstance(dataset, Dataset):
            self.dataset = dataset
        elif isinstance(dataset, str):
            self.dataset = Dataset(dataset)
        else:
            raise TypeError("Invalid dataset of type %s" % type(dataset))
        self.encoder = encoder
        if weighting not in ("tf", "tfidf") or norm not in ("l2", "l1"):
            raise ValueError("weighting must be 'tf' or 'tfidf', norm 'l2' or 'l1'")
        if weighting == "tf" and norm != "l2":
            raise ValueError("the reference's tf scoring is L2 only; use weighting='tfidf' for L1")
        self.weighting, self.norm = weighting, norm
        # multi-GPU (one process per GPU): ``comm`` is a ``dist.Comm``; the images are sharded by id over its ranks,
        # every rank indexes its own range and ``retrieve_batch`` merges the per-rank top-k lists
        self.comm = comm if comm is not None and getattr(comm, "world", 1) > 1 else None
        # all-nodes tf indexes (the reference's semantics): the top tree levels are kept as a dense count block scored
        # by one tcgen05 GEMM per query batch, only the deep levels are posting lists (csrc/vt_dense.cu)
        self.dense_top = bool(dense_top)
        self._dense_index = None
        # strict parity with the reference's side effects (opt-in, host work per image): its ``Database.embedding`` calls
        # ``encoder.embedding(image)``, which PROPAGATES the image -- indexed images and queries alike -- into the tree's
        # per-node ``{image sha1: count}`` postings (vocabulary.py:126-127, :96-100).  With ``record_postings`` the same
        # entries appear in ``encoder.postings()`` / ``encoder.graph`` (and in what ``encoder.save`` writes).  Retrieval
        # results never depend on it, here or in the reference.
        self.record_postings = bool(record_postings)

        # private:
        self._rows = {}           # image path -> row in this rank's indexed CSR
        self._n_images = 0        # images of the whole (global) database
        self._img_base = 0        # global index of this rank's first image
        self._row_off_host = None  # host copy of the row pointers only (8 B per image); entries are fetched on demand
        self._csr = None          # device forward CSR of the indexed images
        self._index = None        # device inverted file
        self._weights = None      # tf-idf modes: dict(idf, img_rnorm, norm), binary64 device tensors
        self._extra = {}          # path -> (nodes, counts) of non-indexed images (host, for embedding())
        self._extra_csr = {}      # path -> (device CSR of the batch it was encoded in, row) of non-indexed queries
        # dense path (encoders without encode_batch)
        self._dense = not hasattr(encoder, "encode_batch")
        self._dense_unit = None   # device float64 [N, D]: L2-normalised embeddings (NaN rows for zero vectors)
        self._dense_raw = {}      # path -> embedding as the encoder returned it (memo, database.py:50-57)
        return

    # ------------------------------------------------------------------ indexing
    def _mode(self):
        if self.weighting == "tf":
            return nat.SCORE_TF_L2
        return nat.SCORE_W_L2 if self.norm == "l2" else nat.SCORE_W_L1

    def _descriptors(self, image_path):
        return np.asarray(self.encoder.descriptor(self.dataset.read_image(image_path)))

    def _adopt_dense(self, dense, sumsq, n_images, img_base):
        """Adopt an all-nodes index whose dense top was accumulated batch by batch (``engine.DenseTopBuilder``): the
        forward CSR of the images is not kept, so indexed images cannot be used as queries by name (descriptor
        arrays can).  Used for databases whose all-nodes forward CSR would not fit (bench.py at 1 M images)."""
        self._dense_index, self._index = dense, dense["index"]
        self._csr = dict(row_off=None, node=None, count=None, sumsq=sumsq, df=None, n_img=dense["n_img"], nnz=0)
        self._rows, self._n_images, self._img_base = {}, int(n_images), int(img_base)
        self._row_off_host, self._weights, self._extra = None, None, {}

    def _load_batch(self, paths, slot):
        """Loader thread: read the images of one batch, run the descriptor and pack the rows into a pinned
        staging buffer (two alternating slots), while the device encodes the previous batch."""
        feats = [self._descriptors(p) for p in paths]
        return self.encoder.stage_host(feats, slot)

    def index(self, batch_images=4096):
        """Encode every dataset image (in ``image_paths`` order) and build the inverted file
        (database.py:36-44).  Images are read, staged and encoded ``batch_images`` at a time: a loader thread
        reads and packs batch b+1 into pinned memory while the device quantises and encodes batch b, so the
        descriptors of the whole dataset are never resident at once and host work overlaps device work; only the
        bag-of-words CSR (8 B per posting) accumulates on the device.  With ``comm`` every rank indexes its own
        contiguous image range and the document frequencies are all-reduced (tf-idf)."""
        from concurrent.futures import ThreadPoolExecutor
        from .dist import shard_range
        paths = list(self.dataset.image_paths)
        if self._dense:
            return self._index_dense(paths)
        lo, hi = (0, len(paths)) if self.comm is None else shard_range(len(paths), self.comm.rank, self.comm.world)
        local = paths[lo:hi]
        step = max(1, int(batch_images))
        batches = [local[b0:b0 + step] for b0 in range(0, max(len(local), 1), step)]
        parts = []
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(self._load_batch, batches[0], 0)
            for bi in range(len(batches)):
                staged = pending.result()
                if bi + 1 < len(batches):
                    pending = pool.submit(self._load_batch, batches[bi + 1], (bi + 1) & 1)
                parts.append(self.encoder.encode_batch(None, want_df=True, staged=staged))
        self._adopt(engine.concat_csr(parts), {p: i for i, p in enumerate(local)}, len(paths), lo)
        self._record(local)
        return

    def _record(self, image_paths):
        """``record_postings``: the visit counts of these images under their content hash, as ``VocabularyTree.propagate``
        leaves them (idempotent per image content)."""
        if not self.record_postings or self._dense or not hasattr(self.encoder, "_propagated"):
            return
        for p in image_paths:
            image_id = utils.get_image_id(self.dataset.read_image(p))
            if image_id not in self.encoder._propagated:
                nodes, counts = self._sparse(p)
                self.encoder._propagated[image_id] = (np.asarray(nodes, np.int32), np.asarray(counts, np.int32))

    def _adopt(self, csr, rows, n_images, img_base, df=None):
        """device forward CSR of this rank's images -> inverted file, weights, bookkeeping."""
        n_ref = self.encoder.flat.n_ref
        self._csr = csr
        self._rows, self._n_images, self._img_base = rows, int(n_images), int(img_base)
        self._row_off_host = csr["row_off"].cpu().numpy()
        self._index, self._dense_index = None, None
        enc = self.encoder
        if (self.dense_top and self.weighting == "tf" and csr["n_img"] > 0 and getattr(enc, "all_nodes", False)
                and getattr(enc, "min_level", 0) == 0):
            tdev = enc.flat.device(csr["row_off"].device)
            layout = engine.dense_layout(tdev, enc.flat.depth, n_ref)
            if layout is not None:
                self._dense_index = engine.DenseTopBuilder(tdev, enc.flat.depth, n_ref, csr["n_img"], layout).add(csr, 0).finish()
        if self._dense_index is not None:
            self._index = self._dense_index["index"]          # posting lists of the deep levels only
        elif csr["n_img"] > 0:
            self._index = engine.build_index(csr, n_ref)
        if self.weighting == "tfidf":
            import torch
            if df is None:
                df = csr["df"]
            if df is None:
                df = torch.bincount(csr["node"].long(), minlength=n_ref).to(torch.int32)
            df = df.clone()
            if self.comm is not None:
                self.comm.all_reduce(df)                          # document frequencies over ALL images (8e row 3)
            self._set_weights(df, n_images)
        else:
            self._weights = None
        self._extra = {}

    def _set_weights(self, df, n_images):
        """tf-idf modes: idf_n = ln(N / df_n) and the per-image weighted norms, binary64, on the device.  The
        inverted file itself stays the integer tf index (csrc/vt_weight.cu)."""
        self._weights = None
        if self.weighting != "tfidf":
            return
        idf = engine.idf_from_df(df, max(int(n_images), 1))
        self._weights = dict(idf=idf, norm=self.norm, img_rnorm=engine.weighted_rnorm(self._csr, idf, self.norm))

    def is_indexed(self, image_path):
        if self.comm is None:
            return image_path in self._rows
        return self._n_images > 0 and image_path in set(self.dataset.image_paths)

    # ------------------------------------------------------------------ dense encoders
    def _device(self):
        import torch
        return torch.device(getattr(self.encoder, "device", "cuda"))

    def _dense_embedding(self, image_path):
        if image_path not in self._dense_raw:
            e = self.encoder.embedding(self.dataset.read_image(image_path))
            self._dense_raw[image_path] = np.asarray(e)
        return self._dense_raw[image_path]

    @staticmethod
    def _unit_rows(m):
        """rows / ||row||_2 in float64 on the device; 0/0 -> NaN like numpy (database.py:66-67)."""
        return m / m.norm(dim=1, keepdim=True)

    def _index_dense(self, paths):
        import torch
        if not torch.cuda.is_available():
            raise nat.NativeError("the dense path needs a CUDA device (no CPU fallback)")
        rows = [np.asarray(self._dense_embedding(p), dtype=np.float64).ravel() for p in paths]
        dim = rows[0].shape[0] if rows else 0
        mat = torch.from_numpy(np.stack(rows) if rows else np.zeros((0, dim))).to(self._device())
        self._dense_unit = self._unit_rows(mat)
        self._rows = {p: i for i, p in enumerate(paths)}
        return

    def _dense_scores(self, q_unit):
        """reference score of every indexed image against unit query rows [Q, D] -> [Q, N] float64:
        ||d/|d| - q/|q|||_2 evaluated directly (exactly 0 for identical vectors), NaN -> 1e6 (``vt_dense_scores``)."""
        import torch
        lib = nat.load()
        n, dim = self._dense_unit.shape
        out = torch.empty((q_unit.shape[0], n), dtype=torch.float64, device=q_unit.device)
        with torch.cuda.device(q_unit.device):
            nat.check(lib.vt_dense_scores(nat.ptr(self._dense_unit), n, dim, nat.ptr(q_unit.contiguous()), q_unit.shape[0],
                                          nat.ptr(out), nat.stream_ptr()), "vt_dense_scores")
        return out

    def _dense_query_rows(self, query_paths):
        import torch
        q = np.stack([np.asarray(self._dense_embedding(p), dtype=np.float64).ravel() for p in query_paths])
        return self._unit_rows(torch.from_numpy(q).to(self._device()))

    def _retrieve_batch_dense(self, query_paths, k, return_all):
        """brute force over the dense embeddings: scores by ``vt_dense_scores``, per-shard selection by
        ``vt_topk_rows``, merge by ``vt_topk_merge_scored`` (ascending score, ties keep ``image_paths`` order like the
        reference's stable sort, database.py:79-80)."""
        import torch
        if self._dense_unit is None:
            self.index()
        lib = nat.load()
        q = self._dense_query_rows(list(query_paths))
        n = self._dense_unit.shape[0]
        k = min(k, n)
        if k > lib.vt_score_max_topk():
            raise ValueError("top-%d requested; at most %d per query" % (k, lib.vt_score_max_topk()))
        scores = self._dense_scores(q)
        nq, dev = q.shape[0], q.device
        n_shards = (n + lib.vt_score_shard_size() - 1) // lib.vt_score_shard_size()
        skey = torch.empty((nq, n_shards, k), dtype=torch.float64, device=dev)
        sid = torch.empty((nq, n_shards, k), dtype=torch.int32, device=dev)
        ssc = torch.empty((nq, n_shards, k), dtype=torch.float64, device=dev)
        okey = torch.empty((nq, k), dtype=torch.float64, device=dev)
        oid = torch.empty((nq, k), dtype=torch.int32, device=dev)
        osc = torch.empty((nq, k), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            s = nat.stream_ptr()
            nat.check(lib.vt_topk_rows(nat.ptr(scores), n, nq, k, nat.ptr(skey), nat.ptr(sid), nat.ptr(ssc), s), "vt_topk_rows")
            nat.check(lib.vt_topk_merge_scored(nat.ptr(skey), nat.ptr(sid), nat.ptr(ssc), nq, n_shards * k, k, nat.ptr(okey),
                                               nat.ptr(oid), nat.ptr(osc), s), "vt_topk_merge_scored")
        ids, sc = oid.cpu().numpy(), osc.cpu().numpy()
        if return_all:
            return ids, sc, scores.cpu().numpy()
        return ids, sc

    def _sparse(self, image_path):
        if image_path in self._rows:
            # one image's entries, fetched from the device CSR on demand (no host mirror of the whole index)
            i = self._rows[image_path]
            b, e = int(self._row_off_host[i]), int(self._row_off_host[i + 1])
            return self._csr["node"][b:e].cpu().numpy(), self._csr["count"][b:e].cpu().numpy()
        if image_path not in self._extra:
            csr = self.encoder.encode_batch([self._descriptors(image_path)])
            self._extra[image_path] = (csr["node"].cpu().numpy(), csr["count"].cpu().numpy())
        return self._extra[image_path]

    def embedding(self, image_path):
        """Dense L2-normalised visit-count vector of an image (database.py:50-57)."""
        if self._dense:
            return self._dense_embedding(image_path)
        nodes, counts = self._sparse(image_path)
        return self.encoder.dense_embedding(nodes, counts)

    def score(self, db_image_path, query_image_path):
        """database.py:59-70 for one pair (not the hot path: two dense vectors on the host)."""
        d = self.embedding(db_image_path)
        q = self.embedding(query_image_path)
        with np.errstate(invalid="ignore", divide="ignore"):
            d = d / np.linalg.norm(d, ord=2)
            q = q / np.linalg.norm(q, ord=2)
            score = np.linalg.norm(d - q, ord=2)
        return score if not np.isnan(score) else 1e6

    # ------------------------------------------------------------------ retrieval
    def _query_csr(self, query_paths, prestaged=None):
        """CSR batch (device) for a list of image paths, plus the permutation that puts the rows back into query
        order.  Images indexed on this rank reuse their rows (gathered on the device); all others are encoded in
        ONE batch (pinned staging, like ``index``)."""
        import torch
        dev = torch.device(self.encoder.device)
        known = [(qi, self._rows[p]) for qi, p in enumerate(query_paths) if p in self._rows]
        fresh = [(qi, p) for qi, p in enumerate(query_paths) if p not in self._rows]
        parts, order = [], []
        if known:
            parts.append(engine.csr_rows(self._csr, torch.tensor([r for _, r in known], dtype=torch.int64, device=dev)))
            order += [qi for qi, _ in known]
        if fresh:
            uniq = {}
            for qi, p in fresh:
                uniq.setdefault(p, []).append(qi)
            missing = [p for p in uniq if p not in self._extra_csr]
            if prestaged is not None and prestaged[0]:
                # packed by the loader thread of ``_retrieve_batch_pipelined`` (a superset of what is still missing)
                staged_paths, staged = prestaged
                csr = self.encoder.encode_batch(None, staged=staged)
                for j, p in enumerate(staged_paths):
                    self._extra_csr[p] = (csr, j)
                missing = [p for p in uniq if p not in self._extra_csr]
            if missing:
                staged = self._load_batch(missing, 0)
                csr = self.encoder.encode_batch(None, staged=staged)
                for j, p in enumerate(missing):
                    self._extra_csr[p] = (csr, j)
            # rows of the (possibly several) cached encode batches, in `fresh` order
            by_batch = {}
            for qi, p in fresh:
                csr, j = self._extra_csr[p]
                by_batch.setdefault(id(csr), (csr, []))[1].append((qi, j))
            for csr, pairs in by_batch.values():
                parts.append(engine.csr_rows(csr, torch.tensor([j for _, j in pairs], dtype=torch.int64, device=dev)))
                order += [qi for qi, _ in pairs]
        if not parts:
            z = torch.zeros(1, dtype=torch.int64, device=dev)
            q = dict(row_off=z, node=torch.zeros(0, dtype=torch.int32, device=dev),
                     count=torch.zeros(0, dtype=torch.int32, device=dev), sumsq=torch.zeros(0, dtype=torch.int64, device=dev),
                     df=None, n_img=0, nnz=0)
        else:
            q = engine.concat_csr(parts)
        back = np.empty(len(order), np.int64)
        back[np.asarray(order, np.int64)] = np.arange(len(order))
        return q, back

    def _score_local(self, q, k, return_all):
        """this rank's top-k over its own image shard (ids local); an empty shard yields empty lists"""
        import torch
        if self._index is None:
            dev = q["row_off"].device
            nq = q["n_img"]
            return dict(key=torch.full((nq, k), -2.0, dtype=torch.float64, device=dev),
                        id=torch.full((nq, k), -1, dtype=torch.int32, device=dev),
                        score=torch.full((nq, k), 1e6, dtype=torch.float64, device=dev),
                        all_key=torch.zeros((nq, 0), dtype=torch.float64, device=dev) if return_all else None)
        if self._dense_index is not None:
            return engine.score_dense(self._dense_index, q, k, want_all=return_all)
        return engine.score(self._index, q, k, self._mode(), want_all=return_all, weights=self._weights)

    def _all_keys_global(self, local_keys):
        """[Q, n_local] rank keys of every rank -> [Q, n_images] in global image order"""
        import torch
        if self.comm is None:
            return local_keys
        from .dist import shard_range
        width = max(shard_range(self._n_images, r, self.comm.world)[1] - shard_range(self._n_images, r, self.comm.world)[0]
                    for r in range(self.comm.world))
        pad = torch.full((local_keys.shape[0], width), -2.0, dtype=torch.float64, device=local_keys.device)
        pad[:, :local_keys.shape[1]] = local_keys
        allk = self.comm.all_gather(pad)
        cols = []
        for r in range(self.comm.world):
            lo, hi = shard_range(self._n_images, r, self.comm.world)
            cols.append(allk[r][:, :hi - lo])
        return torch.cat(cols, dim=1)

    def retrieve_batch(self, query_paths, k=10, return_all=False):
        """Top-k image indices (positions in ``dataset.image_paths``) and scores per query.
        Returns (ids [Q, k] int32, scores [Q, k] float64[, all_keys [Q, n_img]]).  With ``comm`` every rank
        passes the same queries and gets the same, global, result."""
        if self._dense:
            return self._retrieve_batch_dense(query_paths, k, return_all)
        if self._csr is None:
            self.index()
        query_paths = list(query_paths)
        if not return_all and len(query_paths) > 2 * self.query_chunk:
            return self._retrieve_batch_pipelined(query_paths, k)
        q, back = self._query_csr(query_paths)
        self._record(query_paths)
        ids_t, sc_t, out = self._score_global(q, k, return_all)
        ids, sc = ids_t.cpu().numpy()[back], sc_t.cpu().numpy()[back]
        if return_all:
            return ids, sc, self._all_keys_global(out["all_key"]).cpu().numpy()[back]
        return ids, sc

    def _score_global(self, q, k, return_all=False):
        out = self._score_local(q, k, return_all)
        if self.comm is None:
            return out["id"], out["score"], out
        from . import dist as vdist
        ids_t, sc_t, _ = vdist.retrieve_sharded(lambda: out, self._img_base, k, self.comm)
        return ids_t, sc_t, out

    query_chunk = 2048      # queries per pipeline stage of a large ``retrieve_batch``

    def _retrieve_batch_pipelined(self, query_paths, k):
        """Large query batches go through in chunks: two loader threads read, describe and pack the images of chunks c+1
        and c+2 into pinned memory (``_load_batch``, three rotating staging slots) while the device uploads, quantises,
        encodes and scores chunk c -- the host packing, not the device, is what a one-shot batch waits for.  Same results,
        chunk by chunk; the ids and scores stay on the device until the end."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        step = int(self.query_chunk)
        chunks = [query_paths[c0:c0 + step] for c0 in range(0, len(query_paths), step)]

        depth = 2                                            # chunks being packed ahead of the one on the device

        def stage(ci):
            missing = self._missing(chunks[ci])
            return missing, (self._load_batch(missing, "q%d" % (ci % (depth + 1))) if missing else None)

        ids_parts, sc_parts = [], []
        with ThreadPoolExecutor(max_workers=depth) as pool:
            pending = [pool.submit(stage, ci) for ci in range(min(depth, len(chunks)))]
            for ci in range(len(chunks)):
                prestaged = pending.pop(0).result()
                if ci + depth < len(chunks):
                    pending.append(pool.submit(stage, ci + depth))
                q, back = self._query_csr(chunks[ci], prestaged=prestaged)
                self._record(chunks[ci])
                ids_t, sc_t, _ = self._score_global(q, k)
                back_t = torch.from_numpy(back).to(ids_t.device)
                ids_parts.append(ids_t[back_t])
                sc_parts.append(sc_t[back_t])
        return torch.cat(ids_parts).cpu().numpy(), torch.cat(sc_parts).cpu().numpy()

    def _missing(self, query_paths):
        """query images that are neither indexed on this rank nor encoded before (in first-seen order)"""
        seen, out = set(), []
        for p in query_paths:
            if p not in self._rows and p not in self._extra_csr and p not in seen:
                seen.add(p)
                out.append(p)
        return out

    def _scores_from_keys(self, keys, q_sumsq):
        """rank keys of every image -> reference scores (host, float64)."""
        mode = self._mode()
        with np.errstate(invalid="ignore", divide="ignore"):
            if mode == nat.SCORE_TF_L2:
                if q_sumsq == 0:
                    return np.full(keys.shape, 1e6)
                s = np.sqrt(np.maximum(0.0, 2.0 - 2.0 * keys / np.sqrt(float(q_sumsq))))
                sumsq = self._sumsq_global()
                dot = keys * np.sqrt(sumsq.astype(np.float64))
                s[(sumsq == q_sumsq) & (np.abs(dot - q_sumsq) < 0.5)] = 0.0
            elif mode == nat.SCORE_W_L2:
                s = np.sqrt(np.maximum(0.0, 2.0 - 2.0 * keys))
            else:
                s = np.maximum(0.0, 2.0 - keys)
        s[keys < 0] = 1e6                                        # images without descriptors: NaN -> 1e6
        return s

    def _sumsq_global(self):
        import torch
        local = self._csr["sumsq"][:self._csr["n_img"]].to(torch.float64)
        if self.comm is None:
            return local.cpu().numpy().astype(np.int64)
        return self._all_keys_global(local[None, :])[0].cpu().numpy().astype(np.int64)

    def retrieve(self, query_image_path, n=4):
        """All dataset images ordered by ascending score, as a dict path -> score
        (database.py:72-81; ties keep ``image_paths`` order like Python's stable sort)."""
        paths = list(self.dataset.image_paths)
        if self._dense:
            if self._dense_unit is None:
                self.index()
            scores = self._dense_scores(self._dense_query_rows([query_image_path]))[0].cpu().numpy()
            order = np.argsort(scores, kind="stable")
            return {paths[i]: float(scores[i]) for i in order}
        if self._csr is None:
            self.index()
        q, _ = self._query_csr([query_image_path])
        self._record([query_image_path])
        out = self._score_local(q, 1, True)
        keys = self._all_keys_global(out["all_key"])[0].cpu().numpy()
        scores = self._scores_from_keys(keys, int(q["sumsq"][0].item()))
        order = np.argsort(scores, kind="stable")
        return {paths[i]: float(scores[i]) for i in order}

    # ------------------------------------------------------------------ persistence
    @staticmethod
    def reference_key(image_path):
        """The reference's image key (database.py:25-34): Python's salted ``hash`` of the absolute
        path.  Only stable inside one process or under a fixed PYTHONHASHSEED."""
        return hash(os.path.abspath(image_path))

    @staticmethod
    def counts_from_embedding(e):
        """Integer visit counts behind one reference embedding ``c / ||c||`` (vocabulary.py:126-137).
        The smallest non-zero count is tried as 1, 2, ...; any common factor leaves every score
        unchanged.  NaN / all-zero embeddings (images without descriptors) give an empty row."""
        e = np.asarray(e, dtype=np.float64)
        nz = np.flatnonzero(np.nan_to_num(e) > 0)
        if nz.size == 0:
            return nz.astype(np.int32), np.zeros(0, np.int32)
        v = e[nz]
        for m in range(1, 65):
            c = v * (m / v.min())
            r = np.rint(c)
            if np.all(np.abs(c - r) <= 1e-6 * np.maximum(r, 1.0)):
                return nz.astype(np.int32), r.astype(np.int32)
        raise ValueError("embedding is not a normalised integer count vector")

    def _install(self, rows, host):
        """Forward CSR on the host -> device CSR + inverted file."""
        import torch
        ro, node, count = host
        dev = torch.device(self.encoder.device)
        n_img = len(ro) - 1
        # per-image sum of squares from an exact int64 prefix sum (np.add.reduceat raises when trailing images
        # have no descriptors: its last index then equals len(count))
        prefix = np.concatenate([np.zeros(1, np.int64), np.cumsum(count.astype(np.int64) ** 2)])
        sumsq = prefix[ro[1:]] - prefix[ro[:-1]]
        csr = dict(row_off=torch.from_numpy(np.ascontiguousarray(ro, np.int64)).to(dev),
                   node=torch.from_numpy(np.ascontiguousarray(node, np.int32)).to(dev),
                   count=torch.from_numpy(np.ascontiguousarray(count, np.int32)).to(dev),
                   sumsq=torch.from_numpy(sumsq).to(dev), n_img=n_img, nnz=len(node), df=None)
        self._adopt(csr, rows, n_img, 0)

    def _host_csr(self):
        """host copy of the forward CSR (made when saving; not kept)"""
        c = self._csr
        return (c["row_off"].cpu().numpy(), c["node"].cpu().numpy(), c["count"].cpu().numpy())

    def save(self, path=None, reference_format=False, key=None):
        """``index.pickle`` (database.py:83-91).  Default: the forward CSR of the indexed images.
        ``reference_format=True`` writes what the reference writes -- its ``_database`` dict
        ``key(image_path) -> dense embedding`` -- for trees small enough to hold dense rows."""
        if path is None:
            path = "data"
        os.makedirs(path, exist_ok=True)
        if reference_format:
            key = key or self.reference_key
            n_ref = self.encoder.flat.n_ref
            if len(self._rows) * n_ref * 8 > (1 << 31):
                raise ValueError("dense reference-format index would exceed 2 GiB; use the CSR format")
            payload = {key(p): self.embedding(p) for p in self._rows}
        else:
            if self.comm is not None:
                raise NotImplementedError("save() of a sharded database: save per rank with comm=None databases")
            payload = dict(rows=self._rows, host=self._host_csr(), weighting=self.weighting, norm=self.norm)
        with open(os.path.join(path, "index.pickle"), "wb") as f:
            pickle.dump(payload, f)
        return True

    def load(self, path="data", key=None):
        """Loads ``index.pickle`` in either format.  A reference-written file (dict of dense
        embeddings keyed by ``hash(abspath)``, database.py:83-101) is matched to
        ``dataset.image_paths`` through ``key`` (default: the reference's own key, which only
        matches under the PYTHONHASHSEED the file was written with) and converted back to integer
        counts.  Failures print the reference's message and leave the database unchanged."""
        try:
            with open(os.path.join(path, "index.pickle"), "rb") as f:
                st = pickle.load(f)
            if isinstance(st, dict) and "host" in st and "rows" in st:
                self.weighting, self.norm = st["weighting"], st["norm"]
                self._install(st["rows"], st["host"])
                return True
            key = key or self.reference_key
            paths = list(self.dataset.image_paths)
            nodes, counts, ro = [], [], np.zeros(len(paths) + 1, np.int64)
            for i, p in enumerate(paths):
                nd, ct = self.counts_from_embedding(st[key(p)])
                nodes.append(nd)
                counts.append(ct)
                ro[i + 1] = ro[i] + len(nd)
            host = (ro, np.concatenate(nodes).astype(np.int32) if nodes else np.zeros(0, np.int32),
                    np.concatenate(counts).astype(np.int32) if counts else np.zeros(0, np.int32))
            self._install({p: i for i, p in enumerate(paths)}, host)
        except (OSError, KeyError, ValueError, pickle.UnpicklingError):
            print("Cannot load index file from %s/index.pickle" % path)
        return True

    def show_results(self, query_path, scores_dict, n=4, figsize=(10, 4)):
        """Plot the query next to its n best matches; entry 0 of ``scores_dict`` is the query
        itself when it is a database image, so matches start at entry 1 (database.py:103-115)."""
        import matplotlib.pyplot as plt          # optional dependency, plotting only
        ranked = list(scores_dict.items())[1:n + 1]
        panels = [("Query image", query_path)] + [
            ("#%d. %s Score:%.3f" % (r + 1, p, s), p) for r, (p, s) in enumerate(ranked)]
        _, axes = plt.subplots(1, len(panels), figsize=figsize)
        for axis, (title, path) in zip(np.atleast_1d(axes), panels):
            axis.imshow(self.dataset.read_image(path))
            axis.set_title(title)
            axis.axis("off")
        return
