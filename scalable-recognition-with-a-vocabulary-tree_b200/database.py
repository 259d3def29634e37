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
from .dataset import Dataset


class Database(object):
    def __init__(self, dataset, encoder, *, weighting="tf", norm="l2"):
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

        # private:
        self._rows = {}           # image path -> row in the indexed CSR
        self._host = None         # host copy of the forward CSR (row_off, node, count)
        self._csr = None          # device forward CSR of the indexed images
        self._index = None        # device inverted file
        self._weights = None      # tf-idf modes: dict(idf, img_rnorm, norm), binary64 device tensors
        self._extra = {}          # path -> (nodes, counts) of non-indexed images (queries)
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

    def index(self, batch_images=4096):
        """Encode every dataset image (in ``image_paths`` order) and build the inverted file
        (database.py:36-44).  Images are read, staged and encoded ``batch_images`` at a time, so the
        descriptors of the whole dataset are never resident at once; only the bag-of-words CSR
        (8 B per posting) accumulates on the device."""
        import torch
        paths = list(self.dataset.image_paths)
        if self._dense:
            return self._index_dense(paths)
        step = max(1, int(batch_images))
        parts = []
        for b0 in range(0, max(len(paths), 1), step):
            feats = [self._descriptors(p) for p in paths[b0:b0 + step]]
            parts.append(self.encoder.encode_batch(feats, want_df=True))
        csr = engine.concat_csr(parts)
        n_ref = self.encoder.flat.n_ref
        self._csr = csr
        self._index = engine.build_index(csr, n_ref)
        self._set_weights(csr["df"], len(paths))
        self._rows = {p: i for i, p in enumerate(paths)}
        self._host = (csr["row_off"].cpu().numpy(), csr["node"].cpu().numpy(), csr["count"].cpu().numpy())
        self._extra = {}
        return

    def _set_weights(self, df, n_images):
        """tf-idf modes: idf_n = ln(N / df_n) and the per-image weighted norms, binary64, on the device.  The
        inverted file itself stays the integer tf index (csrc/vt_weight.cu)."""
        self._weights = None
        if self.weighting != "tfidf":
            return
        idf = engine.idf_from_df(df, max(int(n_images), 1))
        self._weights = dict(idf=idf, norm=self.norm, img_rnorm=engine.weighted_rnorm(self._csr, idf, self.norm))

    def is_indexed(self, image_path):
        return image_path in self._rows

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
        ||d/|d| - q/|q|||_2 evaluated directly (exactly 0 for identical vectors), NaN -> 1e6."""
        import torch
        out = torch.empty((q_unit.shape[0], self._dense_unit.shape[0]), dtype=torch.float64, device=q_unit.device)
        for i in range(q_unit.shape[0]):
            out[i] = (self._dense_unit - q_unit[i]).norm(dim=1)
        return torch.nan_to_num(out, nan=1e6)

    def _dense_query_rows(self, query_paths):
        import torch
        q = np.stack([np.asarray(self._dense_embedding(p), dtype=np.float64).ravel() for p in query_paths])
        return self._unit_rows(torch.from_numpy(q).to(self._device()))

    def _retrieve_batch_dense(self, query_paths, k, return_all):
        import torch
        if self._dense_unit is None:
            self.index()
        q = self._dense_query_rows(list(query_paths))
        n = self._dense_unit.shape[0]
        k = min(k, n)
        if return_all or n <= 4 * k + 64:
            scores = self._dense_scores(q)
        else:
            # one matrix product ranks everything; the best 4k + 64 are re-scored exactly
            cos = torch.nan_to_num(q, nan=0.0) @ torch.nan_to_num(self._dense_unit, nan=0.0).T
            bad = torch.isnan(self._dense_unit).any(dim=1)
            cos[:, bad] = -2.0
            cand = torch.topk(cos, 4 * k + 64, dim=1).indices.sort(dim=1).values
            scores = torch.full((q.shape[0], n), float("inf"), dtype=torch.float64, device=q.device)
            for i in range(q.shape[0]):
                d = (self._dense_unit[cand[i]] - q[i]).norm(dim=1)
                scores[i, cand[i]] = torch.nan_to_num(d, nan=1e6)
        order = torch.sort(scores, dim=1, stable=True)          # ties keep image_paths order (database.py:79-80)
        ids = order.indices[:, :k].to(torch.int32).cpu().numpy()
        sc = order.values[:, :k].cpu().numpy()
        if return_all:
            return ids, sc, scores.cpu().numpy()
        return ids, sc

    def _sparse(self, image_path):
        if image_path in self._rows:
            ro, node, count = self._host
            i = self._rows[image_path]
            return node[ro[i]:ro[i + 1]], count[ro[i]:ro[i + 1]]
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
    def _query_csr(self, query_paths):
        """CSR batch (device) for a list of image paths; indexed images reuse their rows."""
        import torch
        dev = self._csr["row_off"].device
        rows = [self._sparse(p) for p in query_paths]
        sizes = torch.tensor([len(r[0]) for r in rows], dtype=torch.int64)
        off = torch.zeros(len(rows) + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(sizes, 0)
        node = np.concatenate([r[0] for r in rows]) if rows else np.zeros(0, np.int32)
        count = np.concatenate([r[1] for r in rows]) if rows else np.zeros(0, np.int32)
        sumsq = torch.tensor([int((r[1].astype(np.int64) ** 2).sum()) for r in rows], dtype=torch.int64)
        q = dict(row_off=off.to(dev), node=torch.from_numpy(node.astype(np.int32)).to(dev),
                 count=torch.from_numpy(count.astype(np.int32)).to(dev), sumsq=sumsq.to(dev),
                 n_img=len(rows), nnz=int(off[-1]))
        return q

    def retrieve_batch(self, query_paths, k=10, return_all=False):
        """Top-k image indices (positions in ``dataset.image_paths``) and scores per query.
        Returns (ids [Q, k] int32, scores [Q, k] float64[, all_keys [Q, n_img]])."""
        if self._dense:
            return self._retrieve_batch_dense(query_paths, k, return_all)
        if self._index is None:
            self.index()
        q = self._query_csr(list(query_paths))
        out = engine.score(self._index, q, k, self._mode(), want_all=return_all, weights=self._weights)
        ids, sc = out["id"].cpu().numpy(), out["score"].cpu().numpy()
        if return_all:
            return ids, sc, out["all_key"].cpu().numpy()
        return ids, sc

    def _scores_from_keys(self, keys, q_sumsq):
        """rank keys of every image -> reference scores (host, float64)."""
        mode = self._mode()
        with np.errstate(invalid="ignore", divide="ignore"):
            if mode == nat.SCORE_TF_L2:
                if q_sumsq == 0:
                    return np.full(keys.shape, 1e6)
                s = np.sqrt(np.maximum(0.0, 2.0 - 2.0 * keys / np.sqrt(float(q_sumsq))))
                sumsq = self._index["sumsq"].cpu().numpy()
                dot = keys * np.sqrt(sumsq.astype(np.float64))
                s[(sumsq == q_sumsq) & (np.abs(dot - q_sumsq) < 0.5)] = 0.0
            elif mode == nat.SCORE_W_L2:
                s = np.sqrt(np.maximum(0.0, 2.0 - 2.0 * keys))
            else:
                s = np.maximum(0.0, 2.0 - keys)
        s[keys < 0] = 1e6                                        # images without descriptors: NaN -> 1e6
        return s

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
        if self._index is None:
            self.index()
        q = self._query_csr([query_image_path])
        out = engine.score(self._index, q, 1, self._mode(), want_all=True, weights=self._weights)
        keys = out["all_key"][0].cpu().numpy()
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
        n_ref = self.encoder.flat.n_ref
        # per-image sum of squares from an exact int64 prefix sum (np.add.reduceat raises when trailing images
        # have no descriptors: its last index then equals len(count))
        prefix = np.concatenate([np.zeros(1, np.int64), np.cumsum(count.astype(np.int64) ** 2)])
        sumsq = prefix[ro[1:]] - prefix[ro[:-1]]
        csr = dict(row_off=torch.from_numpy(ro).to(dev), node=torch.from_numpy(node).to(dev),
                   count=torch.from_numpy(count).to(dev), sumsq=torch.from_numpy(sumsq).to(dev),
                   n_img=n_img, nnz=len(node), df=None)
        self._rows, self._host, self._csr = rows, host, csr
        self._index = engine.build_index(csr, n_ref)
        if self.weighting == "tfidf":
            self._set_weights(torch.bincount(csr["node"].long(), minlength=n_ref).to(torch.int32), n_img)
        self._extra = {}

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
            payload = dict(rows=self._rows, host=self._host, weighting=self.weighting, norm=self.norm)
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
