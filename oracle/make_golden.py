"""oracle/make_golden.py -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden [case ...]        # default: every small case
    python -m oracle.make_golden c1                # the full BASELINE config-1 shape (minutes)

Every fixture holds outputs of the reference's own code (``VocabularyTree.fit`` with the seeded
Lloyd patched over MiniBatchKMeans, ``propagate_feature``, ``Database.index`` /
``embedding`` / ``retrieve``) on inputs that ``synth`` regenerates from the stored seeds, plus a
sha256 of the inputs so a drifting generator is detected instead of silently mis-compared.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

synth = importlib.import_module("scalable-recognition-with-a-vocabulary-tree_b200.synth")
GOLDEN = os.path.join(ROOT, "tests", "golden")


# --------------------------------------------------------------------------- case inputs
def case_images(name: str):
    """Returns (list of uint8 [n_i, D] arrays, k, depth, n_iter).  Shared with tests/."""
    if name == "toy":          # notebook cell 54
        return [np.array([[1], [2], [11], [12]], dtype=np.uint8)], 2, 2, 10
    if name == "small":        # config-1 shape, shrunk; ragged tree (early leaves) on purpose
        return synth.image_set(120, 150, d=128, n_centres=300, centres_per_image=12, seed=101,
                               ragged=True), 10, 3, 10
    if name == "orb":          # the literal notebook shape: n_branches=4, depth=4, ORB 32-D
        return synth.image_set(80, 150, d=32, n_centres=200, centres_per_image=10, seed=202,
                               scale=60.0, noise=20.0), 4, 4, 10
    if name == "adversarial":  # duplicates, exact ties, an empty image, identical images
        imgs = synth.image_set(24, 60, d=32, n_centres=40, centres_per_image=6, seed=303,
                               scale=40.0, noise=3.0)
        imgs[5] = imgs[2].copy()                              # identical image -> same sha1
        imgs[7] = np.zeros((0, 32), np.uint8)                 # zero-descriptor image -> NaN -> 1e6
        imgs[9] = np.repeat(imgs[9][:6], 10, axis=0)          # heavy duplication inside an image
        imgs[11][:] = imgs[11][0]                             # one descriptor repeated 60 times
        return imgs, 3, 4, 10
    if name == "c1":           # BASELINE.json configs[0]: 1,491 images x 1,000 SIFT-128, k=10 L=3
        return synth.image_set(1491, 1000, d=128, n_centres=1000, centres_per_image=24,
                               seed=1234), 10, 3, 10
    raise KeyError(name)


def inputs_digest(images) -> str:
    h = hashlib.sha256()
    for a in images:
        h.update(np.ascontiguousarray(a).tobytes())
        h.update(str(a.shape).encode())
    return h.hexdigest()


# --------------------------------------------------------------------------- generation
def run_case(name: str, n_query=None, store_full=True):
    cbir = ref_harness.import_reference()
    SynthDataset = ref_harness.make_dataset_class(cbir)
    images_u8, k, depth, n_iter = case_images(name)
    d = images_u8[0].shape[1]
    images = [np.ascontiguousarray(a, dtype=np.float32) for a in images_u8]
    dataset = SynthDataset(images)
    t0 = time.time()
    features = np.concatenate(images, axis=0)                 # == extract_features() output
    voc = ref_harness.reference_build(cbir, features, k, depth, n_iter)
    t_fit = time.time() - t0

    n_ref = len(voc.nodes)
    ref_nodes = np.stack([np.asarray(voc.nodes[i], dtype=np.float32) for i in range(n_ref)])
    ref_children = np.full((n_ref, k), -1, np.int32)
    for node, kids in voc.tree.items():
        ref_children[node, :len(kids)] = kids

    # literal descent (float32 numpy linalg.norm) for every descriptor
    t0 = time.time()
    ref_leaf = np.empty(len(features), np.int32)
    ref_depth = np.empty(len(features), np.int8)
    for i in range(len(features)):
        p = voc.propagate_feature(features[i])
        ref_leaf[i] = p[-1]
        ref_depth[i] = len(p) - 1
    t_desc = time.time() - t0

    # index + postings
    t0 = time.time()
    db = cbir.Database(dataset, voc)
    with ref_harness.quiet():
        db.index()
    t_index = time.time() - t0
    n_img = len(images)
    sha = [cbir.utils.get_image_id(a) for a in images]
    first_of_sha = {}
    for i, s in enumerate(sha):
        first_of_sha.setdefault(s, i)
    counts = np.zeros((n_img, n_ref), np.int32)
    for node in range(n_ref):
        for s, c in voc.graph.nodes[node].items():
            counts[first_of_sha[s], node] = c
    for i, s in enumerate(sha):                               # identical images share postings
        counts[i] = counts[first_of_sha[s]]
    emb = np.stack([db.embedding(p) for p in dataset.image_paths])

    # retrieve
    t0 = time.time()
    q_idx = np.arange(n_img) if n_query is None else np.linspace(0, n_img - 1, n_query).astype(int)
    order = np.empty((len(q_idx), n_img), np.int32)
    scores = np.empty((len(q_idx), n_img), np.float64)
    pos = {p: i for i, p in enumerate(dataset.image_paths)}
    for qi, q in enumerate(q_idx):
        res = db.retrieve(dataset.image_paths[q])
        order[qi] = [pos[p] for p in res.keys()]
        scores[qi] = list(res.values())
    t_ret = time.time() - t0

    meta = dict(case=name, k=k, depth=depth, n_iter=n_iter, D=d, n_img=n_img, n_ref=n_ref,
                n_desc=int(len(features)), inputs_sha256=inputs_digest(images_u8),
                numpy=np.__version__, t_fit=t_fit, t_descent=t_desc, t_index=t_index,
                t_retrieve=t_ret, n_query=int(len(q_idx)), cores=os.cpu_count())
    out = dict(meta=json.dumps(meta), ref_nodes=ref_nodes, ref_children=ref_children,
               q_idx=q_idx.astype(np.int32))
    if store_full:
        out.update(ref_leaf=ref_leaf, ref_depth=ref_depth, counts=counts, emb=emb, order=order,
                   scores=scores)
    else:   # big case: hashes + heads, top-32 only
        out.update(ref_leaf_sha256=hashlib.sha256(ref_leaf.tobytes()).hexdigest(),
                   ref_leaf=ref_leaf.astype(np.int16) if n_ref < 32768 else ref_leaf,
                   counts_sha256=hashlib.sha256(counts.tobytes()).hexdigest(),
                   counts_head=counts[:16], emb_head=emb[:16],
                   order=order[:, :32].astype(np.int16), scores=scores[:, :32])
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, "%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%s: n_ref=%d fit=%.1fs descent=%.1fs index=%.1fs retrieve=%.1fs -> %s (%.1f KB)" % (
        name, n_ref, t_fit, t_desc, t_index, t_ret, path, os.path.getsize(path) / 1024))


def run_persist(name="adversarial"):
    """tests/golden/persist/: the files the reference's own ``VocabularyTree.save`` and
    ``Database.save`` write (vocabulary.py:148-159, database.py:83-91) for one small case.
    ``nx.write_gpickle`` no longer exists in the installed networkx 3; it is restored for the call
    with its networkx-2 body (``pickle.dump(G, f, HIGHEST_PROTOCOL)``).  ``keys.json`` records the
    salted ``hash(abspath)`` key the reference gave each image in the generating process."""
    import pickle
    import networkx as nx
    cbir = ref_harness.import_reference()
    SynthDataset = ref_harness.make_dataset_class(cbir)
    images_u8, k, depth, n_iter = case_images(name)
    images = [np.ascontiguousarray(a, dtype=np.float32) for a in images_u8]
    dataset = SynthDataset(images)
    voc = ref_harness.reference_build(cbir, np.concatenate(images, axis=0), k, depth, n_iter)
    db = cbir.Database(dataset, voc)
    with ref_harness.quiet():
        db.index()
    out = os.path.join(GOLDEN, "persist")
    os.makedirs(out, exist_ok=True)
    if not hasattr(nx, "write_gpickle"):
        def write_gpickle(G, path, protocol=pickle.HIGHEST_PROTOCOL):
            with open(path, "wb") as f:
                pickle.dump(G, f, protocol)
        nx.write_gpickle = write_gpickle
    voc.save(out)
    db.save(out)
    keys = {p: db.get_image_id(p) for p in dataset.image_paths}
    with open(os.path.join(out, "keys.json"), "w") as f:
        json.dump(dict(case=name, k=k, depth=depth, keys=keys), f)
    print("persist: %s -> %s (%s)" % (name, out, ", ".join(
        "%s %.1f KB" % (fn, os.path.getsize(os.path.join(out, fn)) / 1024) for fn in sorted(os.listdir(out)))))


def main(argv):
    cases = argv or ["toy", "small", "orb", "adversarial", "persist"]
    for c in cases:
        if c == "persist":
            run_persist()
        elif c == "c1":
            run_case("c1", n_query=64, store_full=False)
        else:
            run_case(c)


if __name__ == "__main__":
    main(sys.argv[1:])
