"""Times the tcgen05 descent (stage profile of vt_quantize_mma_profile) on synthetic data; the environment switches
(VT_MMA_PAIR_MIN, VT_PAIR_CTAS, ...) are read once per process, so run one configuration per process.

    python tools/descent_probe.py [rows]        # default 32 M rows, k=10, L=6
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vtree_b200 as vt  # noqa: E402
from vtree_b200 import _native as nat  # noqa: E402

K, L, D = 10, 6, 128
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
dev = torch.device("cuda:0")
lib = nat.load()
cent = vt.synth.hierarchical_tree(K, L, D, seed=7, device=dev)
nh = vt.tree.heap_size(K, L)
internal = np.zeros(nh, np.uint8)
internal[:vt.tree.level_offset(K, L)] = 1
ref_id, n_ref = vt.tree.reference_ids(np.ones(nh, np.uint8), internal, K, L)
internal_t, ref_t = torch.from_numpy(internal).to(dev), torch.from_numpy(ref_id).to(dev)
desc = vt.synth.descriptors_from_tree(cent, K, L, n, seed=100)
work = torch.empty(int(lib.vt_quantize_mma_work_bytes(n, K, L)), dtype=torch.uint8, device=dev)
words = torch.empty(n, dtype=torch.int32, device=dev)
lvl, srt = (C.c_float * L)(), (C.c_float * L)()
acc = []
for rep in range(4):
    nat.check(lib.vt_quantize_mma_profile(nat.ptr(desc), 0, n, D, nat.ptr(cent), nat.ptr(internal_t), nat.ptr(ref_t), K, L,
                                          None, nat.ptr(words), None, nat.ptr(work), work.numel(), lvl, srt, nat.stream_ptr()), "profile")
    if rep:
        acc.append((list(lvl), list(srt)))
lv = np.mean([a[0] for a in acc], 0)
sr = np.mean([a[1] for a in acc], 0)
total = lv.sum() + sr.sum()
print("rows %d env %s" % (n, {k: v for k, v in os.environ.items() if k.startswith("VT_")}))
print(" level ms", np.round(lv, 3).tolist(), "sort ms", np.round(sr, 3).tolist(), "total %.3f ms -> %.2f G rows/s" % (total, n / total / 1e6))
print(" checksum", int(words.to(torch.int64).sum().item()))
