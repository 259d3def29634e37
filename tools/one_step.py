"""Two identical device-resident quantisation steps (k=10, L=6, synthetic data) for `ncu` launch lists: the second step's
launches are one step of bench.py's headline loop.

    python tools/one_step.py one  100000000      # default engine: one tree level per launch
    python tools/one_step.py pair 100000000      # two tree levels per launch (vt_quantize_pair.cu)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vtree_b200 as vt  # noqa: E402
from vtree_b200 import _native as nat, engine  # noqa: E402

K, L, D = 10, 6, 128
mode = sys.argv[1] if len(sys.argv) > 1 else "one"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
dev = torch.device("cuda:0")
lib = nat.load()
cent = vt.synth.hierarchical_tree(K, L, D, seed=7, device=dev)
nh = vt.tree.heap_size(K, L)
internal = np.zeros(nh, np.uint8)
internal[:vt.tree.level_offset(K, L)] = 1
ref_id, n_ref = vt.tree.reference_ids(np.ones(nh, np.uint8), internal, K, L)
tree_dev = dict(centroids=cent, internal=torch.from_numpy(internal).to(dev), ref_id=torch.from_numpy(ref_id).to(dev))
desc = vt.synth.descriptors_from_tree(cent, K, L, n, seed=100)
work = torch.empty(int(lib.vt_quantize_mma_work_bytes(n, K, L)), dtype=torch.uint8, device=dev)
lib.vt_quantize_mma_pair_min((1 << 20) if mode == "pair" else -1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
w = engine.quantize(tree_dev, K, L, desc, method="mma", work=work)
ev[1].record()
w = engine.quantize(tree_dev, K, L, desc, method="mma", work=work)
ev[2].record()
torch.cuda.synchronize()
print("engine %s rows %d: step 1 %.3f ms (incl. plane preparation), step 2 %.3f ms, checksum %d"
      % (mode, n, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), int(w.to(torch.int64).sum().item())))
