"""Per-level timing of engine.build_tree (developer tool, not part of the product path)."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
import vtree_b200 as vt
from vtree_b200 import engine

dev = torch.device("cuda")
cent = vt.synth.hierarchical_tree(10, 6, 128, seed=7, device=dev)
desc = vt.synth.descriptors_from_tree(cent, 10, 6, 10_000_000, seed=100)


class TimedOps(engine.CudaLevelOps):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.t = {}

    def _timed(self, name, fn, *a):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a)
        torch.cuda.synchronize()
        self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0
        return r

    def kmeans_init(self, *a): return self._timed("init", super().kmeans_init, *a)
    def kmeans_assign(self, *a): return self._timed("assign", super().kmeans_assign, *a)
    def kmeans_update(self, *a): return self._timed("update", super().kmeans_update, *a)
    def partition(self, *a): return self._timed("partition", super().partition, *a)
    def column_sums(self, *a): return self._timed("colsum", super().column_sums, *a)


for tc in (True, False):
    ops = TimedOps(use_tensor_cores=tc)
    engine.build_tree(desc[:1_000_000], 128, 10, 6, n_iter=2, ops=TimedOps(use_tensor_cores=tc))   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tree = engine.build_tree(desc, 128, 10, 6, n_iter=10, ops=ops)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print("tensor_cores=%s total %.3f s  ops: %s  other(host glue, zeroing, FlatTree) %.3f s" % (
        tc, total, {k: round(v, 3) for k, v in ops.t.items()}, total - sum(ops.t.values())))
