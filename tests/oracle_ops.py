"""Test double for ``engine.CudaLevelOps``: the same five per-level steps computed by the CPU
oracle on CPU tensors, so that the multi-rank build protocol (``engine.build_tree`` + ``dist.Comm``)
can run under gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

from oracle import vt_oracle as vo


class OracleLevelOps:
    def column_sums(self, desc):
        return desc.to(torch.int64).sum(0)

    def kmeans_init(self, desc, perm, seg_begin, seg_local, seg_global, rank0, n_nodes, k, child_cent):
        for j in range(n_nodes):
            m_local = int(seg_local[j])
            m = int(seg_global[j]) if seg_global is not None else m_local
            if m < k:
                continue
            for c in range(k):
                r = (c * m) // k - (int(rank0[j]) if rank0 is not None else 0)
                if 0 <= r < m_local:
                    child_cent[j * k + c] = desc[int(perm[int(seg_begin[j]) + r])].to(torch.float32)

    def kmeans_assign(self, desc, perm, parent, n_active, node_internal, n_nodes, k, child_cent, label, sums,
                      counts, changed, stats):
        if n_active == 0:
            return
        dp = desc.shape[1]
        x = desc[perm[:n_active].long()].to(torch.float32).numpy()
        par = parent[:n_active].numpy()
        cent = child_cent.numpy()
        new = label[:n_active].numpy().copy()
        s = sums.numpy().reshape(-1, dp)
        cnt = counts.numpy()
        for j in np.unique(par):
            if not node_internal[j]:
                continue
            rows = np.nonzero(par == j)[0]
            lab = vo.assign(x[rows], cent[j * k:(j + 1) * k])
            new[rows] = lab
            np.add.at(s, j * k + lab, x[rows].astype(np.int64))
            np.add.at(cnt, j * k + lab, 1)
        changed += int((new != label[:n_active].numpy()).sum())
        label[:n_active] = torch.from_numpy(new)

    def kmeans_update(self, sums, counts, n_rows, dp, child_cent):
        s = sums.view(n_rows, dp).to(torch.float64)
        c = counts.to(torch.float64)[:, None]
        upd = (s / c).to(torch.float32)
        mask = (counts > 0)[:, None].expand(-1, dp)
        child_cent[mask] = upd[mask]

    def stable_sort(self, keys, vals, n_bits):
        order = torch.sort(keys, stable=True).indices
        return keys[order], vals[order]

    def partition(self, parent, label, n_active, node_internal, k, n_rows, perm):
        par = parent[:n_active].long()
        keys = torch.where(node_internal[par] != 0, par * k + label[:n_active].long(),
                           torch.full_like(par, n_rows))
        order = torch.sort(keys, stable=True).indices
        return keys[order].to(torch.int32), perm[:n_active][order]


def merge_ranked_cpu(keys, ids, scores, topk):
    """Test double for ``dist.merge_ranked`` (the product runs vt_topk_merge_scored on the device): the same
    comparator -- key descending, ties by ascending image index, empty slots (id -1) last -- with torch sorts."""
    w, q, k = keys.shape
    keys = keys.permute(1, 0, 2).reshape(q, w * k)
    ids = ids.permute(1, 0, 2).reshape(q, w * k)
    scores = scores.permute(1, 0, 2).reshape(q, w * k)
    big = torch.iinfo(torch.int32).max
    ids_key = torch.where(ids >= 0, ids, torch.full_like(ids, big))
    keys = torch.where(ids >= 0, keys, torch.full_like(keys, float("-inf")))
    o1 = torch.sort(ids_key, dim=1, stable=True).indices
    k1 = torch.gather(keys, 1, o1)
    o2 = torch.sort(k1, dim=1, descending=True, stable=True).indices
    order = torch.gather(o1, 1, o2)[:, :topk]
    return torch.gather(ids, 1, order), torch.gather(scores, 1, order), torch.gather(keys, 1, order)
