"""Row-sharded search across the GPUs of one box (one process per GPU, torch.distributed over NCCL/NVLink).

The path shards naturally: rank g holds the contiguous rows [offset_g, offset_g + n_g) of
the corpus, every rank sees the whole query block and runs the same single-GPU scan, and there is exactly ONE
exchange step -- an all-gather of each rank's per-query (cosine float64, global row int64) lists, k entries
per query (B*k*16 bytes per rank: 655 KB at B=4096, k=10) -- followed by a k-way merge on every rank.
No all-reduce, no all-to-all.  The collective is NCCL's all-gather; the merge is `sa_merge_shards`.

``index`` is duck-typed (``search(q, k, want_score64=True)``, ``merge_shards(s, i)``): production passes
``engine.VectorIndex``; the CPU tests drive the same code over gloo with an oracle-backed double.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced row ranges: rank g owns [g*N//G, (g+1)*N//G)."""
    return rank * n_total // world, (rank + 1) * n_total // world


class ShardedIndex:
    def __init__(self, index, row_offset: int, group=None):
        self.index = index
        self.row_offset = int(row_offset)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._buf = None

    def _buffers(self, nq: int, k: int, device):
        key = (nq, k, str(device))
        if self._buf is None or self._buf[0] != key:
            self._buf = (key,
                         torch.empty((self.world * nq, k), dtype=torch.float64, device=device),
                         torch.empty((self.world * nq, k), dtype=torch.int64, device=device))
        return self._buf[1], self._buf[2]

    def search(self, q: torch.Tensor, k: int):
        """q: [nq, dim] on this rank's device (identical on every rank).  Returns (score f32 [nq,k],
        global row i64 [nq,k]) -- the same on every rank."""
        s, i, s64 = self.index.search(q, k, want_score64=True)
        gi = torch.where(i >= 0, i.to(torch.int64) + self.row_offset, torch.full_like(i, -1, dtype=torch.int64))
        if self.world == 1:
            return s, gi
        all_s, all_i = self._buffers(q.shape[0], k, s64.device)
        dist.all_gather_into_tensor(all_s, s64.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, gi.contiguous(), group=self.group)
        nq = q.shape[0]  # gathered as [world*nq, k] (the layout gloo and NCCL both accept), viewed per shard
        return self.index.merge_shards(all_s.view(self.world, nq, k), all_i.view(self.world, nq, k))

    def search_host(self, q_f32, k: int):
        """Host buffers in, host buffers out (H2D, shard search, all-gather, merge, D2H)."""
        import numpy as np
        dev = self.index.rows.device if hasattr(self.index, "rows") else torch.device("cpu")
        q = torch.from_numpy(np.ascontiguousarray(q_f32, dtype=np.float32)).to(dev)
        s, gi = self.search(q, k)
        return s.cpu().numpy(), gi.cpu().numpy()
