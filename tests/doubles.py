"""Test doubles: an oracle-backed stand-in for engine.VectorIndex so host logic (operator, serve loop, sharding)
can be exercised on a box without a GPU.  Lives in tests/ -- the product never imports it."""
import numpy as np
import torch

from oracle import bruteforce as bf


class OracleIndex:
    def __init__(self, dim, capacity=1 << 20):
        self.dim = dim
        self.capacity = capacity
        self.bits = np.zeros((0, dim), dtype=np.uint16)

    def __len__(self):
        return len(self.bits)

    def reset(self):
        self.bits = np.zeros((0, self.dim), dtype=np.uint16)

    def append(self, rows_f32):
        first = len(self.bits)
        self.bits = np.concatenate([self.bits, bf.f32_to_bf16_bits(np.asarray(rows_f32, dtype=np.float32))])
        return first

    def delete_rows(self, rows):
        self.bits[list(rows)] = 0

    def search_host(self, q_f32, k):
        s, i = bf.cosine_topk_f64(bf.f32_to_bf16_bits(np.asarray(q_f32, dtype=np.float32)), self.bits, k)
        return s.astype(np.float32), i.astype(np.int32)

    # torch-tensor flavoured API used by sharded.ShardedIndex
    def search(self, q, k, want_score64=False):
        s, i = bf.cosine_topk_f64(bf.f32_to_bf16_bits(q.float().numpy()), self.bits, k)
        out = (torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i.astype(np.int32)))
        return out + (torch.from_numpy(s),) if want_score64 else out

    # exchange-format API (sa_hit = {cosine f64, global row i64}) used by sharded.ShardedIndex
    def search_hits(self, q, k, row_offset=0):
        s, i = bf.cosine_topk_f64(bf.f32_to_bf16_bits(q.float().numpy()), self.bits, k)
        hits = np.zeros(s.shape, dtype=np.dtype([("score", "<f8"), ("row", "<i8")]))
        hits["score"] = s
        hits["row"] = np.where(i >= 0, i.astype(np.int64) + row_offset, -1)
        return torch.from_numpy(hits.view(np.uint8).reshape(s.shape[0], k, 16).copy())

    def merge_hits(self, hits_all):
        g, nq, k, _ = hits_all.shape
        h = hits_all.contiguous().numpy().view(np.dtype([("score", "<f8"), ("row", "<i8")])).reshape(g, nq, k)
        s, i = bf.merge_shard_topk([h["score"][j] for j in range(g)], [h["row"][j] for j in range(g)], [0] * g, k)
        return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i)

    def merge_shards(self, all_s, all_i):
        g = all_s.shape[0]
        s, i = bf.merge_shard_topk([all_s[j].numpy() for j in range(g)], [all_i[j].numpy() for j in range(g)],
                                   [0] * g, all_s.shape[2])
        return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i)


class PipelinedOracleIndex(OracleIndex):
    """Adds the split host call (submit / wait) so the serve loop's software pipelining runs on a CPU box."""

    def __init__(self, dim, capacity=1 << 20):
        super().__init__(dim, capacity)
        self._slots = {}
        self.max_inflight = 0

    def search_host_submit(self, q_f32, k, slot=0):
        assert slot not in self._slots, "slot reused before wait"
        self._slots[slot] = self.search_host(q_f32, k)
        self.max_inflight = max(self.max_inflight, len(self._slots))

    def search_host_wait(self, slot=0, out=None):
        return self._slots.pop(slot)
