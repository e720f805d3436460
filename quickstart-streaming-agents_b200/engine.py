"""VectorIndex -- the host-side handle of one corpus shard on one B200.

Plays the role of the reference's external vector table ``documents_vectordb_lab2`` (connector 'mongodb',
index 'vector_index', cosine, 1536-d: terraform/lab2-vector-search/main.tf:215,
assets/pre-setup/MongoDB-Setup.md:72-83).  torch tensors are used only as device-memory holders; every
operation is a call into libsa_b200.so through ``capi`` (include/sa_api.h).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import capi


@dataclass
class SearchTiming:
    scan_ms: float      # sum of the scan-kernel launches of the last search (CUDA events)
    total_ms: float     # first scan start .. last merge end
    bytes: float        # algorithmic bytes of that search
    flops: float        # algorithmic flops of that search
    launches: int       # scan launches
    kernels: int        # all kernels launched


def _ptr(t: torch.Tensor | None) -> int:
    return 0 if t is None else t.data_ptr()


HIT_DTYPE = np.dtype([("score", "<f8"), ("row", "<i8")])   # sa_hit of include/sa_api.h


def pinned_array(shape, dtype=np.float32) -> np.ndarray:
    """A page-locked numpy array (sa_host_alloc), freed (sa_host_free) when its last view is garbage-collected."""
    import weakref
    lib = capi.load()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    capi.check(lib.sa_host_alloc(C.byref(p), max(nbytes, 1)), "sa_host_alloc")
    buf = (C.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, lib.sa_host_free, C.c_void_p(p.value))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class VectorIndex:
    """A row shard of the corpus resident in HBM: bf16 rows [capacity, dim] + fp32 inverse norms [capacity]."""

    def __init__(self, dim: int = 1536, capacity: int = 1 << 20, max_batch: int = 1024, max_k: int = 10,
                 device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("VectorIndex needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = capi.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dim, self.capacity, self.max_batch, self.max_k = int(dim), int(capacity), int(max_batch), int(max_k)
        dev = torch.device("cuda", self.device)
        # device-memory holders (the engine never copies or frees these)
        self.rows = torch.empty((self.capacity, self.dim), dtype=torch.bfloat16, device=dev)
        self.inv_norm = torch.zeros((self.capacity,), dtype=torch.float32, device=dev)
        h = C.c_void_p()
        capi.check(self.lib.sa_engine_create(C.byref(h), self.device, self.dim, self.capacity, self.max_batch,
                                             self.max_k), "sa_engine_create")
        self._h = h
        self._inflight = {}
        capi.check(self.lib.sa_corpus_bind(self._h, self.rows.data_ptr(), self.inv_norm.data_ptr(), 0),
                   "sa_corpus_bind")

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        """Destroy the engine.  Page-locked arrays handed out by ``pinned_array`` are NOT freed here: each is released
        when its last numpy view is garbage-collected, so a caller still holding one never touches freed memory."""
        if getattr(self, "_h", None):
            self.lib.sa_engine_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.lib.sa_corpus_rows(self._h))

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_option(self, name: str, value: int) -> None:
        capi.check(self.lib.sa_set_option(self._h, name.encode(), int(value)), "sa_set_option")

    def info(self, name: str) -> int:
        v = C.c_int64()
        capi.check(self.lib.sa_get_info(self._h, name.encode(), C.byref(v)), "sa_get_info")
        return int(v.value)

    # ------------------------------------------------------------------ ingest
    def reset(self) -> None:
        """Forget every row (the job of scripts/common/clear_mongodb.py:98-158 in the reference)."""
        capi.check(self.lib.sa_corpus_reset(self._h), "sa_corpus_reset")

    def append(self, rows_f32) -> int:
        """Append fp32 embeddings (host numpy or device tensor) -> bf16 rows + norms.  Returns first row id."""
        first = len(self)
        if isinstance(rows_f32, torch.Tensor) and rows_f32.is_cuda:
            x = rows_f32.to(torch.float32).contiguous()
            assert x.dim() == 2 and x.shape[1] == self.dim
            capi.check(self.lib.sa_corpus_append_f32(self._h, x.data_ptr(), x.shape[0], self._stream()),
                       "sa_corpus_append_f32")
            torch.cuda.current_stream(self.device).synchronize()  # x may be freed by the caller
        else:
            x = np.ascontiguousarray(rows_f32, dtype=np.float32)
            assert x.ndim == 2 and x.shape[1] == self.dim
            torch.cuda.current_stream(self.device).synchronize()  # the *_host calls run on the engine's stream
            capi.check(self.lib.sa_corpus_append_host_f32(self._h, x.ctypes.data, x.shape[0]),
                       "sa_corpus_append_host_f32")
        return first

    def append_bf16_bits(self, bits: np.ndarray) -> int:
        """Append rows given as bf16 bit patterns (uint16 [n, dim]) -- used with the synthetic corpora so the
        device holds exactly the bits the oracle sees."""
        bits = np.ascontiguousarray(bits, dtype=np.uint16)
        assert bits.ndim == 2 and bits.shape[1] == self.dim
        first = len(self)
        n = bits.shape[0]
        if first + n > self.capacity:
            raise capi.SaError(capi.SA_ERR_CAPACITY, "append_bf16_bits", "append past capacity")
        src = torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16)
        self.rows[first:first + n].copy_(src)
        self.commit(first, n)
        return first

    # ------------------------------------------------------------------ checkpoint / resume
    def snapshot(self, path: str) -> int:
        """Write the committed rows (bf16 bits) and their inverse norms to ``path`` (.npz).  Returns the row count.
        (The reference leaves corpus durability to Atlas; here a snapshot + the consumer-group offsets are the
        checkpoint, and replaying `documents_embed` from offset 0 is the fallback.)"""
        n = len(self)
        torch.cuda.current_stream(self.device).synchronize()
        bits = self.rows[:n].view(torch.int16).cpu().numpy().view(np.uint16)
        import os
        path = path if path.endswith(".npz") else path + ".npz"
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:                       # written under a temporary name, then renamed: never half a file
            np.savez(f, rows=bits, inv_norm=self.inv_norm[:n].cpu().numpy(), dim=np.int64(self.dim))
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, path)
        return n

    def restore(self, path: str) -> int:
        """Load a snapshot written by ``snapshot`` into this (empty or not) index, replacing its contents."""
        z = np.load(path if path.endswith(".npz") else path + ".npz")
        if int(z["dim"]) != self.dim:
            raise ValueError(f"snapshot has dim {int(z['dim'])}, index has {self.dim}")
        bits, inv = z["rows"], z["inv_norm"]
        n = bits.shape[0]
        if n > self.capacity:
            raise capi.SaError(capi.SA_ERR_CAPACITY, "restore", "snapshot larger than capacity")
        self.rows[:n].copy_(torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16))
        self.inv_norm[:n].copy_(torch.from_numpy(inv))
        torch.cuda.current_stream(self.device).synchronize()
        capi.check(self.lib.sa_corpus_bind(self._h, self.rows.data_ptr(), self.inv_norm.data_ptr(), n),
                   "sa_corpus_bind")
        return n

    def delete_rows(self, rows) -> None:
        """Tombstone rows: zero the stored vector and its inverse norm -- all-zero rows are never returned."""
        if len(rows) == 0:
            return
        ix = torch.as_tensor(list(rows), dtype=torch.long, device=self.rows.device)
        self.rows.index_fill_(0, ix, 0)
        self.inv_norm.index_fill_(0, ix, 0)

    def commit(self, first: int, n: int) -> None:
        """Rows [first, first+n) were written into ``self.rows`` in place: compute norms and publish them."""
        capi.check(self.lib.sa_corpus_commit(self._h, int(first), int(n), self._stream()), "sa_corpus_commit")

    # ------------------------------------------------------------------ search
    def search(self, q: torch.Tensor, k: int, want_score64: bool = False):
        """Device path.  q: [nq, dim] bf16 or fp32 CUDA tensor.  Returns (score f32 [nq,k], idx i32 [nq,k]
        [, score64 f64 [nq,k]]) as CUDA tensors, asynchronous on the current stream."""
        assert q.is_cuda and q.dim() == 2 and q.shape[1] == self.dim
        q = q.contiguous()
        nq = q.shape[0]
        dev = q.device
        score = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
        s64 = torch.empty((nq, k), dtype=torch.float64, device=dev) if want_score64 else None
        if q.dtype == torch.bfloat16:
            rc = self.lib.sa_search(self._h, q.data_ptr(), nq, k, score.data_ptr(), idx.data_ptr(), _ptr(s64),
                                    self._stream())
        elif q.dtype == torch.float32:
            rc = self.lib.sa_search_f32(self._h, q.data_ptr(), nq, k, score.data_ptr(), idx.data_ptr(), _ptr(s64),
                                        self._stream())
        else:
            raise TypeError("queries must be bf16 or fp32")
        capi.check(rc, "sa_search")
        return (score, idx, s64) if want_score64 else (score, idx)

    def search_host(self, q_f32: np.ndarray, k: int, out=None):
        """End-to-end path with HOST buffers (H2D, search, D2H inside the call).  Returns numpy
        (score f32 [nq,k], idx i32 [nq,k]); ``out=(score, idx)`` reuses caller buffers (e.g. ``pinned_array``)."""
        q = np.ascontiguousarray(q_f32, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.dim
        nq = q.shape[0]
        if out is None:
            score = np.empty((nq, k), dtype=np.float32)
            idx = np.empty((nq, k), dtype=np.int32)
        else:
            score, idx = out
            assert score.shape == (nq, k) and score.dtype == np.float32 and score.flags.c_contiguous
            assert idx.shape == (nq, k) and idx.dtype == np.int32 and idx.flags.c_contiguous
        torch.cuda.current_stream(self.device).synchronize()  # the *_host calls run on the engine's stream
        capi.check(self.lib.sa_search_host(self._h, q.ctypes.data, nq, k, score.ctypes.data, idx.ctypes.data),
                   "sa_search_host")
        return score, idx

    def search_host_submit(self, q_f32: np.ndarray, k: int, slot: int = 0) -> None:
        """First half of ``search_host``: enqueue H2D + search + D2H for ``slot`` (0 or 1) and return at once, so the
        caller can prepare the next batch while the GPU works.  Collect with ``search_host_wait(slot)``."""
        q = np.ascontiguousarray(q_f32, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.dim
        torch.cuda.current_stream(self.device).synchronize()  # the *_host calls run on the engine's own stream
        self._inflight[slot] = (q, q.shape[0], k)  # keeps a pinned source alive until the wait
        capi.check(self.lib.sa_search_host_submit(self._h, slot, q.ctypes.data, q.shape[0], k),
                   "sa_search_host_submit")

    def search_host_wait(self, slot: int = 0, out=None):
        _, nq, k = self._inflight.pop(slot)
        if out is None:
            score = np.empty((nq, k), dtype=np.float32)
            idx = np.empty((nq, k), dtype=np.int32)
        else:
            score, idx = out
            assert score.shape == (nq, k) and score.dtype == np.float32 and score.flags.c_contiguous
            assert idx.shape == (nq, k) and idx.dtype == np.int32 and idx.flags.c_contiguous
        capi.check(self.lib.sa_search_host_wait(self._h, slot, score.ctypes.data, idx.ctypes.data),
                   "sa_search_host_wait")
        return score, idx

    def pinned_array(self, shape, dtype=np.float32) -> np.ndarray:
        """A page-locked numpy array (sa_host_alloc): passing such buffers to ``search_host`` lets the engine DMA
        them directly instead of staging through its own pinned copy.  Freed when the index is closed."""
        return pinned_array(shape, dtype)

    def search_hits(self, q: torch.Tensor, k: int, row_offset: int = 0) -> torch.Tensor:
        """This shard's results in exchange format: uint8 CUDA tensor [nq, k, 16] = sa_hit {cosine f64, global row i64}
        (``hits.view(torch.float64)[..., 0]`` / ``.view(torch.int64)[..., 1]``)."""
        assert q.is_cuda and q.dtype == torch.bfloat16 and q.dim() == 2 and q.shape[1] == self.dim
        q = q.contiguous()
        hits = torch.empty((q.shape[0], k, 16), dtype=torch.uint8, device=q.device)
        capi.check(self.lib.sa_search_hits(self._h, q.data_ptr(), q.shape[0], k, int(row_offset), hits.data_ptr(),
                                           self._stream()), "sa_search_hits")
        return hits

    def merge_hits(self, hits_all: torch.Tensor):
        """hits_all: uint8 [n_shards, nq, k, 16] gathered from all shards.  Returns (score f32 [nq,k], global row i64)."""
        g, nq, k, _ = hits_all.shape
        score = torch.empty((nq, k), dtype=torch.float32, device=hits_all.device)
        idx = torch.empty((nq, k), dtype=torch.int64, device=hits_all.device)
        capi.check(self.lib.sa_merge_hits(self._h, hits_all.contiguous().data_ptr(), g, nq, k, score.data_ptr(),
                                          idx.data_ptr(), self._stream()), "sa_merge_hits")
        return score, idx

    def scan_profile(self) -> dict:
        """Per-CTA role counters of the last scan launch run with option "profile" = 1 (SM cycles): how long the TMA
        producer waited for a free smem slot, the MMA issuer for data / for the epilogue, the epilogue for the MMA, and
        how long the epilogue worked.  Arrays are indexed by CTA."""
        n = self.info("last_grid")
        a = np.zeros((n, 8), dtype=np.int64)
        got = C.c_int()
        capi.check(self.lib.sa_scan_profile(self._h, a.ctypes.data, n, C.byref(got)), "sa_scan_profile")
        a = a[:got.value]
        names = ("prod_wait_empty", "mma_wait_full", "mma_wait_tempty", "epi_wait_tfull", "epi_busy", "epi_slow_chunks",
                 "total", "tiles")
        return {nm: a[:, i] for i, nm in enumerate(names)}

    def merge_shards(self, score64_all: torch.Tensor, gidx_all: torch.Tensor):
        """score64_all / gidx_all: [n_shards, nq, k] (float64 / int64 global rows) gathered from all ranks.
        Returns (score f32 [nq,k], global idx i64 [nq,k])."""
        g, nq, k = score64_all.shape
        score = torch.empty((nq, k), dtype=torch.float32, device=score64_all.device)
        idx = torch.empty((nq, k), dtype=torch.int64, device=score64_all.device)
        capi.check(self.lib.sa_merge_shards(self._h, score64_all.contiguous().data_ptr(),
                                            gidx_all.contiguous().data_ptr(), g, nq, k, score.data_ptr(),
                                            idx.data_ptr(), self._stream()), "sa_merge_shards")
        return score, idx

    def last_timing(self) -> SearchTiming:
        a, b = C.c_float(), C.c_float()
        by, fl = C.c_double(), C.c_double()
        n, kn = C.c_int(), C.c_int()
        capi.check(self.lib.sa_last_timing(self._h, C.byref(a), C.byref(b), C.byref(by), C.byref(fl), C.byref(n),
                                           C.byref(kn)), "sa_last_timing")
        return SearchTiming(a.value, b.value, by.value, fl.value, n.value, kn.value)

    def timing_mean(self, n: int = 16) -> tuple[float, float, int]:
        """(mean scan ms, mean total ms, searches averaged) over the most recent min(n, 16) searches."""
        a, b, m = C.c_float(), C.c_float(), C.c_int()
        capi.check(self.lib.sa_timing_mean(self._h, int(n), C.byref(a), C.byref(b), C.byref(m)), "sa_timing_mean")
        return a.value, b.value, m.value

    def debug_tile_dots(self, q_bf16: torch.Tensor, tile: int, cta_group: int = 1) -> torch.Tensor:
        """Test hook: raw Q.C^T accumulators of one 256-row corpus tile, [padded nq, 256] fp32."""
        nq = q_bf16.shape[0]
        rows = 128 * cta_group
        padded = (nq + rows - 1) // rows * rows
        out = torch.zeros((padded, 256), dtype=torch.float32, device=q_bf16.device)
        capi.check(self.lib.sa_debug_tile_dots(self._h, q_bf16.contiguous().data_ptr(), nq, tile, cta_group,
                                               out.data_ptr(), self._stream()), "sa_debug_tile_dots")
        return out
