"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/`` (incl. ``tests/harness/``), ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path (``quickstart-streaming-agents_b200``) never does and fails loudly when the
CUDA library is missing.

PARITY UNPINNED.  The reference (confluentinc/quickstart-streaming-agents) contains no implementation of
this arithmetic and no test that asserts a score, an index or a recall for it:

* the operator is ``LATERAL TABLE(VECTOR_SEARCH_AGG(documents_vectordb_lab2, DESCRIPTOR(embedding),
  qe.embedding, 3))`` (terraform/lab2-vector-search/main.tf:292; same call LAB3-Walkthrough.md:343-350,
  LAB4-Walkthrough.md:302-309), executed by MongoDB Atlas ``$vectorSearch`` (service, un-versioned; table DDL
  main.tf:215 with 'mongodb.numCandidates'='500') behind Confluent Cloud Flink -- neither is in the tree;
* the metric is fixed by the index definition ``{"type":"vector","path":"embedding","numDimensions":1536,
  "similarity":"cosine"}`` (assets/pre-setup/MongoDB-Setup.md:72-83; scripts/common/validate.py:56-61,167-180);
* the only tests of the path assert "at least one message" / "non-empty response"
  (testing/e2e/test_lab2.py:99-135).

So this file *restates* the published semantics -- k rows of highest cosine similarity, descending, 1..k --
as exact brute force, which is what BASELINE.json's metric ("recall@10 vs numpy") names as the yardstick:

    score(q, d) = <q, d> / (|q| |d|)   over the bf16-rounded values, every sum accumulated in float64,
    result      = the k rows of highest score, ordered by (score descending, row index ascending);
    all-zero corpus rows are never returned; an all-zero query scores 0 against every row.

``cosine_topk_f64`` is the definition.  ``cosine_topk_fast`` (fp32 sgemm prefilter + float64 rescoring of a
margin of candidates) is the same function made cheap enough for 10M-row checks and CPU-baseline timing; the
tests pin it against ``cosine_topk_f64``.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "f32_to_bf16_bits",
    "bf16_bits_to_f32",
    "synth_rows",
    "synth_queries",
    "cosine_topk_f64",
    "cosine_topk_fast",
    "cosine_topk_sgemm",
    "cosine_topk_sgemm_prepared",
    "prepare_chunks_f32",
    "compare_topk",
    "merge_shard_topk",
]

CHUNK_ROWS = 262_144  # generation chunk of the synthetic corpora


# --------------------------------------------------------------------------------------------------
# bf16 <-> fp32 (bit patterns as uint16)
# --------------------------------------------------------------------------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns (NaN stays NaN).
    Works in 32-bit arithmetic over blocks of ~4M elements, so a 262144 x 1536 chunk needs no multi-GB temporaries
    (u + 0x7FFF + lsb cannot wrap for a non-NaN input: the largest finite/inf pattern is 0xFF800000)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).reshape(-1)
    r = np.empty(u.shape, dtype=np.uint16)
    step = 1 << 22
    for lo in range(0, u.size, step):
        b = u[lo:lo + step]
        t = (b >> 16) & 1
        t += 0x7FFF
        nan = (b & 0x7FFFFFFF) > 0x7F800000
        with np.errstate(over="ignore"):
            t += b                                  # wraps only where `nan` is set; those are overwritten below
        out = (t >> 16).astype(np.uint16)
        if nan.any():
            out[nan] = ((b[nan] >> 16) | 0x40).astype(np.uint16)
        r[lo:lo + step] = out
    return r.reshape(x.shape)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint16)
    return (b.astype(np.uint32) << 16).view(np.float32)


# --------------------------------------------------------------------------------------------------
# synthetic data (platform-independent: numpy PCG64, never a device generator)
# --------------------------------------------------------------------------------------------------
def synth_rows(seed: int, chunk: int, rows: int, dim: int) -> np.ndarray:
    """Chunk ``chunk`` of a synthetic corpus as bf16 bits [rows, dim].

    Gaussian rows with a per-row log-uniform scale in [e^-0.7, e^0.7]: rows are NOT pre-normalised, so the
    fused L2 normalisation is exercised.  (Recipe: SURVEY.md section 8d.)
    """
    g = np.random.default_rng(seed + chunk)
    x = g.standard_normal((rows, dim), dtype=np.float32)
    x *= np.exp(g.uniform(-0.7, 0.7, (rows, 1))).astype(np.float32)
    return f32_to_bf16_bits(x)


def synth_queries(seed: int, nq: int, dim: int, corpus_chunk0: np.ndarray | None = None) -> np.ndarray:
    """Query block as bf16 bits [nq, dim].  Odd queries are "planted" near row
    j = (i * 2654435761) mod len(corpus_chunk0) of the corpus' first chunk (known top-1, an oracle-free check)."""
    q = bf16_bits_to_f32(synth_rows(seed, 0, nq, dim)).copy()
    if corpus_chunk0 is not None and len(corpus_chunk0):
        g = np.random.default_rng(seed ^ 0x5EED)
        n0 = len(corpus_chunk0)
        for i in range(1, nq, 2):
            j = (i * 2654435761) % n0
            base = bf16_bits_to_f32(corpus_chunk0[j])
            nrm = float(np.linalg.norm(base.astype(np.float64)))
            eps = g.standard_normal(dim).astype(np.float32)
            q[i] = base + np.float32(0.5 * nrm / np.sqrt(dim)) * eps
    return f32_to_bf16_bits(q)


def planted_row(i: int, n0: int) -> int:
    return (i * 2654435761) % n0


# --------------------------------------------------------------------------------------------------
# the definition
# --------------------------------------------------------------------------------------------------
def _select_topk(scores: np.ndarray, rows: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """Top-k of one query's candidate (score float64, row int64) pairs by (score desc, row asc)."""
    order = np.lexsort((rows, -scores))[:k]
    return scores[order], rows[order]


def cosine_topk_f64(q_bits: np.ndarray, c_bits: np.ndarray, k: int, chunk: int = 32768):
    """Exact cosine top-k.  Returns (score float64 [nq,k], row int64 [nq,k]); unused slots are (-inf, -1)."""
    q = bf16_bits_to_f32(q_bits).astype(np.float64)
    nq = q.shape[0]
    n = c_bits.shape[0]
    qn = np.sqrt((q * q).sum(axis=1))
    best_s = np.full((nq, k), -np.inf)
    best_i = np.full((nq, k), -1, dtype=np.int64)
    for lo in range(0, n, chunk):
        c = bf16_bits_to_f32(c_bits[lo : lo + chunk]).astype(np.float64)
        cn = np.sqrt((c * c).sum(axis=1))
        dots = q @ c.T
        den = qn[:, None] * cn[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(den > 0, dots / den, 0.0)
        s[:, cn == 0] = -np.inf  # all-zero corpus rows are never returned
        m = s.shape[1]
        kk = min(k, m)
        kth = np.partition(s, m - kk, axis=1)[:, m - kk]
        for r in range(nq):
            cand = np.flatnonzero(s[r] >= kth[r])
            cs = np.concatenate([best_s[r], s[r, cand]])
            ci = np.concatenate([best_i[r], cand.astype(np.int64) + lo])
            keep = np.isfinite(cs)
            ts, ti = _select_topk(cs[keep], ci[keep], k)
            best_s[r, : len(ts)] = ts
            best_i[r, : len(ti)] = ti
    return best_s, best_i


def _rescore_f64(q_bits_row: np.ndarray, c_bits_rows: np.ndarray) -> np.ndarray:
    q = bf16_bits_to_f32(q_bits_row).astype(np.float64)
    c = bf16_bits_to_f32(c_bits_rows).astype(np.float64)
    qq = float((q * q).sum())
    dd = (c * c).sum(axis=1)
    dot = c @ q
    den = qq * dd
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(den > 0, dot / np.sqrt(den), 0.0)
    s[dd == 0] = -np.inf
    return s


def cosine_topk_sgemm(q_bits: np.ndarray, c_chunks, k: int):
    """numpy brute force exactly as BASELINE.md section 4 words it: fp32 sgemm over row chunks (bf16 upcast per
    chunk), running top-k via argpartition.  ``c_chunks`` yields (first_row, bits[rows, dim]).  This is the
    timed CPU baseline; its scores are fp32-accurate (no rescoring)."""
    q = bf16_bits_to_f32(q_bits)
    qn = np.sqrt((q.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    qh = q / np.where(qn > 0, qn, 1)[:, None]
    nq = q.shape[0]
    best_s = np.full((nq, k), -np.inf, dtype=np.float32)
    best_i = np.full((nq, k), -1, dtype=np.int64)
    for lo, bits in c_chunks:
        c = bf16_bits_to_f32(bits)
        cn = np.sqrt(np.einsum("ij,ij->i", c, c, dtype=np.float32))
        inv = np.where(cn > 0, 1.0 / np.where(cn > 0, cn, 1), 0).astype(np.float32)
        s = (qh @ c.T) * inv[None, :]
        s[:, cn == 0] = -np.inf
        m = s.shape[1]
        kk = min(k, m)
        part = np.argpartition(s, m - kk, axis=1)[:, m - kk :]
        ps = np.take_along_axis(s, part, axis=1)
        cs = np.concatenate([best_s, ps], axis=1)
        ci = np.concatenate([best_i, part.astype(np.int64) + lo], axis=1)
        order = np.lexsort((ci, -cs), axis=1)[:, :k]
        best_s = np.take_along_axis(cs, order, axis=1)
        best_i = np.take_along_axis(ci, order, axis=1)
    return best_s, best_i


def prepare_chunks_f32(c_chunks):
    """Ingest-time work of a CPU engine, done once and NOT timed by the baseline: upcast each bf16 chunk to fp32
    and L2-normalise its rows (all-zero rows stay zero and are flagged).  Yields (first_row, unit rows f32, zero mask)."""
    out = []
    for lo, bits in c_chunks:
        c = bf16_bits_to_f32(bits)
        cn = np.sqrt(np.einsum("ij,ij->i", c, c, dtype=np.float64)).astype(np.float32)
        zero = cn == 0
        c /= np.where(zero, 1, cn)[:, None]
        out.append((lo, c, zero))
    return out


_TOPK_LIB = None


def _topk_lib():
    """oracle/_build/liboracle_topk.so (oracle/topk.c, built by __graft_entry__.build()), or None when absent."""
    global _TOPK_LIB
    if _TOPK_LIB is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_topk.so")
        _TOPK_LIB = False
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.oracle_topk_merge_f32.restype = None
            lib.oracle_topk_merge_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
            _TOPK_LIB = lib
    return _TOPK_LIB or None


def cosine_topk_sgemm_prepared(q_bits: np.ndarray, prepared, k: int, use_c_topk: bool = True):
    """The timed CPU baseline: brute force over a corpus that already sits in RAM as unit-norm fp32 rows
    (``prepare_chunks_f32``).  Per chunk: one sgemm (all BLAS threads) + top-k selection + running merge; the selection
    runs on all cores through oracle/topk.c when that is built, else numpy argpartition (one thread)."""
    q = bf16_bits_to_f32(q_bits)
    qn = np.sqrt((q.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    qh = q / np.where(qn > 0, qn, 1)[:, None]
    nq = q.shape[0]
    best_s = np.full((nq, k), -np.inf, dtype=np.float32)
    best_i = np.full((nq, k), -1, dtype=np.int64)
    lib = _topk_lib() if use_c_topk else None
    for lo, c, zero in prepared:
        s = qh @ c.T
        if zero.any():
            s[:, zero] = -np.inf
        if lib is not None:
            s = np.ascontiguousarray(s, dtype=np.float32)
            lib.oracle_topk_merge_f32(s.ctypes.data, nq, s.shape[1], k, lo, best_s.ctypes.data, best_i.ctypes.data)
            continue
        m = s.shape[1]
        kk = min(k, m)
        part = np.argpartition(s, m - kk, axis=1)[:, m - kk :]
        ps = np.take_along_axis(s, part, axis=1)
        cs = np.concatenate([best_s, ps], axis=1)
        ci = np.concatenate([best_i, part.astype(np.int64) + lo], axis=1)
        order = np.lexsort((ci, -cs), axis=1)[:, :k]
        best_s = np.take_along_axis(cs, order, axis=1)
        best_i = np.take_along_axis(ci, order, axis=1)
    return best_s, best_i


def cosine_topk_fast(q_bits: np.ndarray, c_chunks, k: int, margin: int = 32):
    """Same result as ``cosine_topk_f64`` at a fraction of the cost: fp32 sgemm prefilter keeping k+margin
    candidates per query per chunk, then float64 rescoring of the survivors.  ``c_chunks`` yields
    (first_row, bits[rows, dim]) and may be a generator (the 10M-row corpus never sits in RAM twice).
    Equality with the definition holds unless more than ``margin`` rows sit within fp32 rounding error
    (~1e-6) of a query's k-th score; tests/test_oracle.py checks it on seeded data."""
    q = bf16_bits_to_f32(q_bits)
    nq = q.shape[0]
    qn = np.sqrt((q.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    qh = q / np.where(qn > 0, qn, 1)[:, None]
    keep = k + margin
    cand_s = np.full((nq, keep), -np.inf, dtype=np.float32)
    cand_i = np.full((nq, keep), -1, dtype=np.int64)
    cand_bits = np.zeros((nq, keep, q.shape[1]), dtype=np.uint16)
    first_rows: list[int] = []  # first k rows with a non-zero norm (the answer for an all-zero query)
    for lo, bits in c_chunks:
        c = bf16_bits_to_f32(bits)
        cn = np.sqrt(np.einsum("ij,ij->i", c, c, dtype=np.float32))
        if len(first_rows) < k:
            first_rows.extend((np.flatnonzero(bits.any(axis=1))[: k - len(first_rows)] + lo).tolist())
        inv = np.where(cn > 0, 1.0 / np.where(cn > 0, cn, 1), 0).astype(np.float32)
        s = (qh @ c.T) * inv[None, :]
        s[:, cn == 0] = -np.inf
        m = s.shape[1]
        kk = min(keep, m)
        part = np.argpartition(s, m - kk, axis=1)[:, m - kk :]
        ps = np.take_along_axis(s, part, axis=1)
        cs = np.concatenate([cand_s, ps], axis=1)
        ci = np.concatenate([cand_i, part.astype(np.int64) + lo], axis=1)
        cb = np.concatenate([cand_bits, bits[part]], axis=1)
        order = np.argsort(-cs, axis=1, kind="stable")[:, :keep]
        cand_s = np.take_along_axis(cs, order, axis=1)
        cand_i = np.take_along_axis(ci, order, axis=1)
        cand_bits = np.take_along_axis(cb, order[:, :, None], axis=1)
    out_s = np.full((nq, k), -np.inf)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        if qn[r] == 0:  # all-zero query: every eligible row scores 0, lowest rows win (handled below)
            continue
        ok = cand_i[r] >= 0
        if not ok.any():
            continue
        s64 = _rescore_f64(q_bits[r], cand_bits[r][ok])
        rows = cand_i[r][ok]
        fin = np.isfinite(s64)
        ts, ti = _select_topk(s64[fin], rows[fin], k)
        out_s[r, : len(ts)] = ts
        out_i[r, : len(ti)] = ti
    for r in np.flatnonzero(qn == 0):
        out_s[r, : len(first_rows)] = 0.0
        out_i[r, : len(first_rows)] = first_rows
    return out_s, out_i


def merge_shard_topk(shard_scores, shard_rows, offsets, k: int):
    """Merge per-shard (score float64 [nq,k], local row [nq,k]) lists into the global top-k; mirrors the
    all-gather + merge of SURVEY.md section 8e.  Rows become global (local + offset)."""
    s = np.concatenate(shard_scores, axis=1)
    i = np.concatenate([np.where(r >= 0, r.astype(np.int64) + o, -1) for r, o in zip(shard_rows, offsets)], axis=1)
    nq = s.shape[0]
    out_s = np.full((nq, k), -np.inf)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        ok = i[r] >= 0
        ts, ti = _select_topk(s[r][ok], i[r][ok], k)
        out_s[r, : len(ts)] = ts
        out_i[r, : len(ti)] = ti
    return out_s, out_i


# --------------------------------------------------------------------------------------------------
# comparator
# --------------------------------------------------------------------------------------------------
def compare_topk(got_idx, got_score, ref_idx, ref_score, tie_tol: float = 0.0) -> dict:
    """Compare an engine result with the oracle's.

    strict_order : fraction of queries whose index list equals the oracle's element for element
    recall       : mean |got ∩ ref| / k
    tie_aware    : like strict_order, but a position may differ when the oracle scores of the two rows
                   differ by <= tie_tol (0 = only exact score ties are forgiven)
    max_abs_dscore : max |got_score - ref_score| over positions with equal indices
    """
    got_idx = np.asarray(got_idx, dtype=np.int64)
    ref_idx = np.asarray(ref_idx, dtype=np.int64)
    got_score = np.asarray(got_score, dtype=np.float64)
    ref_score = np.asarray(ref_score, dtype=np.float64)
    nq, k = ref_idx.shape
    strict = (got_idx == ref_idx).all(axis=1)
    recall = np.mean([len(np.intersect1d(got_idx[r], ref_idx[r])) / k for r in range(nq)]) if nq else 1.0
    same = got_idx == ref_idx
    both = same & np.isfinite(ref_score) & np.isfinite(got_score)
    dmax = float(np.abs(got_score - ref_score)[both].max()) if both.any() else 0.0
    tie_ok = same | (np.abs(got_score - ref_score) <= tie_tol)
    return {
        "strict_order": float(strict.mean()) if nq else 1.0,
        "recall": float(recall),
        "tie_aware": float(tie_ok.all(axis=1).mean()) if nq else 1.0,
        "max_abs_dscore": dmax,
        "mismatched_queries": np.flatnonzero(~strict).tolist()[:16],
    }
