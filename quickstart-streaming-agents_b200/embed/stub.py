"""Stand-ins for ``ML_PREDICT('llm_embedding_model', text)`` (terraform/lab2-vector-search/main.tf:253; model
DDL ``INPUT (text STRING) OUTPUT (embedding ARRAY<FLOAT>)`` terraform/core/main.tf:500,534; 1536-d
amazon.titan-embed-text-v1 / text-embedding-ada-002, terraform/core/main.tf:317,394).

The real call is an HTTPS round trip per row to a cloud model, which cannot sit on a timed path and cannot
run here at all, so (per BASELINE.json's north_star) the stage is stubbed:

* ``StubEmbedder``       deterministic text -> fp32 vector: token-hash bag of Gaussian directions, so texts that
                         share words land near each other (retrieval over the stub is meaningful, not random);
* ``PrecomputedEmbedder`` lookup of vectors computed elsewhere (text -> row of a matrix).
"""
from __future__ import annotations

import hashlib
import re

import numpy as np

_TOKEN = re.compile(r"[a-z0-9_]+")


def _seed(text: str) -> int:
    return int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little")


class StubEmbedder:
    def __init__(self, dim: int = 1536, cache_size: int = 100_000):
        self.dim = dim
        self._tok: dict[str, np.ndarray] = {}
        self._cache_size = cache_size

    def _token_vec(self, tok: str) -> np.ndarray:
        v = self._tok.get(tok)
        if v is None:
            v = np.random.default_rng(_seed("tok:" + tok)).standard_normal(self.dim).astype(np.float32)
            if len(self._tok) < self._cache_size:
                self._tok[tok] = v
        return v

    def embed(self, text: str) -> np.ndarray:
        """fp32 [dim].  Sum of per-token directions (sub-linear in term frequency) + a small text-specific term."""
        toks = _TOKEN.findall(text.lower())
        acc = 0.05 * np.random.default_rng(_seed(text)).standard_normal(self.dim).astype(np.float32)
        if toks:
            uniq, counts = np.unique(np.array(toks), return_counts=True)
            for t, c in zip(uniq.tolist(), counts.tolist()):
                acc = acc + np.float32(1.0 + np.log(c)) * self._token_vec(t)
        return acc.astype(np.float32)

    def embed_many(self, texts) -> np.ndarray:
        return np.stack([self.embed(t) for t in texts]) if len(texts) else np.empty((0, self.dim), np.float32)


class PrecomputedEmbedder:
    def __init__(self, texts, vectors: np.ndarray):
        self.dim = vectors.shape[1]
        self._rows = {t: i for i, t in enumerate(texts)}
        self._vec = np.ascontiguousarray(vectors, dtype=np.float32)

    def embed(self, text: str) -> np.ndarray:
        return self._vec[self._rows[text]]

    def embed_many(self, texts) -> np.ndarray:
        return self._vec[[self._rows[t] for t in texts]]
