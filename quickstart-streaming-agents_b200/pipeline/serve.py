"""Local serve loop for the Lab2 RAG pipeline.

What Confluent Cloud Flink runs as continuous statements, this process runs as consumer/producer stages over
topics with the same names and Avro schemas:

  documents        --embed(stub)-->  documents_embed  --sink-->  VectorTable (HBM)         LAB2-Walkthrough.md:41-51
  queries          --embed(stub)-->  queries_embed                                          main.tf:253
  queries_embed    --VECTOR_SEARCH_AGG(table, DESCRIPTOR(embedding), embedding, k)-->  search_results   main.tf:292
  search_results   --RAG prompt + generator(stub)-->  search_results_response              main.tf:331

Records on ``queries_embed`` / ``documents_embed`` may also come from outside (pre-computed embeddings).
Delivery is at-least-once: each stage commits its consumer offsets after its outputs are flushed.
A poison record (bad magic byte, truncated Avro, wrong embedding length) is quarantined to ``<topic>.dlq``
with the error text in its key, and the stage moves on.
"""
from __future__ import annotations

import logging
import struct
import time

import numpy as np

from ..embed.stub import StubEmbedder
from ..operator import VectorTable, flatten_search_results, rag_prompt, search_results_avro_body, vector_search_agg
from ..transport.filelog import Consumer, Producer
from ..wire import avro, schemas
from ..wire.registry import SchemaRegistry

log = logging.getLogger(__name__)


def stub_generator(prompt: str, rec: dict) -> str:
    """Stand-in for ml_predict('llm_textgen_model', prompt) (main.tf:331): a deterministic extractive answer that
    cites the retrieved sources, so the test shape "response is non-empty" (testing/e2e/test_lab2.py:112-135) and
    a human reading the topic both get something sensible."""
    parts = []
    for i in (1, 2, 3):
        if rec.get(f"document_id_{i}") is not None:
            chunk = (rec.get(f"chunk_{i}") or "").strip().replace("\n", " ")
            parts.append(f"[{rec[f'document_id_{i}']}] (score {rec[f'score_{i}']:.4f}): {chunk[:240]}")
    if not parts:
        return "The search results don't contain relevant information for this query."
    return "Based on the retrieved documents:\n" + "\n".join(parts)


class Codec:
    """Avro + Confluent framing for one log directory (schema ids from the registry stub, compiled codecs)."""

    def __init__(self, log_dir: str):
        self.registry = SchemaRegistry(log_dir)
        self._enc: dict[str, tuple[bytes, avro.CompiledSchema]] = {}
        self._dec: dict[int, avro.CompiledSchema] = {}

    def schema_id(self, topic: str) -> int:
        return struct.unpack(">I", self._encoder(topic)[0][1:5])[0]

    def _encoder(self, topic: str):
        e = self._enc.get(topic)
        if e is None:
            sid = self.registry.register(f"{topic}-value", schemas.TOPIC_SCHEMAS[topic])
            e = self._enc[topic] = (avro.frame(sid, b""), avro.CompiledSchema(schemas.TOPIC_SCHEMAS[topic]))
        return e

    def header(self, topic: str) -> bytes:
        return self._encoder(topic)[0]

    def encode(self, topic: str, record: dict) -> bytes:
        header, cs = self._encoder(topic)
        return cs.encode(record, prefix=header)

    def decode(self, raw: bytes) -> dict:
        if len(raw) < 5:
            raise avro.AvroError(f"Avro payload too short ({len(raw)} bytes)")
        if raw[0] != avro.MAGIC:
            raise avro.AvroError(f"Invalid Avro magic byte: {raw[0]}")
        sid = struct.unpack_from(">I", raw, 1)[0]
        cs = self._dec.get(sid)
        if cs is None:
            cs = self._dec[sid] = avro.CompiledSchema(self.registry.get(sid))
        return cs.decode(raw, 5)


class Lab2Pipeline:
    def __init__(self, log_dir: str, table: VectorTable, embedder=None, k: int = 3, max_batch: int = 1024,
                 group: str = "sa-lab2", generator=stub_generator):
        self.log_dir = log_dir
        self.table = table
        self.embedder = embedder or StubEmbedder(table.index.dim)
        self.k = k
        self.max_batch = max_batch
        self.generator = generator
        self.codec = Codec(log_dir)
        self.producer = Producer({"log.dir": log_dir})
        conf = {"log.dir": log_dir, "group.id": group, "auto.offset.reset": "earliest", "enable.auto.commit": False}
        self.consumers = {}
        for t in ("documents", "documents_embed", "queries", "queries_embed", "search_results"):
            c = Consumer(conf)
            c.subscribe([t])
            self.consumers[t] = c
        self.stats = {"documents": 0, "queries": 0, "searches": 0, "responses": 0, "quarantined": 0,
                      "search_seconds": 0.0}

    # ------------------------------------------------------------------ helpers
    def _decode_all(self, topic: str, msgs):
        good = []
        for m in msgs:
            try:
                good.append((m, self.codec.decode(m.value())))
            except Exception as e:  # poison message: quarantine, keep going
                self.stats["quarantined"] += 1
                log.warning("quarantined %s[%d]@%d: %s", topic, m.partition(), m.offset(), e)
                self.producer.produce(f"{topic}.dlq", key=str(e), value=m.value())
        return good

    def _drain(self, topic: str):
        c = self.consumers[topic]
        msgs = c.consume(self.max_batch, 0.0)
        return c, msgs, self._decode_all(topic, msgs)

    def _check_vec(self, topic, m, vec):
        dim = self.table.index.dim
        if vec is None or len(vec) != dim or not np.isfinite(vec).all():
            self.stats["quarantined"] += 1
            why = f"embedding must be {dim} finite floats"
            log.warning("quarantined %s@%d: %s", topic, m.offset(), why)
            self.producer.produce(f"{topic}.dlq", key=why, value=m.value())
            return False
        return True

    # ------------------------------------------------------------------ stages
    def stage_documents(self) -> int:
        c, msgs, recs = self._drain("documents")
        for m, r in recs:
            text = r.get("document_text") or ""
            vec = self.embedder.embed(text)
            out = {"document_id": r.get("document_id"), "chunk": text, "embedding": vec}
            out.update({c: r.get(c) for c in schemas.METADATA_COLUMNS})   # Lab4-style metadata passthrough
            self.producer.produce("documents_embed", key=m.key(), value=self.codec.encode("documents_embed", out))
        if msgs:
            self.producer.flush()
            c.commit()
        return len(msgs)

    def stage_sink(self) -> int:
        c, msgs, recs = self._drain("documents_embed")
        ids, chunks, vecs, metas = [], [], [], []
        for m, r in recs:
            vec = r.get("embedding")
            if not self._check_vec("documents_embed", m, vec):
                continue
            ids.append(r.get("document_id"))
            chunks.append(r.get("chunk"))
            vecs.append(vec)
            metas.append({c: r.get(c) for c in schemas.METADATA_COLUMNS})
        if ids:
            self.table.upsert_many(ids, chunks, np.stack(vecs), metas)
            self.stats["documents"] += len(ids)
        if msgs:
            self.producer.flush()
            c.commit()
        return len(msgs)

    def stage_queries(self) -> int:
        c, msgs, recs = self._drain("queries")
        for m, r in recs:
            q = r.get("query") or ""
            self.producer.produce("queries_embed", value=self.codec.encode(
                "queries_embed", {"query": q, "embedding": self.embedder.embed(q)}))
        if msgs:
            self.stats["queries"] += len(recs)
            self.producer.flush()
            c.commit()
        return len(msgs)

    def _decode_queries_embed_fast(self, msgs):
        """Batch decode of `queries_embed` records in their usual shape -- non-null query string, non-null single-block
        array of dim non-null floats -- with one strided numpy gather for all embeddings.  Returns (texts, vectors,
        leftovers): messages in any other shape are returned in `leftovers` for the generic per-record path."""
        dim = self.table.index.dim
        sid = self.codec.schema_id("queries_embed")
        header = bytes([0]) + struct.pack(">I", sid)
        block = bytearray()
        avro.write_long(block, dim)
        block = bytes(block)
        span = 5 * dim
        texts, views, leftovers = [], [], []
        for m in msgs:
            raw = m.value()
            try:
                if raw[:5] != header or raw[5] != 2:
                    raise ValueError
                n, pos = avro.read_long(raw, 6)
                end = pos + n
                if raw[end] != 2 or raw[end + 1:end + 1 + len(block)] != block:
                    raise ValueError
                off = end + 1 + len(block)
                if len(raw) != off + span + 1 or raw[-1] != 0:
                    raise ValueError
                text = raw[pos:end].decode("utf-8")
            except (ValueError, IndexError, TypeError):
                leftovers.append(m)
                continue
            texts.append(text)
            views.append(np.frombuffer(raw, dtype=np.uint8, count=span, offset=off))
        if not views:
            return [], np.empty((0, dim), np.float32), leftovers
        flat = np.stack(views).reshape(len(views), dim, 5)
        ok = (flat[:, :, 0] == 2).all(axis=1)
        vecs = np.ascontiguousarray(flat[:, :, 1:]).view("<f4").reshape(len(views), dim)
        ok &= np.isfinite(vecs).all(axis=1)
        if not ok.all():  # a null / non-finite item somewhere: let the generic path judge those records
            bad = set(np.flatnonzero(~ok).tolist())
            fast_msgs = [m for m in msgs if m not in leftovers]
            leftovers.extend(fast_msgs[i] for i in sorted(bad))
            keep = [i for i in range(len(views)) if i not in bad]
            texts = [texts[i] for i in keep]
            vecs = vecs[keep]
        return texts, vecs, leftovers

    def _decode_batch(self, msgs):
        """queries_embed messages -> (texts, vectors): fast batch path, generic codec (and quarantine) for the rest."""
        texts, vecs, leftovers = self._decode_queries_embed_fast(msgs)
        if leftovers:
            slow_t, slow_v = [], []
            for m, r in self._decode_all("queries_embed", leftovers):
                vec = r.get("embedding")
                if self._check_vec("queries_embed", m, vec):
                    slow_t.append(r.get("query"))
                    slow_v.append(vec)
            if slow_v:
                texts = texts + slow_t
                vecs = np.concatenate([vecs, np.stack(slow_v)]) if len(vecs) else np.stack(slow_v)
        return texts, np.ascontiguousarray(vecs, dtype=np.float32)

    def _emit_results(self, texts, score, idx) -> None:
        header = self.codec.header("search_results")
        n_out = schemas.RESULTS_PER_QUERY
        for r, q in enumerate(texts):
            self.producer.produce("search_results",
                                  value=header + search_results_avro_body(self.table, q, score[r], idx[r], n_out))
        self.stats["searches"] += len(texts)

    def stage_search(self) -> int:
        """queries_embed -> VECTOR_SEARCH_AGG -> search_results.  When the index offers the split host call
        (``search_host_submit`` / ``search_host_wait``), batches are software-pipelined: batch i+1 is read and decoded
        while the GPU searches batch i.  Offsets of a batch are committed only after its results are flushed."""
        c = self.consumers["queries_embed"]
        index = self.table.index
        pipelined = hasattr(index, "search_host_submit")
        total = 0
        pending = None  # (messages, texts, slot) of the batch the GPU is working on
        slot = 0
        while True:
            msgs = c.consume(self.max_batch, 0.0)
            batch = None
            if msgs:
                total += len(msgs)
                texts, vecs = self._decode_batch(msgs)
                batch = (msgs, texts, vecs)
            if batch is not None and len(batch[1]) and pipelined:
                t0 = time.perf_counter()
                index.search_host_submit(batch[2], self.k, slot)
                self.stats["search_seconds"] += time.perf_counter() - t0
            if pending is not None:  # collect the previous batch while the new one runs
                p_msgs, p_texts, p_slot = pending
                t0 = time.perf_counter()
                score, idx = index.search_host_wait(p_slot)
                self.stats["search_seconds"] += time.perf_counter() - t0
                self._emit_results(p_texts, score, idx)
                self.producer.flush()
                c.commit_offsets(p_msgs)
                pending = None
            if batch is None:
                break
            msgs, texts, vecs = batch
            if len(texts) and pipelined:
                pending = (msgs, texts, slot)
                slot ^= 1
            else:
                if len(texts):
                    t0 = time.perf_counter()
                    score, idx = index.search_host(vecs, self.k)
                    self.stats["search_seconds"] += time.perf_counter() - t0
                    self._emit_results(texts, score, idx)
                self.producer.flush()  # also carries any quarantined records of this batch
                c.commit_offsets(msgs)
        return total

    def stage_response(self) -> int:
        c, msgs, recs = self._drain("search_results")
        for m, r in recs:
            out = dict(r)
            out["response"] = self.generator(rag_prompt(r), r)
            self.producer.produce("search_results_response", value=self.codec.encode("search_results_response", out))
            self.stats["responses"] += 1
        if msgs:
            self.producer.flush()
            c.commit()
        return len(msgs)

    # ------------------------------------------------------------------ loop
    def run_once(self) -> int:
        """One pass over all stages in topological order; returns the number of records moved.  The ingest stages
        are drained completely first, so a query is searched against every document that was already on the
        log when the pass started (Flink gives no such ordering across topics; this is strictly stronger)."""
        moved = 0
        while True:
            n = self.stage_documents() + self.stage_sink()
            moved += n
            if n == 0:
                break
        return moved + self.stage_queries() + self.stage_search() + self.stage_response()

    def run_until_idle(self, max_passes: int = 1000) -> int:
        total = 0
        for _ in range(max_passes):
            n = self.run_once()
            total += n
            if n == 0:
                break
        return total

    def run_forever(self, idle_sleep: float = 0.05, stop=lambda: False) -> None:
        while not stop():
            if self.run_once() == 0:
                time.sleep(idle_sleep)
