"""Drop-in CLIs + the local Lab2 topic graph, mirroring the acceptance shape of the reference's only test of
this path (testing/e2e/test_lab2.py:74-135): `queries` has >= 1 message -> `search_results` has >= 1 row ->
`search_results_response[0].response` is non-empty.  On a CPU box the vector index is an oracle-backed double
(tests/doubles.py); `test_pipeline_on_gpu` runs the same flow through the CUDA engine."""
import json
import struct

import numpy as np
import pytest

from qsa_b200.embed.stub import StubEmbedder
from qsa_b200.operator import VectorTable, flatten_search_results, rag_prompt, vector_search_agg
from qsa_b200.pipeline.serve import Codec, Lab2Pipeline
from qsa_b200.transport.filelog import Broker, Consumer, Producer, TopicPartition
from qsa_b200.wire import avro, schemas
from scripts import lab2_publish_queries, publish_docs

from doubles import OracleIndex, PipelinedOracleIndex

TOPICS = ["sql window functions tumble hop session", "watermarks and event time late data", "kafka connector properties",
          "state ttl and checkpoints", "user defined functions in java and python", "joins interval temporal lookup",
          "json avro protobuf formats schema registry", "flink sql client and statements"]


def write_docs(d, n=64):
    d.mkdir(parents=True, exist_ok=True)
    for i in range(n):
        topic = TOPICS[i % len(TOPICS)]
        (d / f"flink_doc_{i:03d}_chunk_{i % 3}.md").write_text(
            f"---\ntitle: {topic.title()} {i}\npages: '{i}-{i + 1}'\nsection_reference: S{i}\nchar_count: {200 + i}\n"
            f"policy_keywords: [{topic.split()[0]}, sql]\n---\nThis chunk number {i} explains {topic}: details, "
            f"examples and caveats about {topic}.")
    (d / "plain.md").write_text("No front matter here, just text about watermarks.")


def test_publish_docs_cli_flags_exit_codes_and_records(tmp_path, capsys):
    docs, logd = tmp_path / "docs", tmp_path / "topics"
    write_docs(docs, 30)
    assert publish_docs.main(["--log-dir", str(logd)]) == 1                       # neither --lab2/--lab3 nor --docs-dir
    assert publish_docs.main(["--lab2", "--project-root", str(tmp_path), "--log-dir", str(logd)]) == 1   # no assets/
    assert publish_docs.main(["--docs-dir", str(tmp_path / "missing"), "--log-dir", str(logd)]) == 1
    with pytest.raises(SystemExit):
        publish_docs.main(["--lab2", "--lab3"])                                   # mutually exclusive
    assert publish_docs.main(["--docs-dir", str(docs), "--dry-run", "--log-dir", str(logd)]) == 0
    assert "DRY RUN COMPLETE" in capsys.readouterr().out and Broker(str(logd)).count("documents") == 0
    assert publish_docs.main(["--docs-dir", str(docs), "--workers", "4", "--log-dir", str(logd)]) == 0
    out = capsys.readouterr().out
    assert "PUBLISHING SUMMARY" in out and "Total files:      31" in out and "Failed:           0" in out
    # standard location assets/lab2/flink_docs is auto-detected with --lab2
    write_docs(tmp_path / "assets" / "lab2" / "flink_docs", 3)
    assert publish_docs.main(["--lab2", "--project-root", str(tmp_path), "--log-dir", str(tmp_path / "t2")]) == 0
    assert Broker(str(tmp_path / "t2")).count("documents") == 4

    c = Consumer({"log.dir": str(logd), "group.id": "t"}); c.subscribe(["documents"])
    msgs = c.consume(100, 0.0)
    assert len(msgs) == 31
    codec = Codec(str(logd))
    by_key = {m.key().decode(): codec.decode(m.value()) for m in msgs}
    r = by_key["flink_doc_007_chunk_1.md"]                                        # key = document_id = file name
    assert r["document_id"] == "flink_doc_007_chunk_1.md" and r["title"].endswith(" 7")
    assert r["document_text"].startswith("# " + r["title"] + "\n\nThis chunk number 7")
    assert r["pages"] == "7-8" and r["char_count"] == 207 and r["policy_keywords"] == ["flink", "sql"]
    assert r["fraud_categories"] is None                                          # empty list -> null, as the reference
    p = by_key["plain.md"]
    assert p["title"] == "" and p["document_text"].startswith("No front matter") and p["pages"] is None
    assert msgs[0].value()[0] == 0                                                # Confluent magic byte


def test_publish_queries_cli(tmp_path, capsys, monkeypatch):
    logd = str(tmp_path / "topics")
    assert lab2_publish_queries.main(["How do window functions work?", "--log-dir", logd]) == 0
    assert "✓ Query published successfully!" in capsys.readouterr().out
    assert lab2_publish_queries.main(["azure", "What is watermarking?", "--topic", "queries", "--log-dir", logd]) == 0
    monkeypatch.setattr("builtins.input", lambda prompt="": "  interactive question  ")
    assert lab2_publish_queries.main(["aws", "--log-dir", logd]) == 0             # prompt when QUERY is absent
    monkeypatch.setattr("builtins.input", lambda prompt="": "   ")
    assert lab2_publish_queries.main(["--log-dir", logd]) == 1
    assert "No query provided" in capsys.readouterr().out

    def eof(prompt=""):
        raise EOFError
    monkeypatch.setattr("builtins.input", eof)
    assert lab2_publish_queries.main(["--log-dir", logd]) == 1
    c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["queries"])
    codec = Codec(logd)
    got = [codec.decode(m.value())["query"] for m in c.consume(10, 0.0)]
    assert got == ["How do window functions work?", "What is watermarking?", "interactive question"]
    # byte-exact record: 00 | schema id | 02 (union branch 1) | len | utf-8
    m = Consumer({"log.dir": logd, "group.id": "u"}); m.subscribe(["queries"])
    raw = m.consume(1, 0.0)[0].value()
    assert raw[:1] == b"\x00" and raw[5:7] == bytes([2, 2 * len("How do window functions work?")])


def run_lab2(tmp_path, index, n_docs=64):
    docs, logd = tmp_path / "docs", str(tmp_path / "topics")
    write_docs(docs, n_docs)
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0
    queries = ["How do tumble and hop window functions work?", "What happens to late data and watermarks?",
               "Which formats work with schema registry?"]
    for q in queries:
        assert lab2_publish_queries.main([q, "--log-dir", logd]) == 0
    table = VectorTable(index)
    pipe = Lab2Pipeline(logd, table, embedder=StubEmbedder(index.dim), k=3, max_batch=16)
    moved = pipe.run_until_idle()
    return logd, pipe, table, queries, moved


def check_lab2_outputs(logd, pipe, table, queries, n_docs):
    b = Broker(logd)
    # test_lab2.py:74-79  queries has >= 1 message (counted via watermarks, kafka_helper.py:88-118)
    lo, hi = b.get_watermark_offsets(TopicPartition("queries", 0))
    assert hi - lo == len(queries) >= 1
    assert len(table) == n_docs + 1 and pipe.stats["documents"] == n_docs + 1
    assert b.count("documents_embed") == n_docs + 1 and b.count("queries_embed") == len(queries)
    # test_lab2.py:99-110  search_results has >= 1 row
    assert b.count("search_results") == len(queries)
    c = Consumer({"log.dir": logd, "group.id": "check"}); c.subscribe(["search_results_response"])
    codec = Codec(logd)
    rows = [codec.decode(m.value()) for m in c.consume(10, 0.0)]
    assert [r["query"] for r in rows] == queries
    for r in rows:
        # test_lab2.py:112-135  response non-empty
        assert r["response"] and r["document_id_1"] in r["response"]
        assert r["score_1"] >= r["score_2"] >= r["score_3"] > 0
        assert len({r["document_id_1"], r["document_id_2"], r["document_id_3"]}) == 3
        assert r["chunk_1"].startswith("# ")
    # retrieval over the stub embedder is meaningful: the window question retrieves window chunks
    assert "window functions" in rows[0]["chunk_1"].lower() and "watermarks" in rows[1]["chunk_1"].lower()
    return rows


def test_lab2_pipeline_end_to_end_cpu_plumbing(tmp_path):
    logd, pipe, table, queries, moved = run_lab2(tmp_path, OracleIndex(1536))
    rows = check_lab2_outputs(logd, pipe, table, queries, 64)
    assert moved == 65 * 2 + 3 * 3
    # the operator's answer equals the oracle's brute force over the same stub vectors
    emb = StubEmbedder(1536)
    hits = vector_search_agg(table, "embedding", emb.embed(queries[0]), 10)[0]
    assert [h.document_id for h in hits[:3]] == [rows[0][f"document_id_{i}"] for i in (1, 2, 3)]
    assert [h.score for h in hits] == sorted((h.score for h in hits), reverse=True) and len(hits) == 10
    with pytest.raises(ValueError):
        vector_search_agg(table, "not_a_column", emb.embed("x"), 3)
    # idempotent: nothing pending -> nothing moves; re-publishing a document replaces its row (upsert)
    assert pipe.run_once() == 0
    p = Producer({"log.dir": logd})
    codec = Codec(logd)
    p.produce("documents", key="plain.md", value=codec.encode("documents", {
        "document_id": "plain.md", "document_text": "Completely new text about session windows", "pages": None,
        "section_reference": None, "title": "", "fraud_categories": None, "policy_keywords": None, "char_count": None}))
    p.flush()
    pipe.run_until_idle()
    assert len(table) == 66 and table.document_id.count("plain.md") == 2
    hits = vector_search_agg(table, "embedding", emb.embed("Completely new text about session windows"), 3)[0]
    assert hits[0].document_id == "plain.md" and hits[0].row == 65
    assert all(h.row != table.document_id.index("plain.md") for h in hits)       # the old row is tombstoned
    table.clear()
    assert len(table) == 0 and len(table.index) == 0


def test_poison_records_are_quarantined_not_fatal(tmp_path):
    logd = str(tmp_path / "topics")
    table = VectorTable(OracleIndex(64))
    pipe = Lab2Pipeline(logd, table, embedder=StubEmbedder(64), k=3)
    codec = Codec(logd)
    p = Producer({"log.dir": logd})
    good = codec.encode("documents_embed", {"document_id": "ok", "chunk": "c", "embedding": np.ones(64, np.float32)})
    p.produce("documents_embed", value=good)
    p.produce("documents_embed", value=b"\x07garbage")                                   # bad magic byte
    p.produce("documents_embed", value=good[:40])                                        # truncated Avro
    p.produce("documents_embed", value=codec.encode("documents_embed", {
        "document_id": "short", "chunk": "c", "embedding": np.ones(8, np.float32)}))     # wrong dimension
    p.produce("queries_embed", value=codec.encode("queries_embed", {"query": "q", "embedding": np.ones(64, np.float32)}))
    p.produce("queries_embed", value=codec.encode("queries_embed", {"query": "bad", "embedding": None}))
    p.flush()
    pipe.run_until_idle()
    b = Broker(logd)
    assert len(table) == 1 and pipe.stats["quarantined"] == 4
    assert b.count("documents_embed.dlq") == 3 and b.count("queries_embed.dlq") == 1
    assert b.count("search_results") == 1 and b.count("search_results_response") == 1


def test_flatten_and_prompt_follow_the_sql():
    from qsa_b200.operator import SearchHit
    hits = [SearchHit("a.md", "chunk a", 0.9, 0), SearchHit("b.md", "chunk b", 0.8, 1)]
    rec = flatten_search_results("q?", hits)
    assert list(rec) == [f["name"] for f in schemas.SEARCH_RESULTS_VALUE["fields"]]
    assert rec["document_id_3"] is None and rec["score_2"] == 0.8
    prompt = rag_prompt(rec)
    assert prompt.startswith("Based on the following search results, provide a helpful and comprehensive response")
    assert "USER QUERY: q?\n\nSEARCH RESULTS:\n\nDocument 1 (Similarity Score: 0.9):\nSource: a.md\nContent: chunk a" in prompt
    assert prompt.endswith("- If the search results don't contain relevant information, say so clearly\n\nRESPONSE:")
    body = avro.encode(schemas.SEARCH_RESULTS_VALUE, rec)
    assert avro.decode(schemas.SEARCH_RESULTS_VALUE, body) == rec


def test_lab3_and_lab4_operator_shapes(tmp_path):
    """The other two call sites of the operator: Lab3 (top-3 chunks of an events collection, LAB3-Walkthrough.md:343-350)
    and Lab4 (metadata columns projected next to chunk/score, LAB4-Walkthrough.md:280-309)."""
    from qsa_b200.operator import project_search_results
    docs, logd = tmp_path / "docs", str(tmp_path / "topics")
    write_docs(docs, 24)
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0
    table = VectorTable(OracleIndex(768), name="fema_policies_vectordb")        # 768-d: dim is a runtime parameter
    emb = StubEmbedder(768)
    pipe = Lab2Pipeline(logd, table, embedder=emb, k=5)
    pipe.run_until_idle()
    hits = vector_search_agg(table, "embedding", emb.embed("late data and watermarks in event time"), 5)[0]
    assert len(hits) == 5 and "watermarks" in hits[0].chunk.lower()
    rec = project_search_results(hits, {"chunk": "policy_chunk", "score": "policy_score", "pages": "policy_pages",
                                        "section_reference": "policy_section", "title": "policy_title",
                                        "fraud_categories": "policy_fraud_cats", "policy_keywords": "policy_keywords"}, 3)
    assert set(rec) == {f"{p}_{i}" for i in (1, 2, 3) for p in ("policy_chunk", "policy_score", "policy_pages",
                        "policy_section", "policy_title", "policy_fraud_cats", "policy_keywords")}
    assert rec["policy_title_1"].startswith("Watermarks And Event Time") and rec["policy_keywords_1"] == ["watermarks", "sql"]
    assert rec["policy_pages_1"] is not None and rec["policy_section_1"].startswith("S") and rec["policy_fraud_cats_1"] is None
    assert rec["policy_score_1"] >= rec["policy_score_2"] >= rec["policy_score_3"]
    lab3 = project_search_results(hits, {"chunk": "top_chunk"}, 3)                 # testing/e2e/test_lab3.py:232-268
    assert lab3["top_chunk_1"] and lab3["top_chunk_2"]


@pytest.mark.gpu
def test_pipeline_on_gpu(tmp_path):
    from qsa_b200.engine import VectorIndex
    ix = VectorIndex(dim=1536, capacity=4096, max_batch=64, max_k=10)
    logd, pipe, table, queries, moved = run_lab2(tmp_path, ix)
    rows = check_lab2_outputs(logd, pipe, table, queries, 64)
    # same answers as the oracle-backed run over identical stub vectors
    ref = OracleIndex(1536)
    emb = StubEmbedder(1536)
    ref.append(np.stack([emb.embed(c) for c in table.chunk]))
    for q, r in zip(queries, rows):
        s, i = ref.search_host(emb.embed(q)[None, :], 3)
        assert [table.document_id[j] for j in i[0]] == [r["document_id_1"], r["document_id_2"], r["document_id_3"]]
        assert abs(s[0, 0] - r["score_1"]) < 1e-6
    ix.close()


def test_search_stage_fast_paths_equal_the_generic_codec(tmp_path):
    """The search stage decodes `queries_embed` in batch and emits `search_results` from pre-serialised columns; both
    shortcuts must be byte-for-byte what the generic codec produces, and odd-shaped records must fall back to it."""
    import struct
    from qsa_b200.operator import search_results_avro_body
    logd = str(tmp_path / "topics")
    idx = OracleIndex(64)
    table = VectorTable(idx)
    g = np.random.default_rng(0)
    table.upsert_many([f"doc {i} é.md" for i in range(40)] + [None], [f"chunk {i} " * (i % 7) for i in range(40)] + [None],
                      g.standard_normal((41, 64)).astype(np.float32))
    pipe = Lab2Pipeline(logd, table, embedder=StubEmbedder(64), k=3)
    codec = Codec(logd)
    # (1) emit: concatenation of pre-serialised columns == generic encode of the flattened dict
    q = g.standard_normal((5, 64)).astype(np.float32)
    score, rows = idx.search_host(q, 3)
    hits = vector_search_agg(table, "embedding", q, 3)
    for r in range(5):
        fast = search_results_avro_body(table, f"q{r}", score[r], rows[r], 3)
        assert fast == avro.encode(schemas.SEARCH_RESULTS_VALUE, flatten_search_results(f"q{r}", hits[r], 3))
    short = search_results_avro_body(table, None, score[0][:1], np.array([rows[0][0], -1, -1]), 3)
    assert avro.decode(schemas.SEARCH_RESULTS_VALUE, short)["document_id_2"] is None
    # (2) decode: usual shape goes through the batch path, unusual shapes through the generic one, same answers
    p = Producer({"log.dir": logd})
    vec = g.standard_normal(64).astype(np.float32)
    p.produce("queries_embed", value=codec.encode("queries_embed", {"query": "usual", "embedding": vec}))
    multi = bytearray(codec.header("queries_embed") + b"\x02" + bytes([2 * len("two blocks")]) + b"two blocks" + b"\x02")
    for part in (vec[:10], vec[10:]):                       # the same vector as a two-block array
        avro.write_long(multi, len(part))
        for x in part:
            multi += b"\x02" + struct.pack("<f", float(x))
    multi += b"\x00"
    p.produce("queries_embed", value=bytes(multi))
    p.produce("queries_embed", value=codec.encode("queries_embed", {"query": None, "embedding": vec}))   # null query text
    p.flush()
    assert pipe.stage_search() == 3 and pipe.stats["quarantined"] == 0
    c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results"])
    out = [codec.decode(m.value()) for m in c.consume(10, 0.0)]
    assert [o["query"] for o in out] == ["usual", "two blocks", None]
    assert out[0]["document_id_1"] == out[1]["document_id_1"] == out[2]["document_id_1"]
    assert out[0]["score_1"] == out[1]["score_1"]


def test_search_stage_is_pipelined_and_commits_per_batch(tmp_path):
    """With the split host call, batch i+1 is decoded while batch i is searched; results, order and committed offsets
    are the same as the blocking path, and a crash between batches redelivers only unfinished batches."""
    outs = {}
    for name, cls in (("blocking", OracleIndex), ("pipelined", PipelinedOracleIndex)):
        logd = str(tmp_path / name)
        idx = cls(64)
        table = VectorTable(idx)
        g = np.random.default_rng(3)
        table.upsert_many([f"d{i}" for i in range(200)], [f"c{i}" for i in range(200)], g.standard_normal((200, 64)).astype(np.float32))
        pipe = Lab2Pipeline(logd, table, embedder=StubEmbedder(64), k=3, max_batch=16)
        codec = Codec(logd)
        p = Producer({"log.dir": logd})
        for i in range(70):                                   # 4 full batches + one of 6
            p.produce("queries_embed", value=codec.encode("queries_embed", {"query": f"q{i}", "embedding": g.standard_normal(64).astype(np.float32)}))
        p.produce("queries_embed", value=b"\x01poison")
        p.flush()
        assert pipe.stage_search() == 71 and pipe.stats["searches"] == 70 and pipe.stats["quarantined"] == 1
        assert Broker(logd).committed("sa-lab2")["queries_embed-0"] == 71
        c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results"])
        outs[name] = [codec.decode(m.value()) for m in c.consume(100, 0.0)]
        if name == "pipelined":
            assert idx.max_inflight == 2 and not idx._slots   # batch i+1 is queued before batch i is collected
        assert pipe.stage_search() == 0
    assert [o["query"] for o in outs["blocking"]] == [f"q{i}" for i in range(70)]
    g2 = np.random.default_rng(3)                              # same seeds -> same vectors -> same answers
    assert [(o["query"], o["document_id_1"]) for o in outs["blocking"]] == [(o["query"], o["document_id_1"]) for o in outs["pipelined"]]


@pytest.mark.gpu
def test_sa_serve_cli_once_on_gpu(tmp_path, capsys):
    """The serve CLI end to end: publish with the drop-in CLIs, `sa_serve --once`, read the result topics."""
    from scripts import sa_serve
    docs, logd = tmp_path / "docs", str(tmp_path / "topics")
    write_docs(docs, 40)
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0
    assert lab2_publish_queries.main(["How do tumble windows work?", "--log-dir", logd]) == 0
    capsys.readouterr()
    assert sa_serve.main(["--log-dir", logd, "--once", "--capacity", "1024", "--max-batch", "64", "--k", "3"]) == 0
    stats = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert stats["documents"] == 41 and stats["searches"] == 1 and stats["responses"] == 1 and stats["quarantined"] == 0
    c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results_response"])
    row = Codec(logd).decode(c.consume(1, 0.0)[0].value())
    assert row["query"] == "How do tumble windows work?" and row["response"] and "window functions" in row["chunk_1"].lower()


def test_table_checkpoint_resume_and_atlas_score(tmp_path):
    """Checkpoint / resume of the vector table (side columns here; the HBM half is covered on the GPU), tombstones
    survive, and the optional Atlas score normalisation."""
    from qsa_b200.operator import atlas_score

    class SnapIndex(OracleIndex):
        def snapshot(self, path):
            np.savez(path, rows=self.bits)
            return len(self.bits)

        def restore(self, path):
            self.bits = np.load(path)["rows"]
            return len(self.bits)

    g = np.random.default_rng(5)
    t = VectorTable(SnapIndex(64))
    vecs = g.standard_normal((30, 64)).astype(np.float32)
    t.upsert_many([f"d{i}" for i in range(30)], [f"c{i} é" for i in range(30)], vecs, [{"pages": str(i)} for i in range(30)])
    t.upsert_many(["d7"], ["c7 new"], g.standard_normal((1, 64)).astype(np.float32), [{"pages": "x"}])   # tombstones row 7
    assert t.save(str(tmp_path / "ckpt")) == 31
    t2 = VectorTable(SnapIndex(64))
    assert t2.load(str(tmp_path / "ckpt")) == 31
    assert t2.document_id == t.document_id and t2.chunk == t.chunk and t2.metadata == t.metadata
    assert t2._row_of["d7"] == 30 and t2.avro_chunk == t.avro_chunk
    q = vecs[7:8]
    a, b = vector_search_agg(t, "embedding", q, 5)[0], vector_search_agg(t2, "embedding", q, 5)[0]
    assert [(h.row, h.score) for h in a] == [(h.row, h.score) for h in b] and all(h.row != 7 for h in b)
    t2.upsert_many(["d3"], ["c3 newer"], g.standard_normal((1, 64)).astype(np.float32))    # upserts keep working after resume
    assert t2._row_of["d3"] == 31 and (t2.index.bits[3] == 0).all()
    raw = vector_search_agg(t, "embedding", q, 3)[0]
    atl = vector_search_agg(t, "embedding", q, 3, score_mode="atlas")[0]
    assert [h.row for h in raw] == [h.row for h in atl]
    assert all(abs(x.score - atlas_score(y.score)) < 1e-12 and 0 <= x.score <= 1 for x, y in zip(atl, raw))
    with pytest.raises(ValueError):
        vector_search_agg(t, "embedding", q, 3, score_mode="dot")


def test_sa_serve_cli_snapshot_resume_with_a_double(tmp_path, capsys, monkeypatch):
    """`sa_serve --once --snapshot-dir`: the second run resumes from the checkpoint (no replay of `documents`, offsets
    continue) and serves new queries.  The engine is replaced by an oracle-backed double so this runs without a GPU."""
    import qsa_b200.engine as engine_mod
    from scripts import sa_serve

    class SnapIndex(OracleIndex):
        def __init__(self, dim=1536, capacity=0, max_batch=0, max_k=0, device=None):
            super().__init__(dim, capacity)

        def snapshot(self, path):
            np.savez(path, rows=self.bits)
            return len(self.bits)

        def restore(self, path):
            self.bits = np.load(path)["rows"]
            return len(self.bits)

    monkeypatch.setattr(engine_mod, "VectorIndex", SnapIndex)
    docs, logd, snap = tmp_path / "docs", str(tmp_path / "topics"), str(tmp_path / "ckpt")
    write_docs(docs, 20)
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0
    assert lab2_publish_queries.main(["What about watermarks?", "--log-dir", logd]) == 0
    capsys.readouterr()
    assert sa_serve.main(["--log-dir", logd, "--once", "--snapshot-dir", snap]) == 0
    first = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert first["documents"] == 21 and first["searches"] == 1
    man = json.load(open(tmp_path / "ckpt" / "manifest.json"))
    assert man["rows"] == 21 and man["source_offsets"] == {"documents_embed-0": 21}
    assert all((tmp_path / "ckpt" / f).exists() for f in man["files"].values())
    # second process: nothing new on `documents`, one new query; the table comes from the checkpoint
    assert lab2_publish_queries.main(["How do session windows work?", "--log-dir", logd]) == 0
    capsys.readouterr()
    assert sa_serve.main(["--log-dir", logd, "--once", "--snapshot-dir", snap]) == 0
    cap = capsys.readouterr()
    second = json.loads(cap.out.strip().splitlines()[-1])
    assert "resumed 21 rows" in cap.err
    assert second["documents"] == 0 and second["searches"] == 1 and second["responses"] == 1
    c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results"])
    rows = [Codec(logd).decode(m.value()) for m in c.consume(10, 0.0)]
    assert [r["query"] for r in rows] == ["What about watermarks?", "How do session windows work?"]
    assert all(r["document_id_1"] for r in rows)
    # third process: one more document arrived after the checkpoint -> the sink resumes at the checkpoint's offsets
    (docs / "late.md").write_text("---\ntitle: Late\n---\nSession windows group events by gaps of inactivity.")
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0        # republish: 22 records, 21 upserts
    capsys.readouterr()
    assert sa_serve.main(["--log-dir", logd, "--once", "--snapshot-dir", snap]) == 0
    third = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert third["documents"] == 22
    man2 = json.load(open(tmp_path / "ckpt" / "manifest.json"))
    assert man2["generation"] == man["generation"] + 2 and man2["source_offsets"] == {"documents_embed-0": 43}
    assert not (tmp_path / "ckpt" / man["files"]["columns"]).exists()      # old generations are removed


def test_restart_without_a_checkpoint_rebuilds_the_table_from_the_log(tmp_path, capsys, monkeypatch):
    """the table is volatile (HBM), the log is durable.  A restarted sa_serve without --snapshot-dir
    must rebuild the table from `documents_embed` instead of trusting the sink's committed offsets -- the README
    quick-start runs `sa_serve --once` repeatedly, and the second run used to answer with null documents."""
    import qsa_b200.engine as engine_mod
    from scripts import sa_serve

    class Index(OracleIndex):
        def __init__(self, dim=1536, capacity=0, max_batch=0, max_k=0, device=None):
            super().__init__(dim, capacity)

    monkeypatch.setattr(engine_mod, "VectorIndex", Index)
    docs, logd = tmp_path / "docs", str(tmp_path / "topics")
    write_docs(docs, 12)
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0
    assert lab2_publish_queries.main(["What about watermarks?", "--log-dir", logd]) == 0
    assert sa_serve.main(["--log-dir", logd, "--once"]) == 0
    assert lab2_publish_queries.main(["How do tumble windows work?", "--log-dir", logd]) == 0
    capsys.readouterr()
    assert sa_serve.main(["--log-dir", logd, "--once"]) == 0                       # a NEW process: empty table at start
    second = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert second["documents"] == 13 and second["searches"] == 1                   # table rebuilt, `documents` NOT re-embedded
    c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results"])
    rows = [Codec(logd).decode(m.value()) for m in c.consume(10, 0.0)]
    assert [r["query"] for r in rows] == ["What about watermarks?", "How do tumble windows work?"]
    assert all(r["document_id_1"] is not None and r["chunk_1"] for r in rows)
    c = Consumer({"log.dir": logd, "group.id": "t2"}); c.subscribe(["documents_embed"])
    assert len(c.consume(100, 0.0)) == 13                                          # nothing was produced twice


def _odd_queries_embed_records(codec, dim, g):
    """queries_embed records in every shape the wire allows: the usual one, a null query, a multi-block array, a
    null item, a wrong length, a non-finite value, garbage."""
    from qsa_b200.wire import avro
    sid_hdr = codec.header("queries_embed")
    vec = lambda: g.standard_normal(dim).astype(np.float32)
    usual = [codec.encode("queries_embed", {"query": f"q{i} é", "embedding": vec()}) for i in range(5)]
    null_q = codec.encode("queries_embed", {"query": None, "embedding": vec()})
    v = vec()
    body = bytearray(b"\x02"); avro.write_long(body, 2); body += b"mb"; body += b"\x02"
    half = dim // 2
    for part in (v[:half], v[half:]):                      # two blocks, the second one in the negative-count form
        blk = b"".join(b"\x02" + struct.pack("<f", float(x)) for x in part)
        if part is v:
            avro.write_long(body, len(part))
        else:
            avro.write_long(body, -len(part)); avro.write_long(body, len(blk))
        body += blk
    body += b"\x00"
    multi = sid_hdr + bytes(body)
    null_item = codec.encode("queries_embed", {"query": "ni", "embedding": [1.0, None] + [0.5] * (dim - 2)})
    wrong_len = codec.encode("queries_embed", {"query": "wl", "embedding": vec()[: dim - 3]})
    nonfinite = codec.encode("queries_embed", {"query": "nf", "embedding": np.r_[vec()[:-1], np.float32("inf")]})
    return usual[:2] + [null_q, multi] + usual[2:4] + [null_item, wrong_len, b"\x07garbage", nonfinite, b""] + usual[4:]


@pytest.mark.parametrize("score_mode", ["cosine", "atlas"])
def test_native_search_stage_equals_the_generic_codec_byte_for_byte(tmp_path, score_mode):
    """The native batch path (sa_wire_*: split, decode, encode + framing) and the generic Python codec path must put the
    SAME bytes on `search_results` and quarantine the SAME records, for usual and unusual record shapes alike."""
    import struct as _s
    globals()["struct"] = _s
    g = np.random.default_rng(11)
    dim = 64
    outs, dlqs = {}, {}
    for name in ("generic", "native"):
        logd = str(tmp_path / name)
        idx = PipelinedOracleIndex(dim)
        table = VectorTable(idx)
        gg = np.random.default_rng(12)
        table.upsert_many([f"d{i}" if i % 7 else None for i in range(50)], [f"chunk {i} ü" if i % 5 else None for i in range(50)],
                          gg.standard_normal((50, dim)).astype(np.float32))
        pipe = Lab2Pipeline(logd, table, k=5, max_batch=7, native=(name == "native"), score_mode=score_mode)
        assert (pipe._wire is not None) == (name == "native")
        p = Producer({"log.dir": logd})
        for raw in _odd_queries_embed_records(pipe.codec, dim, np.random.default_rng(13)):
            p.produce("queries_embed", value=raw)
        p.produce("queries_embed", value=None)
        p.flush()
        assert pipe.stage_search() == 13 and pipe.stage_search() == 0
        pipe.producer.flush()
        c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results"])
        outs[name] = [m.value() for m in c.consume(100, 0.0)]
        c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["queries_embed.dlq"])
        dlqs[name] = [(m.key(), m.value()) for m in c.consume(100, 0.0)]
        assert pipe.stats["quarantined"] == len(dlqs[name]) == 6 and pipe.stats["searches"] == 7
    assert outs["native"] == outs["generic"] and len(outs["native"]) == 7
    assert dlqs["native"] == dlqs["generic"]
    rec = Codec(str(tmp_path / "native")).decode(outs["native"][2])
    assert rec["query"] is None and (0 <= rec["score_1"] <= 1 if score_mode == "atlas" else True)


def test_native_sink_stage_builds_the_same_table_as_the_generic_codec(tmp_path):
    """stage_sink through sa_wire_decode_documents_embed vs the generic codec: same rows, same side columns (metadata
    included), same vectors, same quarantine -- for usual records, null ids / chunks, Lab4 metadata, multi-block arrays,
    upserts and poison."""
    from qsa_b200.wire import avro
    dim = 48
    tables, dlqs = {}, {}
    for name in ("generic", "native"):
        g = np.random.default_rng(31)
        logd = str(tmp_path / name)
        table = VectorTable(PipelinedOracleIndex(dim))
        pipe = Lab2Pipeline(logd, table, k=3, max_batch=6, native=(name == "native"))
        enc = lambda **kw: pipe.codec.encode("documents_embed", kw)
        vec = lambda: g.standard_normal(dim).astype(np.float32)
        recs = [enc(document_id=f"d{i}", chunk=f"chunk {i} é", embedding=vec()) for i in range(9)]
        recs += [enc(document_id=None, chunk="no id", embedding=vec()), enc(document_id="nochunk", chunk=None, embedding=vec()),
                 enc(document_id="lab4", chunk="policy text", embedding=vec(), pages="12-14", section_reference="4.2",
                     title="Flood", fraud_categories=["a", None, "c"], policy_keywords=[], char_count=-7),
                 enc(document_id="d3", chunk="d3 again", embedding=vec()),                      # upsert: tombstones row 3
                 enc(document_id="short", chunk="x", embedding=vec()[: dim - 1]),               # wrong length -> quarantine
                 enc(document_id="nullitem", chunk="x", embedding=[None] + [0.5] * (dim - 1)),  # null item -> quarantine
                 b"\x00\x00\x00\x00\x01garbage", b""]
        v = vec()
        body = bytearray(b"\x02"); avro.write_long(body, 2); body += b"mb"; body += b"\x02"; avro.write_long(body, 5); body += b"multi"
        body += b"\x02"
        for part in (v[:7], v[7:]):
            avro.write_long(body, len(part)); body += b"".join(b"\x02" + struct.pack("<f", float(x)) for x in part)
        body += b"\x00" + b"\x00" * 6
        recs.append(pipe.codec.header("documents_embed") + bytes(body))                        # two array blocks: generic path
        for r in recs:
            pipe.producer.produce("documents_embed", value=r)
        pipe.producer.produce("documents_embed", value=None)
        pipe.producer.flush()
        moved = 0
        while (m := pipe.stage_sink()):
            moved += m
        assert moved == len(recs) + 1
        pipe.producer.flush()
        tables[name] = table
        c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["documents_embed.dlq"])
        dlqs[name] = [(m.key(), m.value()) for m in c.consume(100, 0.0)]
        assert pipe.stats["quarantined"] == len(dlqs[name]) == 5 and pipe.stats["documents"] == 14
    a, b = tables["generic"], tables["native"]
    assert a.document_id == b.document_id and a.chunk == b.chunk and a.metadata == b.metadata and len(a) == 14
    assert a.avro_chunk == b.avro_chunk and a._row_of == b._row_of and a._row_of["d3"] == 12
    assert (a.index.bits == b.index.bits).all() and (a.index.bits[3] == 0).all()
    assert bytes(a.arena_chunk.data[:a.arena_chunk.used]) == bytes(b.arena_chunk.data[:b.arena_chunk.used]) == b"".join(a.avro_chunk)
    assert a.metadata[11] == {"pages": "12-14", "section_reference": "4.2", "title": "Flood", "fraud_categories": ["a", None, "c"],
                              "policy_keywords": [], "char_count": -7}
    key = lambda kv: (kv[0], kv[1] or b"")
    assert sorted(dlqs["generic"], key=key) == sorted(dlqs["native"], key=key)     # same records, same reasons (the generic
    # path quarantines a batch's undecodable records before its wrong-length ones; the native path keeps stream order)


def test_serve_metrics_file_has_latency_percentiles(tmp_path):
    g = np.random.default_rng(2)
    logd, mf = str(tmp_path / "topics"), str(tmp_path / "metrics.jsonl")
    idx = PipelinedOracleIndex(32)
    table = VectorTable(idx)
    table.upsert_many([f"d{i}" for i in range(20)], [f"c{i}" for i in range(20)], g.standard_normal((20, 32)).astype(np.float32))
    pipe = Lab2Pipeline(logd, table, k=3, max_batch=8, metrics_file=mf, metrics_every_s=0.0)
    p = Producer({"log.dir": logd})
    for i in range(30):
        p.produce("queries_embed", value=pipe.codec.encode("queries_embed", {"query": f"q{i}", "embedding": g.standard_normal(32).astype(np.float32)}))
    p.flush()
    assert pipe.stage_search() == 30
    rows = [json.loads(l) for l in open(mf)]
    assert sum(r["queries"] for r in rows) == 30 and all(r["batch_latency_ms"]["p99"] >= r["batch_latency_ms"]["p50"] >= 0 for r in rows)


def test_config1_two_thousand_docs_plumbing_on_cpu(tmp_path):
    """BASELINE.json configs[0]: ~2k Lab2-style chunks published with publish_docs, 32 questions with publish_queries,
    numpy cosine top-10 on the CPU (the oracle, standing in for the GPU index in this no-GPU plumbing case), first three
    results on `search_results` in the reference's column layout."""
    from oracle import bruteforce as bf
    docs, logd = tmp_path / "docs", str(tmp_path / "topics")
    write_docs(docs, 2047)                                   # + plain.md = 2048 chunks
    assert publish_docs.main(["--docs-dir", str(docs), "--workers", "8", "--log-dir", logd]) == 0
    questions = [f"Question {i}: tell me about {TOPICS[i % len(TOPICS)]} number {i * 61 % 2047}" for i in range(32)]
    for q in questions:
        assert lab2_publish_queries.main([q, "--log-dir", logd]) == 0
    idx = OracleIndex(1536)
    table = VectorTable(idx)
    emb = StubEmbedder(1536)
    pipe = Lab2Pipeline(logd, table, embedder=emb, k=10, max_batch=256)
    pipe.run_until_idle()
    assert len(table) == 2048 and pipe.stats["searches"] == 32 and pipe.stats["responses"] == 32
    c = Consumer({"log.dir": logd, "group.id": "t"}); c.subscribe(["search_results"])
    rows = [Codec(logd).decode(m.value()) for m in c.consume(100, 0.0)]
    assert [r["query"] for r in rows] == questions
    # independent numpy check of the top-10 -> the topic carries its first three
    qv = bf.f32_to_bf16_bits(np.stack([emb.embed(q) for q in questions]))
    s, i = bf.cosine_topk_f64(qv, idx.bits, 10)
    for r, row in enumerate(rows):
        assert [row[f"document_id_{j}"] for j in (1, 2, 3)] == [table.document_id[x] for x in i[r, :3]]
        assert abs(row["score_1"] - s[r, 0]) < 1e-6 and row["score_1"] >= row["score_2"] >= row["score_3"]
    hits = vector_search_agg(table, "embedding", emb.embed(questions[0]), 10)[0]
    assert [h.row for h in hits] == i[0].tolist()


@pytest.mark.parametrize("native", [True, False])
def test_strings_that_are_not_utf8_are_quarantined_by_both_codecs(tmp_path, native):
    """A record whose string bytes are not UTF-8 (a corrupted producer, a Latin-1 client) is a poison record: both the
    native batch path and the generic codec send it to the DLQ and carry on -- in the sink (chunk, document_id, a metadata
    string) and in the search stage (query text)."""
    logd = str(tmp_path / "topics")
    dim = 64
    table = VectorTable(PipelinedOracleIndex(dim))
    pipe = Lab2Pipeline(logd, table, embedder=StubEmbedder(dim), k=3, max_batch=8, native=native)
    g = np.random.default_rng(3)

    def doc(i, **kw):
        r = {"document_id": f"d{i}", "chunk": f"chunk number {i}", "embedding": g.standard_normal(dim).astype(np.float32)}
        r.update(kw)
        return bytearray(pipe.codec.encode("documents_embed", r))
    docs = [doc(i) for i in range(6)]
    bad_chunk, bad_id, bad_meta = doc(6), doc(7), doc(8, title="Title eight")
    bad_chunk[bytes(bad_chunk).find(b"chunk number 6")] = 0xFF
    bad_id[bytes(bad_id).find(b"d7")] = 0xC0
    bad_meta[bytes(bad_meta).find(b"Title eight") + 2] = 0xED
    for raw in docs[:3] + [bad_chunk, bad_id] + docs[3:] + [bad_meta]:
        pipe.producer.produce("documents_embed", value=bytes(raw))
    q_ok = [bytearray(pipe.codec.encode("queries_embed", {"query": f"question {i}", "embedding": g.standard_normal(dim).astype(np.float32)}))
            for i in range(5)]
    q_bad = bytearray(q_ok[2])
    q_bad[bytes(q_bad).find(b"question 2")] = 0x80
    for raw in q_ok[:2] + [q_bad] + q_ok[3:]:
        pipe.producer.produce("queries_embed", value=bytes(raw))
    pipe.producer.flush()
    pipe.run_until_idle()
    assert len(table) == 6 and sorted(table.document_id) == [f"d{i}" for i in range(6)]
    assert pipe.stats["quarantined"] == 4 and pipe.stats["searches"] == 4
    b = Broker(logd)
    assert b.count("documents_embed.dlq") == 3 and b.count("queries_embed.dlq") == 1 and b.count("search_results") == 4
    c = Consumer({"log.dir": logd, "group.id": "check"}); c.subscribe(["search_results"])
    assert [Codec(logd).decode(m.value())["query"] for m in c.consume(10, 0.0)] == ["question 0", "question 1", "question 3", "question 4"]


@pytest.mark.parametrize("seed", range(8))
def test_random_streams_native_and_generic_pipelines_agree(tmp_path, seed):
    """Differential run on random streams: a random mixture of usual, unusual and mutated records (byte flips,
    truncations, non-UTF-8 strings, null values) on `documents_embed` and `queries_embed`, random batch size -- the native
    batch path and the generic codec must end with the same table, the same `search_results` bytes in the same order and
    the same quarantine."""
    import struct as _s
    globals()["struct"] = _s
    dim = 32
    results = {}
    for name in ("generic", "native"):
        g = np.random.default_rng(1000 + seed)
        logd = str(tmp_path / name)
        table = VectorTable(PipelinedOracleIndex(dim))
        pipe = Lab2Pipeline(logd, table, k=3, max_batch=int(g.choice([1, 2, 3, 8, 64])), native=(name == "native"),
                            score_mode=str(g.choice(["cosine", "atlas"])))
        vec = lambda: g.standard_normal(dim).astype(np.float32)

        def mutate(raw):
            raw = bytearray(raw)
            kind = int(g.integers(0, 6))
            if kind == 0 and len(raw) > 6:
                for _ in range(int(g.integers(1, 3))):
                    raw[int(g.integers(5, len(raw)))] ^= int(g.integers(1, 256))
            elif kind == 1:
                raw = raw[:int(g.integers(0, len(raw) + 1))]
            elif kind == 2:
                raw += bytes(g.integers(0, 256, int(g.integers(1, 5)), dtype=np.uint8))
            return bytes(raw)
        p = Producer({"log.dir": logd})
        n_docs = int(g.integers(5, 40))
        for i in range(n_docs):
            doc_id = None if g.random() < 0.1 else f"d{int(g.integers(0, 25))}"           # repeats = upserts
            rec = {"document_id": doc_id, "chunk": None if g.random() < 0.1 else f"chunk {i} ü", "embedding": vec()}
            if g.random() < 0.3:
                rec.update(title=f"T{i}", pages=str(i), fraud_categories=["x", None], char_count=i)
            raw = pipe.codec.encode("documents_embed", rec)
            if g.random() < 0.25:
                raw = mutate(raw)
            p.produce("documents_embed", value=raw)
        odd = _odd_queries_embed_records(pipe.codec, dim, g)
        n_q = int(g.integers(3, 40))
        for i in range(n_q):
            r = g.random()
            if r < 0.2:
                raw = odd[int(g.integers(0, len(odd)))]
            else:
                raw = pipe.codec.encode("queries_embed", {"query": None if r < 0.3 else f"q{i} é", "embedding": vec()})
                if r > 0.8:
                    raw = mutate(raw)
            p.produce("queries_embed", value=(None if g.random() < 0.03 else raw))
        p.flush()
        pipe.run_until_idle()

        def drain(topic):
            c = Consumer({"log.dir": logd, "group.id": "cmp"}); c.subscribe([topic])
            return [(m.key(), m.value()) for m in c.consume(1000, 0.0)]
        results[name] = {
            "table": (list(table.document_id), list(table.chunk), [dict(m) for m in table.metadata]),
            "rows": table.index.bits.tobytes(),                      # the vectors as ingested (bf16), tombstones included
            "search_results": [v for _, v in drain("search_results")],
            "dlq_q": sorted(drain("queries_embed.dlq"), key=repr), "dlq_d": sorted(drain("documents_embed.dlq"), key=repr),
            "stats": {k: pipe.stats[k] for k in ("documents", "searches", "quarantined")},
        }
    a, b = results["generic"], results["native"]
    assert a["stats"] == b["stats"], (a["stats"], b["stats"])
    assert a["table"] == b["table"]
    assert a["search_results"] == b["search_results"] and a["dlq_q"] == b["dlq_q"] and a["dlq_d"] == b["dlq_d"]
    assert a["rows"] == b["rows"]


@pytest.mark.parametrize("seed", range(6))
def test_vector_table_against_a_dictionary_model_under_random_operations(tmp_path, seed):
    """Random upsert batches (repeated ids inside a batch and across batches, null ids, null chunks, metadata), a clear
    now and then, a checkpoint / restore in the middle: the table must stay equivalent to a plain dictionary -- the live
    row of an id is its last upsert, every older row of it is tombstoned (all-zero vector), the pre-serialised Avro
    columns and arenas describe every row, and searching returns what brute force over the live rows returns."""
    from qsa_b200.operator import _avro_nullable_string
    g = np.random.default_rng(500 + seed)
    dim = 16
    table = VectorTable(OracleIndex(dim))
    model, anon = {}, []                                   # id -> (chunk, vector, metadata); rows published without an id

    def check(t):
        n = len(t)
        assert len(t.chunk) == len(t.metadata) == len(t.avro_document_id) == len(t.avro_chunk) == n == len(t.index)
        vecs = bf_bits_to_f32(t.index.bits)
        live = {d: r for r, d in enumerate(t.document_id) if d is not None and t._row_of.get(d) == r}
        assert set(live) == set(model) and t._row_of == live
        for d, r in live.items():
            chunk, vec, meta = model[d]
            assert t.chunk[r] == chunk and t.metadata[r] == meta and (vecs[r] == bf_round(vec)).all()
        for r, d in enumerate(t.document_id):
            if d is not None and live.get(d) != r:
                assert not vecs[r].any()                   # tombstone
            assert t.avro_document_id[r] == _avro_nullable_string(d) and t.avro_chunk[r] == _avro_nullable_string(t.chunk[r])
            for arena, col in ((t.arena_document_id, t.avro_document_id), (t.arena_chunk, t.avro_chunk)):
                lo, hi = int(arena.off[r]), int(arena.off[r + 1])
                assert bytes(arena.data[lo:hi]) == col[r]
        anon_rows = [r for r, d in enumerate(t.document_id) if d is None]
        assert len(anon_rows) == len(anon)
        for r, (chunk, vec) in zip(anon_rows, anon):
            assert t.chunk[r] == chunk and (vecs[r] == bf_round(vec)).all()

    def bf_bits_to_f32(bits):
        return (bits.astype(np.uint32) << 16).view(np.float32)

    def bf_round(v):
        from oracle import bruteforce as bf
        return bf.bf16_bits_to_f32(bf.f32_to_bf16_bits(np.asarray(v, np.float32)))

    for step in range(30):
        r = g.random()
        if r < 0.06:
            table.clear(); model.clear(); anon.clear()
        elif r < 0.14 and len(table):
            d = str(tmp_path / f"ckpt{step}")
            table.save(d, {"documents_embed-0": step})
            restored = VectorTable(OracleIndex(dim))
            restored.index.bits = table.index.bits.copy()   # the double has no snapshot(): the vectors travel by hand
            assert restored.load(d) == len(table) and restored.source_offsets == {"documents_embed-0": step}
            check(restored)
            table = restored
        else:
            n = int(g.integers(1, 12))
            ids = [None if g.random() < 0.15 else f"d{int(g.integers(0, 20))}" for _ in range(n)]
            chunks = [None if g.random() < 0.1 else f"c{step}-{i} é" for i in range(n)]
            vecs = g.standard_normal((n, dim)).astype(np.float32)
            metas = [({"title": f"t{step}-{i}"} if g.random() < 0.4 else {}) for i in range(n)]
            table.upsert_many(ids, chunks, vecs, metas)
            last = {d: i for i, d in enumerate(ids) if d is not None}
            for i, d in enumerate(ids):
                if d is None:
                    anon.append((chunks[i], vecs[i]))
                elif last[d] == i:
                    model[d] = (chunks[i], vecs[i], metas[i])
        check(table)
    if len(table):
        q = g.standard_normal((5, dim)).astype(np.float32)
        hits = vector_search_agg(table, "embedding", q, 3)
        for h in hits:
            for x in h:
                assert x.document_id is None or table._row_of[x.document_id] == x.row      # never a tombstoned row
