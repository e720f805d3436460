"""The operator joined LATERALly onto an upstream stream -- its Lab3 and Lab4 call sites (LAB3-Walkthrough.md:225-375,
LAB4-Walkthrough.md:251-309) as stages over the topic log (pipeline/lateral.py).  The reference pins only shapes at this
boundary (testing/e2e/test_lab3.py:232-268: top_chunk_1/2 non-empty; test_lab4.py:274-287: statement running), so the
checks are: the output columns are the statement's, carried columns pass through unchanged, the projected columns equal
the operator's answer for the same query vector, and the stage has the delivery semantics of the Lab2 loop."""
import datetime as dt
import struct

import numpy as np
import pytest

from qsa_b200.embed.stub import StubEmbedder
from qsa_b200.operator import VectorTable, project_search_results, vector_search_agg
from qsa_b200.pipeline import lateral
from qsa_b200.pipeline.serve import Lab2Pipeline
from qsa_b200.transport.filelog import Broker, Consumer, Producer
from qsa_b200.wire import avro
from qsa_b200.wire.registry import SchemaRegistry
from scripts import publish_docs

from doubles import OracleIndex
from test_cli_and_pipeline import write_docs


def build_table(tmp_path, dim=1536, n_docs=48, index=None):
    docs, logd = tmp_path / "docs", str(tmp_path / "topics")
    write_docs(docs, n_docs)
    assert publish_docs.main(["--docs-dir", str(docs), "--log-dir", logd]) == 0
    table = VectorTable(index or OracleIndex(dim), name="fema_policies_vectordb")
    pipe = Lab2Pipeline(logd, table, embedder=StubEmbedder(dim), k=3)
    pipe.run_until_idle()
    assert len(table) == n_docs + 1
    return logd, table, pipe


def publish(logd, topic, schema, records, keys=None):
    reg = SchemaRegistry(logd)
    sid = reg.register(f"{topic}-value", schema)
    p = Producer({"log.dir": logd})
    for i, r in enumerate(records):
        p.produce(topic, key=None if keys is None else keys[i], value=avro.frame(sid, avro.encode(schema, r)))
    p.flush()
    return sid


def read_topic(logd, topic, group="check"):
    c = Consumer({"log.dir": logd, "group.id": group})
    c.subscribe([topic])
    reg = SchemaRegistry(logd)
    out = []
    for m in c.consume(1000, 0.0):
        sid = struct.unpack_from(">I", m.value(), 1)[0]
        out.append((m, sid, avro.decode(reg.get(sid), m.value(), 5)))
    return out


def claim(i, narrative):
    ts = int(dt.datetime(2024, 10, 12, 8, i, tzinfo=dt.timezone.utc).timestamp() * 1000)
    return {"claim_id": f"CLM-{i:04d}", "applicant_name": f"Applicant {i}", "city": "Naples", "claim_narrative": narrative,
            "claim_amount": str(1000 + i), "damage_assessed": "roof", "has_insurance": "yes", "insurance_amount": "500",
            "is_primary_residence": "yes", "assessment_date": "2024-10-11", "disaster_date": "2024-10-09",
            "assessment_source": "inspector", "shared_account": None, "shared_phone": "no", "previous_claims_count": "0",
            "last_claim_date": None, "claim_timestamp": ts, "anomaly_window_time": ts + 3_600_000,
            "anomaly_total_amount": 123456.5, "is_anomaly": True}


def test_lab4_claims_are_joined_with_their_policy_chunks(tmp_path):
    logd, table, _ = build_table(tmp_path)
    narratives = ["late data and watermarks in event time", "tumble and hop window functions", "state ttl and checkpoints"]
    claims = [claim(i, n) for i, n in enumerate(narratives)]
    publish(logd, "claims_to_investigate", lateral.CLAIMS_TO_INVESTIGATE_VALUE, claims, keys=[c["claim_id"] for c in claims])
    stage = lateral.lab4_claims_with_policies(logd, table)
    assert stage.run_until_idle() == 3 and stage.stats == {"rows": 3, "searched": 3, "filtered": 0, "quarantined": 0}
    got = read_topic(logd, "claims_to_investigate_with_policies")
    assert [m.key() for m, _, _ in got] == [c["claim_id"].encode() for c in claims]          # keys carried through
    reg = SchemaRegistry(logd)
    sid = reg.latest("claims_to_investigate_with_policies-value")
    names = [f["name"] for f in reg.get(sid)["fields"]]
    carried = names[:20]
    assert carried[:5] == ["claim_id", "applicant_name", "city", "claim_amount", "damage_assessed"] and carried[-1] == "is_anomaly"
    assert names[20:27] == ["policy_chunk_1", "policy_score_1", "policy_pages_1", "policy_section_1", "policy_title_1",
                            "policy_fraud_cats_1", "policy_keywords_1"]                        # LAB4-Walkthrough.md:280-286
    assert len(names) == 20 + 3 * 7 and names[-1] == "policy_keywords_3"
    emb = StubEmbedder(1536)
    for (m, s, rec), c in zip(got, claims):
        assert s == sid and {k: rec[k] for k in carried} == {k: c[k] for k in carried}          # c.* unchanged
        hits = vector_search_agg(table, "embedding", emb.embed(c["claim_narrative"]), 3)[0]
        assert {k: rec[k] for k in names[20:]} == project_search_results(hits, lateral.LAB4_POLICY_COLUMNS, 3)
        assert rec["policy_score_1"] >= rec["policy_score_2"] >= rec["policy_score_3"] > 0
    assert "watermarks" in got[0][2]["policy_chunk_1"].lower() and got[0][2]["policy_title_1"].startswith("Watermarks")
    assert got[0][2]["policy_keywords_1"] == ["watermarks", "sql"]
    # a second instance of the statement (same consumer group) has nothing left to do: offsets were committed
    again = lateral.lab4_claims_with_policies(logd, table)
    assert again.run_until_idle() == 0 and Broker(logd).count("claims_to_investigate_with_policies") == 3


def test_lab4_precomputed_vectors_poison_rows_and_score_mode(tmp_path):
    logd, table, _ = build_table(tmp_path, dim=768)
    emb = StubEmbedder(768)
    schema = {**lateral.CLAIMS_TO_INVESTIGATE_VALUE, "fields": lateral.CLAIMS_TO_INVESTIGATE_VALUE["fields"] + [
        {"name": "narrative_embedding", "type": ["null", {"type": "array", "items": ["null", "float"]}], "default": None}]}
    v = emb.embed("kafka connector properties")
    rows = [dict(claim(0, "something unrelated to the vector"), narrative_embedding=[float(x) for x in v]),     # the vector wins
            dict(claim(1, None), narrative_embedding=None),                                               # nothing to search with
            dict(claim(2, "joins interval temporal lookup"), narrative_embedding=None),
            dict(claim(3, "x"), narrative_embedding=[1.0, 2.0])]                                          # wrong length
    publish(logd, "claims_to_investigate", schema, rows)
    p = Producer({"log.dir": logd}); p.produce("claims_to_investigate", value=b"\x07garbage"); p.flush()   # bad magic byte
    stage = lateral.lab4_claims_with_policies(logd, table, score_mode="atlas")
    assert stage.run_until_idle() == 5
    assert stage.stats["searched"] == 2 and stage.stats["quarantined"] == 3
    got = [r for _, _, r in read_topic(logd, "claims_to_investigate_with_policies")]
    assert [r["claim_id"] for r in got] == ["CLM-0000", "CLM-0002"]
    assert "narrative_embedding" not in got[0]                                  # the vector is not part of SELECT c.*
    hits = vector_search_agg(table, "embedding", v, 3, score_mode="atlas")[0]
    assert got[0]["policy_chunk_1"] == hits[0].chunk and "connector" in got[0]["policy_chunk_1"].lower()
    assert got[0]["policy_score_1"] == hits[0].score and 0.5 < got[0]["policy_score_1"] <= 1.0           # (1 + cos) / 2
    dlq = read_dlq(logd, "claims_to_investigate.dlq")
    assert len(dlq) == 3 and any("neither a query vector nor a text" in k for k in dlq)
    assert any("768 finite floats" in k for k in dlq) and any("magic" in k for k in dlq)


def read_dlq(logd, topic):
    c = Consumer({"log.dir": logd, "group.id": "dlq-check"})
    c.subscribe([topic])
    return [m.key().decode() for m in c.consume(100, 0.0)]


def test_surge_query_text_is_the_statement_s_concat():
    ts = int(dt.datetime(2025, 3, 1, 18, 5, tzinfo=dt.timezone.utc).timestamp() * 1000)
    row = {"pickup_zone": "French Quarter", "window_time": ts, "request_count": 100, "expected_requests": 40.0, "is_surge": True}
    assert lateral.surge_query(row) == (
        "Transportation demand surge in French Quarter at 6:05 PM (18:05) during evening dinner period (5:00 PM - 8:00 PM). "
        "Looking for HIGH demand events occurring between 5:05 PM and 7:05 PM. Expected: 40.0, Actual: 100 (+150.0%). "
        "What HIGH impact events, festivals, or gatherings are active in French Quarter during this time?")
    parts = {0: "late night hours (12:00 AM - 4:00 AM)", 3: "late night hours", 4: "early morning setup period", 8: "morning rush hours",
             11: "late morning period", 13: "lunch service peak", 16: "afternoon hours", 19: "evening dinner period",
             22: "nightlife hours (8:00 PM - 11:00 PM)", 23: "late night period (11:00 PM - 12:00 AM)"}
    for hour, name in parts.items():
        t = int(dt.datetime(2025, 3, 1, hour, 0, tzinfo=dt.timezone.utc).timestamp() * 1000)
        q = lateral.surge_query(dict(row, window_time=t))
        assert f"during {name}" in q
    midnight = lateral.surge_query(dict(row, window_time=int(dt.datetime(2025, 3, 1, 0, 30, tzinfo=dt.timezone.utc).timestamp() * 1000)))
    assert "at 12:30 AM (00:30)" in midnight and "between 11:30 PM and 1:30 AM" in midnight
    central = lateral.surge_query(row, tz=dt.timezone(dt.timedelta(hours=-6)))          # session time zone America/Chicago (CST)
    assert "at 12:05 PM (12:05) during lunch service peak" in central
    prompt = lateral.surge_prompt({"query": "Q", "top_score_1": 0.75, "top_document_1": "d1", "top_chunk_1": "c1",
                                   "top_score_2": None, "top_document_2": None, "top_chunk_2": None,
                                   "top_score_3": None, "top_document_3": None, "top_chunk_3": None})
    assert prompt.startswith("Analyze the retrieved event documents and identify the most likely cause")
    assert "USER QUERY: Q\n\nRETRIEVED DOCUMENTS:\nDocument 1 (Score: 0.75):\nSource: d1\nc1\n\nDocument 2 (Score: ):\nSource: \n\n\n" in prompt
    assert prompt.endswith("Provide only the reason, no additional text.")


def test_lab3_surges_are_enriched_and_the_rest_is_filtered(tmp_path):
    logd, table, _ = build_table(tmp_path)
    base = int(dt.datetime(2025, 3, 1, 21, 0, tzinfo=dt.timezone.utc).timestamp() * 1000)
    rows = [{"pickup_zone": "Marigny", "window_time": base, "request_count": 90, "expected_requests": 30.0, "is_surge": True},
            {"pickup_zone": "Uptown", "window_time": base + 300_000, "request_count": 10, "expected_requests": 30.0, "is_surge": False},
            {"pickup_zone": "CBD", "window_time": base + 600_000, "request_count": 12, "expected_requests": None, "is_surge": None},
            {"pickup_zone": "Treme", "window_time": base + 900_000, "request_count": 75, "expected_requests": 25.0, "is_surge": True}]
    publish(logd, "anomalies_per_zone", lateral.ANOMALIES_PER_ZONE_VALUE, rows)
    stage = lateral.lab3_anomalies_enriched(logd, table)
    assert stage.run_until_idle() == 4
    assert stage.stats == {"rows": 4, "searched": 2, "filtered": 2, "quarantined": 0}           # WHERE is_surge = true
    got = read_topic(logd, "anomalies_enriched")
    reg = SchemaRegistry(logd)
    names = [f["name"] for f in reg.get(got[0][1])["fields"]]
    assert names == ["pickup_zone", "window_time", "request_count", "expected_requests", "anomaly_reason",
                     "top_chunk_1", "top_chunk_2", "top_chunk_3"]                               # LAB3-Walkthrough.md:225-235
    emb = StubEmbedder(1536)
    for (_, _, rec), src in zip(got, (rows[0], rows[3])):
        assert {k: rec[k] for k in names[:4]} == {k: src[k] for k in names[:4]}
        hits = vector_search_agg(table, "embedding", emb.embed(lateral.surge_query(src)), 3)[0]
        assert [rec["top_chunk_1"], rec["top_chunk_2"], rec["top_chunk_3"]] == [h.chunk for h in hits]
        assert rec["top_chunk_1"] and rec["top_chunk_2"]                                        # testing/e2e/test_lab3.py:232-268
        assert rec["anomaly_reason"] and hits[0].document_id in rec["anomaly_reason"]


def test_statement_validation(tmp_path):
    logd, table, _ = build_table(tmp_path, n_docs=8)
    with pytest.raises(ValueError, match="no column 'embedding'"):
        lateral.LateralSearch(logd, table, "a", "b", columns={"embedding": "e"})
    with pytest.raises(ValueError, match="n_out cannot exceed k"):
        lateral.LateralSearch(logd, table, "a", "b", columns={"chunk": "c"}, k=2, n_out=3)
    with pytest.raises(ValueError, match="response_field needs"):
        lateral.LateralSearch(logd, table, "a", "b", columns={"chunk": "c"}, response_field="r")
    with pytest.raises(ValueError, match="nothing to search with"):
        lateral.LateralSearch(logd, table, "a", "b", columns={"chunk": "c"}, vector_field=None)
    # a statement that selects a column it does not produce, or carries one the source lacks, fails per row -> quarantine
    publish(logd, "src", lateral.ANOMALIES_PER_ZONE_VALUE,
            [{"pickup_zone": "Z", "window_time": 0, "request_count": 1, "expected_requests": 1.0, "is_surge": True}])
    st = lateral.LateralSearch(logd, table, "src", "dst", columns={"chunk": "c"}, text_field="pickup_zone", vector_field=None,
                               carry=("pickup_zone", "no_such_column"))
    assert st.run_until_idle() == 1 and st.stats["quarantined"] == 1 and Broker(logd).count("dst") == 0
    # generic use: top-1 document id next to the zone, k = 5 searched, 1 projected
    st = lateral.LateralSearch(logd, table, "src", "dst", columns={"document_id": "doc", "score": "sim"}, k=5, n_out=1,
                               text_field="pickup_zone", vector_field=None, carry=("pickup_zone",), query_field="q", group="g2")
    assert st.run_until_idle() == 1
    (_, _, rec), = read_topic(logd, "dst")
    assert list(rec) == ["pickup_zone", "q", "doc_1", "sim_1"] and rec["q"] == "Z" and rec["doc_1"] is not None


def test_lateral_stage_rides_along_in_the_serve_loop(tmp_path):
    """`sa_serve --lateral lab4`: the statement runs at the end of every pass of the Lab2 loop, over the same table."""
    logd, table, pipe = build_table(tmp_path)
    pipe.extra_stages.append(lateral.lab4_claims_with_policies(pipe.log_dir, table))
    publish(logd, "claims_to_investigate", lateral.CLAIMS_TO_INVESTIGATE_VALUE, [claim(5, "user defined functions in python")])
    assert pipe.run_until_idle() == 1
    (_, _, rec), = read_topic(logd, "claims_to_investigate_with_policies")
    assert rec["claim_id"] == "CLM-0005" and "user defined functions" in rec["policy_chunk_1"].lower()


@pytest.mark.gpu
def test_lab4_statement_on_gpu(tmp_path):
    """The same statement over the CUDA engine (768-d table, as north_star's Lab3/Lab4 variants): answers equal the
    operator's direct answers; the index is the real one, so this is the C ABI's host path under the stage."""
    from qsa_b200.engine import VectorIndex
    ix = VectorIndex(dim=768, capacity=4096, max_batch=64, max_k=10)
    logd, table, _ = build_table(tmp_path, dim=768, index=ix)
    narratives = ["late data and watermarks in event time", "json avro protobuf formats schema registry"]
    claims = [claim(i, n) for i, n in enumerate(narratives)]
    publish(logd, "claims_to_investigate", lateral.CLAIMS_TO_INVESTIGATE_VALUE, claims)
    stage = lateral.lab4_claims_with_policies(logd, table)
    assert stage.run_until_idle() == 2
    emb = StubEmbedder(768)
    for (_, _, rec), c in zip(read_topic(logd, "claims_to_investigate_with_policies"), claims):
        hits = vector_search_agg(table, "embedding", emb.embed(c["claim_narrative"]), 3)[0]
        assert [rec[f"policy_chunk_{i}"] for i in (1, 2, 3)] == [h.chunk for h in hits]
        assert [rec[f"policy_score_{i}"] for i in (1, 2, 3)] == [h.score for h in hits]
    ix.close()


def test_statement_over_the_kafka_adapter_matches_the_file_log(tmp_path, monkeypatch):
    """The stage only speaks the transport interface: over transport.kafka (confluent_kafka stand-in) it writes the same
    bytes as over the file log and commits the source offsets it has covered."""
    import sys

    import fake_confluent_kafka as fck
    from qsa_b200.transport import filelog, kafka
    monkeypatch.setitem(sys.modules, "confluent_kafka", fck)
    fck.reset()
    g = np.random.default_rng(9)
    dim = 32
    outs = {}
    for name, mod, conf in (("file", filelog, None), ("kafka", kafka, {"bootstrap.servers": "fake:9092"})):
        logd = str(tmp_path / name)
        table = VectorTable(OracleIndex(dim))
        g = np.random.default_rng(9)
        table.upsert_many([f"d{i}" for i in range(30)], [f"chunk {i}" for i in range(30)],
                          g.standard_normal((30, dim)).astype(np.float32),
                          metadata=[{"title": f"T{i}", "pages": str(i)} for i in range(30)])
        schema = {**lateral.ANOMALIES_PER_ZONE_VALUE, "fields": lateral.ANOMALIES_PER_ZONE_VALUE["fields"] + [
            {"name": "embedding", "type": ["null", {"type": "array", "items": ["null", "float"]}], "default": None}]}
        sid = SchemaRegistry(logd).register("zones-value", schema)
        prod = mod.Producer(dict(conf or {}, **{"log.dir": logd}))
        for i in range(7):
            rec = {"pickup_zone": f"Z{i}", "window_time": 1000 * i, "request_count": i, "expected_requests": 1.0, "is_surge": True,
                   "embedding": [float(x) for x in g.standard_normal(dim).astype(np.float32)]}
            prod.produce("zones", key=f"k{i}", value=avro.frame(sid, avro.encode(schema, rec)))
        prod.flush()
        st = lateral.LateralSearch(logd, table, "zones", "zones_with_docs", columns={"document_id": "doc", "title": "title", "score": "s"},
                                   k=3, n_out=2, carry=("pickup_zone", "window_time"), max_batch=4, transport=mod, client_conf=conf)
        assert st.run_until_idle() == 7
        c = mod.Consumer(dict(conf or {}, **{"log.dir": logd, "group.id": "check", "enable.auto.commit": False}))
        c.subscribe(["zones_with_docs"])
        outs[name] = [(m.key(), m.value()) for m in c.consume(100, 0.0)]
    assert len(outs["file"]) == 7 and outs["kafka"] == outs["file"]
    assert fck._BROKER["groups"]["sa-lateral"][("zones", 0)] == 7
    rec = avro.decode(SchemaRegistry(str(tmp_path / "file")).get(struct.unpack_from(">I", outs["file"][0][1], 1)[0]), outs["file"][0][1], 5)
    assert list(rec) == ["pickup_zone", "window_time", "doc_1", "title_1", "s_1", "doc_2", "title_2", "s_2"]
    assert rec["title_1"] == "T" + rec["doc_1"][1:] and rec["s_1"] >= rec["s_2"]
