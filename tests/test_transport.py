import os
"""File-log transport: the Kafka semantics the reference's scripts and tests rely on."""
import threading

from qsa_b200.transport.filelog import Broker, Consumer, Producer, TopicPartition


def conf(d, group="g"):
    return {"log.dir": str(d), "group.id": group, "auto.offset.reset": "earliest", "enable.auto.commit": False}


def test_offsets_watermarks_and_message_accessors(tmp_path):
    p = Producer({"log.dir": str(tmp_path)})
    for i in range(5):
        p.produce("queries", key=f"k{i}", value=f"v{i}".encode(), partition=0)
    p.produce("queries", key=None, value=b"x", partition=1)
    p.poll(0)
    p.flush()
    b = Broker(str(tmp_path))
    assert b.list_topics() == {"queries": [0, 1]}
    assert b.get_watermark_offsets(TopicPartition("queries", 0)) == (0, 5)
    assert b.count("queries") == 6
    c = Consumer(conf(tmp_path))
    c.subscribe(["queries"])
    assert c.list_topics("queries") == {"queries": [0, 1]} and c.list_topics("nope") == {}
    got = []
    while (m := c.poll(0.01)) is not None:
        assert m.error() is None
        got.append((m.partition(), m.offset(), m.key(), m.value()))
    assert sorted(got) == [(0, i, f"k{i}".encode(), f"v{i}".encode()) for i in range(5)] + [(1, 0, None, b"x")]
    assert c.poll(0.0) is None                       # idle -> None, like confluent_kafka


def test_consumer_groups_commit_and_at_least_once(tmp_path):
    p = Producer({"log.dir": str(tmp_path)})
    for i in range(10):
        p.produce("t", value=bytes([i]))
    p.flush()
    a = Consumer(conf(tmp_path, "A")); a.subscribe(["t"])
    first = a.consume(4, 0.0)
    assert [m.offset() for m in first] == [0, 1, 2, 3]
    a.close()                                        # crash before commit ...
    a2 = Consumer(conf(tmp_path, "A")); a2.subscribe(["t"])
    again = a2.consume(4, 0.0)
    assert [m.offset() for m in again] == [0, 1, 2, 3]    # ... the same records are redelivered
    a2.commit()
    a3 = Consumer(conf(tmp_path, "A")); a3.subscribe(["t"])
    assert [m.offset() for m in a3.consume(100, 0.0)] == [4, 5, 6, 7, 8, 9]
    b = Consumer(conf(tmp_path, "B")); b.subscribe(["t"])          # another group starts from the beginning
    assert len(b.consume(100, 0.0)) == 10
    latest = Consumer({**conf(tmp_path, "C"), "auto.offset.reset": "latest"}); latest.subscribe(["t"])
    assert latest.consume(100, 0.0) == []
    p.produce("t", value=b"new"); p.flush()
    assert [m.value() for m in latest.consume(100, 0.0)] == [b"new"]


def test_purge_moves_the_low_watermark(tmp_path):
    p = Producer({"log.dir": str(tmp_path)})
    for i in range(7):
        p.produce("ride_requests", value=b"r")
    p.flush()
    b = Broker(str(tmp_path))
    assert b.delete_records("ride_requests") == 7
    assert b.get_watermark_offsets(TopicPartition("ride_requests", 0)) == (7, 7) and b.count("ride_requests") == 0
    p.produce("ride_requests", value=b"after"); p.flush()
    c = Consumer(conf(tmp_path)); c.subscribe(["ride_requests"])
    msgs = c.consume(10, 0.0)
    assert [(m.offset(), m.value()) for m in msgs] == [(7, b"after")]


def test_concurrent_producers_lose_nothing(tmp_path):
    def work(tag):
        p = Producer({"log.dir": str(tmp_path)})
        for i in range(200):
            p.produce("documents", key=f"{tag}-{i}", value=b"d" * (i % 50))
            if i % 7 == 0:
                p.flush()
        p.flush()
    ts = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in ts]; [t.join() for t in ts]
    c = Consumer(conf(tmp_path)); c.subscribe(["documents"])
    msgs = c.consume(10_000, 0.0)
    assert len(msgs) == 1600 and [m.offset() for m in msgs] == list(range(1600))
    assert len({m.key() for m in msgs}) == 1600
    for m in msgs:
        tag, i = m.key().decode().split("-")
        assert m.value() == b"d" * (int(i) % 50)


def test_kafka_adapter_drives_the_same_pipeline(tmp_path, monkeypatch):
    """transport.kafka (confluent_kafka behind the serve loop's transport interface) against an in-memory stand-in of
    the library: the Lab2 graph produces the same `search_results` bytes as over the file log, the sink rebuilds the
    table from the beginning of `documents_embed` whatever the group committed, and offsets are committed per batch."""
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_confluent_kafka as fck
    from doubles import PipelinedOracleIndex
    from qsa_b200.operator import VectorTable
    from qsa_b200.pipeline.serve import Lab2Pipeline
    from qsa_b200.transport import filelog, kafka
    monkeypatch.setitem(sys.modules, "confluent_kafka", fck)
    fck.reset()
    g = np.random.default_rng(4)
    dim = 32
    docs = [(f"d{i}", f"chunk {i}", g.standard_normal(dim).astype(np.float32)) for i in range(40)]
    queries = [(f"q{i}", g.standard_normal(dim).astype(np.float32)) for i in range(25)]
    outs = {}
    for name, mod, conf in (("file", filelog, None), ("kafka", kafka, {"bootstrap.servers": "fake:9092"})):
        logd = str(tmp_path / name)
        for attempt in range(2):                       # the second pipeline object = a restarted process, empty table
            table = VectorTable(PipelinedOracleIndex(dim))
            pipe = Lab2Pipeline(logd, table, k=3, max_batch=8, transport=mod, client_conf=conf)
            if attempt == 0:
                for d, c, v in docs:
                    pipe.producer.produce("documents_embed", key=d, value=pipe.codec.encode(
                        "documents_embed", {"document_id": d, "chunk": c, "embedding": v}))
                pipe.producer.flush()
            for q, v in queries[attempt * 12:(attempt + 1) * 12 + attempt]:
                pipe.producer.produce("queries_embed", value=pipe.codec.encode("queries_embed", {"query": q, "embedding": v}))
            pipe.producer.flush()
            pipe.run_until_idle()
            assert len(table) == 40                    # rebuilt from the log on the restart
        c = mod.Consumer(dict(conf or {}, **{"log.dir": logd, "group.id": "check", "enable.auto.commit": False}))
        c.subscribe(["search_results"])
        outs[name] = [m.value() for m in c.consume(100, 0.0)]
    assert len(outs["file"]) == 25 and outs["kafka"] == outs["file"]
    assert fck._BROKER["groups"]["sa-lab2"][("queries_embed", 0)] == 25


def test_raw_batch_reads_into_a_reusable_buffer(tmp_path):
    """consume_raw(out=buffer): the slice lands in the caller's buffer when it fits (a view of its prefix comes back),
    otherwise a fresh byte string -- the same bytes either way, and positions advance identically."""
    import numpy as np
    from qsa_b200.transport.filelog import Consumer, Producer
    logd = str(tmp_path)
    p = Producer({"log.dir": logd})
    vals = [bytes([i % 251]) * (50 + 13 * i) for i in range(40)]
    for i, v in enumerate(vals):
        p.produce("t", key=(None if i % 3 else f"k{i}".encode()), value=v)
    p.flush()
    plain = Consumer({"log.dir": logd, "group.id": "a"}); plain.subscribe(["t"])
    buffered = Consumer({"log.dir": logd, "group.id": "b"}); buffered.subscribe(["t"])
    big, small = np.empty(1 << 16, np.uint8), bytearray(8)
    for n, out in ((7, big), (9, small), (100, big)):
        a = plain.consume_raw(n)
        b = buffered.consume_raw(n, out)
        assert a[:4] == b[:4] and bytes(a[4]) == bytes(b[4])
        if out is big:
            assert isinstance(b[4], memoryview) and bytes(big[:len(b[4])]) == bytes(a[4])     # it really is the caller's buffer
        else:
            assert isinstance(b[4], bytes)
    assert plain.consume_raw(5) is None and buffered.consume_raw(5, big) is None
