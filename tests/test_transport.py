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
