"""The serve loop's transport interface over a real Kafka cluster (``confluent_kafka``).

The reference talks to Confluent Cloud through ``confluent_kafka`` (scripts/publish_lab3_data.py:201-214,312-317;
testing/helpers/kafka_helper.py:43-52,88-118).  ``transport.filelog`` mirrors the slice of that API the pipeline uses, so
this adapter is thin: it forwards to ``confluent_kafka.Producer`` / ``Consumer`` and adds the few batch helpers the serve
loop calls (``commit_offsets``, ``commit_upto``, ``seek_to_beginning``, ``positions``, ``produce_framed``).  Records are the
same bytes on either transport (Confluent wire format, wire/avro.py).

    from qsa_b200.transport import kafka
    pipe = Lab2Pipeline({"bootstrap.servers": "...", "security.protocol": "SASL_SSL", ...}, table, transport=kafka)

``confluent_kafka`` is not installable in the build image (no network), so this module is exercised in the tests against
an in-memory stand-in of the library (tests/test_transport.py); the import happens on first use and fails loudly.
The batch-at-a-time fast path of the search stage (``consume_raw``) is specific to the file log; over Kafka the stage uses
its per-message path (native batch decode still applies after the messages are collected).
"""
from __future__ import annotations

import struct

from .filelog import TopicPartition  # the same value type: topic, partition, offset

__all__ = ["Producer", "Consumer", "TopicPartition"]


def _ck():
    try:
        import confluent_kafka
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("transport.kafka needs the confluent_kafka package (the reference pins 2.14.0, uv.lock:440)") from e
    return confluent_kafka


def _client_conf(conf: dict) -> dict:
    return {k: v for k, v in conf.items() if k != "log.dir"}


class Producer:
    def __init__(self, conf: dict):
        self._p = _ck().Producer(_client_conf(conf))

    def produce(self, topic, value=None, key=None, partition=None, timestamp=None, on_delivery=None, callback=None):
        kw = {"key": key, "value": value}
        if partition is not None:
            kw["partition"] = int(partition)
        if timestamp is not None:
            kw["timestamp"] = int(timestamp)
        cb = on_delivery or callback
        if cb is not None:
            kw["on_delivery"] = cb
        self._p.produce(topic, **kw)
        self._p.poll(0)

    def produce_framed(self, topic, data, rel_positions, partition=None) -> int:
        """A batch framed for the file log (u32 key_len | key | u32 value_len | value | i64 ts per record): unframe and
        produce record by record."""
        buf = memoryview(data).cast("B")
        n = 0
        for pos in rel_positions:
            p = int(pos)
            (kl,) = struct.unpack_from("<I", buf, p)
            p += 4
            key = None
            if kl != 0xFFFFFFFF:
                key = bytes(buf[p:p + kl])
                p += kl
            (vl,) = struct.unpack_from("<I", buf, p)
            p += 4
            value = None if vl == 0xFFFFFFFF else bytes(buf[p:p + vl])
            self.produce(topic, value=value, key=key, partition=partition)
            n += 1
        return n

    def poll(self, timeout=0):
        return self._p.poll(timeout)

    def flush(self, timeout=None):
        return self._p.flush() if timeout is None else self._p.flush(timeout)

    def __len__(self):
        return len(self._p)


class Consumer:
    def __init__(self, conf: dict):
        ck = _ck()
        c = _client_conf(conf)
        c.setdefault("auto.offset.reset", "earliest")
        c.setdefault("enable.auto.commit", False)
        self._ck = ck
        self._c = ck.Consumer(c)
        self._topics: list[str] = []
        self._from_beginning: set[str] = set()

    def subscribe(self, topics):
        self._topics = list(topics)

        def on_assign(consumer, partitions):
            for tp in partitions:
                if tp.topic in self._from_beginning:
                    tp.offset = self._ck.OFFSET_BEGINNING
            consumer.assign(partitions)
        self._c.subscribe(self._topics, on_assign=on_assign)

    def consume(self, num_messages=1, timeout=0.0):
        return [m for m in self._c.consume(num_messages, max(timeout, 0.0)) if m.error() is None]

    def poll(self, timeout=0.0):
        m = self._c.poll(timeout)
        return m if m is not None and m.error() is None else None

    def commit(self, message=None, asynchronous=False):
        if message is not None:
            self._c.commit(message=message, asynchronous=asynchronous)
        else:
            self._c.commit(asynchronous=asynchronous)

    def commit_offsets(self, messages) -> None:
        best: dict[tuple[str, int], int] = {}
        for m in messages:
            k = (m.topic(), m.partition())
            best[k] = max(best.get(k, 0), m.offset() + 1)
        if best:
            self._c.commit(offsets=[self._ck.TopicPartition(t, p, o) for (t, p), o in best.items()], asynchronous=False)

    def commit_upto(self, topic: str, partition: int, next_offset: int) -> None:
        self._c.commit(offsets=[self._ck.TopicPartition(topic, int(partition), int(next_offset))], asynchronous=False)

    def seek(self, tp: TopicPartition) -> None:
        self._c.seek(self._ck.TopicPartition(tp.topic, tp.partition, tp.offset))

    def seek_to_beginning(self, topic: str) -> None:
        """Read `topic` from its beginning whatever the group committed (applied when partitions are assigned)."""
        self._from_beginning.add(topic)
        for tp in self._c.assignment():
            if tp.topic == topic:
                self._c.seek(self._ck.TopicPartition(tp.topic, tp.partition, self._ck.OFFSET_BEGINNING))

    def positions(self, topic: str) -> dict[int, int]:
        return {tp.partition: tp.offset for tp in self._c.position(self._c.assignment()) if tp.topic == topic and tp.offset >= 0}

    def list_topics(self, topic=None):
        md = self._c.list_topics(topic)
        return {t: sorted(meta.partitions) for t, meta in md.topics.items()}

    def get_watermark_offsets(self, tp: TopicPartition, timeout=None):
        return self._c.get_watermark_offsets(self._ck.TopicPartition(tp.topic, tp.partition))

    def close(self):
        self._c.close()
