"""File-backed topic log with the Kafka semantics the reference's scripts and tests lean on.

There is no broker, no librdkafka and no `confluent` binary in this environment, so the local pipeline keeps
its topics in a directory: one append-only data file + one offset index per topic-partition.  The API mirrors
the slice of ``confluent_kafka`` the reference uses, so code reads the same on either transport:

  Producer.produce(topic, key=, value=, partition=0) / poll(0) / flush()   scripts/publish_lab3_data.py:312-317,385-389
  Consumer.subscribe([topic]) / poll(timeout) -> Message | None / commit()  scripts/capture_lab3_data.py:106-144
  Message.key() .value() .partition() .offset() .timestamp() .error()
  Consumer.list_topics(topic) / get_watermark_offsets(tp) -> (low, high)   testing/helpers/kafka_helper.py:88-118
  Admin.delete_records (purge up to the high watermark)                    scripts/publish_lab3_data.py:216-261

Consumer defaults follow the reference's config: ``auto.offset.reset=earliest``, ``enable.auto.commit=False``
(kafka_helper.py:43-52).  Delivery is at-least-once: a consumer group's committed offset moves only on commit().

Record framing in ``<topic>-<partition>.log``:  u32 key_len (0xFFFFFFFF = null key) | key | u32 value_len | value
| i64 timestamp_ms, all little-endian.  ``<topic>-<partition>.idx`` holds one u64 file position per record, so
offset -> position is a single read and the high watermark is ``size(idx) / 8``.  Writers serialise on flock.
"""
from __future__ import annotations

import fcntl
import json
import os
import struct
import time
from dataclasses import dataclass

NULL_LEN = 0xFFFFFFFF


@dataclass(frozen=True)
class TopicPartition:
    topic: str
    partition: int = 0
    offset: int = -1


class Message:
    __slots__ = ("_t", "_p", "_o", "_k", "_v", "_ts")

    def __init__(self, topic, partition, offset, key, value, ts):
        self._t, self._p, self._o, self._k, self._v, self._ts = topic, partition, offset, key, value, ts

    def topic(self):
        return self._t

    def partition(self):
        return self._p

    def offset(self):
        return self._o

    def key(self):
        return self._k

    def value(self):
        return self._v

    def timestamp(self):
        return (1, self._ts)  # (TIMESTAMP_CREATE_TIME, ms) like confluent_kafka

    def headers(self):
        return None

    def error(self):
        return None


class _Partition:
    def __init__(self, root: str, topic: str, partition: int):
        self.topic, self.partition = topic, partition
        self.log_path = os.path.join(root, f"{topic}-{partition}.log")
        self.idx_path = os.path.join(root, f"{topic}-{partition}.idx")
        self.meta_path = os.path.join(root, f"{topic}-{partition}.meta")

    def exists(self) -> bool:
        return os.path.exists(self.idx_path)

    def create(self) -> None:
        for p in (self.log_path, self.idx_path):
            if not os.path.exists(p):
                open(p, "ab").close()

    def high(self) -> int:
        try:
            return os.path.getsize(self.idx_path) // 8
        except FileNotFoundError:
            return 0

    def low(self) -> int:
        try:
            with open(self.meta_path) as f:
                return int(json.load(f).get("low", 0))
        except (FileNotFoundError, ValueError):
            return 0

    def set_low(self, low: int) -> None:
        tmp = self.meta_path + ".tmp"
        with open(tmp, "w") as f:
            json.dump({"low": int(low)}, f)
        os.replace(tmp, self.meta_path)

    def append_many(self, records) -> int:
        """records: iterable of (key bytes|None, value bytes|None, ts_ms).  Returns the first offset written."""
        self.create()
        with open(self.log_path, "ab") as lf, open(self.idx_path, "ab") as xf:
            fcntl.flock(lf, fcntl.LOCK_EX)
            try:
                lf.seek(0, os.SEEK_END)
                xf.seek(0, os.SEEK_END)
                first = xf.tell() // 8
                pos = lf.tell()
                data, idx = bytearray(), bytearray()
                for key, value, ts in records:
                    idx += struct.pack("<Q", pos + len(data))
                    if key is None:
                        data += struct.pack("<I", NULL_LEN)
                    else:
                        data += struct.pack("<I", len(key)) + key
                    if value is None:
                        data += struct.pack("<I", NULL_LEN)
                    else:
                        data += struct.pack("<I", len(value)) + value
                    data += struct.pack("<q", int(ts))
                lf.write(data)
                lf.flush()
                xf.write(idx)  # index last: a record is visible only once its bytes are in the log
                xf.flush()
                return first
            finally:
                fcntl.flock(lf, fcntl.LOCK_UN)

    def append_framed(self, data, rel_positions) -> int:
        """Append records that are ALREADY in the log's framing (one buffer, e.g. from sa_wire_encode_*):
        rel_positions[i] is the offset of record i inside `data`.  Returns the first offset written."""
        self.create()
        with open(self.log_path, "ab") as lf, open(self.idx_path, "ab") as xf:
            fcntl.flock(lf, fcntl.LOCK_EX)
            try:
                lf.seek(0, os.SEEK_END)
                xf.seek(0, os.SEEK_END)
                first = xf.tell() // 8
                pos = lf.tell()
                lf.write(data)
                lf.flush()
                import numpy as np
                xf.write((np.asarray(rel_positions, dtype="<u8") + np.uint64(pos)).tobytes())
                xf.flush()
                return first
            finally:
                fcntl.flock(lf, fcntl.LOCK_UN)

    def read_raw(self, offset: int, max_records: int, out=None):
        """Records [offset, offset+n) that exist right now as ONE byte string in the log's framing (to be split by
        sa_wire_split_log): returns (n, bytes).  No per-record Python objects are created.  ``out``: a writable buffer
        (bytearray / numpy uint8 array) to read into when the slice fits -- the result is then a memoryview of its
        prefix and no 8 MB allocation (with its page faults) happens per batch."""
        hi = self.high()
        if offset >= hi:
            return 0, b""
        n = min(max_records, hi - offset)
        with open(self.idx_path, "rb") as xf:
            xf.seek(offset * 8)
            raw = xf.read((n + 1) * 8)            # one position past the slice, if it exists, bounds the read
        positions = struct.unpack(f"<{len(raw) // 8}Q", raw)
        with open(self.log_path, "rb", buffering=0) as lf:
            lf.seek(positions[0])
            size = (positions[n] - positions[0]) if len(positions) > n else (os.fstat(lf.fileno()).st_size - positions[0])
            if out is not None and size <= len(out):
                view = memoryview(out).cast("B")[:size]
                got = 0
                while got < size:
                    r = lf.readinto(view[got:])
                    if not r:
                        break
                    got += r
                data = view[:got]
            else:
                data = lf.read(size)
        return n, data

    def read(self, offset: int, max_records: int):
        """Messages [offset, offset+max_records) that exist right now."""
        hi = self.high()
        if offset >= hi:
            return []
        n = min(max_records, hi - offset)
        with open(self.idx_path, "rb") as xf:
            xf.seek(offset * 8)
            raw = xf.read(n * 8)
        positions = struct.unpack(f"<{len(raw) // 8}Q", raw)
        out = []
        with open(self.log_path, "rb") as lf:
            lf.seek(positions[0])
            for i, _ in enumerate(positions):
                (klen,) = struct.unpack("<I", lf.read(4))
                key = None if klen == NULL_LEN else lf.read(klen)
                (vlen,) = struct.unpack("<I", lf.read(4))
                value = None if vlen == NULL_LEN else lf.read(vlen)
                (ts,) = struct.unpack("<q", lf.read(8))
                out.append(Message(self.topic, self.partition, offset + i, key, value, ts))
        return out


class Broker:
    """The log directory: topics, partitions, consumer-group offsets."""

    def __init__(self, root: str):
        self.root = os.path.abspath(root)
        os.makedirs(self.root, exist_ok=True)

    # ---- topics
    def partition(self, topic: str, partition: int = 0) -> _Partition:
        return _Partition(self.root, topic, partition)

    def create_topic(self, topic: str, num_partitions: int = 1) -> None:
        for p in range(num_partitions):
            self.partition(topic, p).create()

    def list_topics(self) -> dict[str, list[int]]:
        topics: dict[str, list[int]] = {}
        for name in os.listdir(self.root):
            if name.endswith(".idx"):
                t, _, p = name[:-4].rpartition("-")
                topics.setdefault(t, []).append(int(p))
        return {t: sorted(ps) for t, ps in topics.items()}

    def get_watermark_offsets(self, tp: TopicPartition) -> tuple[int, int]:
        part = self.partition(tp.topic, tp.partition)
        return part.low(), part.high()

    def count(self, topic: str) -> int:
        return sum(h - l for l, h in (self.get_watermark_offsets(TopicPartition(topic, p))
                                      for p in self.list_topics().get(topic, [])))

    def delete_records(self, topic: str) -> int:
        """Purge: move every partition's low watermark up to its high watermark.  Returns records purged."""
        purged = 0
        for p in self.list_topics().get(topic, []):
            part = self.partition(topic, p)
            lo, hi = part.low(), part.high()
            part.set_low(hi)
            purged += hi - lo
        return purged

    # ---- consumer-group offsets
    def _group_path(self, group: str) -> str:
        return os.path.join(self.root, f"_group.{group}.json")

    def committed(self, group: str) -> dict[str, int]:
        try:
            with open(self._group_path(group)) as f:
                return json.load(f)
        except (FileNotFoundError, ValueError):
            return {}

    def commit(self, group: str, offsets: dict[str, int]) -> None:
        cur = self.committed(group)
        cur.update(offsets)
        tmp = self._group_path(group) + ".tmp"
        with open(tmp, "w") as f:
            json.dump(cur, f)
        os.replace(tmp, self._group_path(group))


class Producer:
    def __init__(self, conf: dict | str):
        root = conf if isinstance(conf, str) else conf["log.dir"]
        self.broker = Broker(root)
        self._pending: dict[tuple[str, int], list] = {}

    def produce(self, topic, value=None, key=None, partition=0, timestamp=None, on_delivery=None, callback=None):
        if isinstance(key, str):
            key = key.encode("utf-8")
        if isinstance(value, str):
            value = value.encode("utf-8")
        ts = int(time.time() * 1000) if timestamp is None else int(timestamp)
        self._pending.setdefault((topic, int(partition)), []).append((key, value, ts))
        cb = on_delivery or callback
        if cb is not None:
            cb(None, Message(topic, partition, -1, key, value, ts))
        if sum(len(v) for v in self._pending.values()) >= 4096:
            self.flush()

    def produce_framed(self, topic, data, rel_positions, partition=0) -> int:
        """Append a batch of records that were framed natively (sa_wire_encode_*) with one write.  Records produced
        earlier on this producer are flushed first, so the topic keeps the producer's order."""
        self.flush()
        return self.broker.partition(topic, int(partition)).append_framed(data, rel_positions)

    def poll(self, timeout=0):
        return 0

    def flush(self, timeout=None):
        for (topic, part), recs in list(self._pending.items()):
            if recs:
                self.broker.partition(topic, part).append_many(recs)
        self._pending.clear()
        return 0

    def __len__(self):
        return sum(len(v) for v in self._pending.values())


class Consumer:
    def __init__(self, conf: dict):
        self.broker = Broker(conf["log.dir"])
        self.group = conf.get("group.id", "default")
        self.reset = conf.get("auto.offset.reset", "earliest")
        self.auto_commit = bool(conf.get("enable.auto.commit", False))
        self._topics: list[str] = []
        self._pos: dict[tuple[str, int], int] = {}
        self._rr = 0

    def subscribe(self, topics):
        self._topics = list(topics)
        self._pos.clear()
        pass

    def list_topics(self, topic=None):
        t = self.broker.list_topics()
        return {topic: t[topic]} if topic is not None and topic in t else ({} if topic is not None else t)

    def get_watermark_offsets(self, tp: TopicPartition, timeout=None):
        return self.broker.get_watermark_offsets(tp)

    def _assign(self):
        committed = self.broker.committed(self.group)
        for t in self._topics:
            for p in self.broker.list_topics().get(t, []):
                if (t, p) in self._pos:
                    continue
                lo, hi = self.broker.get_watermark_offsets(TopicPartition(t, p))
                start = None if t in getattr(self, "_ignore_committed", ()) else committed.get(f"{t}-{p}")
                if start is None:
                    start = lo if self.reset == "earliest" else hi
                self._pos[(t, p)] = max(int(start), lo)

    def consume(self, num_messages=1, timeout=0.0):
        """Up to num_messages messages, waiting at most `timeout` seconds for the first one."""
        deadline = time.monotonic() + max(0.0, timeout)
        out: list[Message] = []
        while True:
            self._assign()
            keys = sorted(self._pos)
            for i in range(len(keys)):
                tp = keys[(self._rr + i) % len(keys)]
                msgs = self.broker.partition(*tp).read(self._pos[tp], num_messages - len(out))
                if msgs:
                    self._pos[tp] = msgs[-1].offset() + 1
                    out.extend(msgs)
                if len(out) >= num_messages:
                    break
            self._rr += 1
            if out or time.monotonic() >= deadline:
                break
            time.sleep(min(0.01, max(0.0, deadline - time.monotonic())))
        if out and self.auto_commit:
            self.commit()
        return out

    def consume_raw(self, num_messages=1, out=None):
        """Batch form without per-record objects: up to num_messages records of ONE partition as
        (topic, partition, first_offset, n, bytes in the log's framing), or None when nothing is pending.
        ``out``: optional reusable buffer the bytes are read into (see _Partition.read_raw)."""
        self._assign()
        keys = sorted(self._pos)
        for i in range(len(keys)):
            tp = keys[(self._rr + i) % len(keys)]
            n, data = self.broker.partition(*tp).read_raw(self._pos[tp], num_messages, out)
            if n:
                first = self._pos[tp]
                self._pos[tp] = first + n
                self._rr += 1
                return tp[0], tp[1], first, n, data
        self._rr += 1
        return None

    def commit_upto(self, topic: str, partition: int, next_offset: int) -> None:
        """Commit `next_offset` as the group's position of one partition (pipelined stages commit batch by batch)."""
        self.broker.commit(self.group, {f"{topic}-{partition}": int(next_offset)})

    def seek(self, tp: TopicPartition) -> None:
        """Continue reading `tp.topic`/`tp.partition` at `tp.offset` (clamped to the low watermark), whatever the
        group has committed."""
        lo, _ = self.broker.get_watermark_offsets(tp)
        self._pos[(tp.topic, tp.partition)] = max(int(tp.offset), lo)

    def seek_to_beginning(self, topic: str) -> None:
        """Re-read `topic` from the low watermark of every partition (a consumer whose state is rebuilt from the log)."""
        for p in self.broker.list_topics().get(topic, []):
            self.seek(TopicPartition(topic, p, 0))
        self._ignore_committed = getattr(self, "_ignore_committed", set()) | {topic}

    def positions(self, topic: str) -> dict[int, int]:
        self._assign()
        return {p: o for (t, p), o in self._pos.items() if t == topic}

    def poll(self, timeout=0.0):
        msgs = self.consume(1, timeout)
        return msgs[0] if msgs else None

    def commit(self, message=None, asynchronous=False):
        if message is not None:
            self.broker.commit(self.group, {f"{message.topic()}-{message.partition()}": message.offset() + 1})
        else:
            self.broker.commit(self.group, {f"{t}-{p}": o for (t, p), o in self._pos.items()})

    def commit_offsets(self, messages) -> None:
        """Commit exactly up to (and including) the given messages -- for pipelined stages that have consumed further
        ahead than they have finished processing."""
        offs: dict[str, int] = {}
        for m in messages:
            key = f"{m.topic()}-{m.partition()}"
            offs[key] = max(offs.get(key, 0), m.offset() + 1)
        if offs:
            self.broker.commit(self.group, offs)

    def position(self, tp: TopicPartition) -> int:
        return self._pos.get((tp.topic, tp.partition), -1)

    def close(self):
        self._pos.clear()
