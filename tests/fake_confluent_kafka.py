"""An in-memory stand-in for the slice of ``confluent_kafka`` that transport/kafka.py uses (the real package cannot be
installed in the build image).  One process-wide broker: topics -> partitions -> records; consumer groups keep committed
offsets; method names, argument names and return shapes follow the library's."""
import time

OFFSET_BEGINNING = -2
OFFSET_END = -1
_BROKER = {"topics": {}, "groups": {}}


def reset():
    _BROKER["topics"].clear()
    _BROKER["groups"].clear()


class TopicPartition:
    def __init__(self, topic, partition=0, offset=-1001):
        self.topic, self.partition, self.offset = topic, partition, offset


class Message:
    def __init__(self, topic, partition, offset, key, value, ts):
        self._t, self._p, self._o, self._k, self._v, self._ts = topic, partition, offset, key, value, ts

    def topic(self): return self._t
    def partition(self): return self._p
    def offset(self): return self._o
    def key(self): return self._k
    def value(self): return self._v
    def timestamp(self): return (1, self._ts)
    def headers(self): return None
    def error(self): return None


class _TopicMeta:
    def __init__(self, n):
        self.partitions = {p: None for p in range(n)}


class _ClusterMeta:
    def __init__(self, topics):
        self.topics = topics


class Producer:
    def __init__(self, conf):
        assert "bootstrap.servers" in conf and "log.dir" not in conf
        self._pending = []

    def produce(self, topic, value=None, key=None, partition=0, timestamp=None, on_delivery=None):
        key = key.encode() if isinstance(key, str) else key
        value = value.encode() if isinstance(value, str) else value
        self._pending.append((topic, partition or 0, key, value, timestamp or int(time.time() * 1000), on_delivery))

    def poll(self, timeout=0):
        return 0

    def flush(self, timeout=None):
        for topic, part, key, value, ts, cb in self._pending:
            log = _BROKER["topics"].setdefault(topic, {}).setdefault(part, [])
            log.append((key, value, ts))
            if cb:
                cb(None, Message(topic, part, len(log) - 1, key, value, ts))
        self._pending = []
        return 0

    def __len__(self):
        return len(self._pending)


class Consumer:
    def __init__(self, conf):
        assert "bootstrap.servers" in conf and conf.get("enable.auto.commit") is False
        self.group = conf["group.id"]
        self.reset = conf.get("auto.offset.reset", "latest")
        self._topics, self._pos, self._on_assign = [], {}, None

    def subscribe(self, topics, on_assign=None):
        self._topics, self._on_assign = list(topics), on_assign

    def assign(self, partitions):
        for tp in partitions:
            off = tp.offset
            if off == OFFSET_BEGINNING:
                off = 0
            elif off < 0:
                off = _BROKER["groups"].get(self.group, {}).get((tp.topic, tp.partition),
                                                                0 if self.reset == "earliest" else len(_BROKER["topics"][tp.topic][tp.partition]))
            self._pos[(tp.topic, tp.partition)] = off

    def _rebalance(self):
        new = [TopicPartition(t, p) for t in self._topics for p in _BROKER["topics"].get(t, {}) if (t, p) not in self._pos]
        if new:
            if self._on_assign:
                self._on_assign(self, new)
            else:
                self.assign(new)

    def assignment(self):
        return [TopicPartition(t, p, o) for (t, p), o in self._pos.items()]

    def position(self, partitions):
        return [TopicPartition(tp.topic, tp.partition, self._pos.get((tp.topic, tp.partition), -1001)) for tp in partitions]

    def consume(self, num_messages=1, timeout=-1):
        self._rebalance()
        out = []
        for (t, p), off in sorted(self._pos.items()):
            log = _BROKER["topics"].get(t, {}).get(p, [])
            while off < len(log) and len(out) < num_messages:
                k, v, ts = log[off]
                out.append(Message(t, p, off, k, v, ts))
                off += 1
            self._pos[(t, p)] = off
        return out

    def poll(self, timeout=None):
        m = self.consume(1, timeout or 0)
        return m[0] if m else None

    def commit(self, message=None, offsets=None, asynchronous=True):
        g = _BROKER["groups"].setdefault(self.group, {})
        if message is not None:
            g[(message.topic(), message.partition())] = message.offset() + 1
        elif offsets is not None:
            for tp in offsets:
                g[(tp.topic, tp.partition)] = tp.offset
        else:
            g.update(self._pos)

    def seek(self, tp):
        self._pos[(tp.topic, tp.partition)] = 0 if tp.offset == OFFSET_BEGINNING else tp.offset

    def list_topics(self, topic=None, timeout=-1):
        t = _BROKER["topics"]
        return _ClusterMeta({k: _TopicMeta(len(v)) for k, v in t.items() if topic is None or k == topic})

    def get_watermark_offsets(self, tp, timeout=None):
        return 0, len(_BROKER["topics"].get(tp.topic, {}).get(tp.partition, []))

    def close(self):
        pass
