"""Shared by the drop-in CLIs: where the local topics live and how a record gets onto one.

The reference resolves Confluent Cloud credentials from Terraform state and shells out to
``confluent kafka topic produce`` (scripts/common/terraform.py:81-170, scripts/publish_docs.py:289-331).  The local
engine needs neither: the "cluster" is a log directory (``--log-dir``, env ``SA_LOG_DIR``, default
``./.sa_topics``), and producing is an in-process append of a Confluent-framed Avro record.
"""
from __future__ import annotations

import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qsa_b200.transport.filelog import Producer  # noqa: E402
from qsa_b200.wire import avro  # noqa: E402
from qsa_b200.wire.registry import SchemaRegistry  # noqa: E402

DEFAULT_LOG_DIR = ".sa_topics"


def resolve_log_dir(arg: str | None) -> str:
    return os.path.abspath(arg or os.environ.get("SA_LOG_DIR") or DEFAULT_LOG_DIR)


def setup_logging(verbose: bool = False, default_level: str = "INFO") -> logging.Logger:
    """Same helper signature as scripts/common/logging_utils.py:11-42."""
    level = logging.DEBUG if verbose else getattr(logging, default_level)
    logging.basicConfig(level=level, format="%(asctime)s - %(levelname)s - %(message)s", force=True)
    return logging.getLogger("scripts")


class AvroJsonProducer:
    """Consumes what the `confluent` CLI would read on stdin -- Avro-JSON with union wrapping, optionally
    ``key:json`` lines (``--parse-key --delimiter :``) -- and appends Confluent-framed Avro binary to a topic."""

    def __init__(self, log_dir: str, topic: str, schema: dict):
        self.topic = topic
        self.schema = avro.parse_schema(schema)
        self.producer = Producer({"log.dir": log_dir})
        self.schema_id = SchemaRegistry(log_dir).register(f"{topic}-value", self.schema)

    def produce_avro_json(self, value_json: dict, key: str | None = None) -> None:
        value = avro.from_avro_json(self.schema, value_json)
        self.producer.produce(self.topic, key=key, value=avro.frame(self.schema_id, avro.encode(self.schema, value)))
        self.producer.flush()

    def produce_line(self, line: str, parse_key: bool = False, delimiter: str = ":") -> None:
        import json
        line = line.rstrip("\n")
        key = None
        if parse_key:
            key, _, line = line.partition(delimiter)  # split on the FIRST delimiter only
        self.produce_avro_json(json.loads(line), key)
