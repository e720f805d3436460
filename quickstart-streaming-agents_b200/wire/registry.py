"""Schema-registry stub: subject -> schema id, persisted next to the topic logs.

Stands in for Confluent Schema Registry, which the reference reaches through the `confluent` CLI
(``--schema-registry-endpoint``, scripts/lab2_publish_queries.py:133-138) or ``SchemaRegistryClient``
(scripts/publish_lab3_data.py:188-196).  Ids start at 100001 like Confluent Cloud's (the captured Lab3 data
uses 100008 / 100009, assets/lab3/data/ride_requests.jsonl).
"""
from __future__ import annotations

import fcntl
import json
import os

from . import avro

FIRST_ID = 100001


class SchemaRegistry:
    def __init__(self, root: str):
        self.path = os.path.join(root, "_schemas.json")
        os.makedirs(root, exist_ok=True)
        self._cache_by_id: dict[int, object] = {}

    def _load(self, f):
        f.seek(0)
        raw = f.read()
        return json.loads(raw) if raw.strip() else {"subjects": {}, "by_id": {}}

    def register(self, subject: str, schema) -> int:
        """Idempotent: the same (subject, schema) always maps to the same id."""
        schema = avro.parse_schema(schema)
        canon = avro.canonical(schema)
        with open(self.path, "a+") as f:
            fcntl.flock(f, fcntl.LOCK_EX)
            try:
                db = self._load(f)
                for sid, rec in db["by_id"].items():
                    if rec["subject"] == subject and rec["schema"] == canon:
                        return int(sid)
                sid = FIRST_ID + len(db["by_id"])
                db["by_id"][str(sid)] = {"subject": subject, "schema": canon}
                db["subjects"].setdefault(subject, []).append(sid)
                f.seek(0)
                f.truncate()
                json.dump(db, f)
                f.flush()
                return sid
            finally:
                fcntl.flock(f, fcntl.LOCK_UN)

    def get(self, schema_id: int):
        if schema_id in self._cache_by_id:
            return self._cache_by_id[schema_id]
        if not os.path.exists(self.path):
            raise KeyError(f"schema id {schema_id} not registered")
        with open(self.path, "r") as f:
            fcntl.flock(f, fcntl.LOCK_SH)
            try:
                db = self._load(f)
            finally:
                fcntl.flock(f, fcntl.LOCK_UN)
        rec = db["by_id"].get(str(schema_id))
        if rec is None:
            raise KeyError(f"schema id {schema_id} not registered")
        schema = json.loads(rec["schema"])
        self._cache_by_id[schema_id] = schema
        return schema

    def latest(self, subject: str):
        if not os.path.exists(self.path):
            return None
        with open(self.path, "r") as f:
            db = self._load(f)
        ids = db["subjects"].get(subject)
        return ids[-1] if ids else None
