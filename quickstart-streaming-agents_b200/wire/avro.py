"""Schema-driven Avro codec (binary and Avro-JSON) plus the Confluent wire framing.

The reference never encodes Avro itself: Lab2 pipes Avro-JSON into ``confluent kafka topic produce
--value-format avro`` (scripts/lab2_publish_queries.py:113-157, scripts/publish_docs.py:260-331) and Lab3/4 use
the ``avro`` / ``confluent_kafka`` packages (scripts/publish_lab3_data.py:96-122,
testing/helpers/kafka_helper.py:70-86).  None of those exist here, so this module implements the subset of
the Avro 1.x specification those paths exercise:

  null, boolean, int, long (zig-zag varint), float, double (IEEE little-endian), bytes, string (length-prefixed),
  record (fields in order), array (blocks, negative counts carry a byte size), union (branch index + value),
  logical timestamp-millis (as long).

Wire framing: byte 0 = 0x00, bytes 1..4 = big-endian schema id, then the Avro body
(scripts/publish_lab3_data.py:96-114).

Embedding vectors (``["null", {"type":"array","items":["null","float"]}]``, the shape Flink derives for
ARRAY<FLOAT> columns) are the only large values on the path; ``decode_float_array_fast`` /
``encode_float_array_fast`` move them with numpy instead of a Python loop.
"""
from __future__ import annotations

import json
import struct
from typing import Any

import numpy as np

MAGIC = 0

PRIMITIVES = {"null", "boolean", "int", "long", "float", "double", "bytes", "string"}


class AvroError(ValueError):
    pass


# ----------------------------------------------------------------------------------------------------
# schema handling
# ----------------------------------------------------------------------------------------------------
def parse_schema(schema: Any) -> Any:
    """Accept a JSON string or a Python structure; return the Python structure (validated lightly)."""
    if isinstance(schema, (bytes, str)) and not (isinstance(schema, str) and schema in PRIMITIVES):
        schema = json.loads(schema)
    _validate(schema)
    return schema


def _validate(s: Any) -> None:
    if isinstance(s, str):
        if s not in PRIMITIVES:
            raise AvroError(f"unsupported named type reference {s!r}")
        return
    if isinstance(s, list):
        if not s:
            raise AvroError("empty union")
        for b in s:
            _validate(b)
        return
    if isinstance(s, dict):
        t = s.get("type")
        if t == "record":
            for f in s["fields"]:
                _validate(f["type"])
        elif t == "array":
            _validate(s["items"])
        elif isinstance(t, (dict, list)):
            _validate(t)
        elif t in PRIMITIVES:
            return
        else:
            raise AvroError(f"unsupported schema type {t!r}")
        return
    raise AvroError(f"bad schema node {s!r}")


def _type_name(s: Any) -> str:
    """Branch name used by Avro-JSON union wrapping."""
    if isinstance(s, str):
        return s
    if isinstance(s, dict):
        t = s["type"]
        if t == "record":
            ns = s.get("namespace")
            return f"{ns}.{s['name']}" if ns else s["name"]
        if isinstance(t, str):
            return t
        return _type_name(t)
    raise AvroError("nested unions are not allowed")


def canonical(schema: Any) -> str:
    return json.dumps(schema, sort_keys=True, separators=(",", ":"))


# ----------------------------------------------------------------------------------------------------
# binary primitives
# ----------------------------------------------------------------------------------------------------
def write_long(out: bytearray, n: int) -> None:
    n = (n << 1) ^ (n >> 63)
    n &= 0xFFFFFFFFFFFFFFFF
    while n > 0x7F:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)


def read_long(buf: bytes, pos: int) -> tuple[int, int]:
    shift = 0
    acc = 0
    while True:
        if pos >= len(buf):
            raise AvroError("truncated varint")
        b = buf[pos]
        pos += 1
        acc |= (b & 0x7F) << shift
        if not (b & 0x80):
            break
        shift += 7
        if shift > 63:
            raise AvroError("varint too long")
    return (acc >> 1) ^ -(acc & 1), pos


# ----------------------------------------------------------------------------------------------------
# encode
# ----------------------------------------------------------------------------------------------------
def _branch_of(value: Any, union: list) -> int:
    for i, b in enumerate(union):
        n = _type_name(b)
        if value is None and n == "null":
            return i
        if value is None:
            continue
        if n == "boolean" and isinstance(value, bool):
            return i
        if n in ("int", "long") and isinstance(value, (int, np.integer)) and not isinstance(value, bool):
            return i
        if n in ("float", "double") and isinstance(value, (float, int, np.floating, np.integer)) and not isinstance(value, bool):
            return i
        if n == "string" and isinstance(value, str):
            return i
        if n == "bytes" and isinstance(value, (bytes, bytearray)):
            return i
        if n == "array" and isinstance(value, (list, tuple, np.ndarray)):
            return i
        if isinstance(b, dict) and b.get("type") == "record" and isinstance(value, dict):
            return i
    raise AvroError(f"value {type(value).__name__} matches no branch of {union}")


def _is_nullable_float_array(s: Any) -> bool:
    return isinstance(s, dict) and s.get("type") == "array" and s.get("items") in (["null", "float"], "float")


def encode_float_array_fast(out: bytearray, vec: np.ndarray, nullable_items: bool) -> None:
    """One-block Avro array of floats: count, items (each 0x02 + 4 B LE when items are ["null","float"]), 0x00."""
    v = np.ascontiguousarray(vec, dtype="<f4")
    n = v.shape[0]
    if n:
        write_long(out, n)
        if nullable_items:
            rec = np.empty(n, dtype=np.dtype([("b", "u1"), ("f", "<f4")]))
            rec["b"] = 2
            rec["f"] = v
            out += rec.tobytes()
        else:
            out += v.tobytes()
    out.append(0)


def _encode(out: bytearray, s: Any, v: Any) -> None:
    if isinstance(s, list):
        i = _branch_of(v, s)
        write_long(out, i)
        _encode(out, s[i], v)
        return
    if isinstance(s, dict):
        t = s["type"]
        if t == "record":
            if not isinstance(v, dict):
                raise AvroError(f"record {s.get('name')} needs a dict")
            for f in s["fields"]:
                if f["name"] in v:
                    _encode(out, f["type"], v[f["name"]])
                elif "default" in f:
                    _encode(out, f["type"], f["default"])
                else:
                    raise AvroError(f"missing field {f['name']}")
            return
        if t == "array":
            if _is_nullable_float_array(s) and isinstance(v, np.ndarray):
                encode_float_array_fast(out, v, isinstance(s["items"], list))
                return
            items = list(v)
            if items:
                write_long(out, len(items))
                for x in items:
                    _encode(out, s["items"], x)
            out.append(0)
            return
        _encode(out, t, v)  # {"type": "long", "logicalType": ...} and friends
        return
    if s == "null":
        if v is not None:
            raise AvroError("null expected")
    elif s == "boolean":
        out.append(1 if v else 0)
    elif s in ("int", "long"):
        write_long(out, int(v))
    elif s == "float":
        out += struct.pack("<f", float(v))
    elif s == "double":
        out += struct.pack("<d", float(v))
    elif s == "string":
        b = v.encode("utf-8")
        write_long(out, len(b))
        out += b
    elif s == "bytes":
        write_long(out, len(v))
        out += bytes(v)
    else:
        raise AvroError(f"unsupported type {s!r}")


def encode(schema: Any, value: Any) -> bytes:
    out = bytearray()
    _encode(out, schema, value)
    return bytes(out)


# ----------------------------------------------------------------------------------------------------
# decode
# ----------------------------------------------------------------------------------------------------
def decode_float_array_fast(buf: bytes, pos: int, nullable_items: bool) -> tuple[np.ndarray, int]:
    """Avro array of (nullable) floats -> float32 vector.  Handles multi-block arrays and blocks with a byte
    size; a null item becomes NaN.  The common single-block all-non-null case is one strided numpy view."""
    parts = []
    while True:
        n, pos = read_long(buf, pos)
        if n == 0:
            break
        if n < 0:
            n = -n
            _, pos = read_long(buf, pos)  # byte size of the block (unused)
        if nullable_items:
            stride = 5
            end = pos + n * stride
            if end <= len(buf):
                rec = np.frombuffer(buf, dtype=np.dtype([("b", "u1"), ("f", "<f4")]), count=n, offset=pos)
                if (rec["b"] == 2).all():
                    parts.append(rec["f"].astype(np.float32))
                    pos = end
                    continue
            vals = np.empty(n, dtype=np.float32)  # slow path: some items are null (branch 0, no payload)
            for i in range(n):
                br, pos = read_long(buf, pos)
                if br == 0:
                    vals[i] = np.nan
                elif br == 1:
                    if pos + 4 > len(buf):
                        raise AvroError("truncated float")
                    vals[i] = struct.unpack_from("<f", buf, pos)[0]
                    pos += 4
                else:
                    raise AvroError("bad union branch in float array")
            parts.append(vals)
        else:
            end = pos + 4 * n
            if end > len(buf):
                raise AvroError("truncated float array")
            parts.append(np.frombuffer(buf, dtype="<f4", count=n, offset=pos).astype(np.float32))
            pos = end
    if not parts:
        return np.empty(0, dtype=np.float32), pos
    return (parts[0] if len(parts) == 1 else np.concatenate(parts)), pos


def _decode(buf: bytes, pos: int, s: Any) -> tuple[Any, int]:
    if isinstance(s, list):
        i, pos = read_long(buf, pos)
        if not 0 <= i < len(s):
            raise AvroError(f"union branch {i} out of range")
        return _decode(buf, pos, s[i])
    if isinstance(s, dict):
        t = s["type"]
        if t == "record":
            rec = {}
            for f in s["fields"]:
                rec[f["name"]], pos = _decode(buf, pos, f["type"])
            return rec, pos
        if t == "array":
            if _is_nullable_float_array(s):
                return decode_float_array_fast(buf, pos, isinstance(s["items"], list))
            items = []
            while True:
                n, pos = read_long(buf, pos)
                if n == 0:
                    break
                if n < 0:
                    n = -n
                    _, pos = read_long(buf, pos)
                for _ in range(n):
                    x, pos = _decode(buf, pos, s["items"])
                    items.append(x)
            return items, pos
        return _decode(buf, pos, t)
    if s == "null":
        return None, pos
    if s == "boolean":
        if pos >= len(buf):
            raise AvroError("truncated boolean")
        return buf[pos] != 0, pos + 1
    if s in ("int", "long"):
        return read_long(buf, pos)
    if s == "float":
        if pos + 4 > len(buf):
            raise AvroError("truncated float")
        return struct.unpack_from("<f", buf, pos)[0], pos + 4
    if s == "double":
        if pos + 8 > len(buf):
            raise AvroError("truncated double")
        return struct.unpack_from("<d", buf, pos)[0], pos + 8
    if s in ("string", "bytes"):
        n, pos = read_long(buf, pos)
        if n < 0 or pos + n > len(buf):
            raise AvroError("truncated string/bytes")
        raw = bytes(buf[pos:pos + n])
        return (raw.decode("utf-8") if s == "string" else raw), pos + n
    raise AvroError(f"unsupported type {s!r}")


def decode(schema: Any, buf: bytes, pos: int = 0, require_all: bool = True) -> Any:
    v, end = _decode(buf, pos, schema)
    if require_all and end != len(buf):
        raise AvroError(f"{len(buf) - end} trailing bytes after Avro datum")
    return v


# ----------------------------------------------------------------------------------------------------
# Confluent framing
# ----------------------------------------------------------------------------------------------------
def frame(schema_id: int, body: bytes) -> bytes:
    return bytes([MAGIC]) + struct.pack(">I", schema_id) + body


def unframe(raw: bytes) -> tuple[int, bytes]:
    if len(raw) < 5:
        raise AvroError(f"Avro payload too short ({len(raw)} bytes)")
    if raw[0] != MAGIC:
        raise AvroError(f"Invalid Avro magic byte: {raw[0]}")
    return struct.unpack(">I", raw[1:5])[0], raw[5:]


# ----------------------------------------------------------------------------------------------------
# Avro-JSON (what `confluent kafka topic produce --value-format avro` reads on stdin)
# ----------------------------------------------------------------------------------------------------
def to_avro_json(schema: Any, v: Any) -> Any:
    """Python value -> Avro-JSON structure with union wrapping ({"string": "x"}, null stays null)."""
    if isinstance(schema, list):
        i = _branch_of(v, schema)
        b = schema[i]
        if _type_name(b) == "null":
            return None
        return {_type_name(b): to_avro_json(b, v)}
    if isinstance(schema, dict):
        t = schema["type"]
        if t == "record":
            return {f["name"]: to_avro_json(f["type"], v.get(f["name"], f.get("default"))) for f in schema["fields"]}
        if t == "array":
            return [to_avro_json(schema["items"], x) for x in (v.tolist() if isinstance(v, np.ndarray) else v)]
        return to_avro_json(t, v)
    if schema in ("float", "double"):
        return float(v)
    if schema in ("int", "long"):
        return int(v)
    return v


def from_avro_json(schema: Any, j: Any) -> Any:
    """Avro-JSON structure -> plain Python value (inverse of ``to_avro_json``)."""
    if isinstance(schema, list):
        if j is None:
            if not any(_type_name(b) == "null" for b in schema):
                raise AvroError("null is not a branch of this union")
            return None
        if not isinstance(j, dict) or len(j) != 1:
            raise AvroError(f"union value must be null or a single-key object, got {j!r}")
        (name, inner), = j.items()
        for b in schema:
            if _type_name(b) == name or _type_name(b).rsplit(".", 1)[-1] == name:
                return from_avro_json(b, inner)
        raise AvroError(f"union branch {name!r} not in schema")
    if isinstance(schema, dict):
        t = schema["type"]
        if t == "record":
            if not isinstance(j, dict):
                raise AvroError("record must be a JSON object")
            return {f["name"]: from_avro_json(f["type"], j[f["name"]]) if f["name"] in j else f.get("default")
                    for f in schema["fields"]}
        if t == "array":
            return [from_avro_json(schema["items"], x) for x in j]
        return from_avro_json(t, j)
    if schema == "null":
        return None
    if schema in ("float", "double"):
        return float(j)
    if schema in ("int", "long"):
        return int(j)
    return j


# ----------------------------------------------------------------------------------------------------
# compiled codecs: one closure tree per schema, built once -- the serve loop's per-record hot path.
# (The generic encode()/decode() above are the readable statement of the rules; tests pin these against them.)
# ----------------------------------------------------------------------------------------------------
_pack_f = struct.Struct("<f").pack
_pack_d = struct.Struct("<d").pack
_unpack_f = struct.Struct("<f").unpack_from
_unpack_d = struct.Struct("<d").unpack_from


def _enc_long(out, n):
    n = ((n << 1) ^ (n >> 63)) & 0xFFFFFFFFFFFFFFFF
    while n > 0x7F:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)


def _enc_string(out, v):
    b = v.encode("utf-8")
    _enc_long(out, len(b))
    out += b


def compile_encoder(schema):
    """Returns f(out: bytearray, value) equivalent to ``_encode(out, schema, value)``."""
    s = schema
    if isinstance(s, list):
        if len(s) == 2 and s[0] == "null":            # the ubiquitous ["null", T]
            inner = compile_encoder(s[1])

            def enc_nullable(out, v):
                if v is None:
                    out.append(0)
                else:
                    out.append(2)
                    inner(out, v)
            return enc_nullable
        encs = [compile_encoder(b) for b in s]

        def enc_union(out, v):
            i = _branch_of(v, s)
            _enc_long(out, i)
            encs[i](out, v)
        return enc_union
    if isinstance(s, dict):
        t = s["type"]
        if t == "record":
            fields = [(f["name"], compile_encoder(f["type"]), "default" in f, f.get("default")) for f in s["fields"]]

            def enc_record(out, v):
                for name, enc, has_default, default in fields:
                    if name in v:
                        enc(out, v[name])
                    elif has_default:
                        enc(out, default)
                    else:
                        raise AvroError(f"missing field {name}")
            return enc_record
        if t == "array":
            item = compile_encoder(s["items"])
            fast = _is_nullable_float_array(s)
            nullable_items = isinstance(s["items"], list)

            def enc_array(out, v):
                if fast and isinstance(v, np.ndarray):
                    encode_float_array_fast(out, v, nullable_items)
                    return
                n = len(v)
                if n:
                    _enc_long(out, n)
                    for x in v:
                        item(out, x)
                out.append(0)
            return enc_array
        return compile_encoder(t)
    if s == "null":
        return lambda out, v: None
    if s == "boolean":
        return lambda out, v: out.append(1 if v else 0)
    if s in ("int", "long"):
        return lambda out, v: _enc_long(out, int(v))
    if s == "float":
        return lambda out, v: out.extend(_pack_f(float(v)))
    if s == "double":
        return lambda out, v: out.extend(_pack_d(float(v)))
    if s == "string":
        return _enc_string
    if s == "bytes":
        def enc_bytes(out, v):
            _enc_long(out, len(v))
            out.extend(v)
        return enc_bytes
    raise AvroError(f"unsupported type {s!r}")


def compile_decoder(schema):
    """Returns f(buf, pos) -> (value, pos) equivalent to ``_decode(buf, pos, schema)``."""
    s = schema
    if isinstance(s, list):
        decs = [compile_decoder(b) for b in s]
        nb = len(decs)
        if nb == 2 and s[0] == "null":
            inner = decs[1]

            def dec_nullable(buf, pos):
                if pos >= len(buf):
                    raise AvroError("truncated union")
                b = buf[pos]
                if b == 0:
                    return None, pos + 1
                if b == 2:
                    return inner(buf, pos + 1)
                raise AvroError(f"union branch {b >> 1} out of range")
            return dec_nullable

        def dec_union(buf, pos):
            i, pos = read_long(buf, pos)
            if not 0 <= i < nb:
                raise AvroError(f"union branch {i} out of range")
            return decs[i](buf, pos)
        return dec_union
    if isinstance(s, dict):
        t = s["type"]
        if t == "record":
            fields = [(f["name"], compile_decoder(f["type"])) for f in s["fields"]]

            def dec_record(buf, pos):
                rec = {}
                for name, dec in fields:
                    rec[name], pos = dec(buf, pos)
                return rec, pos
            return dec_record
        if t == "array":
            if _is_nullable_float_array(s):
                nullable_items = isinstance(s["items"], list)
                return lambda buf, pos: decode_float_array_fast(buf, pos, nullable_items)
            item = compile_decoder(s["items"])

            def dec_array(buf, pos):
                items = []
                while True:
                    n, pos = read_long(buf, pos)
                    if n == 0:
                        return items, pos
                    if n < 0:
                        n = -n
                        _, pos = read_long(buf, pos)
                    for _ in range(n):
                        x, pos = item(buf, pos)
                        items.append(x)
            return dec_array
        return compile_decoder(t)
    if s == "null":
        return lambda buf, pos: (None, pos)
    if s in ("int", "long"):
        return read_long
    if s in ("string", "bytes"):
        is_str = s == "string"

        def dec_str(buf, pos):
            n, pos = read_long(buf, pos)
            end = pos + n
            if n < 0 or end > len(buf):
                raise AvroError("truncated string/bytes")
            raw = bytes(buf[pos:end])
            return (raw.decode("utf-8") if is_str else raw), end
        return dec_str
    if s == "double":
        def dec_double(buf, pos):
            if pos + 8 > len(buf):
                raise AvroError("truncated double")
            return _unpack_d(buf, pos)[0], pos + 8
        return dec_double
    if s == "float":
        def dec_float(buf, pos):
            if pos + 4 > len(buf):
                raise AvroError("truncated float")
            return _unpack_f(buf, pos)[0], pos + 4
        return dec_float
    if s == "boolean":
        def dec_bool(buf, pos):
            if pos >= len(buf):
                raise AvroError("truncated boolean")
            return buf[pos] != 0, pos + 1
        return dec_bool
    raise AvroError(f"unsupported type {s!r}")


class CompiledSchema:
    """encode(value) -> bytes / decode(buf) -> value with the closure trees built once."""

    def __init__(self, schema):
        self.schema = parse_schema(schema)
        self._enc = compile_encoder(self.schema)
        self._dec = compile_decoder(self.schema)

    def encode(self, value, prefix: bytes = b"") -> bytes:
        out = bytearray(prefix)
        self._enc(out, value)
        return bytes(out)

    def decode(self, buf, pos: int = 0, require_all: bool = True):
        v, end = self._dec(buf, pos)
        if require_all and end != len(buf):
            raise AvroError(f"{len(buf) - end} trailing bytes after Avro datum")
        return v
