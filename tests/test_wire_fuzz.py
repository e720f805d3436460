"""Property tests (hypothesis) for the Avro codec: generic and compiled codecs agree on random records of the topic
schemas, round-trip exactly, and never accept a truncated buffer."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from qsa_b200.wire import avro, schemas

text = st.text(max_size=40)
opt_text = st.one_of(st.none(), text)
f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)


@st.composite
def documents(draw):
    arr = st.one_of(st.none(), st.lists(opt_text, max_size=5))
    return {"document_id": draw(opt_text), "document_text": draw(opt_text), "pages": draw(opt_text),
            "section_reference": draw(opt_text), "title": draw(opt_text), "fraud_categories": draw(arr),
            "policy_keywords": draw(arr), "char_count": draw(st.one_of(st.none(), st.integers(-2**31, 2**31 - 1)))}


@st.composite
def search_results(draw):
    rec = {"query": draw(opt_text)}
    for i in (1, 2, 3):
        rec[f"document_id_{i}"] = draw(opt_text)
        rec[f"chunk_{i}"] = draw(opt_text)
        rec[f"score_{i}"] = draw(st.one_of(st.none(), st.floats(allow_nan=False, allow_infinity=False)))
    return rec


@st.composite
def queries_embed(draw):
    n = draw(st.integers(0, 40))
    vec = draw(st.one_of(st.none(), st.lists(f32, min_size=n, max_size=n)))
    return {"query": draw(opt_text), "embedding": None if vec is None else np.asarray(vec, dtype=np.float32)}


def _same(a, b):
    if isinstance(b, np.ndarray) or isinstance(a, np.ndarray):
        return np.array_equal(np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32))
    if isinstance(b, dict):
        return set(a) == set(b) and all(_same(a[k], b[k]) for k in b)
    if isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    return a == b


def _check(schema, rec):
    cs = avro.CompiledSchema(schema)
    ref = avro.encode(schema, rec)
    assert cs.encode(rec) == ref
    assert _same(avro.decode(schema, ref), rec) and _same(cs.decode(ref), rec)
    for cut in {1, len(ref) // 2, len(ref) - 1} - {0, len(ref)}:
        for dec in (lambda b: avro.decode(schema, b), cs.decode):
            try:
                dec(ref[:cut])
            except avro.AvroError:
                continue
            raise AssertionError(f"truncated buffer of {cut}/{len(ref)} bytes was accepted")
    framed = avro.frame(100001, ref)
    assert avro.unframe(framed) == (100001, ref)


@settings(max_examples=150, deadline=None)
@given(documents())
def test_documents_roundtrip(rec):
    _check(schemas.DOCUMENTS_VALUE, rec)
    j = avro.to_avro_json(schemas.DOCUMENTS_VALUE, rec)
    assert avro.from_avro_json(schemas.DOCUMENTS_VALUE, j) == rec


@settings(max_examples=150, deadline=None)
@given(search_results())
def test_search_results_roundtrip(rec):
    _check(schemas.SEARCH_RESULTS_VALUE, rec)


@settings(max_examples=100, deadline=None)
@given(queries_embed())
def test_queries_embed_roundtrip(rec):
    _check(schemas.QUERIES_EMBED_VALUE, rec)


@settings(max_examples=300, deadline=None)
@given(st.integers(-2**63, 2**63 - 1))
def test_varint_roundtrip(n):
    out = bytearray()
    avro.write_long(out, n)
    assert avro.read_long(bytes(out), 0) == (n, len(out)) and len(out) <= 10
