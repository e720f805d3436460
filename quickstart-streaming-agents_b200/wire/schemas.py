"""Topic value schemas of the Lab2 pipeline.

``QUERIES_VALUE`` and ``DOCUMENTS_VALUE`` are the contract the reference spells out and must stay byte-for-byte
compatible (scripts/lab2_publish_queries.py:59-64, scripts/publish_docs.py:63-109).  The other three are what
Flink derives for the tables the Terraform creates -- nullable unions, namespace
``org.apache.flink.avro.generated.record``, record name ``<topic>_value``:

  queries_embed            (query STRING, embedding ARRAY<FLOAT>)            terraform/lab2-vector-search/main.tf:141
  search_results           query + document_id_i, chunk_i, score_i (i=1..3)   main.tf:292
  search_results_response  the ten columns above + response STRING           main.tf:331

The reference never writes those three down, so they are inferred from that convention.
``RIDE_REQUESTS_VALUE`` (scripts/publish_lab3_data.py:68-86) is here only to decode the captured Lab3
records that serve as the codec's known-answer fixture.
"""
NAMESPACE = "org.apache.flink.avro.generated.record"


def _nullable(t):
    return ["null", t]


def _field(name, t):
    return {"name": name, "type": _nullable(t), "default": None}


QUERIES_VALUE = {
    "type": "record",
    "name": "queries_value",
    "namespace": NAMESPACE,
    "fields": [{"name": "query", "type": ["null", "string"], "default": None}],
}

_STRING_ARRAY = {"type": "array", "items": ["null", "string"]}

DOCUMENTS_VALUE = {
    "type": "record",
    "name": "documents_value",
    "namespace": NAMESPACE,
    "fields": [
        _field("document_id", "string"),
        _field("document_text", "string"),
        _field("pages", "string"),
        _field("section_reference", "string"),
        _field("title", "string"),
        _field("fraud_categories", _STRING_ARRAY),
        _field("policy_keywords", _STRING_ARRAY),
        _field("char_count", "int"),
    ],
}

_FLOAT_ARRAY = {"type": "array", "items": ["null", "float"]}

QUERIES_EMBED_VALUE = {
    "type": "record",
    "name": "queries_embed_value",
    "namespace": NAMESPACE,
    "fields": [_field("query", "string"), _field("embedding", _FLOAT_ARRAY)],
}

# Columns of the vector tables: Lab2/Lab3 use (document_id, chunk, embedding) (main.tf:215;
# terraform/lab3-agentic-fleet-management/main.tf:110-124); Lab4's fema_policies_vectordb adds the metadata columns
# of the documents topic (terraform/lab4-pubsec-fraud-agents/main.tf:271-289).  One schema carries them all.
METADATA_COLUMNS = ("pages", "section_reference", "title", "fraud_categories", "policy_keywords", "char_count")

DOCUMENTS_EMBED_VALUE = {
    "type": "record",
    "name": "documents_embed_value",
    "namespace": NAMESPACE,
    "fields": [_field("document_id", "string"), _field("chunk", "string"), _field("embedding", _FLOAT_ARRAY),
               _field("pages", "string"), _field("section_reference", "string"), _field("title", "string"),
               _field("fraud_categories", _STRING_ARRAY), _field("policy_keywords", _STRING_ARRAY),
               _field("char_count", "int")],
}

RESULTS_PER_QUERY = 3  # the reference flattens search_results[1..3] (main.tf:292)

SEARCH_RESULTS_VALUE = {
    "type": "record",
    "name": "search_results_value",
    "namespace": NAMESPACE,
    "fields": [_field("query", "string")] + [
        f for i in range(1, RESULTS_PER_QUERY + 1)
        for f in (_field(f"document_id_{i}", "string"), _field(f"chunk_{i}", "string"), _field(f"score_{i}", "double"))
    ],
}

SEARCH_RESULTS_RESPONSE_VALUE = {
    "type": "record",
    "name": "search_results_response_value",
    "namespace": NAMESPACE,
    "fields": SEARCH_RESULTS_VALUE["fields"] + [_field("response", "string")],
}

RIDE_REQUESTS_VALUE = {
    "type": "record",
    "name": "ride_requests_value",
    "namespace": NAMESPACE,
    "fields": [
        {"name": "request_id", "type": "string"},
        {"name": "customer_email", "type": "string"},
        {"name": "pickup_zone", "type": "string"},
        {"name": "drop_off_zone", "type": "string"},
        {"name": "price", "type": "double"},
        {"name": "number_of_passengers", "type": "int"},
        {"name": "request_ts", "type": {"type": "long", "logicalType": "timestamp-millis"}},
    ],
}
RIDE_REQUESTS_KEY = "string"

TOPIC_SCHEMAS = {
    "queries": QUERIES_VALUE,
    "documents": DOCUMENTS_VALUE,
    "queries_embed": QUERIES_EMBED_VALUE,
    "documents_embed": DOCUMENTS_EMBED_VALUE,
    "search_results": SEARCH_RESULTS_VALUE,
    "search_results_response": SEARCH_RESULTS_RESPONSE_VALUE,
}
