"""Avro binary / Avro-JSON codec, Confluent wire framing, topic schemas, schema-registry stub."""
