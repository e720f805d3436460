#!/usr/bin/env python3
"""publish_queries -- put one user question on the ``queries`` topic.

Drop-in for the reference's scripts/lab2_publish_queries.py: same positional arguments
(``[aws|azure] [QUERY]``), same options (``--topic``, ``--verbose``), same interactive prompt, same Avro value
schema, same messages and exit codes.  The only difference is where the record goes: the reference forks
``confluent kafka topic produce queries --value-format avro`` against Confluent Cloud; this tool appends the
Confluent-framed Avro record to the local topic log that the serve loop (``scripts/sa_serve.py``) consumes.

Usage:
    python -m scripts.lab2_publish_queries "How do I use window functions?"
    python -m scripts.lab2_publish_queries aws "What is watermarking?"
    python -m scripts.lab2_publish_queries --log-dir /data/topics        # interactive
"""
from __future__ import annotations

import argparse
import logging
import sys

try:
    from ._local import AvroJsonProducer, resolve_log_dir, setup_logging as _base_setup_logging
except ImportError:  # executed as a file
    from _local import AvroJsonProducer, resolve_log_dir, setup_logging as _base_setup_logging


def setup_logging(verbose: bool = False) -> logging.Logger:
    return _base_setup_logging(verbose, default_level="ERROR")


class QueryPublisherCLI:
    """Publishes queries; the record layout is the one Flink's ``queries`` table expects."""

    # value schema of topic `queries` -- table DDL: queries (query STRING NOT NULL), main.tf:108
    QUERY_VALUE_SCHEMA = {
        "type": "record",
        "name": "queries_value",
        "namespace": "org.apache.flink.avro.generated.record",
        "fields": [{"name": "query", "type": ["null", "string"], "default": None}],
    }

    def __init__(self, log_dir: str):
        self.log_dir = log_dir
        self.logger = logging.getLogger(__name__)
        self._producers: dict[str, AvroJsonProducer] = {}

    def publish_query(self, query: str, topic: str = "queries") -> bool:
        """True when the record is durably on the topic, False otherwise (the error is printed)."""
        try:
            prod = self._producers.get(topic)
            if prod is None:
                prod = self._producers[topic] = AvroJsonProducer(self.log_dir, topic, self.QUERY_VALUE_SCHEMA)
            self.logger.debug(f"Publishing query to topic '{topic}': {query[:100]}...")
            prod.produce_avro_json({"query": {"string": query}})  # union-wrapped, as the CLI's stdin format
            return True
        except Exception as e:
            print(f"❌ Failed to publish query: {e}")
            return False

    def close(self):
        self._producers.clear()


def main(argv=None) -> int:
    parser = argparse.ArgumentParser(
        description="Publish queries to the local `queries` topic (drop-in for the Confluent CLI based publisher)",
        formatter_class=argparse.RawDescriptionHelpFormatter,
        epilog="""
Examples:
  %(prog)s "How do I use window functions?"
  %(prog)s aws "What is watermarking?"
  %(prog)s azure --verbose
        """,
    )
    # The reference declares choices=["aws", "azure"] here, which makes argparse reject its own documented form
    # `publish_queries "How do window functions work?"` (LAB2-Walkthrough.md:67).  Accept both forms instead.
    parser.add_argument("cloud_provider", nargs="?", metavar="{aws,azure}",
                        help="Accepted for command-line compatibility; the local engine has no cloud to choose.")
    parser.add_argument("query", nargs="?", help="Query to publish. If not provided, interactive mode will be used.")
    parser.add_argument("--topic", default="queries", help="Topic name (default: queries)")
    parser.add_argument("--verbose", action="store_true", help="Enable verbose logging")
    parser.add_argument("--log-dir", default=None, help="Topic log directory (default: $SA_LOG_DIR or ./.sa_topics)")
    args = parser.parse_args(argv)

    logger = setup_logging(args.verbose)
    if args.cloud_provider not in (None, "aws", "azure"):
        if args.query is not None:
            parser.error(f"argument cloud_provider: invalid choice: {args.cloud_provider!r} (choose from 'aws', 'azure')")
        args.query, args.cloud_provider = args.cloud_provider, None   # a lone positional is the query
    if args.cloud_provider:
        logger.debug(f"Ignoring cloud provider argument: {args.cloud_provider}")

    query = args.query
    if not query:
        try:
            query = input("\nEnter your query: ").strip()
            if not query:
                print("❌ No query provided")
                return 1
        except (EOFError, KeyboardInterrupt):
            print("\n❌ Query input cancelled")
            return 1

    publisher = QueryPublisherCLI(resolve_log_dir(args.log_dir))
    try:
        if not publisher.publish_query(query, args.topic):
            return 1
        print("\n✓ Query published successfully!")
        print(f"  Query: {query[:100]}{'...' if len(query) > 100 else ''}")
        print(f"\n  Topic log:  {publisher.log_dir}  (results appear on `search_results` and `search_results_response`)")
        return 0
    finally:
        publisher.close()


if __name__ == "__main__":
    sys.exit(main())
