#!/usr/bin/env python3
"""publish_docs -- turn a directory of Markdown chunks (YAML front matter) into records on the ``documents`` topic.

Drop-in for the reference's scripts/publish_docs.py: same options (``--lab2 | --lab3``, ``--topic``,
``--docs-dir``, ``--dry-run``, ``--verbose``, ``--workers``), same document parsing rules, same Avro value
schema and key (document_id), same progress lines, summary block and exit code (0 iff nothing failed).
The reference starts one ``confluent kafka topic produce ... --parse-key --delimiter :`` process per document
against Confluent Cloud; here each worker appends the Confluent-framed Avro record to the local topic log,
from which the serve loop (``scripts/sa_serve.py``) embeds and indexes it.

Usage:
    python -m scripts.publish_docs --lab2
    python -m scripts.publish_docs --docs-dir my_chunks --topic documents --log-dir /data/topics
    python -m scripts.publish_docs --lab2 --dry-run
"""
from __future__ import annotations

import argparse
import json
import logging
import sys
import threading
from concurrent.futures import ThreadPoolExecutor, as_completed
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import yaml

try:
    from ._local import ROOT, AvroJsonProducer, resolve_log_dir, setup_logging
except ImportError:  # executed as a file
    from _local import ROOT, AvroJsonProducer, resolve_log_dir, setup_logging

_NS = "org.apache.flink.avro.generated.record"


def _opt(name: str, avro_type) -> dict:
    return {"name": name, "type": ["null", avro_type], "default": None}


class FlinkDocsPublisherCLI:
    """Markdown -> ``documents`` records."""

    # value schema of topic `documents` (field order and nullability are part of the wire contract)
    DOCUMENT_VALUE_SCHEMA = {
        "type": "record",
        "name": "documents_value",
        "namespace": _NS,
        "fields": [
            _opt("document_id", "string"),
            _opt("document_text", "string"),
            _opt("pages", "string"),
            _opt("section_reference", "string"),
            _opt("title", "string"),
            _opt("fraud_categories", {"type": "array", "items": ["null", "string"]}),
            _opt("policy_keywords", {"type": "array", "items": ["null", "string"]}),
            _opt("char_count", "int"),
        ],
    }

    def __init__(self, log_dir: str, dry_run: bool = False, max_workers: int = 10):
        self.log_dir = log_dir
        self.dry_run = dry_run
        self.max_workers = max_workers
        self.logger = logging.getLogger(__name__)
        self._lock = threading.Lock()
        self._producers: dict[str, AvroJsonProducer] = {}

    # ------------------------------------------------------------------ parsing
    def parse_markdown_file(self, file_path: Path) -> Optional[Dict[str, Any]]:
        """``---`` front matter (YAML) + body.  document_id defaults to the file name; document_text is
        ``# <title>`` + blank line + body when a title is present (reference rules, publish_docs.py:172-223)."""
        try:
            text = file_path.read_text(encoding="utf-8")
            meta: Dict[str, Any] = {}
            body = text
            if text.startswith("---"):
                pieces = text.split("---", 2)
                if len(pieces) >= 3:
                    meta = yaml.safe_load(pieces[1]) or {}
                    body = pieces[2].strip()
            title = meta.get("title", "")
            return {
                "document_id": meta.get("document_id", file_path.name),
                "document_text": f"# {title}\n\n{body}" if title else body,
                "pages": meta.get("pages"),
                "section_reference": meta.get("section_reference"),
                "title": title,
                "fraud_categories": meta.get("fraud_categories", []),
                "policy_keywords": meta.get("policy_keywords", []),
                "char_count": meta.get("char_count"),
                "metadata": meta,
            }
        except Exception as e:
            self.logger.error(f"Failed to parse file {file_path}: {e}")
            return None

    # ------------------------------------------------------------------ publishing
    @staticmethod
    def avro_json_value(document: Dict[str, Any]) -> dict:
        """The Avro-JSON object the `confluent` CLI would receive: unions wrapped, empty arrays -> null."""
        def s(v):
            return None if v is None else {"string": str(v)}

        def arr(items):
            if not items:
                return None
            return {"array": [None if x is None else {"string": str(x)} for x in items]}

        cc = document.get("char_count")
        return {
            "document_id": {"string": str(document["document_id"])},
            "document_text": {"string": str(document["document_text"])},
            "pages": s(document.get("pages")),
            "section_reference": s(document.get("section_reference")),
            "title": s(document.get("title")),
            "fraud_categories": arr(document.get("fraud_categories")),
            "policy_keywords": arr(document.get("policy_keywords")),
            "char_count": None if cc is None else {"int": int(cc)},
        }

    def _producer(self, topic: str) -> AvroJsonProducer:
        with self._lock:
            p = self._producers.get(topic)
            if p is None:
                p = self._producers[topic] = AvroJsonProducer(self.log_dir, topic, self.DOCUMENT_VALUE_SCHEMA)
            return p

    def publish_document(self, document: Dict[str, Any], topic: str) -> bool:
        try:
            value = self.avro_json_value(document)
            if self.dry_run:
                self.logger.info(f"[DRY RUN] Would publish document: {document['document_id']}")
                self.logger.debug(f"[DRY RUN] Content length: {len(document['document_text'])} chars")
                return True
            # same "key:json" line the reference pipes to the CLI with --parse-key --delimiter :
            line = f"{document['document_id']}:{json.dumps(value)}\n"
            prod = self._producer(topic)
            with self._lock:
                prod.produce_line(line, parse_key=True, delimiter=":")
            self.logger.info(f"Published document: {document['document_id']}")
            return True
        except Exception as e:
            self.logger.error(f"Failed to publish document {document.get('document_id', 'unknown')}: {e}")
            return False

    def _process_single_file(self, file_path: Path, topic: str) -> Tuple[str, bool, Optional[str]]:
        try:
            document = self.parse_markdown_file(file_path)
            if not document:
                return (file_path.name, False, "Failed to parse markdown file")
            if self.publish_document(document, topic):
                return (file_path.name, True, None)
            return (file_path.name, False, "Failed to publish document")
        except Exception as e:
            self.logger.error(f"Error processing {file_path.name}: Unexpected error: {e}")
            return (file_path.name, False, f"Unexpected error: {e}")

    def publish_directory(self, docs_dir: Path, topic: str) -> Dict[str, int]:
        md_files = sorted(docs_dir.glob("*.md"))
        results = {"success": 0, "failed": 0, "total": len(md_files)}
        self.logger.info(f"Found {len(md_files)} markdown files to process")
        self.logger.info(f"Publishing with {self.max_workers} parallel workers")
        done = 0
        with ThreadPoolExecutor(max_workers=self.max_workers) as pool:
            futures = [pool.submit(self._process_single_file, p, topic) for p in md_files]
            for fut in as_completed(futures):
                _, ok, _ = fut.result()
                results["success" if ok else "failed"] += 1
                done += 1
                if done % 10 == 0 or done == results["total"]:
                    self.logger.info(f"Progress: {done}/{results['total']} documents "
                                     f"({results['success']} succeeded, {results['failed']} failed)")
        return results

    def close(self):
        self._producers.clear()


def find_docs_directory(project_root: Path, lab: int, cloud_provider: str | None = None) -> Optional[Path]:
    """Same search order as the reference (publish_docs.py:446-496)."""
    if lab == 2:
        candidates = [project_root / "assets" / "lab2" / "flink_docs",
                      project_root / "assets" / "lab2" / "flink_docs" / "markdown_chunks",
                      project_root / (cloud_provider or "aws") / "lab2-vector-search" / "flink_docs",
                      project_root / "flink_docs"]
    elif lab == 3:
        candidates = [project_root / "assets" / "lab3" / "nola_events_docs",
                      project_root / "assets" / "lab3" / "markdown_chunks"]
    else:
        candidates = []
    for c in candidates:
        if c.exists():
            return c
    return None


def main(argv=None) -> int:
    parser = argparse.ArgumentParser(
        description="Publish documentation to the local `documents` topic (drop-in for the Confluent CLI based publisher)",
        formatter_class=argparse.RawDescriptionHelpFormatter,
        epilog="""
Examples:
  %(prog)s --lab2
  %(prog)s --lab3
  %(prog)s --lab2 --dry-run
  %(prog)s --docs-dir temp_pdf_extraction/output_chunks --topic documents
        """,
    )
    lab_group = parser.add_mutually_exclusive_group(required=False)
    lab_group.add_argument("--lab2", action="store_true", help="Publish Lab2 Flink SQL documentation")
    lab_group.add_argument("--lab3", action="store_true", help="Publish Lab3 New Orleans event documentation")
    parser.add_argument("--topic", default="documents", help="Topic name (default: documents)")
    parser.add_argument("--docs-dir", type=Path, help="Directory containing markdown files (auto-detected if not specified)")
    parser.add_argument("--dry-run", action="store_true", help="Test without actually publishing")
    parser.add_argument("--verbose", action="store_true", help="Enable verbose logging")
    parser.add_argument("--workers", type=int, default=10, help="Number of parallel workers for publishing (default: 10)")
    parser.add_argument("--log-dir", default=None, help="Topic log directory (default: $SA_LOG_DIR or ./.sa_topics)")
    parser.add_argument("--project-root", type=Path, default=None, help="Where assets/ lives (default: this repository)")
    args = parser.parse_args(argv)

    logger = setup_logging(args.verbose)
    if args.docs_dir:
        docs_dir = args.docs_dir
        logger.info(f"Publishing documents from {docs_dir}")
    else:
        if not (args.lab2 or args.lab3):
            logger.error("Either --lab2, --lab3, or --docs-dir must be specified")
            return 1
        lab = 2 if args.lab2 else 3
        lab_name = "Lab2 (Flink SQL documentation)" if lab == 2 else "Lab3 (New Orleans event documentation)"
        logger.info(f"Publishing documents for {lab_name}")
        docs_dir = find_docs_directory(args.project_root or Path(ROOT), lab)
        if not docs_dir:
            logger.error(f"Could not find documentation directory for Lab{lab}. Please specify --docs-dir")
            return 1
        logger.info(f"Found documentation directory: {docs_dir}")
    if not docs_dir.exists():
        logger.error(f"Documentation directory does not exist: {docs_dir}")
        return 1

    publisher = FlinkDocsPublisherCLI(resolve_log_dir(args.log_dir), dry_run=args.dry_run, max_workers=args.workers)
    try:
        logger.info(f"Publishing documents from {docs_dir} to topic '{args.topic}'")
        if args.dry_run:
            logger.info("[DRY RUN MODE - No actual publishing will occur]")
        results = publisher.publish_directory(docs_dir, args.topic)
        print(f"\n{'=' * 60}")
        print("PUBLISHING SUMMARY")
        print(f"{'=' * 60}")
        print(f"Total files:      {results['total']}")
        print(f"Published:        {results['success']}")
        print(f"Failed:           {results['failed']}")
        print(f"{'=' * 60}")
        if args.dry_run:
            print("\n[DRY RUN COMPLETE - No messages were actually published]")
        return 0 if results["failed"] == 0 else 1
    finally:
        publisher.close()


if __name__ == "__main__":
    sys.exit(main())
