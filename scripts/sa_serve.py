#!/usr/bin/env python3
"""sa_serve -- run the Lab2 topic graph locally on one B200.

Consumes ``documents`` / ``queries`` (and pre-embedded ``documents_embed`` / ``queries_embed``) from the topic log,
keeps the vector table in HBM, writes ``search_results`` and ``search_results_response``
(replaces the Flink statements of terraform/lab2-vector-search/main.tf:233-331).  Needs a CUDA device.

    python -m scripts.sa_serve --log-dir .sa_topics --once          # drain what is there, then exit
    python -m scripts.sa_serve --capacity 2000000 --k 3             # run until interrupted
"""
from __future__ import annotations

import argparse
import json
import sys

try:
    from ._local import resolve_log_dir, setup_logging
except ImportError:
    from _local import resolve_log_dir, setup_logging


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--capacity", type=int, default=1 << 20, help="rows of HBM to reserve for the vector table")
    ap.add_argument("--max-batch", type=int, default=1024)
    ap.add_argument("--k", type=int, default=3, help="VECTOR_SEARCH_AGG k (the reference uses 3)")
    ap.add_argument("--once", action="store_true", help="process everything pending, print stats, exit")
    ap.add_argument("--snapshot-dir", default=None,
                    help="checkpoint directory: loaded at start if present, written on exit (with the consumer-group "
                         "offsets in the log directory this makes the loop resumable without replaying `documents`)")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args(argv)
    setup_logging(a.verbose)

    from qsa_b200.engine import VectorIndex  # CUDA only; raises without a device (no CPU fallback)
    from qsa_b200.operator import VectorTable
    from qsa_b200.pipeline.serve import Lab2Pipeline

    index = VectorIndex(dim=a.dim, capacity=a.capacity, max_batch=a.max_batch, max_k=max(a.k, 3))
    import os
    table = VectorTable(index)
    if a.snapshot_dir and os.path.exists(os.path.join(a.snapshot_dir, "columns.jsonl")):
        print(f"resumed {table.load(a.snapshot_dir)} rows from {a.snapshot_dir}", file=sys.stderr)
    pipe = Lab2Pipeline(resolve_log_dir(a.log_dir), table, k=a.k, max_batch=a.max_batch)
    try:
        if a.once:
            pipe.run_until_idle()
            print(json.dumps(pipe.stats))
        else:
            pipe.run_forever()
    except KeyboardInterrupt:
        print(json.dumps(pipe.stats))
    finally:
        if a.snapshot_dir:
            table.save(a.snapshot_dir)
    return 0


if __name__ == "__main__":
    sys.exit(main())
