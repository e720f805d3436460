"""Copies the first records of the reference's captured Lab3 stream into tests/golden/ as the Avro codec's
known-answer fixture (the GPU box has no /root/reference).  These are DATA records captured from Kafka
(base64 Confluent-framed Avro, assets/lab3/data/ride_requests.jsonl), not source code.

    python tests/golden/make_wire_fixture.py            # needs /root/reference
"""
import itertools
import os

SRC = "/root/reference/assets/lab3/data/ride_requests.jsonl"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    with open(SRC) as f, open(os.path.join(HERE, "ride_requests_head.jsonl"), "w") as out:
        lines = list(itertools.islice(f, 20000))
        picked = lines[:120] + lines[5000:5040] + lines[19960:20000]
        out.writelines(picked)
    print("wrote", len(picked), "records")
