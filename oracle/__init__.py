"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see bruteforce.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this package; it is not installed with the product (pyproject.toml)."""
