"""Microbenchmark of the native batch decoder (sa_wire_decode_queries_embed) on this host: ms per batch of 1024 x 1536-d
records, for the SA_WIRE_THREADS of the environment.  No GPU involved.

    for t in 1 2 4 8; do SA_WIRE_THREADS=$t python tools/host_decode_bench.py; done
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from qsa_b200 import capi  # noqa: E402

lib = capi.load()
B, dim = 1024, 1536
vec = np.random.default_rng(0).standard_normal((B, dim)).astype(np.float32)
texts = [f"question {i}".encode() for i in range(B)]
tbuf = b"".join(texts)
tlen = np.array([len(t) for t in texts], np.uint32)
toff = np.concatenate([[0], np.cumsum(tlen[:-1], dtype=np.uint64)]).astype(np.uint64)
rec_off = np.empty(B + 1, np.uint64)
need = C.c_uint64()
lib.sa_wire_encode_queries_embed(B, dim, 100001, tbuf, toff.ctypes.data, tlen.ctypes.data, vec.ctypes.data, 0, None, 0,
                                 rec_off.ctypes.data, C.byref(need))
out = np.empty(int(need.value), np.uint8)
assert lib.sa_wire_encode_queries_embed(B, dim, 100001, tbuf, toff.ctypes.data, tlen.ctypes.data, vec.ctypes.data, 1,
                                        out.ctypes.data, out.size, rec_off.ctypes.data, C.byref(need)) == 0
voff = np.empty(B, np.uint64)
vlen = np.empty(B, np.uint32)
assert lib.sa_wire_split_log(out.ctypes.data, out.size, B, voff.ctypes.data, vlen.ctypes.data, None, None, None) == 0
dst = np.empty((B, dim), np.float32)
to, tl, st, nok = np.empty(B, np.uint64), np.empty(B, np.uint32), np.empty(B, np.uint8), C.c_int()
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(20):
        lib.sa_wire_decode_queries_embed(out.ctypes.data, voff.ctypes.data, vlen.ctypes.data, B, dim, 100001, dst.ctypes.data,
                                         to.ctypes.data, tl.ctypes.data, st.ctypes.data, C.byref(nok))
    best = min(best, (time.perf_counter() - t0) / 20 * 1e3)
assert nok.value == B and (dst == vec).all()
a = np.empty(out.size, np.uint8)
t0 = time.perf_counter()
for _ in range(20):
    a[:] = out
cp = out.size * 20 / (time.perf_counter() - t0) / 1e9
print(f"SA_WIRE_THREADS={os.environ.get('SA_WIRE_THREADS', '(default 4)')}  decode {best:.3f} ms per {B} x {dim} batch "
      f"({out.size / best / 1e6:.1f} GB/s of records; plain memcpy on this host {cp:.1f} GB/s)")
