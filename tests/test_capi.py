"""CPU tests of the boundary: libsa_b200.so builds for sm_100a without a GPU, loads, exports every symbol
include/sa_api.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from qsa_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    names = set()
    for h in ("sa_api.h", "sa_wire.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_and_binding_agree():
    assert header_functions() == sorted(capi.EXPORTS)


def test_library_exports_every_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.sa_version() >= 100
    assert lib.sa_strerror(capi.SA_ERR_DEVICE).decode().startswith("unsupported device")


def test_library_contains_only_sm100a_native_code(lib):
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):   # tcgen05.mma, TMA, tcgen05.ld, tcgen05.commit
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")          # no legacy mma.sync path


def test_argument_validation_without_touching_the_gpu(lib):
    h = C.c_void_p()
    assert lib.sa_engine_create(C.byref(h), 0, 100, 1000, 128, 10) == capi.SA_ERR_ARG       # dim % 64
    assert b"multiple of 64" in lib.sa_last_error()
    assert lib.sa_engine_create(C.byref(h), 0, 128, 0, 128, 10) == capi.SA_ERR_ARG          # capacity
    assert lib.sa_engine_create(C.byref(h), 0, 128, 1000, 128, 99) == capi.SA_ERR_ARG       # max_k
    assert lib.sa_corpus_rows(None) == -1
    assert lib.sa_search(None, None, 1, 1, None, None, None, None) == capi.SA_ERR_ARG


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    h = C.c_void_p()
    rc = lib.sa_engine_create(C.byref(h), 0, 1536, 1000, 128, 10)
    assert rc in (capi.SA_ERR_CUDA, capi.SA_ERR_ARG) and not h.value
    from qsa_b200.engine import VectorIndex
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VectorIndex(dim=1536, capacity=1000)


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "libsa_b200.so"))
    with pytest.raises(capi.SaLibraryMissing, match="no CPU fallback"):
        capi.load()


def plan(lib, num_sms, nq, cg, num_tiles, cap=0):
    out = (C.c_int * (4 * 16))()
    n = C.c_int()
    rc = lib.sa_debug_plan(num_sms, nq, cg, num_tiles, cap, out, 16, C.byref(n))
    assert rc == 0, lib.sa_last_error()
    return [tuple(out[4 * i + j] for j in range(4)) for i in range(n.value)]


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("nq", [1, 37, 128, 129, 256, 700, 1024, 1100, 4096, 5000])
@pytest.mark.parametrize("num_tiles", [0, 1, 5, 79, 39063])
@pytest.mark.parametrize("sms", [148, 160])
def test_launch_planner_invariants(lib, cg, nq, num_tiles, sms):
    """Host logic of the scan: every query is covered exactly once, a launch never needs more CTAs than SMs,
    tile lanes never outnumber tiles nor the merge kernel's 148-lane table, and the machine is used when the batch
    allows it."""
    launches = plan(lib, sms, nq, cg, num_tiles)
    rows = 128 * cg
    nxt = 0
    for q0, n, nqb, tl in launches:
        assert q0 == nxt and n > 0 and q0 % rows == 0
        assert nqb == (n + rows - 1) // rows and tl >= 1
        assert nqb * tl * cg <= sms and tl <= max(num_tiles, 1) and tl <= 148
        nxt += n
    assert nxt == nq
    if num_tiles >= 148 and nq >= rows:
        used = max(nqb * tl * cg for _, _, nqb, tl in launches)
        assert used >= 0.85 * 148, launches


def test_launch_planner_headline_shapes(lib):
    assert plan(lib, 148, 1024, 2, 39063) == [(0, 1024, 4, 18)]             # 4 pair blocks x 18 lanes = 72 pairs
    assert plan(lib, 148, 128, 1, 39063) == [(0, 128, 1, 148)]              # HBM-bound: every SM its own lane
    assert plan(lib, 148, 512, 2, 39063) == [(0, 512, 2, 37)]               # already fills the machine
    p = plan(lib, 148, 4096, 2, 4883)                                       # config 4 per GPU: 2 launches of 8 x 9
    assert [x[2:] for x in p] == [(8, 9), (8, 9)]
    assert len(plan(lib, 148, 1100, 2, 118, cap=2)) == 3                    # forced small launches
