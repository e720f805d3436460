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
    src = open(os.path.join(ROOT, "include", "sa_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", src)))


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
