"""The drop-in boundary: `csrc/libcast_b200.so` must load without a GPU and export every function that
include/cast_b200.h declares; the ctypes table in `_lib.py` must cover exactly those; and the product package must
never import the oracle (test infrastructure) or fall back to another backend."""
import ctypes
import os
import re

import pytest

from cast_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cast_b200.h")
PKG = os.path.join(ROOT, "context-aware-sequential-recommendation_b200")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cast_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    names = declared_functions()
    assert len(names) >= 30 and "cast_attn_fwd" in names and "cast_scatter_rows" in names


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.fail("csrc/libcast_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(_lib.LIB_PATH)  # loads without a GPU (no CUDA call at load time)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.cast_version() >= 1


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_missing_library_is_a_loud_error(tmp_path):
    with pytest.raises(_lib.CastError):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_product_never_touches_the_oracle_or_other_backends():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "import triton" in txt \
                        or "torch.compile" in txt:
                    bad.append(f)
    assert not bad, bad


def test_no_cuda_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    import cast_b200
    from helpers import make_args
    with pytest.raises(cast_b200.CastError):
        cast_b200.SASRec(10, 20, make_args())
