"""CPU: the C-ABI library loads, exports every symbol include/revers_o_b200.h declares, answers its
size queries, and fails LOUDLY (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from revers_o_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "revers_o_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rvo_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "python prototypes out of sync with the header"


def test_version_and_size_queries():
    lib = _lib.load()
    assert lib.rvo_version() >= 100
    assert lib.rvo_search_workspace_bytes(1_000_000, 1024, 256, 100) > 0
    assert lib.rvo_search_workspace_bytes(1_000_000, 1024, 1, 10) > 0
    assert lib.rvo_search_workspace_bytes(10, 1024, 1, 0) == 0          # k out of range
    assert lib.rvo_search_workspace_bytes(10, 1024, 1, _lib.RVO_MAX_K + 1) == 0
    assert lib.rvo_mask_pool_workspace_bytes(256, 64, 576, 1024) > 256 * 64 * 576 * 2
    assert lib.rvo_packed_result_bytes(256, 100) == 256 * 100 * 12 + 256 * 4
    assert lib.rvo_padded_queries(1, 1024) == 16 and lib.rvo_padded_queries(256, 1024) == 256
    assert lib.rvo_padded_queries(4096, 1280) == 4096 and lib.rvo_padded_queries(300, 1024) == 320


def test_bad_option_and_error_text():
    lib = _lib.load()
    assert lib.rvo_set_option(b"no_such_option", 1) < 0
    assert b"no_such_option" in lib.rvo_last_error()
    with pytest.raises(_lib.RvoError):
        _lib.set_option("no_such_option", 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from revers_o_b200 import ops
    from revers_o_b200.vector_db import B200VectorDB
    with pytest.raises(_lib.RvoError):
        ops.normalize_rows(torch.zeros(2, 8))
    with pytest.raises(_lib.RvoError):
        B200VectorDB()
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    rc = lib.rvo_normalize_rows(ctypes.addressof(buf), 1, 8, 8, ctypes.addressof(buf), 8, -1, None, None)
    assert rc < 0 and lib.rvo_last_error() != b""


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "revers_o_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f
