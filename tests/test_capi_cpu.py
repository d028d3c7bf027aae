"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/syzgy_b200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

import syzgydb_b200
from syzgydb_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "syzgy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(szg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.load()
    declared = _declared()
    assert declared, "header parse found nothing"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/syzgy_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == declared


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "syzgy_b200.h")).read()
    assert len(re.findall(r"[a-z]+\.go:\d+", src)) >= 15


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "syzgy_b200.h")).read()
    assert "torch" not in src and "at::" not in src and "std::" not in src


def _gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_gpu(), reason="needs a box without a GPU")
def test_fails_loudly_without_gpu():
    with pytest.raises(syzgydb_b200.SzgError) as e:
        syzgydb_b200.Index(8, 8, syzgydb_b200.COSINE)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_bad_arguments_are_statuses_not_crashes():
    L = _capi.load()
    h = ctypes.c_void_p()
    assert L.szg_create(8, 7, 0, 0, ctypes.byref(h)) == -1  # quantization 7: the reference panics (collection.go:809)
    assert b"quantization" in L.szg_last_error()
    assert L.szg_create(8, 8, 5, 0, ctypes.byref(h)) == -1  # unsupported distance method (collection.go:281-282)
    assert L.szg_create(0, 8, 0, 0, ctypes.byref(h)) == -1
    assert L.szg_destroy(None) == 0
    assert L.szg_count(None, None) == -1


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under syzgydb_b200/ may reference it
    pkg = os.path.join(ROOT, "syzgydb_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in text and "libsyzgy_oracle" not in text and "import oracle" not in text, f
