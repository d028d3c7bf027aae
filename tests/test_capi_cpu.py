"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/syzgy_b200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess
import tempfile

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


def test_ctypes_structs_match_the_header_layout(tmp_path=None):
    # the binding mirrors four structs by hand: compile the header with gcc and compare sizes, field offsets and constants
    probe = r"""
#include <stddef.h>
#include <stdio.h>
#include "syzgy_b200.h"
#define F(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
    printf("sizeof.szg_stats %zu\n", sizeof(szg_stats));
    printf("sizeof.szg_meta_value %zu\n", sizeof(szg_meta_value));
    printf("sizeof.szg_filter_op %zu\n", sizeof(szg_filter_op));
    printf("sizeof.szg_spanfile_info %zu\n", sizeof(szg_spanfile_info));
    F(szg_stats, batch_queries); F(szg_stats, rowbytes); F(szg_stats, scan_smem_bytes); F(szg_stats, combined_queries);
    F(szg_meta_value, kind); F(szg_meta_value, str_len); F(szg_meta_value, num); F(szg_meta_value, str);
    F(szg_filter_op, op); F(szg_filter_op, arg); F(szg_filter_op, num); F(szg_filter_op, str); F(szg_filter_op, str_len);
    F(szg_filter_op, table_len); F(szg_filter_op, table);
    printf("const.SZG_FOP_NOT_EXISTS %u\n", SZG_FOP_NOT_EXISTS);
    printf("const.SZG_FOP_STR_TABLE %u\n", SZG_FOP_STR_TABLE);
    printf("const.SZG_OPT_COMBINE %d\n", SZG_OPT_COMBINE);
    printf("const.SZG_MV_ERROR %u\n", SZG_MV_ERROR);
    return 0;
}
"""
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "probe.c"), os.path.join(d, "probe")
        open(src, "w").write(probe)
        subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        out = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    structs = {"szg_stats": _capi.Stats, "szg_meta_value": _capi.MetaValue, "szg_filter_op": _capi.FilterOp,
               "szg_spanfile_info": _capi.SpanFileInfo}
    for name, cls in structs.items():
        assert ctypes.sizeof(cls) == int(out[f"sizeof.{name}"]), name
    for key, val in out.items():
        if key.startswith(("sizeof.", "const.")):
            continue
        sname, field = key.split(".")
        assert getattr(structs[sname], field).offset == int(val), key
    assert _capi.FOP_NOT_EXISTS == int(out["const.SZG_FOP_NOT_EXISTS"]) and _capi.FOP_STR_TABLE == int(out["const.SZG_FOP_STR_TABLE"])
    assert _capi.OPT_COMBINE == int(out["const.SZG_OPT_COMBINE"])
    from syzgydb_b200 import filter as hf
    assert hf.MV_ERROR == int(out["const.SZG_MV_ERROR"])
