"""Runs the C++ host mirror's test binary (tests/cpp/test_collection.cpp: the reference's collection_test.go
restated against syzgydb_b200/host, plus oracle parity of the LSH replay and of exact search)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_collection")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "syzgydb_b200", "csrc"), "-j", str(os.cpu_count() or 2)],
                          stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "syzgydb_b200", "host")], stdout=subprocess.DEVNULL)


def test_host_mirror_codec_and_no_gpu_behaviour():
    _build()
    out = subprocess.run([BIN, "--cpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS TestCodec" in out.stdout


@pytest.mark.gpu
def test_host_mirror_collection_tests_on_gpu():
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    for name in ("TestEuclideanDistance", "TestCollectionSearch", "TestExhaustiveSearch",
                 "TestVectorSearchWith4BitQuantization", "TestDocumentRoundTripUpdateRemove", "TestListModePagination",
                 "TestSearchExactVsLSH", "TestLSHQuantized", "TestLSHRadiusWithFilter"):
        assert f"PASS {name}" in out.stdout, out.stdout[-4000:]
