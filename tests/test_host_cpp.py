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


@pytest.mark.gpu
def test_host_mirror_cfg3_at_size_lsh_radius_filter_against_oracle():
    """BASELINE.json configs[2] at its stated size: 1 M x 384 float64 cosine documents in the C++ host mirror (5 LSH trees built
    on 5 threads), medium-precision Search with radius 0.46 + the bucket < 3 filter and with K = 10: the GPU-rescored
    result must equal the oracle's replay of `consider` over the same visit sequence (lshtree.go:283-351,
    collection.go:598-619), and the exact search the oracle's scan of all rows."""
    _build()
    out = subprocess.run([BIN, "--cfg3", "1000000"], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert "PASS cfg3 k=0 radius=0.46 filter=1" in out.stdout and "PASS cfg3 k=10 radius=0 filter=0" in out.stdout, out.stdout[-4000:]


@pytest.mark.gpu
def test_host_mirror_opens_a_collection_file_and_batches(tmp_path):
    """Collection::Open over a span file written by the restated reference writer (updates and removals included),
    then SearchBatch: the C++ mirror must return what the oracle returns, with the metadata of the file."""
    import json

    import numpy as np

    from oracle import pyoracle as o
    from oracle import spanfile as sfo
    _build()
    n, dims, bits, metric, k, nq = 600, 16, 8, 1, 5, 6
    w = sfo.SpanFileWriter()
    w.write_header("hosttest", metric, dims, bits)
    codes = o.synth_rows(33, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64) * 2 + 5
    for i in range(n):
        w.add_document(int(ids[i]), codes[i].tobytes(), json.dumps({"i": int(ids[i])}).encode())
    new = o.synth_rows(34, 0, 20, dims, bits)
    for j in range(20):                       # updates move the records and leave FREE spans behind
        w.add_document(int(ids[j * 7]), new[j].tobytes(), json.dumps({"i": int(ids[j * 7]), "v": 2}).encode())
        codes[j * 7] = new[j]
    gone = set(int(x) for x in ids[100:130])
    for x in gone:
        w.remove_document(x)
    keep = np.array([int(x) not in gone for x in ids])
    path = os.path.join(tmp_path, "host.dat")
    with open(path, "wb") as f:
        f.write(w.tobytes())
    qs = o.synth_queries(35, 0, nq, dims)
    argv = [BIN, "--open", path, str(k)] + [repr(float(x)) for x in qs.reshape(-1)]
    out = subprocess.run(argv, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert f"OPTIONS hosttest {metric} {dims} {bits}" in lines and f"COUNT {int(keep.sum())}" in lines
    got = {}
    for ln in lines:
        if ln.startswith("RESULT "):
            _, q, i, d, meta = ln.split(" ", 4)
            got.setdefault(int(q), []).append((int(i), float.fromhex(d), meta))
    for qi in range(nq):
        ri, rd, _ = o.search_exact(codes[keep], ids[keep], dims, bits, metric, qs[qi], k=k)
        assert [g[0] for g in got[qi]] == ri.tolist(), (qi, got[qi], ri)
        assert np.allclose([g[1] for g in got[qi]], rd, rtol=1e-12, atol=0)
        for g in got[qi]:
            assert json.loads(g[2])["i"] == g[0]
    single = [ln.split(" ", 4) for ln in lines if ln.startswith("SINGLE ")]
    assert [int(s[2]) for s in single] == [g[0] for g in got[0]]
    assert "PERCENT 100.000000" in lines
    lsh = [ln for ln in lines if ln.startswith("LSH ")][0].split()
    assert int(lsh[1]) == k
