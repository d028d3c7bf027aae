"""CPU tests of the oracle (oracle/syzgy_oracle.c): the reference's own known-answer and
behavioural tests restated, the committed golden vectors, and self-consistency."""
import math

import numpy as np
import pytest

from oracle import pyoracle as o
from tests.common import golden_case_inputs, load_golden


def test_reference_kat_euclidean():
    # collection_test.go:12-21 -- the only numeric known-answer the reference holds on this path
    assert o.euclidean([1.0, 2.0, 3.0], [4.0, 5.0, 6.0]) == 5.196152422706632


def test_codec_tables():
    g = load_golden()["restated_codec"]
    for bits in (4, 8, 16):
        assert [o.quantize(v, bits) for v in g["inputs"]] == g[str(bits)]
    assert o.dequantize(128, 8) == float.fromhex(g["dequantize_128_8"])
    # 0.0 does not round-trip through 8 bits
    assert o.dequantize(o.quantize(0.0, 8), 8) != 0.0


@pytest.mark.parametrize("bits", [32, 64])
def test_float_codec_passthrough(bits):
    # quantization.go:6-10: no clamping for 32/64
    for v in (-3.5, 0.1, 1e10, -1e-20):
        back = o.dequantize(o.quantize(v, bits), bits)
        assert back == (np.float64(np.float32(v)) if bits == 32 else v)


def test_vector_size():
    # collection.go:796-811
    assert [o.vector_size(b, 7) for b in (4, 8, 16, 32, 64)] == [4, 7, 14, 28, 56]
    assert o.vector_size(12, 7) == -1


def test_packing_layout():
    # collection.go:713-744: 4-bit high nibble first, odd dim leaves the last low nibble 0; big-endian 16/32/64
    assert o.encode([-1, 1, 0.5], 4).tolist() == [0x0F, 0xB0]
    assert o.encode([1.0], 16).tolist() == [0xFF, 0xFF]
    assert o.encode([0.0], 16).tolist() == [0x80, 0x00]
    assert o.encode([1.0], 32).tolist() == [0x3F, 0x80, 0x00, 0x00]
    assert o.encode([1.0], 64).tolist() == [0x3F, 0xF0, 0, 0, 0, 0, 0, 0]
    for bits in (4, 8, 16, 32, 64):
        v = np.linspace(-1, 1, 9)
        assert np.allclose(o.decode(o.encode(v, bits), 9, bits), v, atol=2.0 / ((1 << min(bits, 20)) - 1))


def test_fp64_roundtrip_exact():
    # collection_test.go:422-436, 537-547: fp64 vectors round-trip bit-exactly
    v = np.random.default_rng(1).normal(size=33) * 100
    assert np.array_equal(o.decode(o.encode(v, 64), 33, 64), v)


def test_angular_distance_properties():
    # collection.go:821-832
    a = np.array([1.0, 0.0]); b = np.array([0.0, 2.0])
    assert o.angular(a, b) == 0.5
    assert o.angular(a, -a) == 1.0
    assert o.angular(a, np.zeros(2)) == 1.0 and o.angular(np.zeros(2), a) == 1.0
    # ratio rounding above 1 gives NaN (Go math.Acos), SURVEY.md appendix B-10
    rng = np.random.default_rng(0)
    vs = rng.random((2000, 3))
    nans = sum(math.isnan(o.angular(v, v)) for v in vs)
    assert 0 < nans < 2000


def test_lex_order():
    # spanfile.go:540-560: sort.Strings over decimal ids
    ids = np.array([1, 2, 10, 20, 3, 100, 19, 0], dtype=np.uint64)
    want = sorted(ids.tolist(), key=lambda x: str(x))
    assert ids[o.lex_order(ids)].tolist() == want


def test_exhaustive_search_tiny():
    # collection_test.go:549-612: K=3 exact over 3 docs returns ids {1,2,3}, PercentSearched == 100
    vecs = np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]])
    codes = o.encode_rows(vecs, 64)
    ids = np.array([1, 2, 3], dtype=np.uint64)
    rid, rd, pct = o.search_exact(codes, ids, 3, 64, o.EUCLIDEAN, [1.0, 2, 3], k=3)
    assert rid.tolist() == [1, 2, 3] and pct == 100.0
    assert rd[0] == 0.0 and rd[1] == 5.196152422706632


def test_collection_search_properties():
    # collection_test.go:283-382
    rng = np.random.default_rng(5)
    empty = o.search_exact(np.zeros((0, 16), np.uint8), np.zeros(0, np.uint64), 2, 64, o.EUCLIDEAN, [50.0, 50], k=5)
    assert empty[0].size == 0 and empty[2] == 0.0
    vecs = rng.random((10, 2)) * 100
    codes = o.encode_rows(vecs, 64)
    ids = np.arange(10, dtype=np.uint64)
    assert o.search_exact(codes, ids, 2, 64, o.EUCLIDEAN, [50.0, 50], k=3)[0].size == 3
    rid, rd, _ = o.search_exact(codes, ids, 2, 64, o.EUCLIDEAN, [50.0, 50], radius=30.0)
    truth = np.sqrt(((vecs - 50) ** 2).sum(1))
    assert sorted(rid.tolist()) == sorted(np.nonzero(truth <= 30.0)[0].tolist()) and np.all(rd <= 30.0)
    # radius overrides K (collection.go:598-605)
    rid2, _, _ = o.search_exact(codes, ids, 2, 64, o.EUCLIDEAN, [50.0, 50], k=1, radius=30.0)
    assert rid2.size == rid.size
    # filter: only even ids, but filtered rows still count as searched (collection.go:589 precedes 592)
    rid, _, pct = o.search_exact(codes, ids, 2, 64, o.EUCLIDEAN, [50.0, 50], k=5, passmask=(ids % 2 == 0))
    assert all(i % 2 == 0 for i in rid.tolist()) and pct == 100.0


def test_first_seen_wins_ties():
    # strict '>' replacement (collection.go:608): among equal distances the earlier in scan order survives
    vecs = np.zeros((6, 2)); vecs[:, 0] = 1.0
    codes = o.encode_rows(vecs, 64)
    ids = np.array([30, 4, 100, 2, 11, 5], dtype=np.uint64)  # lexicographic: 100, 11, 2, 30, 4, 5
    rid, rd, _ = o.search_exact(codes, ids, 2, 64, o.EUCLIDEAN, [0.0, 0.0], k=3)
    assert sorted(rid.tolist()) == [2, 11, 100] and np.all(rd == 1.0)


def test_nan_policy():
    # NaN never satisfies radius and enters the k-heap only while it is not full (collection.go:598, 608)
    rng = np.random.default_rng(0)
    vs = rng.random((400, 3))
    q = next(v for v in vs if math.isnan(o.angular(v, v)))
    rows = np.vstack([rng.random((50, 3)), q[None, :]])  # the NaN row is scanned last (ids ascending, same width)
    codes = o.encode_rows(rows, 64)
    ids = np.arange(100, 151, dtype=np.uint64)
    rid, rd, _ = o.search_exact(codes, ids, 3, 64, o.COSINE, q, k=5)
    assert 150 not in rid.tolist() and not np.isnan(rd).any()
    rid, rd, _ = o.search_exact(codes, ids, 3, 64, o.COSINE, q, radius=0.9)
    assert 150 not in rid.tolist()


def test_golden_vectors_reproduce():
    for case in load_golden()["cases"]:
        codes, ids, queries, passmask = golden_case_inputs(case)
        for q, want in zip(queries, case["results"]):
            rid, rd, pct = o.search_exact(codes, ids, case["dims"], case["bits"], case["metric"], q, k=case["k"],
                                          radius=case["radius"], passmask=passmask)
            assert rid.tolist() == want["ids"], case["name"]
            assert [float.hex(float(x)) for x in rd] == want["dist"], case["name"]
            assert pct == want["percent"]


def test_search_matches_numpy_bruteforce():
    n, d = 1500, 24
    for bits, metric in ((8, o.COSINE), (4, o.EUCLIDEAN), (16, o.COSINE), (64, o.EUCLIDEAN)):
        codes = o.synth_rows(3, 0, n, d, bits)
        q = o.synth_queries(4, 0, 1, d)[0]
        X = np.stack([o.decode(c, d, bits) for c in codes])
        if metric == o.COSINE:
            dist = np.arccos(np.clip(X @ q / np.linalg.norm(X, axis=1) / np.linalg.norm(q), -1, 1)) / np.pi
        else:
            dist = np.sqrt(((X - q) ** 2).sum(1))
        rid, rd, _ = o.search_exact(codes, np.arange(n, dtype=np.uint64), d, bits, metric, q, k=8)
        assert rid.tolist() == np.argsort(dist, kind="stable")[:8].tolist()
        assert np.allclose(rd, np.sort(dist)[:8], rtol=1e-12)


def test_lsh_search_matches_replay_and_exact_subset():
    # collection_test.go:23-103 restated: same count, PercentSearched < 100, and the LSH result equals
    # `consider` replayed over the visit sequence (appendix B-13)
    n, d = 4000, 16
    codes = o.synth_rows(11, 0, n, d, 64)
    ids = np.arange(n, dtype=np.uint64)
    tree = o.LshTree(codes, d, 64, o.COSINE, seed=7)
    tree.add_all_decoded(n)
    q = o.synth_queries(12, 0, 1, d)[0]
    lid, ld, pct, visit = tree.search(ids, q, k=10)
    eid, ed, _ = o.search_exact(codes, ids, d, 64, o.COSINE, q, k=10)
    assert lid.size == eid.size == 10 and pct < 100 and len(visit) == round(pct / 100 * n)
    rid, rd, ps = o.replay(codes, ids, d, 64, o.COSINE, q, visit, k=10)
    assert rid.tolist() == lid.tolist() and np.array_equal(rd, ld) and ps == len(visit)
    assert np.all(ld >= ed - 1e-15)
    tree.close()


def test_synth_generator_properties():
    a = o.synth_rows(1, 0, 64, 33, 8)
    b = o.synth_rows(1, 32, 32, 33, 8)
    assert np.array_equal(a[32:], b)  # counter-based: any row range reproduces
    c4 = o.synth_rows(1, 0, 8, 7, 4)
    assert np.all(c4[:, -1] & 0x0F == 0)  # odd dims: last low nibble stays 0 (encodeDocument)
    f = o.synth_rows(2, 0, 16, 5, 32)
    v = np.stack([o.decode(r, 5, 32) for r in f])
    assert np.all(np.abs(v) <= 1)


def test_go_math_restatement_against_libm():
    """Go's math.Atan / math.Acos (standard library, Cephes algorithm) are restated in the oracle from the published
    algorithm; the coefficients were written down without access to the Go source, so the restatement is pinned
    against libm: a wrong coefficient would show up as an error of many ulp."""
    import math
    rng = np.random.default_rng(0)

    def ulps(a, b):
        return 0.0 if a == b else abs(a - b) / math.ulp(max(abs(a), abs(b)))

    worst = max(ulps(o.go_atan(float(x)), math.atan(float(x)))
                for x in np.concatenate([rng.uniform(-50, 50, 20000), np.logspace(-10, 10, 2000), -np.logspace(-10, 10, 2000)]))
    assert worst <= 1.0, f"atan restatement is {worst} ulp from libm"
    # Acos = Pi/2 - Asin cancels near +1: the ABSOLUTE error stays at one ulp of Pi/2, the relative one does not
    # (that is a property of the reference's library, not of this restatement)
    xs = np.concatenate([rng.uniform(-0.9, 0.9, 20000), np.linspace(-0.9, 0.9, 2001)])
    assert max(abs(o.go_acos(float(x)) - math.acos(float(x))) for x in xs) <= 4.5e-16
    # towards +-1 the reference's Asin computes Sqrt(1 - x*x) from an already rounded product and Acos = Pi/2 - Asin
    # cancels: measured against libm, Go's Acos is off by up to 6e-16 absolute on (0.9, 0.999), 2e-14 on
    # (0.999, 1 - 1e-6) and 2.3e-13 closer to 1 (1e-11 .. 2e-9 relative).  The GPU path reproduces exactly this.
    for lo, hi, tol in ((0.9, 0.999, 2e-15), (0.999, 1 - 1e-6, 1e-13), (1 - 1e-6, 1.0, 1e-12)):
        zone = rng.uniform(lo, hi, 5000)
        assert max(abs(o.go_acos(float(x)) - math.acos(float(x))) for x in zone) <= tol, (lo, hi)
    mid = rng.uniform(-0.9, 0.9, 20000)
    assert max(ulps(o.go_acos(float(x)), math.acos(float(x))) for x in mid) <= 4.0  # small results near x = 0.9: 1 ulp of Pi/2 is 4 ulp of 0.45
    assert o.go_acos(1.0) == 0.0 and o.go_acos(-1.0) == math.pi and o.go_acos(0.0) == math.pi / 2
    assert math.isnan(o.go_acos(1.0000000000000002)) and math.isnan(o.go_acos(float("nan")))  # collection.go:831 -> NaN policy
    # the angular distance follows: libm mode and Go mode agree to a few ulp on random vectors
    a, b = rng.normal(size=64), rng.normal(size=64)
    d_go = o.angular(a, b)
    o.libm_acos_mode(True)
    try:
        d_libm = o.angular(a, b)
    finally:
        o.libm_acos_mode(False)
    assert abs(d_go - d_libm) <= 4 * math.ulp(d_go)


def test_faithful_getdocument_variant_reads_the_same_records():
    """The faithful CPU variant (decimal-string key -> index -> parseSpan -> CRC-32 over the span -> decode with
    an allocation per record, collection.go:470-484 / spanfile.go:730-849) must find every record the lean variant reads."""
    import zlib
    assert o.crc32(b"123456789") == 0xCBF43926
    blob = bytes(range(256)) * 37 + b"tail"
    assert o.crc32(blob) == zlib.crc32(blob)
    for bits, dims in ((8, 96), (4, 33), (64, 5)):
        n = 3000
        codes = o.synth_rows(9, 0, n, dims, bits)
        ids = np.arange(n, dtype=np.uint64) * 7 + 3
        sp = o.Spans(codes, ids)
        q = o.synth_queries(10, 0, 1, dims)[0]
        for kw in (dict(k=10), dict(radius=0.49 if bits != 4 else 4.0)):
            metric = o.COSINE if bits != 4 else o.EUCLIDEAN
            a = o.search_exact(codes, ids, dims, bits, metric, q, **kw)
            b = o.search_exact(codes, ids, dims, bits, metric, q, spans=sp, **kw)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True) and a[2] == b[2] == 100.0
        sp.close()
