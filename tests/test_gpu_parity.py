"""GPU parity tests: the CUDA path, called through the C ABI (ctypes over libsyzgy_b200.so),
against the CPU oracle on the same seeded inputs and against the committed golden vectors.

Bar (BASELINE.json north_star): ids and order identical to the reference scan in
lexicographic-decimal-id order, except among results whose float64 distances are within 1e-5
relative of each other; returned distances are the fp64 re-score and are compared bit for bit (every
operation IEEE-exact and in the reference's order, math.Acos restated on both sides).
"""
import math

import numpy as np
import pytest

import syzgydb_b200 as szg
from oracle import pyoracle as o
from syzgydb_b200 import _capi
from tests.common import (assert_radius_match, assert_results_match, golden_case_inputs, load_golden)

pytestmark = pytest.mark.gpu


def _build(codes, ids, dims, bits, metric):
    ix = szg.Index(dims, bits, metric)
    ix.upsert(ids, codes)
    return ix


def _true_dist(codes, ids, dims, bits, metric, q):
    row_of = {int(i): r for r, i in enumerate(ids.tolist())}
    return lambda i: float(o.row_distances(codes, dims, bits, metric, q, [row_of[i]])[0])


# ------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("case", load_golden()["cases"], ids=lambda c: c["name"])
def test_golden_cases(case):
    codes, ids, queries, passmask = golden_case_inputs(case)
    dims, bits, metric = case["dims"], case["bits"], case["metric"]
    with _build(codes, ids, dims, bits, metric) as ix:
        assert ix.count() == case["n"]
        mask = -1 if passmask is None else ix.mask_create(ids, passmask)
        for q, want in zip(queries, case["results"]):
            ref_ids = np.array(want["ids"], dtype=np.uint64)
            ref_dist = np.array([float.fromhex(x) for x in want["dist"]])
            if case["radius"] > 0:
                gi, gd, scanned = ix.search_radius(q, case["radius"], mask_id=mask)
                assert_radius_match(gi, gd, ref_ids, ref_dist, case["radius"], case["name"])
            else:
                gi, gd, n, scanned = ix.search_topk(q, case["k"], mask_id=mask)
                assert n[0] == ref_ids.size
                assert_results_match(gi[0, :n[0]], gd[0, :n[0]], ref_ids, ref_dist,
                                     _true_dist(codes, ids, dims, bits, metric, q), case["name"])
            # PercentSearched: filtered rows count as searched (collection.go:589 precedes 592)
            assert scanned == case["n"] and want["percent"] == 100.0


# ------------------------------------------------------------------ every quantization x metric vs the oracle
@pytest.mark.parametrize("bits", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("metric", [szg.EUCLIDEAN, szg.COSINE])
@pytest.mark.parametrize("dims", [1, 31, 128, 385])
def test_topk_matches_oracle(bits, metric, dims):
    n, k, seed = 5000, 10, 77 + bits + dims
    codes = o.synth_rows(seed, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64)
    queries = o.synth_queries(seed + 1, 0, 4, dims)
    with _build(codes, ids, dims, bits, metric) as ix:
        gi, gd, gn, _ = ix.search_topk(queries, k)
        for qi, q in enumerate(queries):
            ri, rd, _ = o.search_exact(codes, ids, dims, bits, metric, q, k=k)
            if np.isnan(rd).any():
                continue  # NaN policy is tested separately
            assert gn[qi] == ri.size
            assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd,
                                 _true_dist(codes, ids, dims, bits, metric, q), f"b{bits} m{metric} d{dims} q{qi}")


@pytest.mark.parametrize("bits,metric,dims", [(8, szg.COSINE, 96), (4, szg.EUCLIDEAN, 64), (16, szg.EUCLIDEAN, 40),
                                              (32, szg.COSINE, 20), (64, szg.EUCLIDEAN, 24)])
@pytest.mark.parametrize("k", [1, 33, 100, 224])
def test_large_k(bits, metric, dims, k):
    n, seed = 4000, 5
    codes = o.synth_rows(seed, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64) + 1000
    q = o.synth_queries(seed + 1, 0, 1, dims)[0]
    with _build(codes, ids, dims, bits, metric) as ix:
        gi, gd, gn, _ = ix.search_topk(q, k)
        ri, rd, _ = o.search_exact(codes, ids, dims, bits, metric, q, k=k)
        assert gn[0] == ri.size == k
        assert_results_match(gi[0], gd[0], ri, rd, _true_dist(codes, ids, dims, bits, metric, q), f"k={k}")


def test_gaussian_normalised_rows_q8_cosine():
    # all-minilm-like rows: L2-normalised Gaussians pushed through the restated quantize
    rng = np.random.default_rng(3)
    n, d = 20000, 384
    x = rng.normal(size=(n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    codes = o.encode_rows(x, 8)
    ids = np.arange(n, dtype=np.uint64)
    qs = rng.normal(size=(3, d))
    with _build(codes, ids, d, 8, szg.COSINE) as ix:
        gi, gd, gn, _ = ix.search_topk(qs, 10)
        for qi, q in enumerate(qs):
            ri, rd, _ = o.search_exact(codes, ids, d, 8, szg.COSINE, q, k=10)
            assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd,
                                 _true_dist(codes, ids, d, 8, szg.COSINE, q), f"gauss q{qi}")


# ------------------------------------------------------------------ reference test behaviours
def test_reference_kat_and_tiny_exact_search():
    # collection_test.go:12-21 and 549-612
    vecs = np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]])
    codes = o.encode_rows(vecs, 64)
    ids = np.array([1, 2, 3], dtype=np.uint64)
    with _build(codes, ids, 3, 64, szg.EUCLIDEAN) as ix:
        gi, gd, gn, scanned = ix.search_topk([1.0, 2, 3], 3)
        assert gi[0].tolist() == [1, 2, 3] and gn[0] == 3 and scanned == 3
        assert gd[0, 0] == 0.0 and gd[0, 1] == 5.196152422706632


def test_collection_search_properties():
    # collection_test.go:283-382: empty -> 0, len <= K, radius hits <= r, filter honoured, K > N
    rng = np.random.default_rng(5)
    with szg.Index(2, 64, szg.EUCLIDEAN) as ix:
        gi, gd, gn, scanned = ix.search_topk([50.0, 50.0], 5)
        assert gn[0] == 0 and scanned == 0
        ri, rd, s = ix.search_radius([50.0, 50.0], 10.0)
        assert ri.size == 0 and s == 0
        vecs = rng.random((10, 2)) * 100
        ids = np.arange(10, dtype=np.uint64)
        ix.upsert(ids, o.encode_rows(vecs, 64))
        assert ix.search_topk([50.0, 50.0], 3)[2][0] == 3
        gi, gd, gn, _ = ix.search_topk([50.0, 50.0], 25)  # K > N returns all rows (appendix B-7)
        assert gn[0] == 10 and sorted(gi[0, :10].tolist()) == list(range(10))
        ri, rd, _ = ix.search_radius([50.0, 50.0], 30.0)
        truth = np.sqrt(((vecs - 50) ** 2).sum(1))
        assert sorted(ri.tolist()) == sorted(np.nonzero(truth <= 30.0)[0].tolist()) and np.all(rd <= 30.0)
        m = ix.mask_create(ids, (ids % 2 == 0))
        gi, gd, gn, scanned = ix.search_topk([50.0, 50.0], 5, mask_id=m)
        assert gn[0] == 5 and all(i % 2 == 0 for i in gi[0].tolist()) and scanned == 10


def test_first_seen_wins_ties_in_lexicographic_order():
    # strict '>' (collection.go:608) + sorted scan (spanfile.go:540-560): equal distances keep the
    # lexicographically smallest decimal ids
    vecs = np.zeros((6, 2)); vecs[:, 0] = 1.0
    ids = np.array([30, 4, 100, 2, 11, 5], dtype=np.uint64)
    with _build(o.encode_rows(vecs, 64), ids, 2, 64, szg.EUCLIDEAN) as ix:
        gi, gd, gn, _ = ix.search_topk([0.0, 0.0], 3)
        assert gi[0].tolist() == [100, 11, 2] and np.all(gd[0] == 1.0)


def test_zero_norm_rows_and_query_cosine():
    # collection.go:828-830: distance exactly 1.0 when either norm is 0
    vecs = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0]])
    ids = np.array([7, 8, 9, 10], dtype=np.uint64)
    with _build(o.encode_rows(vecs, 64), ids, 3, 64, szg.COSINE) as ix:
        gi, gd, gn, _ = ix.search_topk([1.0, 0, 0], 4)
        want = {8: 0.0, 9: 0.5, 7: 1.0, 10: 1.0}
        assert gn[0] == 4 and {int(i): float(d) for i, d in zip(gi[0], gd[0])} == want
        assert gi[0].tolist()[:2] == [8, 9] and gi[0].tolist()[2:] == [10, 7]  # tie at 1.0: "10" < "7"
        gi, gd, gn, _ = ix.search_topk([0.0, 0, 0], 4)
        assert np.all(gd[0] == 1.0) and gi[0].tolist() == [10, 7, 8, 9]
        assert np.all(ix.rescore([0.0, 0, 0], ids) == 1.0)


def test_nan_rows_are_never_returned():
    # Deliberate, documented deviation: the reference lets a NaN
    # distance (cosine ratio rounded above 1) into the heap only while it is not full; the GPU path
    # never returns NaN.  With the NaN row scanned after the heap is full both agree.
    rng = np.random.default_rng(0)
    vs = rng.random((400, 3))
    q = next(v for v in vs if math.isnan(o.angular(v, v)))
    rows = np.vstack([rng.random((50, 3)), q[None, :]])
    codes = o.encode_rows(rows, 64)
    ids = np.arange(100, 151, dtype=np.uint64)
    with _build(codes, ids, 3, 64, szg.COSINE) as ix:
        gi, gd, gn, _ = ix.search_topk(q, 5)
        ri, rd, _ = o.search_exact(codes, ids, 3, 64, szg.COSINE, q, k=5)
        assert 150 not in gi[0].tolist() and not np.isnan(gd).any()
        assert_results_match(gi[0], gd[0], ri, rd)
        ri2, rd2, _ = ix.search_radius(q, 0.9)
        assert 150 not in ri2.tolist()
        assert math.isnan(ix.rescore(q, [150])[0])  # the raw distance is NaN, like Go's math.Acos


def test_values_outside_unit_range_and_float_passthrough():
    # quantization.go:6-17: 4/8/16 clamp to [-1,1]; 32/64 pass through
    rng = np.random.default_rng(9)
    x = rng.normal(size=(3000, 12)) * 3.0
    q = rng.normal(size=12) * 2.0
    ids = np.arange(3000, dtype=np.uint64)
    for bits in (4, 16, 32, 64):
        codes = o.encode_rows(x, bits)
        with _build(codes, ids, 12, bits, szg.EUCLIDEAN) as ix:
            gi, gd, gn, _ = ix.search_topk(q, 10)
            ri, rd, _ = o.search_exact(codes, ids, 12, bits, szg.EUCLIDEAN, q, k=10)
            assert_results_match(gi[0], gd[0], ri, rd, _true_dist(codes, ids, 12, bits, szg.EUCLIDEAN, q), f"b{bits}")


# ------------------------------------------------------------------ mirror maintenance
@pytest.mark.parametrize("bits", [4, 8, 16, 32, 64])
def test_mirror_roundtrip_and_synthetic_generator(bits):
    dims = 37
    codes = o.synth_rows(21, 0, 300, dims, bits)
    ids = np.arange(300, dtype=np.uint64) * 7
    with _build(codes, ids, dims, bits, szg.COSINE) as ix:
        assert np.array_equal(ix.fetch_codes(ids[::-1]), codes[::-1])  # stream-1 bytes come back verbatim
    with szg.Index(dims, bits, szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(21, 0, 300)  # device generator == oracle generator
        assert np.array_equal(ix.fetch_codes(np.arange(300, dtype=np.uint64)), codes)
        ix.fill_synthetic(21, 1000, 50)
        assert np.array_equal(ix.fetch_codes(np.arange(1000, 1050, dtype=np.uint64)),
                              o.synth_rows(21, 1000, 50, dims, bits))


@pytest.mark.parametrize("bits", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("dims", [1, 7, 37, 384])
def test_device_encode_is_encodeDocument(bits, dims):
    # szg_encode = encodeDocument (collection.go:713-744) + quantize (quantization.go:5-23) on the device: byte-identical
    # to the oracle's restatement, including the clamp, the exact .5 ties of math.Round and odd 4-bit tails
    rng = np.random.default_rng(100 + bits + dims)
    n = 257
    x = rng.uniform(-1.3, 1.3, size=(n, dims))
    if bits <= 16:  # values whose scaled image is an exact half (ties round away from zero), and the a7 check values
        M = (1 << bits) - 1
        ties = (np.arange(n * dims) % (M + 1) + 0.5) / M * 2 - 1
        x[::3] = ties.reshape(n, dims)[::3]
        kat = np.array([-1.5, -1, -.5, 0, .1, .5, 1, 2])
        x[1, :min(dims, 8)] = kat[:min(dims, 8)]
    want = o.encode_rows(x, bits)
    ids = np.arange(n, dtype=np.uint64) * 3 + 1
    with szg.Index(dims, bits, szg.EUCLIDEAN) as ix:
        got = ix.encode(x)
        assert ix.count() == 0  # encode only
        assert np.array_equal(got, want)
        got2 = ix.encode(x, ids=ids, upsert=True)
        assert np.array_equal(got2, want) and ix.count() == n
        assert np.array_equal(ix.fetch_codes(ids), want)  # the mirror holds the same rows
        q = rng.uniform(-1, 1, size=dims)
        gi, gd, gn, _ = ix.search_topk(q, 5)
        ri, rd, _ = o.search_exact(want, ids, dims, bits, szg.EUCLIDEAN, q, k=5)
        assert_results_match(gi[0, :gn[0]], gd[0, :gn[0]], ri, rd, what=f"encode b{bits}")
    if bits == 8 and dims >= 8:
        assert got[1, :8].tolist() == [0, 0, 64, 128, 140, 191, 255, 255]


def test_device_encode_large_batch_spans_staging_chunks():
    # more rows than one 64 MB staging chunk of float64 input holds (4-bit: 16 input bytes per code byte)
    dims, n = 768, 30000
    rng = np.random.default_rng(5)
    x = rng.normal(size=(n, dims)) * 0.5
    with szg.Index(dims, 4, szg.COSINE) as ix:
        got = ix.encode(x, ids=np.arange(n, dtype=np.uint64), upsert=True)
        rows = np.arange(0, n, 611)
        assert np.array_equal(got[rows], o.encode_rows(x[rows], 4))
        assert ix.count() == n
        assert np.array_equal(ix.fetch_codes(rows.astype(np.uint64)), got[rows])


def test_upsert_replaces_and_remove_hides():
    # appendix B-14: re-adding an id replaces its row (spanfile.go:459-472); removeDocument hides it
    dims, bits = 16, 8
    codes = o.synth_rows(31, 0, 2000, dims, bits)
    ids = np.arange(2000, dtype=np.uint64)
    q = o.synth_queries(32, 0, 1, dims)[0]
    with _build(codes, ids, dims, bits, szg.EUCLIDEAN) as ix:
        gi, gd, gn, _ = ix.search_topk(q, 5)
        best = int(gi[0, 0])
        assert ix.remove([best, 999999]) == 1 and ix.count() == 1999
        keep = ids != best
        ri, rd, _ = o.search_exact(codes[keep], ids[keep], dims, bits, szg.EUCLIDEAN, q, k=5)
        gi, gd, gn, scanned = ix.search_topk(q, 5)
        assert scanned == 1999
        assert_results_match(gi[0], gd[0], ri, rd)
        # replace row 5 by the code closest to the query: it becomes the best hit
        newcode = o.encode(q, bits)[None, :]
        ix.upsert([5], newcode)
        codes2 = codes.copy(); codes2[5] = newcode[0]
        ri, rd, _ = o.search_exact(codes2[keep], ids[keep], dims, bits, szg.EUCLIDEAN, q, k=5)
        gi, gd, gn, _ = ix.search_topk(q, 5)
        assert gi[0, 0] == 5 and ix.count() == 1999
        assert_results_match(gi[0], gd[0], ri, rd)
        # re-adding the removed id brings it back (slot reuse)
        ix.upsert([best], codes[best][None, :])
        ri, rd, _ = o.search_exact(codes2, ids, dims, bits, szg.EUCLIDEAN, q, k=5)
        gi, gd, gn, _ = ix.search_topk(q, 5)
        assert_results_match(gi[0], gd[0], ri, rd)
        # duplicate ids inside one batch: the last one wins
        ix.upsert([7, 7], np.stack([codes[100], codes[200]]))
        assert np.array_equal(ix.fetch_codes([7])[0], codes[200])


def test_filter_mask_density_and_radius():
    dims, bits, n = 48, 64, 6000
    codes = o.synth_rows(41, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64)
    q = o.synth_queries(42, 0, 1, dims)[0]
    passmask = (ids % 10 < 3).astype(np.uint8)  # cfg3: bucket < 3 => 30 % density
    with _build(codes, ids, dims, bits, szg.COSINE) as ix:
        m = ix.mask_create(ids, passmask)
        ri, rd, _ = o.search_exact(codes, ids, dims, bits, szg.COSINE, q, radius=0.46, passmask=passmask)
        gi, gd, scanned = ix.search_radius(q, 0.46, mask_id=m)
        assert scanned == n
        assert_radius_match(gi, gd, ri, rd, 0.46, "radius+filter")
        ri, rd, _ = o.search_exact(codes, ids, dims, bits, szg.COSINE, q, k=10, passmask=passmask)
        gi, gd, gn, _ = ix.search_topk(q, 10, mask_id=m)
        assert_results_match(gi[0], gd[0], ri, rd)
        ix.mask_destroy(m)
        with pytest.raises(szg.SzgError):
            ix.search_topk(q, 10, mask_id=m)


def test_radius_overflowing_first_compaction_buffer():
    # a radius that accepts most rows forces the compaction buffer to be re-sized and the scan re-run
    dims, bits, n = 8, 8, 20000
    codes = o.synth_rows(51, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64)
    q = o.synth_queries(52, 0, 1, dims)[0]
    with _build(codes, ids, dims, bits, szg.EUCLIDEAN) as ix:
        ri, rd, _ = o.search_exact(codes, ids, dims, bits, szg.EUCLIDEAN, q, radius=2.6)
        assert ri.size > 4096
        gi, gd, _ = ix.search_radius(q, 2.6)
        assert_radius_match(gi, gd, ri, rd, 2.6, "wide radius")


# ------------------------------------------------------------------ rescoring (LSH candidate path)
@pytest.mark.parametrize("bits", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("metric", [szg.EUCLIDEAN, szg.COSINE])
def test_rescore_matches_oracle_distances(bits, metric):
    dims, n = 45, 3000
    codes = o.synth_rows(61, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64) * 2 + 5
    q = o.synth_queries(62, 0, 1, dims)[0]
    rows = np.random.default_rng(1).integers(0, n, size=777)
    with _build(codes, ids, dims, bits, metric) as ix:
        got = ix.rescore(q, np.concatenate([ids[rows], [4]]).astype(np.uint64))  # id 4 does not exist
        want = o.row_distances(codes, dims, bits, metric, q, rows)
        assert got[-1] == _capi.MISSING_DISTANCE
        if metric == szg.EUCLIDEAN:
            assert np.array_equal(got[:-1], want)  # every op IEEE-exact: bit-identical
        else:
            assert np.allclose(got[:-1], want, rtol=1e-13, atol=0)


def test_lsh_candidate_replay_on_gpu_distances():
    # appendix B-13: given the visit sequence of the (restated) lshTree.search, replaying `consider`
    # over GPU-rescored distances gives the oracle's LSH result
    n, d = 4000, 16
    codes = o.synth_rows(11, 0, n, d, 64)
    ids = np.arange(n, dtype=np.uint64)
    tree = o.LshTree(codes, d, 64, o.COSINE, seed=7)
    tree.add_all_decoded(n)
    q = o.synth_queries(12, 0, 1, d)[0]
    lid, ld, pct, visit = tree.search(ids, q, k=10)
    with _build(codes, ids, d, 64, szg.COSINE) as ix:
        dist = ix.rescore(q, ids[visit])
    # host replay of collection.go:606-619 over the returned distances (strict '>' replacement)
    import heapq
    heap = []  # max-heap via negation; ties: first seen wins
    for seq, (row, dd) in enumerate(zip(visit, dist)):
        if len(heap) < 10:
            heapq.heappush(heap, (-dd, -seq, int(ids[row])))
        elif -heap[0][0] > dd:
            heapq.heapreplace(heap, (-dd, -seq, int(ids[row])))
    got = sorted(((-a, -b, c) for a, b, c in heap))
    assert [g[2] for g in got] == lid.tolist()
    assert np.allclose([g[0] for g in got], ld, rtol=1e-13, atol=0)
    tree.close()


# ------------------------------------------------------------------ BASELINE configs at (near) full size
def test_cfg2_1m_x128_q4_euclid_k10():
    # BASELINE.json configs[1], full size: oracle scan of 1M x 128 takes ~1 s
    n, d, bits = 1_000_000, 128, 4
    codes = o.synth_rows(0x5A590002, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    qs = o.synth_queries(0x5A590003, 0, 3, d)
    with szg.Index(d, bits, szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(0x5A590002, 0, n)
        gi, gd, gn, scanned = ix.search_topk(qs, 10)
        assert scanned == n
        for qi, q in enumerate(qs):
            ri, rd, _ = o.search_exact(codes, ids, d, bits, szg.EUCLIDEAN, q, k=10)
            assert_results_match(gi[qi], gd[qi], ri, rd, _true_dist(codes, ids, d, bits, szg.EUCLIDEAN, q), "cfg2")
        st = ix.stats()
        assert st["uncertain_results"] == 0


def test_cfg1_100k_x384_q8_cosine_k10():
    n, d, bits = 100_000, 384, 8
    codes = o.synth_rows(0x5A590001, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    qs = o.synth_queries(0x5A590011, 0, 4, d)
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590001, 0, n)
        gi, gd, gn, _ = ix.search_topk(qs, 10)
        for qi, q in enumerate(qs):
            ri, rd, _ = o.search_exact(codes, ids, d, bits, szg.COSINE, q, k=10)
            assert_results_match(gi[qi], gd[qi], ri, rd, _true_dist(codes, ids, d, bits, szg.COSINE, q), "cfg1")


def test_cfg4_shape_properties_at_scale():
    """10M x 768 8-bit cosine is too large for the oracle to scan in seconds; check size-independent
    properties instead: (a) every returned distance equals the oracle's distance of that very row,
    (b) no row of a 200k-row oracle-scanned sample beats the k-th result unless it is in the result,
    (c) the result equals the merge of the results of disjoint row shards (associativity, what the
    multi-GPU path relies on), (d) ascending order."""
    d, bits, seed, k = 768, 8, 0x5A590004, 10
    n = 2_000_000  # 1.5 GB: > L2, exercises 32-bit chunk indexing beyond 2^32 bytes? no: see test below
    q = o.synth_queries(seed + 1, 0, 1, d)[0]
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(seed, 0, n)
        gi, gd, gn, _ = ix.search_topk(q, k)
        assert gn[0] == k and np.all(np.diff(gd[0]) >= 0)
        for i, dist in zip(gi[0].tolist(), gd[0].tolist()):
            row = o.synth_rows(seed, i, 1, d, bits)
            want = o.row_distances(row, d, bits, szg.COSINE, q, [0])[0]
            assert abs(dist - want) <= 1e-13 * want
        s0 = 700_000
        sample = o.synth_rows(seed, s0, 200_000, d, bits)
        sids = np.arange(s0, s0 + 200_000, dtype=np.uint64)
        ri, rd, _ = o.search_exact(sample, sids, d, bits, szg.COSINE, q, k=k)
        got = set(gi[0].tolist())
        for i, dist in zip(ri.tolist(), rd.tolist()):
            assert i in got or dist >= gd[0, -1] * (1 - 1e-5)
    # (c) shards
    parts_i, parts_d = [], []
    for r0 in range(0, n, 500_000):
        with szg.Index(d, bits, szg.COSINE) as sh:
            sh.fill_synthetic(seed, r0, 500_000)
            si, sd, sn, _ = sh.search_topk(q, k)
            parts_i.append(si[0]); parts_d.append(sd[0])
    ai, ad = np.concatenate(parts_i), np.concatenate(parts_d)
    order = np.lexsort((ai, ad))[:k]
    assert ai[order].tolist() == gi[0].tolist() and np.array_equal(ad[order], gd[0])


def test_more_than_4gib_of_codes():
    # 6M x 768 bytes = 4.6 GB: chunk addressing must be 64-bit
    d, bits, seed, n = 768, 8, 99, 6_000_000
    q = o.synth_queries(seed + 1, 0, 1, d)[0]
    with szg.Index(d, bits, szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(seed, 0, n)
        tail = np.arange(n - 3, n, dtype=np.uint64)
        assert np.array_equal(ix.fetch_codes(tail), o.synth_rows(seed, n - 3, 3, d, bits))
        # plant the best possible row at the very end of the mirror
        ix.upsert([n + 5], o.encode(q, bits)[None, :])
        gi, gd, gn, _ = ix.search_topk(q, 3)
        assert gi[0, 0] == n + 5


# ------------------------------------------------------------------ device-resident + shard merge entry points
def test_device_resident_search_and_merge():
    import torch
    d, bits, k, nq, seed = 64, 8, 10, 5, 123
    n = 30000
    codes = o.synth_rows(seed, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    qs = o.synth_queries(seed + 1, 0, nq, d)
    dev = torch.device("cuda:0")
    tq = torch.from_numpy(qs).to(dev)
    G = 3
    bounds = [0, 9000, 21000, n]
    g_ids = torch.zeros((G, nq, k), dtype=torch.int64, device=dev)
    g_dist = torch.zeros((G, nq, k), dtype=torch.float64, device=dev)
    g_n = torch.zeros((G, nq), dtype=torch.int32, device=dev)
    shards = []
    stream = torch.cuda.current_stream().cuda_stream
    for g in range(G):
        sh = szg.Index(d, bits, szg.COSINE)
        sh.fill_synthetic(seed, bounds[g], bounds[g + 1] - bounds[g])
        sh.search_topk_dev(tq.data_ptr(), nq, k, g_ids[g].data_ptr(), g_dist[g].data_ptr(), g_n[g].data_ptr(), stream)
        shards.append(sh)
    out_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_dist = torch.zeros((nq, k), dtype=torch.float64, device=dev)
    out_n = torch.zeros(nq, dtype=torch.int32, device=dev)
    shards[0].merge_topk_dev(g_ids.data_ptr(), g_dist.data_ptr(), g_n.data_ptr(), G, nq, k, out_ids.data_ptr(),
                             out_dist.data_ptr(), out_n.data_ptr(), stream)
    torch.cuda.synchronize()
    for qi, q in enumerate(qs):
        ri, rd, _ = o.search_exact(codes, ids, d, bits, szg.COSINE, q, k=k)
        assert out_n[qi].item() == k
        assert_results_match(out_ids[qi].cpu().numpy().astype(np.uint64), out_dist[qi].cpu().numpy(), ri, rd)
    for sh in shards:
        sh.close()


def test_no_fp64_verify_flag_stays_within_tolerance():
    d, bits, n = 128, 8, 50000
    codes = o.synth_rows(71, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    q = o.synth_queries(72, 0, 1, d)[0]
    for metric in (szg.EUCLIDEAN, szg.COSINE):
        with _build(codes, ids, d, bits, metric) as ix:
            gi, gd, gn, _ = ix.search_topk(q, 10, flags=_capi.F_NO_FP64_VERIFY)
            ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, q, k=10)
            assert np.allclose(gd[0], rd, rtol=2e-4)  # surrogate distances: 2-digit fixed-point query


def test_concurrent_searches_on_one_handle():
    # Search holds only the RLock (collection.go:570): concurrent searches must be safe
    import threading
    d, bits, n = 96, 8, 40000
    codes = o.synth_rows(81, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    qs = o.synth_queries(82, 0, 8, d)
    want = [o.search_exact(codes, ids, d, bits, szg.COSINE, q, k=10) for q in qs]
    with _build(codes, ids, d, bits, szg.COSINE) as ix:
        errs = []

        def work(t):
            try:
                for rep in range(5):
                    qi = (t + rep) % len(qs)
                    gi, gd, gn, _ = ix.search_topk(qs[qi], 10)
                    assert_results_match(gi[0], gd[0], want[qi][0], want[qi][1])
                    if t % 2 == 0:  # the batched tensor-core path is a search too: same lock, same safety
                        bi, bd, bn, _ = ix.search_batch(qs, 10)
                        for j in range(len(qs)):
                            assert_results_match(bi[j], bd[j], want[j][0], want[j][1])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs[0]


def test_concurrent_single_query_calls_are_combined_and_identical():
    # SZG_OPT_COMBINE: calls that arrive while a launch is running share the next launch; every caller still gets
    # exactly what a call of its own returns (ids, fp64 distances, counts), also with different k / masks in the mix
    import threading
    d, bits, n = 128, 8, 600000
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(91, 0, n)
        ids = np.arange(n, dtype=np.uint64)
        m = ix.mask_create(ids, (ids % 3 == 0).astype(np.uint8))
        qs = o.synth_queries(92, 0, 48, d)
        ix.set_option(_capi.OPT_COMBINE, 0)
        alone = {(qi, k, mk): ix.search_topk(qs[qi], k, mask_id=mk) for qi in range(48) for k, mk in ((10, -1), (3, m))}
        assert ix.stats()["combined_queries"] == 0
        ix.set_option(_capi.OPT_COMBINE, 1)
        errs = []

        def work(t):
            try:
                for rep in range(40):
                    qi = (t * 7 + rep) % 48
                    k, mk = ((10, -1), (3, m))[(t + rep) % 3 == 0]
                    got = ix.search_topk(qs[qi], k, mask_id=mk)
                    want = alone[(qi, k, mk)]
                    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
                    assert got[3] == n
                with pytest.raises(_capi.SzgError):  # an error reaches the caller that caused it
                    ix.search_topk(qs[0], 10, mask_id=12345)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        for attempt in range(4):  # whether two calls overlap is up to the scheduler: give it a few rounds
            th = [threading.Thread(target=work, args=(t,)) for t in range(12)]
            [t.start() for t in th]
            [t.join() for t in th]
            assert not errs, errs[0]
            if ix.stats()["combined_queries"] > 0:
                break
        assert ix.stats()["combined_queries"] > 0, "no two calls ever shared a launch"


def test_dimension_mismatch_is_a_status():
    # appendix B-12: Search does not validate len(query); the binding must reject it
    with szg.Index(8, 8, szg.COSINE) as ix:
        with pytest.raises(ValueError):
            ix.search_topk(np.zeros(7), 3)
        with pytest.raises(szg.SzgError):
            ix.search_topk(np.zeros(8), 0)
        with pytest.raises(szg.SzgError):
            ix.search_topk(np.zeros(8), 1000)


# ------------------------------------------------------------------ fast (2-digit) / precise (3-digit) surrogate
@pytest.mark.parametrize("digits", [2, 3])
@pytest.mark.parametrize("bits,metric", [(4, szg.EUCLIDEAN), (8, szg.COSINE), (8, szg.EUCLIDEAN), (16, szg.COSINE),
                                         (16, szg.EUCLIDEAN)])
def test_forced_digit_count_matches_oracle(digits, bits, metric):
    n, d, k = 30000, 200, 10
    codes = o.synth_rows(91, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    qs = o.synth_queries(92, 0, 3, d)
    with _build(codes, ids, d, bits, metric) as ix:
        ix.set_option(_capi.OPT_DIGITS, digits)
        gi, gd, gn, _ = ix.search_topk(qs, k)
        for qi, q in enumerate(qs):
            ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, q, k=k)
            assert_results_match(gi[qi], gd[qi], ri, rd, _true_dist(codes, ids, d, bits, metric, q), f"nd={digits}")


@pytest.mark.parametrize("metric", [szg.COSINE, szg.EUCLIDEAN])
def test_near_duplicate_rows_force_escalation(metric):
    """Rows that differ from one another by a single code step: their distances to the query differ far less
    than the 2-digit surrogate's error bound, so the fast pass cannot certify the candidate set; the library
    must notice (rigorous bound in the prepared-query header), re-run precisely and still match the oracle."""
    rng = np.random.default_rng(17)
    n, d, k = 6000, 96, 10
    base = rng.integers(60, 200, size=d, dtype=np.int64)
    codes = np.tile(base, (n, 1))
    for r in range(n):  # one or two coordinates nudged by +-1 code
        j = rng.integers(0, d, size=2)
        codes[r, j] += rng.integers(-1, 2, size=2)
    codes = codes.astype(np.uint8)
    ids = np.arange(n, dtype=np.uint64)
    q = o.decode(base.astype(np.uint8), d, 8) + rng.normal(scale=0.05, size=d)
    with _build(codes, ids, d, 8, metric) as ix:
        gi, gd, gn, _ = ix.search_topk(q, k)
        st = ix.stats()
        ri, rd, _ = o.search_exact(codes, ids, d, 8, metric, q, k=k)
        assert st["escalations"] >= 1, "the 2-digit pass should not have been certified on near-duplicates"
        if not st["uncertain_results"]:
            assert_results_match(gi[0], gd[0], ri, rd, _true_dist(codes, ids, d, 8, metric, q), "near-duplicates")
        else:  # even 256 candidates of the precise surrogate were not separable: distances must still be true ones
            assert np.allclose(gd[0], rd, rtol=1e-5)


def test_device_variant_reports_uncertified_queries():
    import torch
    rng = np.random.default_rng(3)
    n, d, k = 4000, 64, 5
    base = rng.integers(60, 200, size=d, dtype=np.int64)
    codes = np.tile(base, (n, 1)).astype(np.uint8)
    codes[np.arange(n), rng.integers(0, d, size=n)] += 1
    q = o.decode(base.astype(np.uint8), d, 8)
    q2 = o.synth_queries(5, 0, 1, d)[0]
    dev = torch.device("cuda:0")
    tq = torch.from_numpy(np.stack([q + 0.01, q2])).to(dev)
    out_i = torch.zeros((2, k), dtype=torch.int64, device=dev)
    out_d = torch.zeros((2, k), dtype=torch.float64, device=dev)
    out_n = torch.zeros(2, dtype=torch.int32, device=dev)
    out_f = torch.full((2,), 7, dtype=torch.int32, device=dev)
    with szg.Index(d, 8, szg.COSINE) as ix:
        ix.upsert(np.arange(n, dtype=np.uint64), codes)
        ix.search_topk_dev(tq.data_ptr(), 2, k, out_i.data_ptr(), out_d.data_ptr(), out_n.data_ptr(),
                           torch.cuda.current_stream().cuda_stream, d_out_flags=out_f.data_ptr())
        torch.cuda.synchronize()
    assert out_f[0].item() == 1  # thousands of rows within the surrogate error of each other
    assert out_n.tolist() == [k, k]


# ------------------------------------------------------------------ batched queries on the tensor cores (szg_search_batch)
def _batch_vs_oracle(ix, codes, ids, dims, metric, queries, k, mask_id=-1, passmask=None, what=""):
    before = ix.stats()["batch_queries"]
    gi, gd, gn, scanned = ix.search_batch(queries, k, mask_id=mask_id)
    assert ix.stats()["batch_queries"] == before + len(queries), "the tensor-core path did not run"
    flt = None if passmask is None else (lambda i, m: bool(passmask[int(i)]))
    for qi, q in enumerate(queries):
        if flt is None:
            ri, rd, _ = o.search_exact(codes, ids, dims, 8, metric, q, k=k)
        else:
            keep = passmask.astype(bool)
            ri, rd, _ = o.search_exact(codes[keep], ids[keep], dims, 8, metric, q, k=k)
        if np.isnan(rd).any():
            continue
        assert gn[qi] == ri.size, f"{what} q{qi}: {gn[qi]} results, oracle {ri.size}"
        assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd, _true_dist(codes, ids, dims, 8, metric, q),
                             f"{what} q{qi}")
    return gi, gd, gn, scanned


@pytest.mark.parametrize("metric", [szg.COSINE, szg.EUCLIDEAN])
@pytest.mark.parametrize("dims,n,nq,k", [(768, 3000, 70, 10), (128, 9000, 130, 10), (384, 5001, 64, 50), (96, 4100, 3, 100),
                                         (32, 700, 65, 1)])
def test_batch_matches_oracle(metric, dims, n, nq, k):
    """szg_search_batch (tcgen05 contraction + fused top-k) returns what nq single Search calls return."""
    seed = 400 + dims + k
    codes = o.synth_rows(seed, 0, n, dims, 8)
    ids = np.arange(n, dtype=np.uint64) * 7 + 3
    queries = o.synth_queries(seed + 1, 0, nq, dims)
    with _build(codes, ids, dims, 8, metric) as ix:
        gi, gd, gn, scanned = _batch_vs_oracle(ix, codes, ids, dims, metric, queries, k, what=f"batch m{metric} d{dims} k{k}")
        assert scanned == n
        # and bit-identical to the streaming scan (same candidates, same fp64 re-score)
        si, sd, sn, _ = ix.search_topk(queries, k)
        assert np.array_equal(gn, sn) and np.array_equal(gi, si) and np.array_equal(gd, sd)


def test_batch_with_filter_mask_and_tombstones():
    n, dims, nq, k = 6000, 256, 96, 10
    codes = o.synth_rows(909, 0, n, dims, 8)
    ids = np.arange(n, dtype=np.uint64)
    queries = o.synth_queries(910, 0, nq, dims)
    with _build(codes, ids, dims, 8, szg.COSINE) as ix:
        dead = ids[(ids % 5 == 1)]
        ix.remove(dead)
        alive = (ids % 5 != 1)
        passmask = ((ids % 10 < 3) & alive).astype(np.uint8)   # 30 % density (cfg3's filter), minus tombstones
        mask = ix.mask_create(ids[alive], passmask[alive])
        gi, gd, gn, scanned = _batch_vs_oracle(ix, codes, ids, dims, szg.COSINE, queries, k, mask_id=mask, passmask=passmask,
                                               what="batch filtered")
        assert scanned == int(alive.sum())  # filtered rows count as searched, removed ones do not
        # a filter that leaves whole 128-row tiles empty and fewer than k rows in total
        few = np.zeros(n, dtype=np.uint8)
        few[[10, 4000, 4001, 5999]] = 1
        few &= alive.astype(np.uint8)
        mask2 = ix.mask_create(ids[alive], few[alive])
        gi, gd, gn, _ = ix.search_batch(queries, k, mask_id=mask2)
        assert np.all(gn == int(few.sum()))
        assert set(gi[0, :gn[0]].tolist()) == set(ids[few.astype(bool)].tolist())


def test_batch_zero_and_degenerate_queries():
    n, dims = 2500, 64
    codes = o.synth_rows(31, 0, n, dims, 8)
    codes[100] = 128  # decodes to +0.0039..., fine; a true zero-norm row cannot exist with 8-bit codes
    ids = np.arange(n, dtype=np.uint64)
    queries = o.synth_queries(32, 0, 66, dims)
    queries[5] = 0.0                      # zero query: every cosine distance is exactly 1.0 (collection.go:828-830)
    queries[6] = queries[7]               # duplicate queries in one batch
    queries[8] *= 1e-9                    # tiny norm
    for metric in (szg.COSINE, szg.EUCLIDEAN):
        with _build(codes, ids, dims, 8, metric) as ix:
            gi, gd, gn, _ = _batch_vs_oracle(ix, codes, ids, dims, metric, queries, 10, what=f"batch degenerate m{metric}")
            assert np.array_equal(gi[6], gi[7]) and np.array_equal(gd[6], gd[7])
            if metric == szg.COSINE:
                assert np.all(gd[5] == 1.0)  # 2500-way exact tie: any members (north_star tie rule)


def test_batch_falls_back_to_the_scan_when_the_geometry_does_not_fit():
    """32/64-bit collections, odd chunk counts and k beyond the list sizes are served by the streaming scan, on the GPU."""
    for bits, dims, k in [(4, 40, 10), (16, 40, 10), (32, 48, 10), (64, 32, 10), (8, 40, 10), (8, 64, 200)]:
        n = 1500
        codes = o.synth_rows(5, 0, n, dims, bits)
        ids = np.arange(n, dtype=np.uint64)
        queries = o.synth_queries(6, 0, 5, dims)
        with _build(codes, ids, dims, bits, szg.EUCLIDEAN) as ix:
            gi, gd, gn, _ = ix.search_batch(queries, k)
            assert ix.stats()["batch_queries"] == 0
            si, sd, sn, _ = ix.search_topk(queries, k)
            assert np.array_equal(gi, si) and np.array_equal(gd, sd) and np.array_equal(gn, sn)


def test_batch_near_duplicate_rows_escalate_like_the_scan():
    """Rows closer to each other than the 2-digit surrogate can resolve: the batch path must flag and re-run them."""
    n, dims = 4000, 128
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, size=dims, dtype=np.uint8)
    codes = np.tile(base, (n, 1))
    flip = rng.integers(0, dims, size=n)
    codes[np.arange(n), flip] ^= 1  # every row differs from the base in one least-significant bit
    ids = np.arange(n, dtype=np.uint64)
    queries = o.synth_queries(44, 0, 64, dims)
    with _build(codes, ids, dims, 8, szg.COSINE) as ix:
        _batch_vs_oracle(ix, codes, ids, dims, szg.COSINE, queries[:8], 10, what="batch near-dup")


def test_batch_cfg5_shape_against_streaming_scan():
    """1024 queries, k = 100 (BASELINE.json configs[4] shape, 8-bit, one shard-sized slice): identical to the scan."""
    rows, dims, nq, k = 200000, 768, 1024, 100
    qs = np.random.default_rng(11).uniform(-1, 1, size=(nq, dims))
    with szg.Index(dims, 8, szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(0x5A590005, 0, rows)
        bi, bd, bn, _ = ix.search_batch(qs, k)
        assert ix.stats()["batch_queries"] == nq
        si, sd, sn, _ = ix.search_topk(qs, k)
        assert np.array_equal(bn, sn) and np.array_equal(bi, si) and np.array_equal(bd, sd)
        assert np.all(np.diff(bd, axis=1) >= 0)


def test_batch_gaussian_normalised_rows_cosine():
    """all-MiniLM-like data (L2-normalised Gaussian rows use only ~+-14 codes around 128): candidate margins are far smaller
    than with uniform codes; the batch must certify or escalate exactly like the scan."""
    n, dims, nq, k = 20000, 384, 80, 10
    rng = np.random.default_rng(21)
    x = rng.normal(size=(n, dims))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    codes = o.encode_rows(x, 8)
    ids = np.arange(n, dtype=np.uint64)
    q = rng.normal(size=(nq, dims))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    with _build(codes, ids, dims, 8, szg.COSINE) as ix:
        _batch_vs_oracle(ix, codes, ids, dims, szg.COSINE, q[:6], k, what="batch gaussian")
        bi, bd, bn, _ = ix.search_batch(q, k)
        si, sd, sn, _ = ix.search_topk(q, k)
        assert np.array_equal(bi, si) and np.array_equal(bd, sd) and np.array_equal(bn, sn)


# ------------------------------------------------------------------ 16-bit collections on the tensor cores (byte planes)
@pytest.mark.parametrize("metric", [szg.COSINE, szg.EUCLIDEAN])
@pytest.mark.parametrize("dims,n,nq,k", [(768, 3000, 70, 10), (96, 9000, 130, 100), (24, 5001, 64, 30), (32, 700, 3, 1)])
def test_batch_16bit_matches_oracle(metric, dims, n, nq, k):
    """16-bit rows: the batch is contracted as a high-byte and a low-byte plane of the uncentred codes (a secondary
    byte-planar copy in HBM); results are what nq single Search calls return."""
    seed = 900 + dims + k
    codes = o.synth_rows(seed, 0, n, dims, 16)
    ids = np.arange(n, dtype=np.uint64) * 5 + 2
    queries = o.synth_queries(seed + 1, 0, nq, dims)
    with _build(codes, ids, dims, 16, metric) as ix:
        gi, gd, gn, scanned = ix.search_batch(queries, k)
        assert ix.stats()["batch_queries"] == nq, "the tensor-core path did not run"
        assert scanned == n
        for qi in range(min(nq, 6)):
            ri, rd, _ = o.search_exact(codes, ids, dims, 16, metric, queries[qi], k=k)
            assert gn[qi] == ri.size
            assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd, _true_dist(codes, ids, dims, 16, metric, queries[qi]),
                                 f"batch16 m{metric} d{dims} k{k} q{qi}")
        si, sd, sn, _ = ix.search_topk(queries, k)
        assert np.array_equal(gn, sn) and np.array_equal(gi, si) and np.array_equal(gd, sd)


def test_batch_16bit_follows_mutations_and_masks():
    """The byte-planar copy is rebuilt lazily after upsert / fill; removals and filter masks act through the live words."""
    n, dims, nq, k = 4000, 64, 66, 10
    codes = o.synth_rows(61, 0, n, dims, 16)
    ids = np.arange(n, dtype=np.uint64)
    queries = o.synth_queries(62, 0, nq, dims)
    with _build(codes[:3000], ids[:3000], dims, 16, szg.EUCLIDEAN) as ix:
        a = ix.search_batch(queries, k)
        ix.upsert(ids[3000:], codes[3000:])                    # new rows: the copy must be rebuilt
        newc = o.synth_rows(63, 0, 10, dims, 16)
        ix.upsert(ids[:10], newc)                              # replaced rows
        ix.remove(ids[100:200])
        cur = codes.copy(); cur[:10] = newc
        alive = np.ones(n, dtype=bool); alive[100:200] = False
        passmask = ((ids % 4 != 0) & alive).astype(np.uint8)
        mask = ix.mask_create(ids[alive], passmask[alive])
        gi, gd, gn, scanned = ix.search_batch(queries, k, mask_id=mask)
        assert ix.stats()["batch_queries"] == 2 * nq and scanned == int(alive.sum())
        keep = passmask.astype(bool)
        for qi in range(5):
            ri, rd, _ = o.search_exact(cur[keep], ids[keep], dims, 16, szg.EUCLIDEAN, queries[qi], k=k)
            assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd, None, f"batch16 mutated q{qi}")
        si, sd, sn, _ = ix.search_topk(queries, k, mask_id=mask)
        assert np.array_equal(gi, si) and np.array_equal(gd, sd)


def test_batch_cfg5_shape_16bit_euclid_k100():
    """BASELINE.json configs[4] shape on one shard-sized slice: 1024 queries, k = 100, 16-bit, euclidean."""
    rows, dims, nq, k = 150000, 768, 1024, 100
    qs = np.random.default_rng(12).uniform(-1, 1, size=(nq, dims))
    with szg.Index(dims, 16, szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(0x5A590005, 0, rows)
        bi, bd, bn, _ = ix.search_batch(qs, k)
        assert ix.stats()["batch_queries"] == nq
        si, sd, sn, _ = ix.search_topk(qs, k)
        assert np.array_equal(bn, sn) and np.array_equal(bi, si) and np.array_equal(bd, sd)


@pytest.mark.parametrize("metric", [szg.COSINE, szg.EUCLIDEAN])
@pytest.mark.parametrize("dims,n,nq,k", [(128, 9000, 130, 10), (31, 3000, 70, 50), (768, 2500, 64, 10)])
def test_batch_4bit_matches_oracle(metric, dims, n, nq, k):
    """4-bit rows go to the tensor cores through a one-byte-per-code copy (tcgen05 has no 4-bit integer kind); odd
    dimension counts leave the last low nibble unused (collection.go:774-779)."""
    seed = 1300 + dims + k
    codes = o.synth_rows(seed, 0, n, dims, 4)
    ids = np.arange(n, dtype=np.uint64) * 3 + 7
    queries = o.synth_queries(seed + 1, 0, nq, dims)
    with _build(codes, ids, dims, 4, metric) as ix:
        gi, gd, gn, scanned = ix.search_batch(queries, k)
        assert ix.stats()["batch_queries"] == nq, "the tensor-core path did not run"
        for qi in range(min(nq, 6)):
            ri, rd, _ = o.search_exact(codes, ids, dims, 4, metric, queries[qi], k=k)
            assert gn[qi] == ri.size
            assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd, _true_dist(codes, ids, dims, 4, metric, queries[qi]),
                                 f"batch4 m{metric} d{dims} k{k} q{qi}")
        si, sd, sn, _ = ix.search_topk(queries, k)
        assert np.array_equal(gn, sn) and np.array_equal(gi, si) and np.array_equal(gd, sd)


@pytest.mark.parametrize("bits", [4, 8, 16])
def test_batch_tiny_and_emptied_collections(bits):
    """Fewer rows than one 128-row tile, k larger than the collection, everything removed, nothing ever added."""
    dims, nq = 32, 5
    queries = o.synth_queries(71, 0, nq, dims)
    for n in (1, 5, 127, 129):
        codes = o.synth_rows(70 + n, 0, n, dims, bits)
        ids = np.arange(n, dtype=np.uint64) + 100
        with _build(codes, ids, dims, bits, szg.EUCLIDEAN) as ix:
            gi, gd, gn, scanned = ix.search_batch(queries, 10)
            assert scanned == n and np.all(gn == min(n, 10))
            for qi in range(nq):
                ri, rd, _ = o.search_exact(codes, ids, dims, bits, szg.EUCLIDEAN, queries[qi], k=10)
                assert_results_match(gi[qi, :gn[qi]], gd[qi, :gn[qi]], ri, rd, None, f"tiny b{bits} n{n} q{qi}")
            if n == 129:
                ix.remove(ids)                       # emptied: K > N_pass returns nothing, PercentSearched 0 (collection.go:706-709)
                gi, gd, gn, scanned = ix.search_batch(queries, 10)
                assert scanned == 0 and np.all(gn == 0)
    with szg.Index(dims, bits, szg.COSINE) as ix:   # never filled
        gi, gd, gn, scanned = ix.search_batch(queries, 3)
        assert scanned == 0 and np.all(gn == 0)


# ------------------------------------------------------------------ scan_small.cuh (rows up to 48 chunks, k <= 24)
@pytest.mark.parametrize("bits,dims", [(4, 64), (4, 128), (8, 64), (8, 128), (4, 384), (16, 96), (8, 256), (16, 128), (8, 384),
                                       (4, 768), (16, 256), (8, 512), (8, 768), (16, 384), (4, 1536)])
def test_short_row_kernel_matches_general_kernel_and_oracle(bits, dims):
    # chunk counts 2, 4, 8, 12, 16, 24, 32, 48; query counts that give one part per SM (1 query), several parts per query and
    # several queries per CTA; a filter mask, tombstones and a ragged last block.  The general kernel is forced by setting a
    # scan geometry option (the short-row kernel only runs with the automatic geometry).
    metric = szg.COSINE if (bits + dims) % 3 else szg.EUCLIDEAN
    n = 4000 + dims % 37
    codes = o.synth_rows(300 + bits + dims, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64) * 3
    keep = np.ones(n, dtype=bool)
    keep[5:n:11] = False
    passmask = (np.arange(n) % 4 != 1).astype(np.uint8)
    with szg.Index(dims, bits, metric) as ix, szg.Index(dims, bits, metric) as gx:
        for x in (ix, gx):
            x.upsert(ids, codes)
            x.remove(ids[~keep])
        gx.set_option(_capi.OPT_SCAN_WARPS, 16)  # general kernel
        m1, m2 = ix.mask_create(ids, passmask), gx.mask_create(ids, passmask)
        for nq, k, masked in ((1, 10, False), (3, 24, True), (37, 10, False), (150, 5, True), (300, 1, False)):
            qs = o.synth_queries(900 + nq, 0, nq, dims)
            a = ix.search_topk(qs, k, mask_id=m1 if masked else -1)
            b = gx.search_topk(qs, k, mask_id=m2 if masked else -1)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
            live = keep & (passmask.astype(bool) if masked else True)
            for qi in (0, nq - 1):
                ri, rd, _ = o.search_exact(codes[keep], ids[keep], dims, bits, metric, qs[qi], k=k,
                                           passmask=passmask[keep] if masked else None)
                assert_results_match(a[0][qi, :a[2][qi]], a[1][qi, :a[2][qi]], ri, rd, what=f"short rows b{bits} d{dims} nq{nq}")
                assert a[2][qi] == min(k, int(live.sum()))


@pytest.mark.parametrize("metric", [szg.COSINE, szg.EUCLIDEAN])
@pytest.mark.parametrize("nq,k", [(20, 10), (20, 40), (5, 100), (64, 100), (130, 60)])
def test_batch_every_sm_has_a_range_and_lists_are_wider_than_a_warp(metric, nq, k):
    """Enough rows that a query group is spread over all 148 CTAs (row tiles dealt on demand), with candidate lists of 32 / 64 /
    128 keys: the shared bound is then built from 32 / 64 / 128 groups of published keys (more groups than lanes of the polling
    warp for the wider lists).  Results are what nq single Search calls return, bit for bit."""
    dims, n = 64, 40_000
    seed = 7100 + nq + k
    codes = o.synth_rows(seed, 0, n, dims, 8)
    ids = np.arange(n, dtype=np.uint64) * 3 + 1
    queries = o.synth_queries(seed + 1, 0, nq, dims)
    with _build(codes, ids, dims, 8, metric) as ix:
        gi, gd, gn, scanned = _batch_vs_oracle(ix, codes, ids, dims, metric, queries, k, what=f"148 ranges m{metric} nq{nq} k{k}")
        assert scanned == n
        ix.set_option(_capi.OPT_BATCH_TENSOR, 0)
        si, sd, sn, _ = ix.search_topk(queries, k)
        assert np.array_equal(gn, sn) and np.array_equal(gi, si) and np.array_equal(gd, sd)


def test_batch_sparse_filter_over_all_ranges():
    """A 1 % filter over a collection that fills all 148 ranges: most ranges cannot seed the shared bound from their first tile
    (its best row is filtered out) and publish "no bound"; the others' keys must still never cut a passing row."""
    dims, n, nq, k = 64, 40_000, 24, 10
    codes = o.synth_rows(7301, 0, n, dims, 8)
    ids = np.arange(n, dtype=np.uint64)
    queries = o.synth_queries(7302, 0, nq, dims)
    rng = np.random.default_rng(7303)
    with _build(codes, ids, dims, 8, szg.COSINE) as ix:
        for density in (0.01, 0.0005):
            passmask = (rng.random(n) < density).astype(np.uint8)
            mask = ix.mask_create(ids, passmask)
            _batch_vs_oracle(ix, codes, ids, dims, szg.COSINE, queries, k, mask_id=mask, passmask=passmask,
                             what=f"sparse filter {density}")
