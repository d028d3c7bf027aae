"""GPU parity tests, second set: the BASELINE.json configurations at their stated sizes against the oracle, the
multi-device handle (szg_create_sharded), dispatch by batch size, captured launch sequences, radius search ordered on the
device, batched re-scoring, and the reference-held end-to-end cases (rest_test.go:503-569).

Same bar as tests/test_gpu_parity.py: ids and order identical to the oracle's scan except among distances within 1e-5
relative; returned fp64 distances bit-identical to the oracle's.
"""
import json
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import syzgydb_b200 as szg
from oracle import pyoracle as o
from syzgydb_b200 import _capi
from syzgydb_b200 import filter as szf
from tests.common import assert_radius_match, assert_results_match

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _true_dist_synth(seed, d, bits, metric, q):
    def f(i):
        row = o.synth_rows(seed, int(i), 1, d, bits)
        return float(o.row_distances(row, d, bits, metric, q, [0])[0])
    return f


def E(op, left, right):
    return ("expr", op, left, right)


def I(name):  # noqa: E743
    return ("ident", name)


def V(v):
    return ("value", v)


def _upsert_meta(ix, ids, raws, fields):
    """What the shim does at AddDocument time: one typed value per mirrored field (syzgydb_b200/filter.py)."""
    kinds, vals = zip(*[szf.column_values(r, fields) for r in raws])
    ix.meta_upsert(ids, kinds, list(range(len(fields))), vals)
    return {f: i for i, f in enumerate(fields)}


def _device_lists():
    """Shard layouts every box can run (several shards on GPU 0) plus real multi-GPU ones when the box has them."""
    out = [[0, 0], [0, 0, 0]]
    n = _ngpu()
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    if n >= 8:
        out.append(list(range(8)))
    return out


# ------------------------------------------------------------------ dispatch by batch size, captured launch sequences
def test_topk_call_with_many_queries_takes_the_tensor_path_and_equals_the_scan():
    d, bits, n, k = 768, 8, 60000, 10
    qs = o.synth_queries(7, 0, 32, d)
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(6, 0, n)
        b0 = ix.stats()["batch_queries"]
        gi, gd, gn, _ = ix.search_topk(qs, k)                 # 32 queries >= SZG_OPT_BATCH_MIN_QUERIES: one contraction
        assert ix.stats()["batch_queries"] - b0 == 32
        ix.set_option(_capi.OPT_BATCH_MIN_QUERIES, 4096)      # the same call as 32 streaming scans
        si, sd, sn, _ = ix.search_topk(qs, k)
        assert ix.stats()["batch_queries"] - b0 == 32
        assert np.array_equal(gi, si) and np.array_equal(gd, sd) and np.array_equal(gn, sn)
        codes = o.synth_rows(6, 0, n, d, bits)
        ids = np.arange(n, dtype=np.uint64)
        for qi in (0, 13, 31):
            ri, rd, _ = o.search_exact(codes, ids, d, bits, szg.COSINE, qs[qi], k=k)
            assert_results_match(gi[qi], gd[qi], ri, rd)
        # below the threshold nothing goes to the tensor cores
        ix.set_option(_capi.OPT_BATCH_MIN_QUERIES, 4)
        ix.set_option(_capi.OPT_COMBINE, 0)
        b1 = ix.stats()["batch_queries"]
        ti, td, tn, _ = ix.search_topk(qs[:3], k)
        assert ix.stats()["batch_queries"] == b1
        assert np.array_equal(ti, gi[:3]) and np.array_equal(td, gd[:3])


@pytest.mark.parametrize("bits,d,nq", [(8, 768, 1), (8, 768, 32), (4, 128, 1), (64, 48, 2), (16, 96, 40)])
def test_repeated_call_shapes_are_replayed_as_graphs_with_identical_results(bits, d, nq):
    n, k = 30000, 10
    metric = szg.COSINE if bits != 4 else szg.EUCLIDEAN
    with szg.Index(d, bits, metric) as ix:
        ix.fill_synthetic(17, 0, n)
        ix.set_option(_capi.OPT_COMBINE, 0)
        ix.set_option(_capi.OPT_GRAPHS, 0)
        qsets = [o.synth_queries(100 + r, 0, nq, d) for r in range(5)]
        plain = [ix.search_topk(q, k) for q in qsets]
        assert ix.stats()["graph_launches"] == 0
        ix.set_option(_capi.OPT_GRAPHS, 1)
        for rnd in range(2):
            for q, want in zip(qsets, plain):
                got = ix.search_topk(q, k)
                assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        assert ix.stats()["graph_launches"] >= 6, ix.stats()
        # a mutation invalidates the captured sequence: the new row must be found
        target = qsets[0][0]
        new_id = np.array([10 ** 9], dtype=np.uint64)
        if bits <= 16:
            code = o.encode(np.clip(target / np.abs(target).max(), -1, 1), bits)[None, :]
        else:  # not exactly parallel: a cosine ratio rounded above 1 would be NaN (Go's Acos) and never returned
            code = o.encode(target + 1e-3 * np.random.default_rng(1).normal(size=d), bits)[None, :]
        ix.upsert(new_id, code)
        for _ in range(3):
            gi, gd, gn, _ = ix.search_topk(qsets[0], k)
            assert gi[0, 0] == 10 ** 9
        ix.remove(new_id)
        for _ in range(3):
            got = ix.search_topk(qsets[0], k)
            assert np.array_equal(got[0], plain[0][0]) and np.array_equal(got[1], plain[0][1])


@pytest.mark.parametrize("bits,d,k,metric", [(64, 40, 10, szg.COSINE), (8, 40, 10, szg.COSINE), (8, 96, 150, szg.EUCLIDEAN),
                                             (32, 24, 10, szg.EUCLIDEAN)])
def test_concurrent_callers_on_shapes_the_tensor_path_rejects(bits, d, k, metric):
    """a combined batch of 4..16 queries on a collection the tensor-core path cannot serve (float rows,
    an odd number of 16-byte chunks, k above the candidate lists) must fall back to the scan, not queue behind itself."""
    n = 20000
    codes = o.synth_rows(81, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    qs = o.synth_queries(82, 0, 12, d)
    want = [o.search_exact(codes, ids, d, bits, metric, q, k=k) for q in qs]
    with szg.Index(d, bits, metric) as ix:
        ix.upsert(ids, codes)
        errs = []

        def work(t):
            try:
                for rep in range(12):
                    qi = (t + rep) % len(qs)
                    gi, gd, gn, _ = ix.search_topk(qs[qi], k)
                    if not np.isnan(want[qi][1]).any():
                        assert_results_match(gi[0, :gn[0]], gd[0, :gn[0]], want[qi][0], want[qi][1])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(t,), daemon=True) for t in range(14)]
        [t.start() for t in th]
        for t in th:
            t.join(timeout=120)
            assert not t.is_alive(), "a caller never returned (combining leader deadlock)"
        assert not errs, errs[0]


def test_masks_can_be_built_while_searches_run():
    """Search applies its filter under the RLock, so szg_mask_create / szg_filter_mask / szg_mask_destroy
    run next to other searches on the same handle."""
    d, bits, n = 64, 8, 50000
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(5, 0, n)
        ids = np.arange(n, dtype=np.uint64)
        codes = o.synth_rows(5, 0, n, d, bits)
        qs = o.synth_queries(6, 0, 6, d)
        errs = []

        def work(t):
            try:
                for rep in range(8):
                    mod = 2 + (t + rep) % 5
                    pm = (ids % mod == 0).astype(np.uint8)
                    m = ix.mask_create(ids, pm)
                    q = qs[(t + rep) % len(qs)]
                    gi, gd, gn, _ = ix.search_topk(q, 5, mask_id=m)
                    ri, rd, _ = o.search_exact(codes, ids, d, bits, szg.COSINE, q, k=5, passmask=pm)
                    assert_results_match(gi[0, :gn[0]], gd[0, :gn[0]], ri, rd)
                    ix.mask_destroy(m)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs[0]


# ------------------------------------------------------------------ radius on the device, batched; batched re-scoring
@pytest.mark.parametrize("bits,d,metric,radius,n", [(8, 8, szg.EUCLIDEAN, 2.6, 60000),      # > 2048 hits: the multi-pass sort
                                                    (64, 48, szg.COSINE, 0.46, 20000),      # cfg3's shape, a few dozen hits
                                                    (4, 64, szg.EUCLIDEAN, 6.1, 30000),
                                                    (16, 24, szg.COSINE, 0.40, 30000),
                                                    (32, 16, szg.EUCLIDEAN, 2.0, 30000)])
def test_radius_results_are_filtered_and_ordered_on_the_device(bits, d, metric, radius, n):
    codes = o.synth_rows(51, 0, n, d, bits)
    ids = (np.arange(n, dtype=np.uint64) * 7 + 3)
    qs = o.synth_queries(52, 0, 3, d)
    with szg.Index(d, bits, metric) as ix:
        ix.upsert(ids, codes)
        for q in qs:
            ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, q, radius=radius)
            gi, gd, scanned = ix.search_radius(q, radius)
            assert scanned == n
            assert_radius_match(gi, gd, ri, rd, radius, f"radius b{bits}")
            # exact (distance, lexicographic id) order, as the oracle's drain produces it
            if not np.isnan(rd).any() and gi.size == ri.size:
                assert np.array_equal(gd, rd)
        # the same queries as one batched call, each with its own radius
        radii = [radius, radius * 0.97, radius * 1.02]
        res, scanned = ix.search_radius_batch(qs, radii)
        for q, r, (gi, gd) in zip(qs, radii, res):
            ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, q, radius=r)
            assert_radius_match(gi, gd, ri, rd, r, f"radius batch b{bits}")


def test_radius_ties_are_ordered_by_lexicographic_id():
    d, n = 8, 5000
    base = o.synth_rows(3, 0, 50, d, 8)
    codes = np.repeat(base, n // 50, axis=0)                      # every row 100 times: big tie groups
    ids = np.random.default_rng(0).permutation(np.arange(1, n + 1, dtype=np.uint64))
    q = o.synth_queries(4, 0, 1, d)[0]
    with szg.Index(d, 8, szg.EUCLIDEAN) as ix:
        ix.upsert(ids, codes)
        ri, rd, _ = o.search_exact(codes, ids, d, 8, szg.EUCLIDEAN, q, radius=2.3)
        gi, gd, _ = ix.search_radius(q, 2.3)
        assert gi.size == ri.size and gi.size > 200
        assert np.array_equal(gd, rd)
        # within a tie group the oracle's heap drain is not the scan order; the device orders ties by lexicographic id
        for dist in np.unique(gd):
            grp = gi[gd == dist]
            assert sorted(grp.tolist(), key=str) == grp.tolist()
            assert set(grp.tolist()) == set(ri[rd == dist].tolist())


@pytest.mark.parametrize("bits,metric", [(64, szg.COSINE), (32, szg.EUCLIDEAN), (8, szg.COSINE), (4, szg.EUCLIDEAN), (16, szg.EUCLIDEAN)])
def test_rescore_batch_matches_oracle(bits, metric):
    d, n = 100, 9000
    codes = o.synth_rows(61, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64) * 3 + 11
    qs = o.synth_queries(62, 0, 5, d)
    rng = np.random.default_rng(2)
    lists = [rng.integers(0, n, size=s) for s in (0, 1, 37, 200, 1500)]
    with szg.Index(d, bits, metric) as ix:
        ix.upsert(ids, codes)
        got = ix.rescore_batch(qs, [ids[r] for r in lists])
        for q, rows, g in zip(qs, lists, got):
            want = o.row_distances(codes, d, bits, metric, q, rows)
            assert np.array_equal(g, want, equal_nan=True)  # bit-identical, both metrics (math.Acos restated on both sides)
        big = rng.integers(0, n, size=40000)                  # the wide-CTA variant
        g = ix.rescore(qs[0], np.concatenate([ids[big], [5]]).astype(np.uint64))
        assert g[-1] == _capi.MISSING_DISTANCE
        assert np.array_equal(g[:-1], o.row_distances(codes, d, bits, metric, qs[0], big), equal_nan=True)


# ------------------------------------------------------------------ one handle over several devices
@pytest.mark.parametrize("devices", _device_lists() if True else [], ids=lambda v: "dev" + "".join(map(str, v)))
@pytest.mark.parametrize("bits,d,metric", [(8, 768, szg.COSINE), (64, 48, szg.COSINE), (4, 128, szg.EUCLIDEAN), (16, 96, szg.EUCLIDEAN)])
def test_sharded_handle_returns_what_one_device_returns(devices, bits, d, metric):
    """results independent of the number of devices -- top-k (scan and tensor path), radius,
    re-scoring, masks, mutations, against a single-device handle (bit for bit) and the oracle."""
    n, k = 24000, 10
    codes = o.synth_rows(21, 0, n, d, bits)
    ids = np.random.default_rng(5).permutation(np.arange(1, 4 * n, 4, dtype=np.uint64))[:n]
    qs = o.synth_queries(22, 0, 20, d)
    pm = (ids % 10 < 3).astype(np.uint8)
    with szg.Index(d, bits, metric) as one, szg.Index(d, bits, metric, devices=devices) as sh:
        one.upsert(ids, codes)
        sh.upsert(ids, codes)
        assert sh.count() == n and sh.stats()["shards"] == len(devices)
        assert np.array_equal(sh.fetch_codes(ids[:500]), codes[:500])
        for nq in (1, 3, 20):                                   # scan for 1-3 queries, tensor path for 20 (quantized rows)
            a = one.search_topk(qs[:nq], k)
            b = sh.search_topk(qs[:nq], k)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and b[3] == n
        ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, qs[0], k=k)
        gi, gd, gn, _ = sh.search_topk(qs[0], k)
        if not np.isnan(rd).any():
            assert_results_match(gi[0, :gn[0]], gd[0, :gn[0]], ri, rd)
        # batched entry point, larger k
        a = one.search_batch(qs, 50)
        b = sh.search_batch(qs, 50)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        # masks
        m1, ms = one.mask_create(ids, pm), sh.mask_create(ids, pm)
        a = one.search_topk(qs[:4], k, mask_id=m1)
        b = sh.search_topk(qs[:4], k, mask_id=ms)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, qs[1], k=k, passmask=pm)
        assert_results_match(b[0][1, :b[2][1]], b[1][1, :b[2][1]], ri, rd)
        # radius (sharded: per-device ordered lists merged by (distance, lexicographic id))
        radius = {(8, szg.COSINE): 0.47, (64, szg.COSINE): 0.45, (4, szg.EUCLIDEAN): 8.6, (16, szg.EUCLIDEAN): 7.4}[(bits, metric)]
        ai, ad, _ = one.search_radius(qs[2], radius, mask_id=m1)
        bi, bd, scanned = sh.search_radius(qs[2], radius, mask_id=ms)
        assert scanned == n and np.array_equal(ai, bi) and np.array_equal(ad, bd)
        ri, rd, _ = o.search_exact(codes, ids, d, bits, metric, qs[2], radius=radius, passmask=pm)
        assert ri.size > 0
        assert_radius_match(bi, bd, ri, rd, radius, "sharded radius")
        # re-scoring: ids are routed to their owners, distances come back in visit order
        rows = np.random.default_rng(1).integers(0, n, size=3000)
        want = o.row_distances(codes, d, bits, metric, qs[3], rows)
        got = sh.rescore(qs[3], np.concatenate([ids[rows], [2]]).astype(np.uint64))
        assert got[-1] == _capi.MISSING_DISTANCE and np.array_equal(got[:-1], want, equal_nan=True)
        # mutations: replace and remove
        sh.upsert(ids[:100], codes[100:200])
        one.upsert(ids[:100], codes[100:200])
        assert sh.remove(ids[200:260]) == 60 and one.remove(ids[200:260]) == 60 and sh.count() == n - 60
        a = one.search_topk(qs[:5], k)
        b = sh.search_topk(qs[:5], k)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("devices", _device_lists(), ids=lambda v: "dev" + "".join(map(str, v)))
def test_sharded_handle_device_filters_and_synthetic_ranges(devices):
    d, bits, n = 64, 8, 30000
    with szg.Index(d, bits, szg.COSINE, devices=devices) as sh:
        sh.fill_synthetic(31, 0, n)                              # cut in contiguous pieces, one per device
        assert sh.count() == n
        ids = np.arange(n, dtype=np.uint64)
        codes = o.synth_rows(31, 0, n, d, bits)
        assert np.array_equal(sh.fetch_codes(ids[::997]), codes[::997])
        docs = [{"bucket": int(i % 10), "tag": "t%d" % (i % 7)} for i in range(n)]
        cols = _upsert_meta(sh, ids, [json.dumps(x).encode() for x in docs], ["bucket", "tag"])
        prog = szf.lower(E("AND", E("<", I("bucket"), V(3.0)), E("==", I("tag"), V("t2"))), cols)
        m = sh.filter_mask(prog)
        pm = np.array([x["bucket"] < 3 and x["tag"] == "t2" for x in docs], dtype=np.uint8)
        q = o.synth_queries(32, 0, 1, d)[0]
        ri, rd, _ = o.search_exact(codes, ids, d, bits, szg.COSINE, q, k=10, passmask=pm)
        gi, gd, gn, scanned = sh.search_topk(q, 10, mask_id=m)
        assert scanned == n
        assert_results_match(gi[0, :gn[0]], gd[0, :gn[0]], ri, rd)
        # a row of the synthetic range is replaced in place on the device that owns it
        sh.upsert(ids[12345:12346], codes[0:1])
        assert np.array_equal(sh.fetch_codes(ids[12345:12346]), codes[0:1])
        sh.mask_destroy(m)


def test_sharded_device_resident_call_and_concurrency():
    import torch
    devices = [0, 1] if _ngpu() >= 2 else [0, 0]
    d, bits, n, k, nq = 256, 8, 80000, 10, 32
    dev = torch.device("cuda", 0)
    with szg.Index(d, bits, szg.COSINE, devices=devices) as sh, szg.Index(d, bits, szg.COSINE) as one:
        sh.fill_synthetic(41, 0, n)
        one.fill_synthetic(41, 0, n)
        qs = o.synth_queries(42, 0, nq, d)
        want = one.search_topk(qs, k)
        tq = torch.from_numpy(qs).to(dev)
        oi = torch.zeros((nq, k), dtype=torch.int64, device=dev)
        od = torch.zeros((nq, k), dtype=torch.float64, device=dev)
        on = torch.zeros(nq, dtype=torch.int32, device=dev)
        of = torch.zeros(nq, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        for _ in range(3):
            sh.search_topk_dev(tq.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), on.data_ptr(), st, d_out_flags=of.data_ptr())
        torch.cuda.synchronize(dev)
        assert np.array_equal(oi.cpu().numpy().astype(np.uint64), want[0]) and np.array_equal(od.cpu().numpy(), want[1])
        # concurrent host-buffer searches on the sharded handle
        errs = []

        def work(t):
            try:
                for rep in range(6):
                    j = (t + rep) % nq
                    gi, gd, gn, _ = sh.search_topk(qs[j], k)
                    assert np.array_equal(gi[0], want[0][j]) and np.array_equal(gd[0], want[1][j])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs[0]


@pytest.mark.parametrize("devices", _device_lists(), ids=lambda v: "dev" + "".join(map(str, v)))
def test_sharded_repeated_calls_are_replayed_as_one_graph_over_all_devices(devices):
    d, bits, n, k = 768, 8, 40000, 10
    with szg.Index(d, bits, szg.COSINE, devices=devices) as sh, szg.Index(d, bits, szg.COSINE) as one:
        sh.fill_synthetic(77, 0, n)
        one.fill_synthetic(77, 0, n)
        for nq in (1, 32):
            qsets = [o.synth_queries(200 + r, 0, nq, d) for r in range(4)]
            want = [one.search_topk(q, k) for q in qsets]
            g0 = sh.stats()["graph_launches"]
            for rnd in range(3):
                for q, w in zip(qsets, want):
                    got = sh.search_topk(q, k)
                    assert np.array_equal(got[0], w[0]) and np.array_equal(got[1], w[1]) and np.array_equal(got[2], w[2])
            assert sh.stats()["graph_launches"] - g0 >= 8, sh.stats()
        # a mutation through the router makes the captured sequences stale
        q = o.synth_queries(200, 0, 1, d)[0]
        code = o.encode(np.clip(q / np.abs(q).max(), -1, 1), bits)[None, :]
        sh.upsert(np.array([10 ** 9 + 1], dtype=np.uint64), code)
        for _ in range(3):
            gi, gd, gn, _ = sh.search_topk(q, k)
            assert gi[0, 0] == 10 ** 9 + 1
        sh.remove(np.array([10 ** 9 + 1], dtype=np.uint64))
        for _ in range(3):
            got = sh.search_topk(q, k)
            w = one.search_topk(q, k)
            assert np.array_equal(got[0], w[0]) and np.array_equal(got[1], w[1])


def test_sharded_near_duplicates_escalate_on_every_shard():
    # rows closer together than the 2-digit surrogate resolves: the sharded path must walk the same escalation ladder
    d, n, k = 64, 6000, 10
    rng = np.random.default_rng(9)
    base = rng.uniform(-1, 1, size=d)
    vecs = base[None, :] + rng.normal(scale=2e-6, size=(n, d))
    codes = o.encode_rows(vecs, 64)
    ids = np.arange(n, dtype=np.uint64)
    q = base + rng.normal(scale=1e-3, size=d)
    with szg.Index(d, 64, szg.EUCLIDEAN, devices=[0, 0]) as sh:
        sh.upsert(ids, codes)
        ri, rd, _ = o.search_exact(codes, ids, d, 64, szg.EUCLIDEAN, q, k=k)
        gi, gd, gn, _ = sh.search_topk(q, k)
        assert_results_match(gi[0], gd[0], ri, rd)
    # quantized rows: identical codes everywhere (every distance ties) plus a few distinct ones
    codes8 = np.tile(o.synth_rows(1, 0, 1, d, 8), (n, 1))
    codes8[::500] = o.synth_rows(2, 0, n // 500, d, 8)
    with szg.Index(d, 8, szg.COSINE, devices=[0, 0, 0]) as sh:
        sh.upsert(ids, codes8)
        q = o.synth_queries(3, 0, 1, d)[0]
        ri, rd, _ = o.search_exact(codes8, ids, d, 8, szg.COSINE, q, k=k)
        gi, gd, gn, _ = sh.search_topk(q, k)
        assert_results_match(gi[0], gd[0], ri, rd)


# ------------------------------------------------------------------ BASELINE.json configurations at their stated sizes
def test_cfg4_10m_x768_q8_cosine_k10_windows_against_oracle():
    """configs[3] at full size.  The oracle cannot scan 10 M x 768 in seconds, so it scans disjoint 250 k-row windows: the best
    rows of every window must be in the result or no closer than the k-th result, every returned distance is the oracle's
    distance of that row, and a sharded handle (several devices, or several shards on this one) returns the same bits."""
    d, bits, seed, k, n = 768, 8, 0x5A590004, 10, 10_000_000
    qs = o.synth_queries(seed + 1, 0, 3, d)
    devs = list(range(min(_ngpu(), 8))) if _ngpu() >= 2 else [0, 0]
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(seed, 0, n)
        ix.set_option(_capi.OPT_COMBINE, 0)
        gi, gd, gn, scanned = ix.search_topk(qs, k)
        bi, bd, bn, _ = ix.search_batch(qs, k)                          # the tensor-core path over the same 10 M rows
        assert scanned == n and (gn == k).all()
        assert np.array_equal(gi, bi) and np.array_equal(gd, bd)
    windows = [0, 2_400_000, 5_100_000, 7_300_000, 9_750_000]
    with ThreadPoolExecutor(max_workers=8) as ex:
        def scan(w0, qi):
            rows = o.synth_rows(seed, w0, 250_000, d, bits)
            wid = np.arange(w0, w0 + 250_000, dtype=np.uint64)
            return o.search_exact(rows, wid, d, bits, szg.COSINE, qs[qi], k=k, order=np.arange(250_000, dtype=np.int64))
        futs = {(w0, qi): ex.submit(scan, w0, qi) for w0 in windows for qi in range(3)}
        for (w0, qi), f in futs.items():
            ri, rd, _ = f.result()
            got = set(gi[qi].tolist())
            for i, dist in zip(ri.tolist(), rd.tolist()):
                assert i in got or dist >= gd[qi, -1], f"row {i} of window {w0} (d={dist!r}) beats the k-th result {gd[qi, -1]!r}"
                if i in got:
                    assert dist == gd[qi, gi[qi].tolist().index(i)]    # bit-identical fp64 distance
    for qi in range(3):
        assert np.all(np.diff(gd[qi]) >= 0)
        td = _true_dist_synth(seed, d, bits, szg.COSINE, qs[qi])
        for i, dist in zip(gi[qi].tolist(), gd[qi].tolist()):
            assert td(i) == dist
    with szg.Index(d, bits, szg.COSINE, devices=devs) as sh:            # appendix B-15 at size
        sh.fill_synthetic(seed, 0, n)
        si, sd, sn, scanned = sh.search_topk(qs, k)
        assert scanned == n and np.array_equal(si, gi) and np.array_equal(sd, gd)
        s1 = sh.search_topk(qs[0], k)
        assert np.array_equal(s1[0][0], gi[0]) and np.array_equal(s1[1][0], gd[0])


def test_cfg5_1m_x768_q16_euclid_k100_batch_against_oracle():
    """configs[4]'s shape on 1 M rows: a 1024-query batch, k = 100, 16-bit euclidean, through the tensor-core contraction;
    8 of the queries against the oracle's exact scan (not against the streaming scan)."""
    d, bits, seed, k, n, nq = 768, 16, 0x5A590005, 100, 1_000_000, 1024
    qs = o.synth_queries(seed + 1, 0, nq, d)
    with szg.Index(d, bits, szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(seed, 0, n)
        b0 = ix.stats()["batch_queries"]
        gi, gd, gn, scanned = ix.search_batch(qs, k)
        assert ix.stats()["batch_queries"] - b0 == nq and scanned == n and (gn == k).all()
    codes = o.synth_rows(seed, 0, n, d, bits)
    ids = np.arange(n, dtype=np.uint64)
    order = np.arange(n, dtype=np.int64)
    check = [0, 1, 63, 64, 500, 777, 1022, 1023]
    with ThreadPoolExecutor(max_workers=8) as ex:
        refs = list(ex.map(lambda qi: o.search_exact(codes, ids, d, bits, szg.EUCLIDEAN, qs[qi], k=k, order=order), check))
    for qi, (ri, rd, _) in zip(check, refs):
        assert_results_match(gi[qi], gd[qi], ri, rd, what=f"cfg5 q{qi}")
        assert np.array_equal(gd[qi], rd)


def test_cfg3_1m_x384_f64_cosine_radius_with_device_filter_against_oracle():
    """configs[2] at full size, exact half: 1 M x 384 float64 rows, cosine, radius 0.46, metadata filter `bucket < 3` evaluated
    on the device (szg_meta_upsert + szg_filter_mask), against the oracle's scan of all 1 M rows; plus the candidate
    re-scoring of 200 k gathered rows (the LSH half at this size runs in tests/test_host_cpp.py)."""
    d, bits, seed, n = 384, 64, 0x5A590003, 1_000_000
    qs = o.synth_queries(seed + 1, 0, 2, d)
    ids = np.arange(n, dtype=np.uint64)
    with szg.Index(d, bits, szg.COSINE) as ix:
        ix.fill_synthetic(seed, 0, n)
        step = 100_000
        for r0 in range(0, n, step):
            sub = ids[r0:r0 + step]
            cols = _upsert_meta(ix, sub, [b'{"bucket": %d}' % (int(i) % 10) for i in sub], ["bucket"])
        m = ix.filter_mask(szf.lower(E("<", I("bucket"), V(3.0)), cols))
        got = [ix.search_radius(q, 0.46, mask_id=m) for q in qs]
        gk = ix.search_topk(qs, 10, mask_id=m)
        rows = np.random.default_rng(3).integers(0, n, size=200_000)
        gr = ix.rescore(qs[0], ids[rows])
    codes = o.synth_rows(seed, 0, n, d, bits)
    pm = (ids % 10 < 3).astype(np.uint8)
    order = np.arange(n, dtype=np.int64)
    with ThreadPoolExecutor(max_workers=4) as ex:
        fr = [ex.submit(o.search_exact, codes, ids, d, bits, szg.COSINE, q, 0, 0.46, pm, order) for q in qs]
        fk = [ex.submit(o.search_exact, codes, ids, d, bits, szg.COSINE, q, 10, 0.0, pm, order) for q in qs]
        for (gi, gd, scanned), f in zip(got, fr):
            ri, rd, _ = f.result()
            assert scanned == n and ri.size > 100
            assert_radius_match(gi, gd, ri, rd, 0.46, "cfg3 radius+filter")
            assert np.array_equal(gd, rd)
        for qi, f in enumerate(fk):
            ri, rd, _ = f.result()
            assert_results_match(gk[0][qi], gk[1][qi], ri, rd, what="cfg3 top-k+filter")
    want = o.row_distances(codes, d, bits, szg.COSINE, qs[0], rows)
    assert np.array_equal(gr, want, equal_nan=True)


# ------------------------------------------------------------------ reference-held end-to-end cases
def test_rest_test_filter_and_cosine_search_returns_exactly_id_1():
    """rest_test.go:503-569 (TestSearchRecordsWithFilter) through the library: three float64 cosine documents
       1: [0.1, 0.2, 0.3, 0.4, 0.5] {"category": "A", "score": 85}
       2: [0.6, 0.7, 0.8, 0.9, 1.0] {"category": "B", "score": 90}
       3: [0.2, 0.3, 0.4, 0.5, 0.6] {"category": "A", "score": 75}
    query [0.1, 0.2, 0.3, 0.4, 0.5], k = 3, filter `category == "A" AND score > 75`  =>  exactly one result, id 1
    (szg_meta_upsert -> szg_filter_mask -> szg_search_topk)."""
    vecs = np.array([[0.1, 0.2, 0.3, 0.4, 0.5], [0.6, 0.7, 0.8, 0.9, 1.0], [0.2, 0.3, 0.4, 0.5, 0.6]])
    metas = [b'{"category": "A", "score": 85}', b'{"category": "B", "score": 90}', b'{"category": "A", "score": 75}']
    ids = np.array([1, 2, 3], dtype=np.uint64)
    for devices in (None, [0, 0]):
        with szg.Index(5, 64, szg.COSINE, devices=devices) as ix:
            ix.encode(vecs, ids, upsert=True)                      # AddDocument: encodeDocument + mirror
            cols = _upsert_meta(ix, ids, metas, ["category", "score"])
            tree = E("AND", E("==", I("category"), V("A")), E(">", I("score"), V(75.0)))
            m = ix.filter_mask(szf.lower(tree, cols))
            gi, gd, gn, scanned = ix.search_topk(vecs[0], 3, mask_id=m)
            assert gn[0] == 1 and gi[0, 0] == 1 and scanned == 3   # PercentSearched counts the filtered rows (collection.go:589)
            # the stored vector against itself: dot / (sqrt(m1) sqrt(m2)) is exactly 1 for these values, Acos(1) = 0
            assert gd[0, 0] == o.angular(vecs[0], o.decode(o.encode(vecs[0], 64), 5, 64)) == 0.0
            # without the filter all three come back, nearest first, with the oracle's distances
            gi, gd, gn, _ = ix.search_topk(vecs[0], 3)
            assert gn[0] == 3 and gi[0].tolist() == [1, 3, 2]
            assert gd[0].tolist() == [0.0, o.angular(vecs[0], vecs[2]), o.angular(vecs[0], vecs[1])]
