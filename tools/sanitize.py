"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from oracle import pyoracle as o
n, dims = 3000, 96
for bits, metric in ((8, szg.COSINE), (4, szg.EUCLIDEAN), (16, szg.COSINE), (64, szg.EUCLIDEAN)):
    codes = o.synth_rows(3 + bits, 0, n, dims, bits)
    ids = np.arange(n, dtype=np.uint64) * 2 + 1
    q = o.synth_queries(9, 0, 70, dims)
    with szg.Index(dims, bits, metric) as ix:
        ix.upsert(ids, codes)
        ix.remove(ids[::7])
        m = ix.mask_create(ids, (ids % 3 == 0).astype(np.uint8))
        a = ix.search_topk(q[:5], 10)
        b = ix.search_topk(q[:3], 50, mask_id=m)
        c = ix.search_radius(q[0], 0.45 if metric == szg.COSINE else 7.5, mask_id=m)
        d = ix.rescore(q[1], ids[:200])
        e = ix.search_batch(q, 10)
        f = ix.search_batch(q[:66], 100, mask_id=m)
        print(bits, metric, a[2][:2], b[2][:2], len(c[0]), float(d[3]), e[2][:2], f[2][:2], ix.stats()["batch_queries"], flush=True)
# short-row kernel (scan_small.cuh) at several chunk counts, the device encoder and the device filter
from syzgydb_b200 import filter as hf
for bits, d2 in ((4, 128), (8, 64), (8, 384), (16, 96), (8, 768)):
    rows = np.random.default_rng(bits + d2).uniform(-1.2, 1.2, size=(2500, d2))
    ids2 = np.arange(2500, dtype=np.uint64) + 10
    with szg.Index(d2, bits, szg.COSINE) as ix:
        ix.encode(rows, ids=ids2, upsert=True)
        ix.remove(ids2[::9])
        docs = [('{"bucket": %d, "tag": "t%d"}' % (i % 10, i % 7)).encode() for i in range(2500)]
        kinds, vals = zip(*[hf.column_values(dd, ["bucket", "tag"]) for dd in docs])
        live = np.ones(2500, dtype=bool); live[::9] = False
        ix.meta_upsert(ids2[live], np.array(kinds)[live], [0, 1], [v for v, l in zip(vals, live) if l])
        tree = ("expr", "AND", ("expr", "<", ("ident", "bucket"), ("value", 4.0)), ("expr", "STARTS_WITH", ("ident", "tag"), ("value", "t")))
        m = ix.filter_mask(hf.lower(tree, {"bucket": 0, "tag": 1}))
        qq = o.synth_queries(11, 0, 40, d2)
        r1 = ix.search_topk(qq[:1], 10)
        r2 = ix.search_topk(qq, 10, mask_id=m)
        r3 = ix.search_topk(qq[:7], 24)
        print("short rows", bits, d2, r1[2][:1], r2[2][:2], r3[2][:2], flush=True)
print("done")
