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
print("done")
