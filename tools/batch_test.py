"""Quick check of the tensor-core batched path against the streaming scan (and timing)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi
rows = int(os.environ.get("ROWS", "100000")); dims = int(os.environ.get("DIMS", "768")); nq = int(os.environ.get("NQ", "130"))
k = int(os.environ.get("K", "10")); metric = szg.COSINE if os.environ.get("METRIC", "cosine") == "cosine" else szg.EUCLIDEAN
qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
ix = szg.Index(dims, 8, metric)
ix.fill_synthetic(7, 0, rows)
t0 = time.time(); bi, bd, bn, _ = ix.search_batch(qs, k); t1 = time.time()
bt = ix.last_scan_times_ms()
print("batch call", round(t1 - t0, 4), "s; kernel launches ms:", bt, "stats", {k_: v for k_, v in ix.stats().items() if k_ in ("batch_queries", "escalations", "uncertain_results")})
si, sd, sn, _ = ix.search_topk(qs, k)
ok = np.array_equal(bi, si) and np.array_equal(bd, sd) and np.array_equal(bn, sn)
print("identical to the streaming scan:", ok)
if not ok:
    bad = np.nonzero((bi != si).any(axis=1))[0]
    print("queries differing:", bad[:10], "of", nq)
    q = bad[0]; print(bi[q], si[q]); print(bd[q], sd[q])
for rep in range(3):
    ix.search_batch(qs, k)
    bt = ix.last_scan_times_ms()
    flops = 2.0 * 2 * rows * dims * (-(-nq // 64) * 64)  # 2 digit planes
    print(f"rep {rep}: batch kernel {bt.sum():.3f} ms -> {flops / bt.sum() / 1e9:.1f} TOP/s int8 (2 planes), {nq / bt.sum() * 1e3:.0f} queries/s kernel-only")
