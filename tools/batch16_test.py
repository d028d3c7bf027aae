"""16-bit batched path (byte planes on the tensor cores) against the streaming scan, plus timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
rows = int(os.environ.get("ROWS", "100000")); dims = int(os.environ.get("DIMS", "768")); nq = int(os.environ.get("NQ", "130"))
k = int(os.environ.get("K", "10")); metric = szg.COSINE if os.environ.get("METRIC", "euclid") == "cosine" else szg.EUCLIDEAN
qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
ix = szg.Index(dims, 16, metric)
ix.fill_synthetic(7, 0, rows)
bi, bd, bn, _ = ix.search_batch(qs, k)
print("stats", {k_: v for k_, v in ix.stats().items() if k_ in ("batch_queries", "escalations", "uncertain_results")}, flush=True)
si, sd, sn, _ = ix.search_topk(qs, k)
ok = np.array_equal(bi, si) and np.array_equal(bd, sd) and np.array_equal(bn, sn)
print("identical to the streaming scan:", ok)
if not ok:
    bad = np.nonzero((bi != si).any(axis=1))[0]
    print("queries differing:", bad[:10], "of", nq); q = bad[0]; print(bi[q], si[q]); print(bd[q], sd[q])
for rep in range(3):
    ix.search_batch(qs, k, flags=1)
    bt = ix.last_scan_times_ms()
    print(f"rep {rep}: batch kernel {bt.sum():.3f} ms -> {nq / bt.sum() * 1e3:.0f} queries/s kernel-only", flush=True)
t0 = time.time(); ix.search_topk(qs[:64], k); print(f"streaming scan: {(time.time() - t0) / 64 * 1e3:.3f} ms per query")
