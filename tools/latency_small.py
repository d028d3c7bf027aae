"""Per-call latency of a single-query szg_search_topk (host buffers in and out) on small collections: wall time per call,
and the scan launch's share of it."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
for rows, dims, quant, metric in [(100_000, 384, 8, szg.COSINE), (1_000_000, 128, 4, szg.EUCLIDEAN), (1_000_000, 384, 8, szg.COSINE), (10_000, 384, 8, szg.COSINE)]:
    ix = szg.Index(dims, quant, metric)
    ix.fill_synthetic(7, 0, rows)
    qs = np.random.default_rng(1).uniform(-1, 1, size=(64, dims))
    for i in range(5):
        ix.search_topk(qs[i], 10)
    t, scan = [], []
    for i in range(50):
        t0 = time.perf_counter(); ix.search_topk(qs[i], 10); t.append(time.perf_counter() - t0)
        scan.append(float(np.sum(ix.last_scan_times_ms())))
    t.sort()
    print(json.dumps(dict(rows=rows, dims=dims, quant=quant, call_us_median=round(t[25] * 1e6, 1), call_us_min=round(t[0] * 1e6, 1),
                          scan_launch_us=round(float(np.median(scan)) * 1e3, 1))), flush=True)
    ix.close()
