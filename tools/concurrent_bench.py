"""Concurrent single-query Search calls on one handle (what the Go shim sees: one query per Collection.Search, many
goroutines): aggregate QPS with and without SZG_OPT_COMBINE, T caller threads, host buffers in and out."""
import json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi
rows, dims = int(os.environ.get("ROWS", "10000000")), int(os.environ.get("DIMS", "768"))
ix = szg.Index(dims, 8, szg.COSINE)
ix.fill_synthetic(7, 0, rows)
qs = np.random.default_rng(1).uniform(-1, 1, size=(256, dims))
ix.search_topk(qs[:4], 10)
for combine in (0, 1):
    ix.set_option(_capi.OPT_COMBINE, combine)
    for T in (1, 4, 16, 64):
        per = max(8, 256 // T)
        c0 = ix.stats()["combined_queries"]
        def work(t):
            for r in range(per):
                ix.search_topk(qs[(t * per + r) % 256], 10)
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        dt = time.perf_counter() - t0
        print(json.dumps(dict(rows=rows, dims=dims, combine=combine, threads=T, calls=T * per, qps=round(T * per / dt, 1),
                              ms_per_call=round(dt / per * 1e3, 3), combined_queries=ix.stats()["combined_queries"] - c0)), flush=True)
