"""Where the tensor-core path starts to win over per-query streaming scans: wall time of szg_search_topk (host buffers, graphs on) with
2 / 3 / 4 queries per call, SZG_OPT_BATCH_MIN_QUERIES = 2 (tensor path) against 99 (scans), on collections of different sizes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

rng = np.random.default_rng(1)
for rows, dims, bits in ((100_000, 384, 8), (1_000_000, 128, 4), (1_000_000, 768, 8), (10_000_000, 768, 8)):
    with szg.Index(dims, bits, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590004, 0, rows)
        ix.set_option(_capi.OPT_COMBINE, 0)
        for nq in (2, 3, 4, 6, 8):
            q = rng.uniform(-1, 1, size=(nq, dims))
            out = []
            for bm in (2, 99, 0):
                ix.set_option(_capi.OPT_BATCH_MIN_QUERIES, bm)
                for _ in range(5):
                    ix.search_topk(q, 10)
                t = []
                for _ in range(20):
                    t0 = time.perf_counter()
                    ix.search_topk(q, 10)
                    t.append(time.perf_counter() - t0)
                out.append(np.median(t) * 1e6)
            print(f"rows {rows} x {dims} {bits}-bit ({rows * dims * bits // 8 / 1e6:.0f} MB) nq {nq}: tensor path {out[0]:8.1f} us, scans {out[1]:8.1f} us, default {out[2]:8.1f} us", flush=True)
