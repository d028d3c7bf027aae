"""Kernel time of the many-query shapes (1024 queries, 8-bit cosine k=10 and 16-bit euclidean k=100) on a 1.25 M-row shard."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

rng = np.random.default_rng(1)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
for quant, metric, k in ((8, szg.COSINE, 10), (16, szg.EUCLIDEAN, 100)):
    with szg.Index(768, quant, metric) as ix:
        ix.fill_synthetic(0x5A590004, 0, rows)
        ix.set_option(_capi.OPT_TIMING, 2)
        ix.set_option(_capi.OPT_COMBINE, 0)
        for nq in (256, 1024):
            q = rng.uniform(-1, 1, size=(nq, 768))
            for _ in range(2):
                ix.search_topk(q, k)
            ix.last_scan_times_ms()
            for _ in range(4):
                ix.search_topk(q, k)
            t = ix.last_scan_times_ms()
            print(f"rows {rows} {quant}-bit k={k} nq {nq}: batch_kernel launches {len(t) // 4} per call, {np.sum(t) / 4:.4f} ms per call", flush=True)
