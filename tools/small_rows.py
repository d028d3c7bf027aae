"""Small-row scans (rows of a few 16-byte chunks): per-query device time for a few
streaming geometries.  Usage: python tools/small_rows.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi

nq = int(os.environ.get("NQ", "64"))
shapes = [(1_000_000, 128, 4, szg.EUCLIDEAN), (1_000_000, 64, 8, szg.COSINE), (2_000_000, 64, 4, szg.EUCLIDEAN),
          (1_000_000, 32, 16, szg.EUCLIDEAN), (4_000_000, 128, 8, szg.COSINE), (100_000, 128, 4, szg.EUCLIDEAN)]
for rows, dims, quant, metric in shapes:
    qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
    ix = szg.Index(dims, quant, metric)
    ix.fill_synthetic(7, 0, rows)
    ix.set_option(_capi.OPT_STREAMS, 1)
    for warps, stages in ((16, 2), (16, 3), (8, 2), (8, 4)):
        ix.set_option(_capi.OPT_SCAN_WARPS, warps)
        ix.set_option(_capi.OPT_SCAN_STAGES, stages)
        ix.search_topk(qs[:2], 10)
        ids, dd, n, _ = ix.search_topk(qs, 10)
        ms = float(np.sum(ix.last_scan_times_ms())) / nq
        st = ix.stats()
        print(json.dumps(dict(rows=rows, dims=dims, quant=quant, warps=warps, stages=stages,
                              smem=st["scan_smem_bytes"], us_per_query=round(ms * 1e3, 2), qps=round(1e3 / ms),
                              gbs=round(rows * ix.rowbytes / ms / 1e6, 1), check=int(ids.sum() % 1000003))), flush=True)
    ix.close()
