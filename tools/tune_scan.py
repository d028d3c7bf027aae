"""Sweeps the streaming geometry of the scan kernel (warps x stages x tile chunks) on one B200 and prints the
mean device time of a scan launch and the achieved GB/s.  Usage: python tools/tune_scan.py [rows dims quant metric]"""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dims = int(sys.argv[2]) if len(sys.argv) > 2 else 768
quant = int(sys.argv[3]) if len(sys.argv) > 3 else 8
metric = szg.COSINE if (sys.argv[4] if len(sys.argv) > 4 else "cosine") == "cosine" else szg.EUCLIDEAN
nq = int(os.environ.get("NQ", "12"))
qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
ix = szg.Index(dims, quant, metric)
ix.fill_synthetic(7, 0, rows)
ix.set_option(_capi.OPT_STREAMS, 1)
ref = None
out = []
for warps, stages, tc in itertools.product((8, 16), (2, 3, 4, 6), (4, 6, 8, 12, 16)):
    ix.set_option(_capi.OPT_SCAN_WARPS, warps)
    ix.set_option(_capi.OPT_SCAN_STAGES, stages)
    ix.set_option(_capi.OPT_SCAN_TILE_CHUNKS, tc)
    st = ix.stats()
    ix.search_topk(qs[:2], 10)
    ids, dd, n, _ = ix.search_topk(qs, 10)
    if ref is None:
        ref = ids.copy()
    assert np.array_equal(ids, ref), "results changed with the geometry"
    t = ix.last_scan_times_ms()
    ms = float(np.mean(t)) / nq  # one launch scans all nq queries
    gbs = rows * ix.rowbytes / ms / 1e6
    rec = dict(warps=warps, stages=stages, tile_chunks=tc, tile_bytes=st["scan_tile_bytes"], eff_stages=st["scan_stages"],
               smem=st["scan_smem_bytes"], ms=round(ms, 4), gbs=round(gbs, 1))
    out.append(rec)
    print(json.dumps(rec), flush=True)
best = max(out, key=lambda r: r["gbs"])
print("BEST", json.dumps(best))
