"""Small drivers for the round-2 ncu captures (one kernel each; see profiles/README.md):
   python tools/prof_r02.py single   -- 1 query per call on 10M x 768 8-bit (scan_small + finalize)
   python tools/prof_r02.py batch32  -- 32 queries per call (batch_kernel, the headline launch)
   python tools/prof_r02.py rescore  -- 200 k gathered fp64 rows (rescore_kernel) and 200-candidate calls
   python tools/prof_r02.py radius   -- cfg3 radius search
   python tools/prof_r02.py batch1024 [rows] [bits] -- 1024 queries per call (batch_kernel, 16 query groups per launch)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

mode = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else None
rng = np.random.default_rng(1)
if mode in ("single", "batch32"):
    n = rows or 10_000_000
    with szg.Index(768, 8, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590004, 0, n)
        ix.set_option(_capi.OPT_GRAPHS, 0)
        ix.set_option(_capi.OPT_COMBINE, 0)
        nq = 1 if mode == "single" else 32
        for _ in range(4):
            ix.search_topk(rng.uniform(-1, 1, size=(nq, 768)), 10)
elif mode == "batch1024":
    n = rows or 1_250_000
    bits = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    with szg.Index(768, bits, szg.COSINE if bits == 8 else szg.EUCLIDEAN) as ix:
        ix.fill_synthetic(0x5A590004, 0, n)
        ix.set_option(_capi.OPT_GRAPHS, 0)
        ix.set_option(_capi.OPT_COMBINE, 0)
        q = rng.uniform(-1, 1, size=(1024, 768))
        for _ in range(3):
            ix.search_topk(q, 10 if bits == 8 else 100)
elif mode == "rescore":
    n = rows or 1_000_000
    with szg.Index(384, 64, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590003, 0, n)
        q = rng.uniform(-1, 1, size=384)
        ids = rng.integers(0, n, size=200_000).astype(np.uint64)
        for _ in range(3):
            ix.rescore(q, ids)
        for _ in range(3):
            ix.rescore(q, ids[:200])
elif mode == "radius":
    n = rows or 1_000_000
    with szg.Index(384, 64, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590003, 0, n)
        for _ in range(3):
            ix.search_radius(rng.uniform(-1, 1, size=384), 0.46)
print("ok")
