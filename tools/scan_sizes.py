"""Scan time vs collection size (fit of streaming rate and fixed per-launch overhead)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi
dims, quant = 768, 8
qs = np.random.default_rng(1).uniform(-1, 1, size=(16, dims))
for flags in (0, 1):
    pts = []
    for rows in [int(x) for x in os.environ.get('SIZES', '250000,500000,1000000,2000000,4000000,8000000,16000000').split(',')]:
        ix = szg.Index(dims, quant, szg.COSINE)
        ix.fill_synthetic(7, 0, rows)
        ix.set_option(_capi.OPT_STREAMS, 1)
        ix.search_topk(qs[:2], 10, flags=flags)
        nq = int(os.environ.get('NQ', '16'))
        ix.search_topk(qs[:nq], 10, flags=flags)
        ms = float(np.mean(ix.last_scan_times_ms())) / nq  # one launch scans all nq queries
        pts.append((rows * 768 / 1e9, ms))
        print(json.dumps(dict(flags=flags, rows=rows, gb=rows * 768 / 1e9, ms=round(ms, 4), gbs=round(rows * 768 / ms / 1e6, 1))), flush=True)
        ix.close()
    x = np.array([p[0] for p in pts]); y = np.array([p[1] for p in pts])
    A = np.vstack([x, np.ones_like(x)]).T
    slope, icpt = np.linalg.lstsq(A, y, rcond=None)[0]
    print(f"flags={flags}: rate {1 / slope:.1f} GB/ms-> {1000 / slope / 1000:.0f} GB/s, fixed overhead {icpt * 1000:.1f} us")
