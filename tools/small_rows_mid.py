"""Mid-size rows (12, 16, 24 chunks): general kernel vs scan_small.cuh (SZG_SCAN_SMALL_MAXC=24), per-query scan time."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi
shapes = [(100_000, 384, 8, szg.COSINE), (1_000_000, 384, 8, szg.COSINE), (4_000_000, 384, 8, szg.COSINE), (1_000_000, 256, 8, szg.COSINE),
          (1_000_000, 384, 4, szg.EUCLIDEAN), (1_000_000, 96, 16, szg.EUCLIDEAN), (100_000, 192, 16, szg.COSINE)]
for rows, dims, quant, metric in shapes:
    ix = szg.Index(dims, quant, metric)
    ix.fill_synthetic(7, 0, rows)
    ix.set_option(_capi.OPT_STREAMS, 1)
    for nq in (1, 8, 64, 256):
        qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
        ix.search_topk(qs[:2], 10)
        ids, dd, n, _ = ix.search_topk(qs, 10)
        ms = float(np.sum(ix.last_scan_times_ms())) / nq
        print(json.dumps(dict(maxc=os.environ.get("SZG_SCAN_SMALL_MAXC", "8"), rows=rows, dims=dims, quant=quant, nq=nq,
                              us_per_query=round(ms * 1e3, 2), qps=round(1e3 / ms), gbs=round(rows * ix.rowbytes / ms / 1e6, 1),
                              check=int(ids.sum() % 1000003))), flush=True)
    ix.close()
