"""Timing of the batched kernel with parts of it switched off (SZG_BATCH_DEBUG) and different K slices."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
rows, dims, nq = 1250000, 768, 1024
qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
ix = szg.Index(dims, 8, szg.COSINE)
ix.fill_synthetic(7, 0, rows)
for slc in [int(x) for x in os.environ.get("SLICES", "8,16,24,48").split(",")]:
    os.environ["SZG_BATCH_SLICE"] = str(slc)
    for dbg in (0, 1, 4, 5):
        os.environ["SZG_BATCH_DEBUG"] = str(dbg)
        for rep in range(2):
            ix.search_batch(qs, 10, flags=1)
            bt = ix.last_scan_times_ms()
        print(f"slice={slc} debug={dbg}: batch kernel {bt.sum():.3f} ms", flush=True)
