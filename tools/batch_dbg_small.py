import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
rows, dims = 1250000, 768
ix = szg.Index(dims, 8, szg.COSINE)
ix.fill_synthetic(7, 0, rows)
for nq in (64, 128):
    qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
    for dbg in (0, 1, 4, 5):
        os.environ["SZG_BATCH_DEBUG"] = str(dbg)
        for rep in range(3):
            ix.search_batch(qs, 10, flags=1)
            bt = ix.last_scan_times_ms()
        print(f"nq={nq} debug={dbg}: batch kernel {bt.sum():.3f} ms", flush=True)
