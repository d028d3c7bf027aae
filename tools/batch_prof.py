import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
rows, dims, nq = int(os.environ.get("ROWS", "400000")), 768, 1024
qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
ix = szg.Index(dims, 8, szg.COSINE)
ix.fill_synthetic(7, 0, rows)
for rep in range(2):
    ix.search_batch(qs, 10, flags=1)
print("ok", ix.last_scan_times_ms())
