"""Host-buffer 32-query steps per second with 1 and 2 calls in flight on one GPU (two caller threads, each with its own workspace
and stream), at shard sizes of the 8- and 1-GPU jobs: what the overlap of one call's finalize with the next call's batch_kernel
is worth.  SZG_BATCH_FIXED_RANGES=1 in a second process gives the fixed-range kernel for comparison."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

rng = np.random.default_rng(1)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_250_000, 10_000_000]
for rows in sizes:
    with szg.Index(768, 8, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590004, 0, rows)
        ix.set_option(_capi.OPT_COMBINE, 0)
        qs = [rng.uniform(-1, 1, size=(32, 768)) for _ in range(2)]
        for callers in (1, 2):
            n = 300

            def work(i):
                for _ in range(n):
                    ix.search_topk(qs[i], 10)

            for i in range(callers):
                for _ in range(5):
                    ix.search_topk(qs[i], 10)
            ts = [threading.Thread(target=work, args=(i,)) for i in range(callers)]
            t0 = time.perf_counter()
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            dt = time.perf_counter() - t0
            print(f"rows {rows} callers {callers}: {dt / (n * callers) * 1e3:.4f} ms/step, fixed_ranges={os.environ.get('SZG_BATCH_FIXED_RANGES', '0')}", flush=True)
