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
        qs = [rng.uniform(-1, 1, size=(32, 768)) for _ in range(3)]
        for callers in (1, 2, 3):
            n = 300

            lat = [[] for _ in range(callers)]

            def work(i):
                for _ in range(n):
                    t = time.perf_counter()
                    ix.search_topk(qs[i], 10)
                    lat[i].append(time.perf_counter() - t)

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
            a = np.sort(np.concatenate(lat)) * 1e3
            print(f"rows {rows} callers {callers}: {dt / (n * callers) * 1e3:.4f} ms/step; call latency ms p50 {a[len(a) // 2]:.3f} p90 {a[int(len(a) * 0.9)]:.3f} "
                  f"p99 {a[int(len(a) * 0.99)]:.3f} max {a[-1]:.3f}, calls > 1.5 x p50: {int((a > 1.5 * a[len(a) // 2]).sum())}; "
                  f"fixed_ranges={os.environ.get('SZG_BATCH_FIXED_RANGES', '0')} priority={os.environ.get('SZG_PRIORITY', 'default')}", flush=True)
