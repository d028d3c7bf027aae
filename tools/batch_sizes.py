"""batch_kernel time against shard size and batch size (what bends the 1 -> 8 GPU curve of the 32-query step): per-launch kernel
time from the library's CUDA events, with the epilogue arithmetic on and off (SZG_BATCH_DEBUG=1 in a second process)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

rng = np.random.default_rng(1)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1_250_000, 2_500_000, 5_000_000, 10_000_000]
for rows in sizes:
    with szg.Index(768, 8, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590004, 0, rows)
        ix.set_option(_capi.OPT_TIMING, 2)
        ix.set_option(_capi.OPT_COMBINE, 0)
        for nq in (4, 32, 64, 128):
            for _ in range(3):
                ix.search_topk(rng.uniform(-1, 1, size=(nq, 768)), 10)
            ix.last_scan_times_ms()
            for _ in range(8):
                ix.search_topk(rng.uniform(-1, 1, size=(nq, 768)), 10)
            t = ix.last_scan_times_ms()
            floor = rows * 768 / 6556.2e9 * 1e3
            print(f"rows {rows} nq {nq}: batch_kernel {np.mean(t):.4f} ms (min {np.min(t):.4f}), HBM floor {floor:.4f} ms, debug={os.environ.get('SZG_BATCH_DEBUG', '0')}", flush=True)
