"""Phase stamps of batch_kernel's first epilogue warp (SZG_OPT_TRACE_BUFFER, words 8..15): where the fixed cost of a launch goes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

buf = torch.zeros(16, dtype=torch.int64, device="cuda:0")
rng = np.random.default_rng(1)
for rows in (1_250_000, 10_000_000):
    with szg.Index(768, 8, szg.COSINE) as ix:
        ix.fill_synthetic(0x5A590004, 0, rows)
        ix.set_option(_capi.OPT_GRAPHS, 0)
        ix.set_option(_capi.OPT_COMBINE, 0)
        ix.set_option(_capi.OPT_TIMING, 2)
        ix.set_option(_capi.OPT_TRACE_BUFFER, buf.data_ptr())
        for nq in (32, 64):
            for _ in range(4):
                ix.search_topk(rng.uniform(-1, 1, size=(nq, 768)), 10)
                torch.cuda.synchronize()
                t = buf.cpu().numpy()[8:]
            ms = ix.last_scan_times_ms()
            print(f"rows {rows} nq {nq}: kernel {ms[-1]:.4f} ms; epilogue warp cycles: setup {t[0]-t[7]}, first tile (seeding) {t[1]-t[0]}, wait for all ranges {t[2]-t[1]} "
                  f"({t[5]} sleeps), rest {t[3]-t[2]}; inserts of the warp's 8 queries {t[4]}, polls {t[6]}", flush=True)
