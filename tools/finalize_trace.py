"""Phase times of finalize_kernel (clock64 stamps of its first CTA, SZG_OPT_TRACE_BUFFER): one query per call."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg  # noqa: E402
from syzgydb_b200 import _capi  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
buf = torch.zeros(8, dtype=torch.int64, device="cuda:0")
rng = np.random.default_rng(1)
for bits, dims, metric in ((8, 768, szg.COSINE), (4, 128, szg.EUCLIDEAN), (64, 384, szg.COSINE)):
    with szg.Index(dims, bits, metric) as ix:
        ix.fill_synthetic(7, 0, rows if bits != 64 else rows // 4)
        ix.set_option(_capi.OPT_GRAPHS, 0)
        ix.set_option(_capi.OPT_COMBINE, 0)
        ix.set_option(_capi.OPT_TRACE_BUFFER, buf.data_ptr())
        out = []
        for _ in range(6):
            ix.search_topk(rng.uniform(-1, 1, size=(nq, dims)), 10)
            torch.cuda.synchronize()
            t = buf.cpu().numpy()
            out.append([int(t[i + 1] - t[i]) for i in range(5)] + [int(t[6] - t[2]), int(t[7] >> 32), int(t[7] & 0xFFFFFFFF)])
        print(f"q{bits} d{dims} nq{nq}: cycles [merge lists, block merge, exact_staged, ranking, certify | fetch+stage, thread0 in chains, thread0 at barriers] =", out[-3:])
