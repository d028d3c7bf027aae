"""Ingest-side quantization: szg_encode (float64 vectors -> stream-1 bytes on the device, optionally into the mirror)
against the oracle's encodeDocument restatement on one host core.  Host buffers, copies inside the timed call."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from oracle import pyoracle as o

n, dims = int(os.environ.get("ROWS", "200000")), int(os.environ.get("DIMS", "768"))
x = np.random.default_rng(1).normal(size=(n, dims)) * 0.4
for bits in (4, 8, 16):
    with szg.Index(dims, bits, szg.COSINE) as ix:
        ix.encode(x[:1000])
        t0 = time.perf_counter(); codes = ix.encode(x); t_enc = time.perf_counter() - t0
        ids = np.arange(n, dtype=np.uint64)
        t0 = time.perf_counter(); ix.encode(x, ids=ids, upsert=True); t_up = time.perf_counter() - t0
        m = 2000
        t0 = time.perf_counter(); ref = o.encode_rows(x[:m], bits); t_cpu = (time.perf_counter() - t0) / m
        assert np.array_equal(codes[:m], ref)
        print(json.dumps({"bits": bits, "rows": n, "dims": dims, "encode_ms": round(t_enc * 1e3, 1), "encode_rows_per_s": round(n / t_enc),
                          "input_GBps": round(n * dims * 8 / t_enc / 1e9, 2), "encode_and_upsert_ms": round(t_up * 1e3, 1),
                          "oracle_one_core_rows_per_s": round(1 / t_cpu), "identical_to_oracle_on_sample": True}), flush=True)
