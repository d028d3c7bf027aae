"""Device-side metadata filter at cfg3's scale (1M documents, {"bucket": id % 10, "status": ..., "name": ...}):
wall time of szg_meta_upsert (column load), of szg_filter_mask for three programs, and of the path it replaces on the
library side (szg_mask_create from a host-evaluated pass[] array); plus the Python oracle's per-document rate on a
sample, as a reminder of what "evaluate the predicate per document" costs even before the Go json.Unmarshal."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from oracle import filter as of
from syzgydb_b200 import _capi
from syzgydb_b200 import filter as hf

n = int(os.environ.get("ROWS", "1000000"))
I, V, E, A = (lambda x: ("ident", x)), (lambda v: ("value", v)), (lambda op, l, r: ("expr", op, l, r)), (lambda *e: ("array", list(e)))
ix = szg.Index(8, 8, szg.EUCLIDEAN)
ix.fill_synthetic(1, 0, n)
ids = np.arange(n, dtype=np.uint64)
L = ix._L
# columns: 0 bucket (number), 1 status (string, 5 distinct), 2 name (string, 5000 distinct)
statuses = [b"active", b"pending", b"inactive", b"suspended", b"deleted"]
names = [("user_%05d@example.%s" % (i, "com" if i % 3 else "org")).encode() for i in range(5000)]
bufs = [C.create_string_buffer(s, len(s) + 1) for s in statuses + names]
arr = (_capi.MetaValue * (n * 3))()
t0 = time.time()
for i in range(n):
    a, b, c = arr[3 * i], arr[3 * i + 1], arr[3 * i + 2]
    a.kind, a.num = 3, float(i % 10)
    s = bufs[(i * 7) % 5]
    b.kind, b.str, b.str_len = 4, C.cast(s, C.c_void_p), len(s.value)
    s = bufs[5 + (i * 13) % 5000]
    c.kind, c.str, c.str_len = 4, C.cast(s, C.c_void_p), len(s.value)
print(f"(python staging of {n} x 3 values: {time.time() - t0:.1f} s, not part of any measurement)")
dk = np.ones(n, dtype=np.uint8)
cols = np.arange(3, dtype=np.uint32)
t0 = time.perf_counter()
_capi._check(L.szg_meta_upsert(ix._h, ids.ctypes.data_as(C.POINTER(C.c_uint64)), n, dk.ctypes.data_as(C.POINTER(C.c_uint8)),
                               cols.ctypes.data_as(C.POINTER(C.c_uint32)), 3, arr))
t_up = time.perf_counter() - t0
print(json.dumps({"step": "szg_meta_upsert", "rows": n, "columns": 3, "ms": round(t_up * 1e3, 2), "rows_per_s": round(n / t_up)}))
colmap = {"bucket": 0, "status": 1, "name": 2}
programs = {
    "bucket < 3": E("<", I("bucket"), V(3.0)),
    "bucket >= 2 AND status IN ['active','pending'] AND NOT (bucket == 7)":
        E("AND", E("AND", E(">=", I("bucket"), V(2.0)), E("IN", I("status"), A(V("active"), V("pending")))),
          E("NOT", None, E("==", I("bucket"), V(7.0)))),
    "name ENDS_WITH '.org' OR (name > 'user_04000' AND status != 'deleted')":
        E("OR", E("ENDS_WITH", I("name"), V(".org")), E("AND", E(">", I("name"), V("user_04000")), E("!=", I("status"), V("deleted")))),
}
q = np.zeros(8)
for label, tree in programs.items():
    prog = hf.lower(tree, colmap)
    ix.mask_destroy(ix.filter_mask(prog))  # warm-up
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        m = ix.filter_mask(prog)
        ts.append(time.perf_counter() - t0)
        if _ < 4:
            ix.mask_destroy(m)
    passing = len(ix.search_radius(q, 1e9, mask_id=m)[0])
    ix.mask_destroy(m)
    print(json.dumps({"step": "szg_filter_mask", "filter": label, "ops": len(prog), "ms_min": round(min(ts) * 1e3, 3),
                      "ms_median": round(sorted(ts)[2] * 1e3, 3), "rows_per_s": round(n / min(ts)), "passing": passing}))
# the path it replaces inside the library: a host-evaluated pass[] array turned into a mask
passed = (ids % 10 < 3).astype(np.uint8)
ix.mask_destroy(ix.mask_create(ids, passed))
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    m = ix.mask_create(ids, passed)
    ts.append(time.perf_counter() - t0)
    ix.mask_destroy(m)
print(json.dumps({"step": "szg_mask_create (pass[] already evaluated on the host)", "ms_min": round(min(ts) * 1e3, 2)}))
sample = 20000
docs = [json.dumps({"bucket": i % 10, "status": statuses[(i * 7) % 5].decode(), "name": names[(i * 13) % 5000].decode()}).encode() for i in range(sample)]
tree = programs["bucket < 3"]
t0 = time.perf_counter()
cnt = sum(of.filter_document(tree, d) for d in docs)
t = time.perf_counter() - t0
print(json.dumps({"step": "python oracle (json parse + tree walk per document)", "docs": sample, "us_per_doc": round(t / sample * 1e6, 2),
                  "extrapolated_ms_for_all_rows": round(t / sample * n * 1e3)}))
