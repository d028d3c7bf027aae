"""cfg3-shaped measurement of the candidate rescoring path (szg_rescore: LSH leaf id lists -> gather kernel, fp64, the
reference's operation order) and of the filtered radius scan: 1M x 384 fp64 cosine rows (3.07 GB), metadata filter of 30 %
density as a bitmask, radius 0.46.  Times are end to end through the host-buffer C ABI (ids H2D, distances D2H)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
rows, dims = int(os.environ.get("ROWS", "1000000")), 384
ix = szg.Index(dims, 64, szg.COSINE)
ix.fill_synthetic(0x5A590003, 0, rows)
rng = np.random.default_rng(3)
q = rng.uniform(-1, 1, size=dims)
rowbytes = dims * 8
for m in (200, 2000, 20000, 200000):
    ids = rng.choice(rows, size=m, replace=False).astype(np.uint64)
    ix.rescore(q, ids)
    t = []
    for rep in range(10):
        t0 = time.perf_counter(); d = ix.rescore(q, ids); t.append(time.perf_counter() - t0)
    best = min(t)
    print(f"rescore m={m:7d}: {best * 1e6:9.1f} us  -> {m / best / 1e6:7.2f} M candidates/s, {m * rowbytes / best / 1e9:7.1f} GB/s of gathered rows", flush=True)
ids_all = np.arange(rows, dtype=np.uint64)
mask = ix.mask_create(ids_all, (ids_all % 10 < 3).astype(np.uint8))
for name, mk in (("no filter", -1), ("30 % filter mask", mask)):
    ix.search_radius(q, 0.46, mask_id=mk)
    t = []
    for rep in range(10):
        t0 = time.perf_counter(); gi, gd, sc = ix.search_radius(q, 0.46, mask_id=mk); t.append(time.perf_counter() - t0)
    best = min(t)
    print(f"radius 0.46, {name}: {best * 1e3:7.3f} ms, {len(gi)} hits of {sc} scanned -> {rows * rowbytes / best / 1e9:7.1f} GB/s", flush=True)
t = []
for rep in range(5):
    t0 = time.perf_counter(); ix.search_topk(q, 10, mask_id=mask); t.append(time.perf_counter() - t0)
print(f"top-10 with the filter mask: {min(t) * 1e3:7.3f} ms")
