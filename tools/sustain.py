"""Time series of scan launches under sustained load (per-launch ms/query + clocks)."""
import os, subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from syzgydb_b200 import _capi
rows, dims, nq, calls = int(os.environ.get("ROWS", "10000000")), 768, int(os.environ.get("NQ", "8")), int(os.environ.get("CALLS", "60"))
qs = np.random.default_rng(1).uniform(-1, 1, size=(nq, dims))
ix = szg.Index(dims, 8, szg.COSINE)
ix.fill_synthetic(7, 0, rows)
ix.set_option(_capi.OPT_TIMING, 2)
for k, v in ((_capi.OPT_SCAN_WARPS, "WARPS"), (_capi.OPT_SCAN_STAGES, "STAGES"), (_capi.OPT_SCAN_TILE_CHUNKS, "TC")):
    if v in os.environ:
        ix.set_option(k, int(os.environ[v]))
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append((time.perf_counter(), l.strip())) for l in p.stdout], daemon=True).start()
ix.search_topk(qs, 10)
ix.last_scan_times_ms()
t0 = time.perf_counter()
marks = []
for c in range(calls):
    ix.search_topk(qs, 10)
    marks.append(time.perf_counter() - t0)
t = ix.last_scan_times_ms()
time.sleep(0.2)
p.terminate()
per = t / nq
print("ms/query per launch:", " ".join(f"{x:.3f}" for x in per))
print("GB/s per launch    :", " ".join(f"{rows * dims / x / 1e6:.0f}" for x in per))
print("stats", ix.stats())
for ts, l in lines:
    print(f"{ts - t0:7.3f}s {l}")
