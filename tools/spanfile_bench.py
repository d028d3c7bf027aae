"""Load-path measurement of the direct span-file reader: writes a collection file the way the
reference does (oracle/spanfile.py), then times szg_spanfile_open (walk + CRC32 + parse + merge, all host threads)
and szg_spanfile_load (gather + bulk upsert into HBM)."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import syzgydb_b200 as szg
from oracle import pyoracle as o
from oracle import spanfile as sfo

n = int(os.environ.get("ROWS", "200000")); dims = int(os.environ.get("DIMS", "768")); bits = 8
t0 = time.time()
w = sfo.SpanFileWriter()
w.write_header("bench", 1, dims, bits)
codes = o.synth_rows(11, 0, n, dims, bits)
meta = b'{"bucket": 3}'
for i in range(n):
    w.add_document(i, codes[i].tobytes(), meta)
data = w.tobytes()
path = os.path.join(tempfile.gettempdir(), "szg_bench.dat")
with open(path, "wb") as f:
    f.write(data)
print(f"fixture: {n} x {dims} 8-bit documents, {len(data) / 1e6:.1f} MB, written in {time.time() - t0:.1f} s (python restatement of the writer)")
for rep in range(3):
    t0 = time.perf_counter()
    sf = szg.SpanFile(path)
    t1 = time.perf_counter()
    info = sf.info()
    ix = szg.Index(dims, bits, szg.COSINE)
    t2 = time.perf_counter()
    loaded = sf.load_into(ix)
    ix.count()
    t3 = time.perf_counter()
    print(f"rep {rep}: open+scan {1e3 * (t1 - t0):.1f} ms ({len(data) / (t1 - t0) / 1e9:.2f} GB/s of file, {info['records'] / (t1 - t0) / 1e6:.2f} M records/s), "
          f"load into HBM {1e3 * (t3 - t2):.1f} ms ({loaded / (t3 - t2) / 1e6:.2f} M rows/s, {loaded * dims / (t3 - t2) / 1e9:.2f} GB/s of codes); "
          f"host threads {os.cpu_count()}")
    q = o.synth_queries(5, 0, 1, dims)
    gi, gd, gn, _ = ix.search_topk(q, 10)
    ri, rd, _ = o.search_exact(codes[:20000], np.arange(20000, dtype=np.uint64), dims, bits, szg.COSINE, q[0], k=10) if rep == 0 else (None, None, None)
    ix.close(); sf.close()
print("first result ids", gi[0].tolist())
