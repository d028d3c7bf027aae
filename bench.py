#!/usr/bin/env python
"""bench.py -- exact k=10 search throughput of the B200 search path (BASELINE.json metric).

Headline workload (all N): BASELINE.json configs[3] -- 10M x 768, 8-bit quantization, cosine (angular) distance, exact
k=10, synthetic uniform codes, the collection row-sharded over the N GPUs of the box (strong scaling: total rows fixed).
One "step" = one szg_search_topk call with --nq (32) queries on ONE handle that spans the N GPUs (szg_create_sharded: one
process, per-device mirrors, peer-store merge on GPU 0 -- what the Go host of the north star binds through cgo).  A call
with >= 4 queries is a batch: every device answers it with one pass of the tensor-core contraction over its shard.

  value        QPS with the query batch already resident in HBM (szg_search_topk_dev), CUDA events on GPU 0's stream,
               which the merge of every step runs on
  e2e          QPS through the host-buffer call (szg_search_topk: H2D of the queries, the searches on all devices, the
               merge, D2H of the results), copies inside the timed region
  roofline     the dominant kernel of the step.  32 queries share ONE pass over the shard, so the launch is HBM-bound:
               achieved = rows_per_gpu x 768 B / mean launch duration (CUDA events around every launch, whole timed
               region) vs MEASURED_PEAKS.json hbm_gbs.  The tensor-pipe view of the same launch is in roofline.tensor.
  single_query the same search one query per call: the memory-bound GEMV of the north star, its own HBM roofline
  batched      1024 queries per call (tensor-pipe bound), flops in SURVEY 8d units (2 B N d)
  cfg2 / cfg5 / cfg3   the other BASELINE configurations (cfg2 with an L2 flush between iterations)
  sustained    >= 2 s of back-to-back headline steps with clocks
  cpu_baseline the CPU restatement of the reference's Go scan (oracle/, "port"), bounded sample; its result on that
               sample is also the checker of the GPU result of the same queries

Under torchrun (N > 1, one rank per GPU) the ranks first run the multi-process variant of the same step (per-rank mirrors,
ONE NCCL all-gather, merge kernel: syzgydb_b200/sharded.py) as a cross-check -> "nccl_multiprocess"; then rank 0 alone
measures the in-library handle over the N GPUs (the other ranks wait on a CPU barrier) and prints the line.

`--impl reference` times the CPU restatement alone, with all host threads (the reference is Go; this image has no Go
toolchain, so oracle/_ref cannot exist -- DESIGN.md section 2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "exact_k10_qps"
UNIT = "queries/s"
SEED = 0x5A590004
CONFIGS = {
    (10_000_000, 768, 8, "cosine", 10): "BASELINE.json configs[3]",
    (10_000_000, 768, 16, "euclidean", 100): "BASELINE.json configs[4]",
    (1_000_000, 128, 4, "euclidean", 10): "BASELINE.json configs[1]",
    (1_000_000, 384, 64, "cosine", 10): "BASELINE.json configs[2]",
    (100_000, 384, 8, "cosine", 10): "BASELINE.json configs[0]",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dims", type=int, default=768)
    ap.add_argument("--quant", type=int, default=8)
    ap.add_argument("--metric", default="cosine", choices=["cosine", "euclidean"])
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=32, help="queries per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--cpu-sample-rows", type=int, default=250_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg2 / cfg5 / cfg3 / sustained legs")
    ap.add_argument("--no-nccl-crosscheck", action="store_true")
    ap.add_argument("--batch-nq", type=int, default=1024,
                    help="queries of the secondary, batched (tensor-pipe bound) measurement; 0 = skip")
    ap.add_argument("--batch-steps", type=int, default=4)
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
    ap.add_argument("--dist", default="uniform", choices=["uniform", "gaussian"],
                    help="uniform: codes uniform over the code range (device-generated, the headline); gaussian: L2-normalised "
                         "Gaussian rows and queries through the reference's quantize (all-MiniLM-like: only ~+-14 codes around "
                         "128 are used at d = 768), host-generated and uploaded; 8-bit, one GPU")
    return ap.parse_args()


def workload_name(rows, dims, quant, metric, k):
    tag = CONFIGS.get((rows, dims, quant, metric, k))
    return f"{rows}x{dims} {quant}-bit {metric} exact k={k}" + (f" ({tag})" if tag else " (not a BASELINE.json configuration)")


def gaussian_codes(seed, row0, nrows, dims):
    """Rows ~ N(0, I) L2-normalised, then quantize(v, 8) = round-half-away((v + 1) / 2 * 255) (quantization.go:5-23),
    vectorised; the same function feeds the GPU mirror and the CPU arm, so both see identical bytes."""
    import numpy as np
    v = np.random.default_rng([seed, row0]).standard_normal((nrows, dims))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    x = (np.clip(v, -1.0, 1.0) + 1.0) / 2.0 * 255.0
    r = np.floor(x)
    r += (x - r) >= 0.5
    return r.astype(np.uint8)


def rowbytes(quant, dims):
    return (dims + 1) // 2 if quant == 4 else dims * (quant // 8)


def load_json(*path):
    try:
        with open(os.path.join(ROOT, *path)) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------------- CPU arm
def cpu_arm(a, seconds: float, threads: int, steps: int | None = None, warmup: int = 0, gpu_check=None):
    """Times the oracle's restatement of Search(Precision="exact") (collection.go:569-711) on a bounded
    sample: the first S rows of the same synthetic collection, `threads` host threads each running
    independent queries (legal under the RLock, collection.go:570).  QPS over the full collection is
    the sample QPS scaled by S/rows (the scan is linear in rows).  Returns (qps_full, detail).

    gpu_check(queries) -> (ids, dist) of the GPU path over the FULL collection for the same queries: the oracle's best rows
    of the sample must be in that result or no closer than its k-th distance, with bit-identical distances."""
    import numpy as np

    from oracle import pyoracle as o
    metric = o.COSINE if a.metric == "cosine" else o.EUCLIDEAN
    S = min(a.cpu_sample_rows, a.rows)
    gauss = getattr(a, "dist", "uniform") == "gaussian"
    codes = gaussian_codes(SEED, 0, S, a.dims) if gauss else o.synth_rows(SEED, 0, S, a.dims, a.quant)
    ids = np.arange(S, dtype=np.uint64)
    order = np.arange(S, dtype=np.int64)  # ids == row index: lexicographic order precomputed once is not timed
    queries = o.synth_queries(SEED + 1, 0, max(64, threads), a.dims)
    if gauss:
        queries = np.random.default_rng(SEED + 1).standard_normal(queries.shape)
        queries /= np.linalg.norm(queries, axis=1, keepdims=True)
    o.search_exact(codes[:1000], ids[:1000], a.dims, a.quant, metric, queries[0], k=a.k, order=order[:1000])

    spans = o.Spans(codes, ids)  # the faithful variant's span-file image of the sample (built outside the timed region)

    def one_round(faithful=False):
        done = [0] * threads

        def work(t):
            o.search_exact(codes, ids, a.dims, a.quant, metric, queries[t % len(queries)], k=a.k, order=order,
                           spans=spans if faithful else None)
            done[t] = 1
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        return sum(done), time.perf_counter() - t0

    step_times = []
    if steps is None:  # time-budgeted (cpu_baseline leg)
        nqs, t = one_round()
        step_times.append(t)
        while sum(step_times) + t < seconds:
            nqs, t = one_round()
            step_times.append(t)
    else:
        for _ in range(warmup):
            one_round()
        for _ in range(steps):
            nqs, t = one_round()
            step_times.append(t)
    per_step = threads
    qps_sample = per_step * len(step_times) / sum(step_times)
    qps_full = qps_sample * S / a.rows
    # the faithful variant (span parse + CRC32 + id -> string -> map + allocation per record, like getDocument), one round
    _, tf = one_round(faithful=True)
    # one core = what a single Search call costs in the reference (one goroutine per query)
    t1 = time.perf_counter()
    o.search_exact(codes, ids, a.dims, a.quant, metric, queries[0], k=a.k, order=order)
    one_core_qps = 1.0 / (time.perf_counter() - t1) * S / a.rows
    parity = None
    if gpu_check is not None:
        nchk = 4
        gi, gd = gpu_check(queries[:nchk])
        ok = True
        for qi in range(nchk):
            ri, rd, _ = o.search_exact(codes, ids, a.dims, a.quant, metric, queries[qi], k=a.k, order=order)
            got = {int(i): float(d) for i, d in zip(gi[qi].tolist(), gd[qi].tolist())}
            kth = float(gd[qi][-1])
            for i, d in zip(ri.tolist(), rd.tolist()):
                ok = ok and ((int(i) in got and got[int(i)] == d) or d >= kth)  # bit-identical distances
            ok = ok and bool(np.all(np.diff(gd[qi]) >= 0))
        parity = (f"{nchk} queries: every one of the oracle's best {a.k} rows of the {S}-row sample is in the GPU result over all "
                  f"{a.rows} rows with a bit-identical fp64 distance, or is no closer than its k-th result") if ok else "MISMATCH"
    detail = {
        "value": qps_full, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{threads} concurrent queries x {len(step_times)} rounds over the first {S} rows "
                  f"({S * rowbytes(a.quant, a.dims) / 1e6:.0f} MB) of the same synthetic collection, oracle "
                  f"orc_search_exact (scalar fp64 restatement of the Go scan, lean variant: decode+distance+heap, no span "
                  f"parse/CRC/alloc); QPS scaled by {S}/{a.rows} rows (an extrapolation: the scan is linear in rows)",
        "qps_on_sample": qps_sample,
        "faithful_variant_qps": threads / tf * S / a.rows,
        "one_core_qps": one_core_qps,
        "gpu_parity_on_sample": parity,
        "ms_per_step": 1e3 * sum(step_times) / len(step_times),
    }
    return qps_full, detail


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    qps, det = cpu_arm(a, 0.0, threads, steps=a.steps, warmup=a.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": det["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.rows, a.dims, a.quant, a.metric, a.k), "rows": a.rows, "dims": a.dims,
                   "quantization": a.quant, "distance": a.metric, "k": a.k,
                   "note": "CPU restatement (oracle/) of the reference's Go scan; Go toolchain absent, no oracle/_ref. Each step "
                           f"scans a {min(a.cpu_sample_rows, a.rows)}-row sample with every host thread; value is scaled to the full "
                           "row count (same metric and unit, NOT the same amount of work per step as the GPU arm)"},
        "cpu_baseline": {k: det[k] for k in ("value", "unit", "cores", "kind", "sample", "faithful_variant_qps", "one_core_qps")},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.rows = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=8.0):
        """nvidia-smi takes a while to start on an 8-GPU box: the timed regions begin once it delivers samples"""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def window(self, t0: float, t1: float):
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if not (t0 <= ts <= t1 + 0.1):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}

    def stop(self):
        if not self.proc:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


# ------------------------------------------------------------------------------- GPU arm
IN_FLIGHT = int(os.environ.get("SZG_BENCH_IN_FLIGHT", "3"))  # device-resident steps in flight: consecutive steps go to alternating streams, like requests of concurrent callers


class DevRunner:
    """Device-resident calls of one handle on GPU 0, with reusable output buffers.  Steps are issued on IN_FLIGHT alternating
    streams, so that the merge / exact re-score of one step overlaps the scan of the next (each step's own launches stay
    ordered on its stream)."""

    def __init__(self, ix, torch, dev):
        self.ix, self.torch, self.dev = ix, torch, dev
        self.bufs = {}
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(IN_FLIGHT)]

    def out(self, nq, k, slot=0):
        key = (nq, k, slot)
        if key not in self.bufs:
            t = self.torch
            self.bufs[key] = (t.zeros((nq, k), dtype=t.int64, device=self.dev), t.zeros((nq, k), dtype=t.float64, device=self.dev),
                              t.zeros(nq, dtype=t.int32, device=self.dev), t.zeros(nq, dtype=t.int32, device=self.dev))
        return self.bufs[key]

    def topk(self, tq, k, batched=False, slot=0):
        nq = tq.shape[0]
        oi, od, on, of = self.out(nq, k, slot)
        st = self.streams[slot].cuda_stream
        fn = self.ix.search_batch_dev if batched else self.ix.search_topk_dev
        fn(tq.data_ptr(), nq, k, oi.data_ptr(), od.data_ptr(), on.data_ptr(), st, d_out_flags=of.data_ptr())
        return oi, od, on, of


def timed(torch, dev, fn, n, run, in_flight=IN_FLIGHT):
    """Times n steps fn(i, slot) with CUDA events; steps alternate over `in_flight` of the runner's streams (all of them start
    after the start event and the end event waits for all of them)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    w0 = time.perf_counter()
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    streams = run.streams[:in_flight]
    for s in streams:
        s.wait_event(e0)
    for i in range(n):
        fn(i, i % len(streams))
    for s in streams:
        if s is not cur:
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)
    e1.record(cur)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), w0, time.perf_counter()


def nccl_crosscheck(a, torch, dist, rank, world, local, dev):
    """The multi-process variant of the headline step: one rank per GPU, per-rank mirrors, one NCCL all-gather of the packed
    local lists, merge kernel on every rank (syzgydb_b200/sharded.py).  Short; reported beside the headline."""
    import numpy as np

    import syzgydb_b200 as szg
    from syzgydb_b200.sharded import ShardedIndex, unpack_record
    metric = szg.COSINE if a.metric == "cosine" else szg.EUCLIDEAN
    sh = ShardedIndex(a.dims, a.quant, metric, rank, world, local)
    sh.fill_synthetic(SEED, a.rows)
    steps = min(a.steps, 20)
    hq = np.random.default_rng(SEED + 1).uniform(-1.0, 1.0, size=(a.warmup + steps, a.nq, a.dims))
    dq = torch.from_numpy(hq).to(dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
    for s in range(a.warmup):
        sh.search_topk_dev(dq[s], a.k)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for s in range(a.warmup, a.warmup + steps):
        last = sh.search_topk_dev(dq[s], a.k)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ids, dd, n = unpack_record(last.cpu().numpy(), a.nq, a.k)
    chk = torch.from_numpy(ids.astype(np.int64)).to(dev)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    agree = bool(torch.equal(lo, hi))
    unc = torch.tensor([sh.uncertain_total if hasattr(sh, "uncertain_total") else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(unc, op=dist.ReduceOp.SUM)
    sh.close()
    torch.cuda.synchronize(dev)
    return {"value": a.nq * steps / (float(t.item()) / 1e3), "unit": UNIT, "steps": steps, "ranks": world,
            "ms_per_step": float(t.item()) / steps, "ranks_agree": agree, "last_ids": ids, "last_dist": dd,
            "uncertain_results_all_ranks": int(unc.item()),
            "what": "one process per GPU (torchrun), szg_search_topk_dev per rank + ONE ncclAllGather of the packed lists + "
                    "szg_merge_topk_dev on every rank; queries resident in HBM; max over ranks"}


def run_b200(a):
    import numpy as np
    import torch

    import syzgydb_b200 as szg
    from oracle import pyoracle as o  # checker + cpu_baseline leg only
    from syzgydb_b200 import _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback; use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = a.gpus
    if world > 1:
        N = a.gpus = world
    if N > torch.cuda.device_count():
        raise SystemExit(f"bench.py --gpus {N}: the box has {torch.cuda.device_count()} GPUs")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")
        if not a.no_nccl_crosscheck and a.dist == "uniform":
            nccl = nccl_crosscheck(a, torch, dist, rank, world, local, dev)
        if rank != 0:
            dist.barrier(group=cpu_group)  # rank 0 measures the in-library handle over all N GPUs meanwhile
            dist.destroy_process_group()
            return 0
    metric = szg.COSINE if a.metric == "cosine" else szg.EUCLIDEAN
    devices = list(range(N))

    def make_index(dims, quant, met):
        return szg.Index(dims, quant, met, devices=devices) if N > 1 else szg.Index(dims, quant, met, device=0)

    ix = make_index(a.dims, a.quant, metric)
    if a.dist == "gaussian":
        if N != 1 or a.quant != 8:
            raise SystemExit("bench.py --dist gaussian: 8-bit, one GPU")
        chunk = 250_000  # chunk 0 is exactly the CPU arm's sample
        ix.reserve(a.rows)
        for c0 in range(0, a.rows, chunk):
            n = min(chunk, a.rows - c0)
            ix.upsert(np.arange(c0, c0 + n, dtype=np.uint64), gaussian_codes(SEED, c0, n, a.dims))
    else:
        ix.fill_synthetic(SEED, 0, a.rows)
    rows_gpu = -(-a.rows // N)
    rb = rowbytes(a.quant, a.dims)
    peaks = load_json("MEASURED_PEAKS.json")
    mypeaks = load_json("profiles", "r02_peaks.json")
    traffic_tab = load_json("profiles", "scan_traffic.json")
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "6650 GB/s (profiling guide fallback)"
    run = DevRunner(ix, torch, dev)
    sampler = ClockSampler(0)

    # distinct queries every step, uniform(-1,1)^d, never copied from the rows
    total_steps = a.warmup + a.steps
    hq = np.random.default_rng(SEED + 1).uniform(-1.0, 1.0, size=(total_steps, a.nq, a.dims))
    if a.dist == "gaussian":
        hq = np.random.default_rng(SEED + 1).standard_normal((total_steps, a.nq, a.dims))
        hq /= np.linalg.norm(hq, axis=2, keepdims=True)
    dq = torch.from_numpy(hq).to(dev)

    def kernel_times():
        return ix.last_scan_times_ms(65536)
    sampler.wait_first()

    # ---- value: queries resident in HBM.  The library's per-launch event pairs (SZG_OPT_TIMING = 2) are off in the timed region: an
    # event record between a scan and its finalize kernel would undo their programmatic dependent launch.
    # Kernel durations for the roofline: the timed steps ONE at a time, with the library's CUDA event pair around every scan
    # launch, BEFORE the timed region and followed by a second of idle, so that both run in the same power state (a long run of
    # these steps hits the board's power cap and drops the SM clock -- the sustained leg below reports that regime).  With several
    # steps in flight the scans of consecutive steps share the SMs (row tiles are dealt to CTAs on demand), so an event pair
    # around one launch inside the timed region would also span part of its neighbour's work.
    ix.set_option(_capi.OPT_TIMING, 2)
    for s in range(max(a.warmup, 3)):
        run.topk(dq[s % total_steps], a.k, slot=0)
    torch.cuda.synchronize(dev)
    kernel_times()
    ms_serial, _, _ = timed(torch, dev, lambda i, slot: run.topk(dq[a.warmup + i], a.k, slot=slot), a.steps, run, 1)
    scan_ms = kernel_times()
    time.sleep(1.0)
    ix.set_option(_capi.OPT_TIMING, 0)
    for s in range(max(a.warmup, IN_FLIGHT)):
        run.topk(dq[s % total_steps], a.k, slot=s % IN_FLIGHT)
    torch.cuda.synchronize(dev)
    st0 = ix.stats()
    ms, w0, w1 = timed(torch, dev, lambda i, slot: run.topk(dq[a.warmup + i], a.k, slot=slot), a.steps, run)
    clocks = sampler.window(w0, w1)
    if clocks.get("sm_mhz") is None:  # the timed region is shorter than the 100 ms sampling interval: the nearest samples
        clocks = sampler.window(w0 - 0.3, w1 + 0.3)
        clocks["note"] = "timed region shorter than the 100 ms sampling interval: samples within +-0.3 s of it; sustained.clocks covers >= 2 s of the same steps"
    st1 = ix.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    tensor_served = st1["batch_queries"] - st0["batch_queries"]
    qps = a.nq * a.steps / (ms / 1e3)
    mean_scan_ms = float(np.mean(scan_ms)) if len(scan_ms) else float("nan")
    oi, od, on, of = run.out(a.nq, a.k, (a.steps - 1) % IN_FLIGHT)
    last_ids, last_dist, last_n = oi.cpu().numpy().astype(np.uint64), od.cpu().numpy(), on.cpu().numpy()
    assert (last_n == min(a.k, a.rows)).all() and (np.diff(last_dist, axis=1) >= 0).all(), "bench result check failed"
    uncertified_dev = int((of.cpu().numpy() & 1).sum())
    if nccl is not None:  # the multi-process path must have produced the same bits for its last step (same queries: same seed)
        s_last = a.warmup + nccl["steps"] - 1
        ref = run.topk(dq[s_last], a.k)
        torch.cuda.synchronize(dev)
        nccl["identical_to_in_library_handle"] = bool(np.array_equal(ref[0].cpu().numpy().astype(np.uint64), nccl.pop("last_ids")) and
                                                      np.array_equal(ref[1].cpu().numpy(), nccl.pop("last_dist")))
    ix.set_option(_capi.OPT_TIMING, 0)

    # ---- e2e: host buffers through the public call, copies inside the timed region (captured launch sequences on).
    # CALLERS host threads issue the calls, like the goroutines of concurrent Search requests under the RLock
    # (collection.go:570) -- and like the CPU arm, which runs one query per host thread; the single-caller figure is beside it.
    e2e = None
    if not a.no_e2e:
        # every e2e leg starts like the value leg: a second of idle, then warm-up calls, then the timed calls -- the same power state
        # (back-to-back legs push a single GPU into its power cap within ~0.2 s; that regime is the sustained leg's subject)
        time.sleep(1.0)
        for s in range(max(a.warmup, 3)):
            ix.search_topk(hq[s % total_steps], a.k)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for s in range(a.warmup, total_steps):
            e2e_last = ix.search_topk(hq[s], a.k)
        el1 = time.perf_counter() - t0
        assert np.array_equal(e2e_last[0], last_ids) and np.array_equal(e2e_last[1], last_dist), "host and device paths disagree"

        def concurrent(ncallers):
            """a.steps calls from ncallers host threads.  Every thread first makes untimed calls next to the others, so that each of
            the library's per-call workspaces has seen (and captured) the call shape before the clock starts."""
            results = [None] * a.steps
            gate = threading.Barrier(ncallers + 1)

            def caller(c):
                for s in range(6):
                    ix.search_topk(hq[(c + s) % total_steps], a.k)
                gate.wait()
                for s in range(c, a.steps, ncallers):
                    results[s] = ix.search_topk(hq[a.warmup + s], a.k)
            time.sleep(1.0)
            th = [threading.Thread(target=caller, args=(c,)) for c in range(ncallers)]
            [t.start() for t in th]
            gate.wait()
            t0 = time.perf_counter()
            [t.join() for t in th]
            el = time.perf_counter() - t0
            assert np.array_equal(results[-1][0], last_ids) and np.array_equal(results[-1][1], last_dist), "concurrent callers disagree"
            return el
        els = {1: el1, 2: concurrent(2), 3: concurrent(3)}
        ncall = min(els, key=els.get)
        best = els[ncall]
        e2e = {"value": a.nq * a.steps / best, "unit": UNIT, "h2d_bytes_per_step": a.nq * a.dims * 8,
               "d2h_bytes_per_step": a.nq * a.k * 16 + a.nq * 8, "ms_per_step": 1e3 * best / a.steps, "callers": ncall,
               "single_caller": {"value": a.nq * a.steps / el1, "ms_per_step": 1e3 * el1 / a.steps},
               "two_callers": {"value": a.nq * a.steps / els[2], "ms_per_step": 1e3 * els[2] / a.steps},
               "three_callers": {"value": a.nq * a.steps / els[3], "ms_per_step": 1e3 * els[3] / a.steps},
               "graph_launches": ix.stats()["graph_launches"]}

    # ---- roofline of the step's dominant kernel
    nl = max(len(scan_ms), 1)
    launches_per_step = nl / a.steps
    chunks = -(-rb // 16)
    on_tensor = tensor_served >= a.nq * a.steps
    tensor = None
    if on_tensor:
        kernel = (f"batch_kernel<{a.metric}> (tcgen05.mma kind::i8: {a.nq} queries x 2 digit planes on M = 128, 128 rows per MMA; TMA "
                  f"3-D tile loads; fused threshold top-k epilogue); one launch per device and step, {rows_gpu} rows each")
        alg_bytes = rows_gpu * rb / max(launches_per_step, 1e-9)  # 16-bit: the two byte planes together are the row's 2 bytes per code
        alg_note = ("algorithmic bytes of this launch = rows_per_gpu x getVectorSize: the queries of a step share ONE pass over the "
                    "shard (SURVEY 8d's per-query figure would count the same bytes once per query)")
        tkey = f"batch_q{a.quant}_d{a.dims}_nq{a.nq}"
        issued = 2.0 * 2 * (2 if a.quant == 16 else 1) * rows_gpu * a.dims * (-(-a.nq // 64) * 64)
        ipk = mypeaks.get("int8_tops_m128_n128")
        tensor = {"flops_8d_units": 2.0 * a.nq * rows_gpu * a.dims, "int8_ops_issued": issued,
                  "int8_tops_issued": issued / (mean_scan_ms / 1e3) / 1e12 if mean_scan_ms == mean_scan_ms else None,
                  "int8_peak_tops_measured_n128": ipk,
                  "pipe_frac": (issued / (mean_scan_ms / 1e3) / 1e12 / ipk) if (ipk and mean_scan_ms == mean_scan_ms) else None,
                  "note": "M is half empty at 32 queries (64 queries x 2 digit planes fill it): the launch is bound by HBM, not by the pipe"}
    else:
        small = a.quant <= 16 and a.k <= 24 and chunks in (2, 4, 8, 12, 16, 24, 32, 48)
        kernel = (f"scan_small_kernel<Q{a.quant}, 2 digits, {chunks} chunks>" if small else f"scan_kernel<Q{a.quant}, top-k>") + \
                 f" ({a.nq * a.steps // nl} queries x shard per launch)"
        alg_bytes = rows_gpu * rb * a.nq * a.steps / nl
        alg_note = "every query of the launch scans the shard on its own"
        tkey = f"q{a.quant}_d{a.dims}_nq{a.nq}"
    traffic, traffic_src = None, None
    if tkey in traffic_tab:
        traffic = traffic_tab[tkey]["dram_bytes_per_row_per_launch"] * rows_gpu
        traffic_src = traffic_tab[tkey]["source"]
    achieved = alg_bytes / (mean_scan_ms / 1e3) / 1e9 if mean_scan_ms == mean_scan_ms else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel, "alg_bytes_per_launch": alg_bytes,
                "alg_bytes_note": alg_note, "mean_launch_ms": mean_scan_ms, "launches_timed": int(len(scan_ms)),
                "serial_ms_per_step": ms_serial / a.steps,
                "launch_ms_note": ("mean_launch_ms: CUDA events around every scan launch of the timed steps, run one step at a time right "
                                   "before the timed region (same queries, then 1 s idle; serial_ms_per_step is that pass's step time).  In the timed region "
                                   "itself steps are in flight together and neighbouring launches share the SMs, so a per-launch event pair "
                                   "there would span part of the neighbour's work; step_floor_gbs = algorithmic bytes / ms_per_step is the "
                                   "throughput the timed region itself proves"),
                "step_floor_gbs": alg_bytes * launches_per_step / (ms / a.steps / 1e3) / 1e9,
                "launches_per_step_per_device": launches_per_step, "timed_on": "device 0's shard" if N > 1 else "the device",
                "peak_source": peak_src, "tensor": tensor}

    # ---- the same search one query per call: the memory-bound single-query case of the north star
    single_query = None
    if a.nq > 1:
        ix.set_option(_capi.OPT_TIMING, 2)
        for s in range(4):
            run.topk(dq[s % total_steps][:1], a.k, slot=s % IN_FLIGHT)
        torch.cuda.synchronize(dev)
        kernel_times()
        ms1_serial, _, _ = timed(torch, dev, lambda i, slot: run.topk(dq[a.warmup + i][:1], a.k, slot=slot), a.steps, run, 1)
        sq_scan = kernel_times()  # scan launches of the one-at-a-time run: nothing else on the GPU while they run
        ms1, _, _ = timed(torch, dev, lambda i, slot: run.topk(dq[a.warmup + i][:1], a.k, slot=slot), a.steps, run)
        kernel_times()
        ix.set_option(_capi.OPT_TIMING, 0)
        sq_ms = float(np.mean(sq_scan)) if len(sq_scan) else float("nan")
        sq_ach = rows_gpu * rb / (sq_ms / 1e3) / 1e9 if sq_ms == sq_ms else None
        lat = []
        for s in range(a.warmup + a.steps):  # host-buffer call latency (one captured launch sequence per call)
            t0 = time.perf_counter()
            ix.search_topk(hq[s % total_steps][0], a.k)
            lat.append(time.perf_counter() - t0)
        lat = lat[a.warmup:]
        single_query = {"value": a.steps / (ms1 / 1e3), "unit": UNIT, "queries_per_call": 1, "ms_per_query": ms1 / a.steps,
                        "scan_launch_ms": sq_ms, "steps_in_flight": IN_FLIGHT,
                        "one_step_at_a_time": {"value": a.steps / (ms1_serial / 1e3), "ms_per_query": ms1_serial / a.steps,
                                               "fixed_cost_us_over_scan": (ms1_serial / a.steps - sq_ms) * 1e3},
                        "host_call_latency_us_median": 1e6 * statistics.median(lat), "host_call_qps": len(lat) / sum(lat),
                        "aggregate_hbm_frac": (a.rows * rb * a.steps / (ms1 / 1e3) / 1e9) / (peak * N),
                        "roofline": {"bound": "hbm", "achieved": sq_ach, "peak": peak, "unit": "GB/s",
                                     "frac": (sq_ach / peak) if sq_ach else None,
                                     "kernel": "scan_small_kernel / scan_kernel, one query per launch: rows_per_gpu x rowbytes per launch"}}

    # ---- 1024 queries per call: the tensor-pipe-bound case
    batched = None
    if a.batch_nq > 0 and a.quant in (4, 8, 16):
        batched = batch_leg(a, np, torch, dev, ix, run, a.batch_nq, a.k, rows_gpu, a.dims, a.quant, peaks, mypeaks, a.batch_steps, o, _capi)

    # ---- sustained: back-to-back headline steps for >= --sustain-seconds
    sustained = None
    if not a.no_extras and a.sustain_seconds > 0:
        per = ms / a.steps / 1e3
        n = max(a.steps, int(a.sustain_seconds / max(per, 1e-6)) + 1)
        mss, w0, w1 = timed(torch, dev, lambda i, slot: run.topk(dq[i % total_steps], a.k, slot=slot), n, run)
        sustained = {"value": a.nq * n / (mss / 1e3), "unit": UNIT, "steps": n, "seconds": mss / 1e3, "ms_per_step": mss / n,
                     "clocks": sampler.window(w0, w1)}

    # ---- CPU restatement beside it, and its sample as the checker of the GPU result over the full collection
    cpu = None
    if not a.no_cpu_baseline:
        def gpu_check(queries):
            gi, gd, gn, _ = ix.search_topk(queries, a.k)
            return gi, gd
        _, cpu = cpu_arm(a, a.cpu_seconds, os.cpu_count() or 1, gpu_check=gpu_check)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "faithful_variant_qps", "one_core_qps",
                                   "gpu_parity_on_sample")}
    stats = ix.stats()
    ix.close()

    extras = {}
    if not a.no_extras and a.dist == "uniform":
        extras["cfg5"] = cfg5_leg(a, np, torch, dev, make_index, rows_gpu, peaks, mypeaks, o, _capi, szg)
        if N == 1:
            extras["cfg2"] = cfg2_leg(a, np, torch, dev, szg, _capi, peak)
            extras["cfg3"] = cfg3_leg(a, np, torch, dev, szg, _capi, peak, o)
    sampler.stop()

    line = {
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": N, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8" if a.quant == 8 else f"q{a.quant}",
        "data": "synthetic" if a.dist == "uniform" else "synthetic (L2-normalised Gaussian rows through the reference's quantize)",
        "config": {"workload": workload_name(a.rows, a.dims, a.quant, a.metric, a.k), "rows": a.rows, "dims": a.dims,
                   "quantization": a.quant, "distance": a.metric, "k": a.k, "queries_per_step": a.nq, "rows_per_gpu": rows_gpu,
                   "steps_in_flight": IN_FLIGHT,
                   "parallelism": (f"one handle over {N} GPUs (szg_create_sharded, one process): rows dealt to the devices, peer-store "
                                   "merge on GPU 0") if N > 1 else "one GPU",
                   "l2": f"shard payload {rows_gpu * rb / 1e6:.0f} MB per pass vs 126 MB L2: inputs larger than L2, no flush"
                         if rows_gpu * rb > 2 * 126e6 else "shard fits L2: see cfg2 for the flushed measurement"},
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "single_query": single_query, "batched": batched, "sustained": sustained,
        "cpu_baseline": cpu, "clocks": clocks, "nccl_multiprocess": nccl,
        "library": {"escalations": stats["escalations"], "uncertain_results": stats["uncertain_results"],
                    "uncertified_in_last_device_resident_step": uncertified_dev, "graph_launches": stats["graph_launches"],
                    "shards": stats["shards"], "queries_on_tensor_path_in_timed_region": int(tensor_served)},
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    return 0


def batch_leg(a, np, torch, dev, ix, run, nq, k, rows_gpu, dims, quant, peaks, mypeaks, steps, o, _capi, check_rows=None):
    """nq queries per call through szg_search_batch_dev; flops in SURVEY 8d units (2 B N d; digit and byte planes are the
    implementation's business) against the measured bf16 peak, issued int8 operations against the measured int8 peak."""
    bq_h = np.random.default_rng(SEED + 2).uniform(-1.0, 1.0, size=(nq, dims))
    if a.dist == "gaussian":
        bq_h = np.random.default_rng(SEED + 2).standard_normal((nq, dims))
        bq_h /= np.linalg.norm(bq_h, axis=1, keepdims=True)
    bq = torch.from_numpy(bq_h).to(dev)
    ix.set_option(_capi.OPT_TIMING, 2)
    for w in range(4):
        run.topk(bq, k, batched=True, slot=w % IN_FLIGHT)
    torch.cuda.synchronize(dev)
    ix.last_scan_times_ms(65536)
    b0 = ix.stats()["batch_queries"]
    # one batch at a time for the kernel's own duration (two in flight time-slice the tensor pipe) ...
    bms1, w0, w1 = timed(torch, dev, lambda i, slot: run.topk(bq, k, batched=True, slot=slot), steps, run, 1)
    kern_ms = ix.last_scan_times_ms(65536)
    ix.set_option(_capi.OPT_TIMING, 0)
    # ... and two in flight for the throughput: one batch's exact pass and merge overlap the next one's contraction
    bms, _, _ = timed(torch, dev, lambda i, slot: run.topk(bq, k, batched=True, slot=slot), steps, run)
    if bms1 < bms:
        bms = bms1
    served = ix.stats()["batch_queries"] - b0
    oi, od, on, _ = run.out(nq, k, 0)
    bi, bd = oi.cpu().numpy().astype(np.uint64), od.cpu().numpy()
    served = served // 2 if served > nq * steps else served  # both passes were counted
    # the batch returns what single-query scans return: compare a few (tensor path off for them)
    nchk = min(3, nq)
    ix.set_option(_capi.OPT_BATCH_MIN_QUERIES, 4096)
    si, sd, sn, _ = ix.search_topk(bq_h[:nchk], k)
    ix.set_option(_capi.OPT_BATCH_MIN_QUERIES, 4)
    same = bool((bi[:nchk] == si).all() and (bd[:nchk] == sd).all())
    kern_s = float(np.sum(kern_ms)) / 1e3 / max(steps, 1)  # kernel time per batch on device 0
    planes = 2 if quant == 16 else 1
    flops_8d = 2.0 * nq * rows_gpu * dims
    issued = 2.0 * 2 * planes * rows_gpu * dims * (-(-nq // 64) * 64)
    bf16 = float(peaks.get("bf16_tflops_sustained", 1405.0))
    ipk = mypeaks.get("int8_tops_m128_n128")
    rb = rowbytes(quant, dims)
    return {
        "metric": "exact_k%d_qps_batched" % k, "value": nq * steps / (bms / 1e3), "unit": UNIT, "queries_per_batch": nq,
        "ms_per_batch": bms / steps, "ms_per_batch_one_at_a_time": bms1 / steps, "steps": steps,
        "served_by_tensor_path": int(served) == nq * steps,
        "identical_to_single_query_scan": same,
        "roofline": {"bound": "tensor", "achieved": flops_8d / kern_s / 1e12 if kern_s > 0 else None, "peak": bf16, "unit": "TFLOP/s",
                     "frac": (flops_8d / kern_s / 1e12 / bf16) if kern_s > 0 else None,
                     "flops_definition": "SURVEY 8d: 2 x B x N x d per batch, per device; digit planes (2) and byte planes "
                                         "(16-bit: 2) of the exact integer formulation are not counted",
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained",
                     "int8_tops_issued": issued / kern_s / 1e12 if kern_s > 0 else None, "int8_peak_tops_measured_n128": ipk,
                     "pipe_frac": (issued / kern_s / 1e12 / ipk) if (ipk and kern_s > 0) else None,
                     "pipe_peak_source": "profiles/r02_peaks.json (tools/peaks.cu: back-to-back tcgen05.mma kind::i8 M128 N128 K32 on every SM)",
                     "kernel": "batch_kernel (tcgen05.mma kind::i8, 2 digit planes x 64 queries x 128 rows per MMA"
                               + (", high-byte and low-byte planes of the 16-bit codes)" if quant == 16 else ")"),
                     "kernel_ms_per_batch": kern_s * 1e3,
                     "hbm_floor_ms": rows_gpu * rb * planes / (float(peaks.get("hbm_gbs", 6650.0)) * 1e9) * 1e3},
    }


def cfg5_leg(a, np, torch, dev, make_index, rows_gpu, peaks, mypeaks, o, _capi, szg):
    """BASELINE.json configs[4]: 10M x 768, 16-bit, euclidean, 1024-query batch, k = 100, on the N GPUs of the run."""
    rows, dims, quant, k, nq = a.rows, 768, 16, 100, 1024
    ix = make_index(dims, quant, szg.EUCLIDEAN)
    try:
        ix.fill_synthetic(0x5A590005, 0, rows)
        run = DevRunner(ix, torch, dev)

        class A:
            dist = "uniform"
        out = batch_leg(A, np, torch, dev, ix, run, nq, k, -(-rows // a.gpus), dims, quant, peaks, mypeaks, 3, o, _capi)
        out["config"] = {"workload": workload_name(rows, dims, quant, "euclidean", k), "n_gpus": a.gpus}
        # host-buffer call (copies inside)
        hq = np.random.default_rng(SEED + 2).uniform(-1.0, 1.0, size=(nq, dims))
        ix.search_batch(hq, k)
        t0 = time.perf_counter()
        for _ in range(3):
            ix.search_batch(hq, k)
        out["e2e"] = {"value": 3 * nq / (time.perf_counter() - t0), "unit": UNIT, "h2d_bytes_per_step": nq * dims * 8,
                      "d2h_bytes_per_step": nq * k * 16 + nq * 8}
        return out
    finally:
        ix.close()


def cfg2_leg(a, np, torch, dev, szg, _capi, peak):
    """BASELINE.json configs[1]: 1M x 128, 4-bit, euclidean, single-query exact k = 10 on one GPU.  The 64 MB collection fits
    the 126 MB L2, so every timed search is preceded by a 256 MB write that evicts it; the unflushed figure is
    reported beside it and labelled for what it is."""
    rows, dims, quant, k, iters = 1_000_000, 128, 4, 10, 60
    hq = np.random.default_rng(SEED + 7).uniform(-1.0, 1.0, size=(iters + 3, dims))
    dq = torch.from_numpy(hq).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with szg.Index(dims, quant, szg.EUCLIDEAN, device=0) as ix:
        ix.fill_synthetic(0x5A590002, 0, rows)
        run = DevRunner(ix, torch, dev)
        ix.set_option(_capi.OPT_TIMING, 2)
        for s in range(3):
            run.topk(dq[s:s + 1], k)
        torch.cuda.synchronize(dev)
        ix.last_scan_times_ms(65536)
        pairs = []
        st = run.streams[0]
        with torch.cuda.stream(st):
            for s in range(iters):
                flush.fill_(s & 0xFF)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                run.topk(dq[3 + s:4 + s], k)
                e1.record(st)
                pairs.append((e0, e1))
        torch.cuda.synchronize(dev)
        per = [x.elapsed_time(y) for x, y in pairs]
        scan = ix.last_scan_times_ms(65536)
        # without the flush: L2-resident
        ms_hot, _, _ = timed(torch, dev, lambda i, slot: run.topk(dq[3 + i:4 + i], k, slot=slot), iters, run, 1)
        scan_hot = ix.last_scan_times_ms(65536)
        ix.set_option(_capi.OPT_TIMING, 0)
        lat = []
        for s in range(iters):
            flush.fill_(s & 0xFF)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            ix.search_topk(hq[3 + s], k)
            lat.append(time.perf_counter() - t0)
        lat = lat[5:]  # the first calls of the shape run launch by launch and capture the launch sequence
        payload = rows * rowbytes(quant, dims)
        sm = float(np.mean(scan))
        return {"config": {"workload": workload_name(rows, dims, quant, "euclidean", k),
                           "l2": "256 MB written to HBM before every timed search (L2 is 126 MB): the collection is read from DRAM"},
                "value": 1e3 / float(np.mean(per)), "unit": UNIT, "ms_per_query": float(np.mean(per)), "iterations": iters,
                "e2e": {"value": len(lat) / sum(lat), "unit": UNIT, "host_call_latency_us_median": 1e6 * statistics.median(lat),
                        "h2d_bytes_per_step": dims * 8, "d2h_bytes_per_step": k * 16 + 8},
                "roofline": {"bound": "hbm", "achieved": payload / (sm / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": payload / (sm / 1e3) / 1e9 / peak, "mean_launch_ms": sm,
                             "kernel": "scan_small_kernel<Q4, 2 digits, 4 chunks>, one query per launch, 64 MB per launch"},
                "l2_resident_not_hbm": {"value": iters / (ms_hot / 1e3), "unit": UNIT, "scan_launch_ms": float(np.mean(scan_hot)),
                                        "note": "no flush: the 64 MB stay in L2; NOT an HBM figure"}}


def cfg3_leg(a, np, torch, dev, szg, _capi, peak, o):
    """BASELINE.json configs[2]: 1M x 384 float64, cosine: radius 0.46 with the bucket < 3 filter (exact scan), and the gather
    re-scoring the LSH candidate path uses (m = 200 k rows for bandwidth, m = 200 for the latency of one speculative batch)."""
    rows, dims, quant = 1_000_000, 384, 64
    ids = np.arange(rows, dtype=np.uint64)
    hq = np.random.default_rng(SEED + 9).uniform(-1.0, 1.0, size=(24, dims))
    with szg.Index(dims, quant, szg.COSINE, device=0) as ix:
        ix.fill_synthetic(0x5A590003, 0, rows)
        m = ix.mask_create(ids, (ids % 10 < 3).astype(np.uint8))
        ix.set_option(_capi.OPT_TIMING, 2)
        for s in range(3):
            ix.search_radius(hq[s], 0.46, mask_id=m)
        ix.last_scan_times_ms(65536)
        t0 = time.perf_counter()
        hits = 0
        for s in range(3, 23):
            gi, gd, _ = ix.search_radius(hq[s], 0.46, mask_id=m)
            hits += gi.size
        rad = (time.perf_counter() - t0) / 20
        scan = ix.last_scan_times_ms(65536)
        res, _ = ix.search_radius_batch(hq[3:11], [0.46] * 8, mask_id=m)
        t0 = time.perf_counter()
        res, _ = ix.search_radius_batch(hq[3:11], [0.46] * 8, mask_id=m)
        rad8 = (time.perf_counter() - t0) / 8
        ix.set_option(_capi.OPT_TIMING, 0)
        payload = rows * rowbytes(quant, dims)
        sm = float(np.mean(scan))
        # gather re-scoring
        big = np.random.default_rng(1).integers(0, rows, size=200_000).astype(np.uint64)
        ix.rescore(hq[0], big)
        t0 = time.perf_counter()
        for _ in range(3):
            ix.rescore(hq[0], big)
        t_big = (time.perf_counter() - t0) / 3
        small = big[:200]
        for _ in range(5):
            ix.rescore(hq[0], small)
        lat = []
        for _ in range(50):
            t0 = time.perf_counter()
            ix.rescore(hq[0], small)
            lat.append(time.perf_counter() - t0)
        lsh = None
        try:  # Precision "medium" through the C++ host mirror (LSH walk on the host + GPU re-scoring) next to "exact": SURVEY 8f-2
            binp = os.path.join(ROOT, "tests", "cpp", "test_collection")
            out = subprocess.run([binp, "--cfg3-bench", str(rows)], capture_output=True, text=True, timeout=600)
            lsh = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            lsh["what"] = ("Collection.Search through the C++ host mirror on 1 M x 384 float64 cosine documents: Precision medium = "
                           "lshTree.search restated on the host (5 trees, leaf 100), leaf candidates re-scored by szg_rescore in "
                           "speculative batches, consider / k_counter replayed; Precision exact = one GPU scan")
        except Exception as e:  # noqa: BLE001
            lsh = {"unavailable": repr(e)[:200]}
        return {"config": {"workload": workload_name(rows, dims, quant, "cosine", 10), "radius": 0.46, "filter": "bucket < 3 (30 %)"},
                "lsh_medium_vs_exact": lsh,
                "radius_search": {"value": 1.0 / rad, "unit": UNIT, "ms_per_query_host_call": 1e3 * rad, "hits_per_query": hits / 20,
                                  "ms_per_query_in_a_batch_of_8": 1e3 * rad8,
                                  "roofline": {"bound": "hbm", "achieved": payload / (sm / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                               "frac": payload / (sm / 1e3) / 1e9 / peak, "mean_launch_ms": sm,
                                               "kernel": "scan_kernel<F64, radius>: 3.07 GB per launch",
                                               "end_to_end_frac": payload / rad / 1e9 / peak}},
                "rescore": {"m_200k_ms_host_call": 1e3 * t_big, "m_200k_gathered_gbs_host_call": 200_000 * 3072 / t_big / 1e9,
                            "m_200_call_latency_us_median": 1e6 * statistics.median(lat),
                            "note": "host-call times include the copy of ids/slots and distances; kernel-only figures: profiles/"}}


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
