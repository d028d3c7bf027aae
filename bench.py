#!/usr/bin/env python
"""bench.py -- exact k=10 search throughput of the B200 search path (BASELINE.json metric).

Workload (all N): BASELINE.json configs[3] -- 10M x 768, 8-bit quantization, cosine (angular) distance,
exact k=10, synthetic uniform codes, the collection row-sharded over the N GPUs (strong scaling:
total rows fixed).  One "step" = one batch of --nq independent single-query scans (each query streams
the rank's whole shard from HBM once; one persistent scan launch walks the queries of the step back to
back), one finalize launch (merge + fp64 re-score), one all-gather of the packed local top-k lists
(N > 1) and one merge launch.

  value    QPS with the query batch already resident in HBM            (device-timed, max over ranks)
  e2e      QPS through the host-buffer call (ShardedIndex.search_topk: pinned host queries -> H2D ->
           scan -> all-gather -> merge -> D2H of the results), copies inside the timed region
  roofline achieved = rows_per_rank * 768 B / mean scan-kernel duration (CUDA events around every
           launch on the launching stream, whole timed region) vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU restatement of the reference's Go scan (oracle/, "port"), bounded sample

`--impl reference` times that CPU restatement alone, with all host threads (the reference is Go; this
image has no Go toolchain, so oracle/_ref cannot exist -- DESIGN.md section 2).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
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
    ap.add_argument("--batch-nq", type=int, default=1024,
                    help="queries of the secondary, batched (tensor-core) measurement; 0 = skip")
    ap.add_argument("--batch-steps", type=int, default=4)
    ap.add_argument("--dist", default="uniform", choices=["uniform", "gaussian"],
                    help="uniform: codes uniform over the code range (device-generated, the headline); gaussian: L2-normalised "
                         "Gaussian rows and queries through the reference's quantize (all-MiniLM-like: only ~+-14 codes around "
                         "128 are used at d = 768), host-generated and uploaded; 8-bit, one GPU")
    return ap.parse_args()


def workload_name(a):
    return f"{a.rows}x{a.dims} {a.quant}-bit {a.metric} exact k={a.k} (BASELINE.json configs[3])"


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


# ------------------------------------------------------------------------------- CPU arm
def cpu_arm(a, seconds: float, threads: int, steps: int | None = None, warmup: int = 0, gpu_check=None):
    """Times the oracle's restatement of Search(Precision="exact") (collection.go:569-711) on a bounded
    sample: the first S rows of the same synthetic collection, `threads` host threads each running
    independent queries (legal under the RLock, collection.go:570).  QPS over the full collection is
    the sample QPS scaled by S/rows (the scan is linear in rows).  Returns (qps_full, detail)."""
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

    def one_round(faithful=False):
        done = [0] * threads

        def work(t):
            o.search_exact(codes, ids, a.dims, a.quant, metric, queries[t % len(queries)], k=a.k, order=order,
                           faithful=faithful)
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
    # the faithful variant (per-record allocation like decodeVector's make([]float64)), one round
    _, tf = one_round(faithful=True)
    # one core = what a single Search call costs in the reference (one goroutine per query)
    t1 = time.perf_counter()
    o.search_exact(codes, ids, a.dims, a.quant, metric, queries[0], k=a.k, order=order)
    one_core_qps = 1.0 / (time.perf_counter() - t1) * S / a.rows
    parity = None
    if gpu_check is not None:  # the oracle as checker: the GPU scan of the same sample returns the same neighbours
        with gpu_check.Index(a.dims, a.quant, metric) as ix:
            if gauss:
                ix.upsert(ids, codes)
            else:
                ix.fill_synthetic(SEED, 0, S)
            gi, gd, gn, _ = ix.search_topk(queries[:4], a.k)
        ok = True
        for qi in range(4):
            ri, rd, _ = o.search_exact(codes, ids, a.dims, a.quant, metric, queries[qi], k=a.k, order=order)
            ok = ok and gi[qi, :gn[qi]].tolist() == ri.tolist() and bool(np.array_equal(gd[qi, :gn[qi]], rd))  # bit for bit
        parity = "ids and fp64 distances bit-identical on 4 queries" if ok else "MISMATCH"
    detail = {
        "value": qps_full, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{threads} concurrent queries x {len(step_times)} rounds over the first {S} rows "
                  f"({S * rowbytes(a.quant, a.dims) / 1e6:.0f} MB) of the same synthetic collection, oracle "
                  f"orc_search_exact (scalar fp64 restatement of the Go scan, lean variant: decode+distance+heap, no span "
                  f"parse/CRC/alloc); QPS scaled by {S}/{a.rows} rows",
        "qps_on_sample": qps_sample,
        "faithful_alloc_variant_qps": threads / tf * S / a.rows,
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
        "config": {"workload": workload_name(a), "rows": a.rows, "dims": a.dims, "quantization": a.quant,
                   "distance": a.metric, "k": a.k,
                   "note": "CPU restatement (oracle/) of the reference's Go scan; Go toolchain absent, no oracle/_ref"},
        "cpu_baseline": {k: det[k] for k in ("value", "unit", "cores", "kind", "sample")},
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

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
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


# ------------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import syzgydb_b200 as szg
    from syzgydb_b200 import _capi
    from syzgydb_b200.sharded import ShardedIndex, record_layout, unpack_record

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback; use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        a.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    metric = szg.COSINE if a.metric == "cosine" else szg.EUCLIDEAN
    sh = ShardedIndex(a.dims, a.quant, metric, rank, world, local)
    if a.dist == "gaussian":
        if world != 1 or a.quant != 8:
            raise SystemExit("bench.py --dist gaussian: 8-bit, one GPU")
        r0, r1, chunk = 0, a.rows, 250_000  # chunk 0 is exactly the CPU arm's sample
        sh.shard.index.reserve(a.rows)
        for c0 in range(0, a.rows, chunk):
            n = min(chunk, a.rows - c0)
            sh.shard.index.upsert(np.arange(c0, c0 + n, dtype=np.uint64), gaussian_codes(SEED, c0, n, a.dims))
        sh.total_rows = a.rows
    else:
        r0, r1 = sh.fill_synthetic(SEED, a.rows)
    my_rows = r1 - r0
    ix = sh.shard.index
    ix.set_option(_capi.OPT_TIMING, 2)
    rb = rowbytes(a.quant, a.dims)

    # distinct queries every step, uniform(-1,1)^d, never copied from the rows (same on every rank)
    total_steps = a.warmup + a.steps
    hq = np.random.default_rng(SEED + 1).uniform(-1.0, 1.0, size=(total_steps, a.nq, a.dims))
    if a.dist == "gaussian":
        hq = np.random.default_rng(SEED + 1).standard_normal((total_steps, a.nq, a.dims))
        hq /= np.linalg.norm(hq, axis=2, keepdims=True)
    dq = torch.from_numpy(hq).to(dev)
    hq_pinned = torch.from_numpy(hq).pin_memory()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- value: queries resident in HBM
    for s in range(a.warmup):
        sh.search_topk_dev(dq[s], a.k)
    sync_all()
    ix.last_scan_times_ms()  # drain warm-up events
    launches0 = ix.stats()["kernel_launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    last = None
    for s in range(a.warmup, total_steps):
        last = sh.search_topk_dev(dq[s], a.k)
    e1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms = e0.elapsed_time(e1)
    scan_ms = ix.last_scan_times_ms(65536)
    launches = ix.stats()["kernel_launches"] - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    qps = a.nq * a.steps / (ms_max / 1e3)
    mean_scan_ms = float(np.mean(scan_ms)) if len(scan_ms) else float("nan")
    stats = ix.stats()

    # sanity: the last step's results are k ascending distances per query, identical on every rank
    ids, dd, n = unpack_record(last.cpu().numpy(), a.nq, a.k)
    assert (n == min(a.k, a.rows)).all() and (np.diff(dd, axis=1) >= 0).all(), "bench result check failed"
    if world > 1:
        chk = torch.from_numpy(ids.astype(np.int64)).to(dev)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the merged result"

    # ---- e2e: host buffers through the public call, copies inside the timed region
    e2e = None
    if not a.no_e2e:
        for s in range(a.warmup):
            sh.search_topk(hq[s], a.k)
        sync_all()
        t0 = time.perf_counter()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for s in range(a.warmup, total_steps):
            sh.search_topk(hq_pinned[s].numpy(), a.k)
        c1.record()
        torch.cuda.synchronize(dev)
        el = max(c0.elapsed_time(c1) / 1e3, time.perf_counter() - t0)
        t = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        _, _, _, words = record_layout(a.nq, a.k)
        e2e = {"value": a.nq * a.steps / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": a.nq * a.dims * 8,
               "d2h_bytes_per_step": (a.nq * a.k * 16 + a.nq * 8) if world == 1 else words * 8}

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass

    # ---- secondary: the same searches as ONE batch on the tensor cores (szg_search_batch_dev: tcgen05 kind::i8
    #      contraction + fused top-k), then the same all-gather + merge.  Reported beside the headline, not as it.
    batched = None
    if a.batch_nq > 0 and a.quant in (8, 16):
        bq_h = np.random.default_rng(SEED + 2).uniform(-1.0, 1.0, size=(a.batch_nq, a.dims))
        if a.dist == "gaussian":
            bq_h = np.random.default_rng(SEED + 2).standard_normal((a.batch_nq, a.dims))
            bq_h /= np.linalg.norm(bq_h, axis=1, keepdims=True)
        bq = torch.from_numpy(bq_h).to(dev)
        for _ in range(3):
            outb = sh.search_topk_dev(bq, a.k, batched=True)
        sync_all()
        ix.last_scan_times_ms(65536)
        bq0 = ix.stats()["batch_queries"]
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(a.batch_steps):
            outb = sh.search_topk_dev(bq, a.k, batched=True)
        b1.record()
        sync_all()
        bms = b0.elapsed_time(b1)
        kern_ms = ix.last_scan_times_ms(65536)
        tb = torch.tensor([bms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        bms = float(tb.item())
        served = ix.stats()["batch_queries"] - bq0
        # the batch returns what single queries return: compare a few with the streaming scan
        nchk = min(8, a.batch_nq)
        single = sh.search_topk_dev(bq[:nchk].contiguous(), a.k)
        torch.cuda.synchronize(dev)
        bi, bd, bn = unpack_record(outb.cpu().numpy(), a.batch_nq, a.k)
        si, sd, sn = unpack_record(single.cpu().numpy(), nchk, a.k)
        same = bool((bi[:nchk] == si).all() and (bd[:nchk] == sd).all() and (bn[:nchk] == sn).all())
        kern_s = float(np.sum(kern_ms)) / 1e3 / max(a.batch_steps, 1)  # batch_kernel time per batch on this rank
        byte_planes = 2 if a.quant == 16 else 1  # 16-bit rows are contracted as a high-byte and a low-byte plane
        ops = 2.0 * 2 * byte_planes * my_rows * a.dims * (-(-a.batch_nq // 64) * 64)  # 2 digit planes, M padded to 64-query groups
        tpeak = 2.0 * float(peaks.get("bf16_tflops_sustained", 1405.0))
        batched = {
            "metric": "exact_k%d_qps_batched" % a.k, "value": a.batch_nq * a.batch_steps / (bms / 1e3), "unit": UNIT,
            "queries_per_batch": a.batch_nq, "ms_per_batch": bms / a.batch_steps, "steps": a.batch_steps,
            "served_by_tensor_path": int(served) == a.batch_nq * a.batch_steps,
            "identical_to_single_query_scan": same,
            "roofline": {"bound": "tensor", "achieved": ops / kern_s / 1e12 if kern_s > 0 else None, "peak": tpeak,
                         "unit": "TOP/s (int8)", "frac": (ops / kern_s / 1e12 / tpeak) if kern_s > 0 else None,
                         "kernel": "batch_kernel (tcgen05.mma kind::i8, 2 digit planes x 64 queries x 128 rows per MMA"
                                   + (", high-byte and low-byte planes of the 16-bit codes)" if a.quant == 16 else ")"),
                         "kernel_ms_per_batch": kern_s * 1e3,
                         "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops_sustained (int8 dense = 2 x bf16 dense on B200)",
                         "hbm_floor_ms": my_rows * rb / (float(peaks.get("hbm_gbs", 6650.0)) * 1e9) * 1e3},
        }
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    # algorithmic bytes of one scan launch: every query of the launch scans the shard's payload once
    nl = max(len(scan_ms), 1)
    alg_bytes = my_rows * rb * a.nq * a.steps / nl
    chunks = -(-rb // 16)
    small = (a.quant <= 16 and a.k <= 24 and chunks in (2, 4, 8, 12, 16, 24, 32, 48) and os.environ.get("SZG_SCAN_SMALL", "1") != "0")
    kernel = (f"scan_small_kernel<Q{a.quant}, 2 digits, {chunks} chunks> (queries dealt to CTA groups; {a.nq * a.steps // nl} queries x shard per launch)"
              if small else f"scan_kernel<Q{a.quant}, top-k> ({a.nq * a.steps // nl} queries x shard per launch)")
    # DRAM traffic of that launch from the committed ncu --set full capture of the same kernel, row shape and queries per
    # launch (dram__bytes_read.sum + dram__bytes_write.sum), scaled by rows
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            tr = json.load(f)
        key = f"q{a.quant}_d{a.dims}_nq{a.nq}" if small else f"q{a.quant}_d{a.dims}"
        if key in tr and "dram_bytes_per_row_per_launch" in tr[key]:
            traffic = tr[key]["dram_bytes_per_row_per_launch"] * my_rows
            traffic_src = tr[key]["source"]
        elif key in tr:
            traffic = tr[key]["dram_bytes_per_row_per_query"] * my_rows * a.nq
            traffic_src = tr[key]["source"]
    except Exception:
        pass
    achieved = alg_bytes / (mean_scan_ms / 1e3) / 1e9 if mean_scan_ms == mean_scan_ms else None
    dram_gbs = traffic / (mean_scan_ms / 1e3) / 1e9 if (traffic and mean_scan_ms == mean_scan_ms) else None
    note = None
    if small and a.nq > 1:
        note = ("achieved counts every query's scan of the shard (algorithmic bytes); the queries of a launch run on different "
                "CTA groups at the same time and find each other's rows in L2, so the DRAM traffic of the launch (traffic, "
                "dram_achieved) is a fraction of it and frac can exceed 1.  The launch is bound by the L1 data pipe, not by HBM "
                "(profiles/r01b_scan_small_q8_ncu_full.csv); single_query below is the HBM-bound case")

    # ---- the same scan one query per launch: the memory-bound single-query case of the north star
    single_query = None
    if a.nq > 1:
        for s in range(min(3, total_steps)):
            sh.search_topk_dev(dq[s][:1].contiguous(), a.k)
        sync_all()
        ix.last_scan_times_ms(65536)
        e0.record()
        for s in range(a.warmup, total_steps):
            sh.search_topk_dev(dq[s][:1].contiguous(), a.k)
        e1.record()
        sync_all()
        t1q = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t1q, op=dist.ReduceOp.MAX)
        sq_scan = ix.last_scan_times_ms(65536)
        sq_ms = float(np.mean(sq_scan)) if len(sq_scan) else float("nan")
        sq_ach = my_rows * rb / (sq_ms / 1e3) / 1e9 if sq_ms == sq_ms else None
        single_query = {"value": a.steps / (float(t1q.item()) / 1e3), "unit": UNIT, "queries_per_launch": 1,
                        "ms_per_query": float(t1q.item()) / a.steps, "scan_launch_ms": sq_ms,
                        "roofline": {"bound": "hbm", "achieved": sq_ach, "peak": peak, "unit": "GB/s",
                                     "frac": (sq_ach / peak) if sq_ach else None}}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        _, cpu = cpu_arm(a, a.cpu_seconds, os.cpu_count() or 1, gpu_check=szg)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "faithful_alloc_variant_qps", "one_core_qps",
                                   "gpu_parity_on_sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8" if a.quant == 8 else f"q{a.quant}",
            "data": "synthetic" if a.dist == "uniform" else "synthetic (L2-normalised Gaussian rows through the reference's quantize)",
            "config": {"workload": workload_name(a), "rows": a.rows, "dims": a.dims, "quantization": a.quant,
                       "distance": a.metric, "k": a.k, "queries_per_step": a.nq, "rows_per_gpu": my_rows,
                       "parallelism": f"row-sharded x{world}, one all-gather + merge per step",
                       "l2": f"shard payload {my_rows * rb / 1e6:.0f} MB per query vs 126 MB L2: inputs larger than L2, no flush"
                             if my_rows * rb > 2 * 126e6 else "WARNING: shard fits L2"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                         "dram_achieved": dram_gbs, "dram_frac": (dram_gbs / peak) if dram_gbs else None,
                         "kernel": kernel, "alg_bytes_per_launch": alg_bytes,
                         "mean_launch_ms": mean_scan_ms, "launches_timed": int(len(scan_ms)), "peak_source": peak_src,
                         "note": note},
            "single_query": single_query,
            "cpu_baseline": cpu, "clocks": clocks, "batched": batched,
            "library": {"escalations": stats["escalations"], "uncertain_results": stats["uncertain_results"],
                        "scan_grid": stats["scan_grid"], "scan_block": stats["scan_block"]},
        }
        print(json.dumps(line), flush=True)
    sh.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
