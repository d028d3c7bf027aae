// search.cu -- host side of every search entry point of include/syzgy_b200.h on ONE device: launch planning, the top-k
// pipeline (prep -> scan | tensor-core batch -> finalize), certification / escalation, captured launch sequences (CUDA
// graphs), combining of concurrent callers, radius search, candidate re-scoring and the cross-shard merge.  Host logic
// only; the kernels live in scan_*.cu, batch_q8.cu, kernels.cu.  A handle made by szg_create_sharded is routed to
// sharded.cu at the top of every entry point.
#include <cmath>

#include "index_internal.h"
#include "scan_small.cuh"
#include "sharded.h"

using namespace szg;

namespace szg {

size_t pq_stride(const szg_index *h, int nd) {
    size_t payload = ((size_t)h->C * pq_bytes_per_chunk(h->qt, nd) + 15) / 16 * 16;
    return sizeof(PQHeader) + payload;
}
// digits of the first pass: 2 (fast) for quantized rows unless SZG_OPT_DIGITS forces 3
int first_digits(const szg_index *h) { return (h->qt <= Q16 && h->digits != 3) ? 2 : 3; }

constexpr size_t kCandBytes = 64u << 20;       // candidate lists of one scan launch (bounds queries per launch)
constexpr size_t kScanSmemLimit = 224 * 1024; // dynamic; + ~3 KB static stays under the 227 KB CTA limit

// Persistent launch: one CTA per SM (fewer when the collection has fewer row blocks than warps).
int plan_scan(szg_index *h, int nd, ScanPlan *p, int *grid) {
    // measured on B200 (profiles/r01_tune_scan_*): 16 warps x 4 KB tiles win on multi-GB shards (7.29 vs 7.04 TB/s at
    // 7.7 GB), 8 warps x 8 KB tiles on ~1 GB shards (7.41 vs 7.04 TB/s at 0.96 GB: half as many per-warp lists per query)
    uint32_t warps = (uint32_t)h->scan_warps, tile_chunks = (uint32_t)h->scan_tile_chunks;
    if (!h->scan_geometry_set && h->qt == Q8 && (uint64_t)h->nslots * h->rowbytes < 1500000000ull && h->C >= 16) {
        warps = 8;
        tile_chunks = 16;
    }
    if (!scan_plan(h->C, warps, (uint32_t)h->scan_stages, tile_chunks, pq_stride(h, nd), kScanSmemLimit, p, h->qt >= F32 ? 8u : 1u))
        return fail(SZG_EINTERNAL, "scan geometry does not fit shared memory");
    const uint32_t nblk = (h->nslots + 31) / 32;
    uint32_t g = (nblk + p->warps - 1) / p->warps;
    g = std::max<uint32_t>(1, std::min<uint32_t>(g, (uint32_t)h->sm_count));
    *grid = (int)g;
    return SZG_OK;
}

void fill_scan_args(szg_index *h, ScanArgs &a, const uint32_t *mask) {
    memset(&a, 0, sizeof a);
    a.codes = h->codes.p;
    a.aux = h->aux.p;
    a.live = h->live.p;
    a.mask = mask;
    a.ids = h->ids.p;
    a.lut = h->lut.p;
    a.C = h->C;
    a.nblk = (h->nslots + 31) / 32;
    a.dims = (uint32_t)h->dim;
    a.metric = (uint32_t)h->metric;
}

// ---- CUDA events around the scan launches of a call (SZG_OPT_TIMING: 1 keeps the last call's, 2 accumulates).  The events
// belong to the workspace (one call at a time); finished ones are moved to h->last_times under the handle's mutex.
static int timing_reserve(szg_index *h, Workspace *ws, uint32_t nlaunch, uint32_t *tbase, bool *on) {
    *tbase = 0;
    *on = h->timing != 0 && !ws->capturing;
    if (!*on) return SZG_OK;
    if (h->timing == 2 && ws->timed + nlaunch <= 65536) *tbase = ws->timed;
    while (ws->t0.size() < *tbase + nlaunch) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        ws->t0.push_back(a);
        ws->t1.push_back(b);
    }
    return SZG_OK;
}

void drain_timing(szg_index *h, Workspace *ws) {
    if (!ws->timed) return;
    std::vector<float> ms(ws->timed, 0.f);
    for (uint32_t i = 0; i < ws->timed; ++i) {
        if (cudaEventSynchronize(ws->t1[i]) != cudaSuccess) { ms.resize(i); break; }
        cudaEventElapsedTime(&ms[i], ws->t0[i], ws->t1[i]);
    }
    ws->timed = 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->timing != 2) h->last_times.clear();
    if (h->last_times.size() + ms.size() <= 65536) h->last_times.insert(h->last_times.end(), ms.begin(), ms.end());
}

static void fill_finalize_args(szg_index *h, FinalizeArgs &f, size_t stride, uint32_t k, uint32_t flags, const PeerSink *sink) {
    memset(&f, 0, sizeof f);
    f.codes = h->codes.p; f.ids = h->ids.p; f.lut = h->lut.p;
    f.pq_stride = stride;
    f.C = h->C; f.dims = (uint32_t)h->dim; f.metric = (uint32_t)h->metric; f.k = k;
    f.flags = flags & SZG_F_NO_FP64_VERIFY;
    f.done_cnt = sink ? sink->done_cnt : nullptr;
    f.trace = h->trace;
}

// run_topk for short rows: the launch is cut in (query, part) items handled by one CTA each (scan_small.cuh)
int run_topk_small(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags,
                   int nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags,
                   const PeerSink *sink) {
    const size_t stride = pq_stride(h, nd);
    int rc;
    if ((rc = ws->d_pq.ensure(stride * nq))) return rc;
    const uint32_t nblk = (h->nslots + 31) / 32;
    const uint32_t sms = (uint32_t)h->sm_count;
    cudaStream_t main = ws->main;
    PrepArgs pa;
    pa.queries = d_q; pa.pq = ws->d_pq.p; pa.pq_stride = stride;
    pa.dims = (uint32_t)h->dim; pa.C = h->C; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
    pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 0; pa.radius = 0.0;
    CK(launch_prep(nq, main, pa));
    h->launches++;
    // queries per launch: bounded by the candidate buffer (parts <= SM count lists of 16 warps x 32 keys per query)
    // ... and by the constant window the prepared queries of a launch go through (SZG_SMALL_CONST=0: shared memory instead)
    // measured (profiles/r01b_scan_small_vs_general.log): +5..7 % at 48 chunks, -5..10 % on rows of <= 8 chunks, nothing at 24.
    // The window is one per quantization and device (launches that use it are chained by an event), so it is used only where
    // it pays: launches of several queries (a launch of one or two is HBM-bound either way), never while a launch sequence
    // is being captured.
    static const int const_env = getenv("SZG_SMALL_CONST") ? atoi(getenv("SZG_SMALL_CONST")) : -1;
    const bool const_ok = !ws->capturing && nq >= 4 && (const_env >= 0 ? const_env != 0 : h->C >= 32);
    const size_t window_q = std::max<size_t>(1, (size_t)kConstSlots * 16 / stride);
    const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>({(size_t)nq, (size_t)4096, kCandBytes / ((size_t)sms * kSmallWarps * 32 * 8),
                                                                             const_ok ? window_q : (size_t)4096}));
    const uint32_t nlaunch = (nq + chunk - 1) / chunk;
    uint32_t tbase = 0;
    bool timing = false;
    if ((rc = timing_reserve(h, ws, nlaunch, &tbase, &timing))) return rc;
    ScanArgs a;
    fill_scan_args(h, a, mask);
    a.pq_stride = stride;
    FinalizeArgs f;
    fill_finalize_args(h, f, stride, k, flags, sink);
    for (uint32_t l = 0, q0 = 0; q0 < nq; q0 += chunk, ++l) {
        const uint32_t m = std::min(chunk, nq - q0);
        // parts per query: as many as keep every CTA busy, but no part shorter than one block per warp
        const uint32_t qper = 1; // queries per warp (pairs were measured slower: see scan_small.cuh)
        // warp groups per CTA, each on another query of the same row part (they share the rows in L1)
        static const int wg_env = getenv("SZG_SMALL_WG") ? atoi(getenv("SZG_SMALL_WG")) : 0;
        // measured (profiles/r01b_scan_small_vs_general.log): two groups win at 48 chunks (cfg4 2390 -> 2795 QPS, and 340 W
        // instead of 460 W: the board no longer throttles) and on collections of a few MB; one group wins in between
        uint32_t wgroups = wg_env == 1 || wg_env == 2 || wg_env == 4 ? (uint32_t)wg_env : ((h->C >= 32 || nblk < 8192) ? 2u : 1u);
        while (wgroups > 1 && m < 2 * wgroups) wgroups >>= 1; // too few queries to fill the groups of several CTAs
        const uint32_t gw = kSmallWarps / wgroups;
        const uint32_t groups = (m + wgroups - 1) / wgroups;
        uint32_t parts = small_parts(groups, sms);
        parts = std::max<uint32_t>(1, std::min<uint32_t>(parts, (nblk + gw - 1) / gw));
        const uint32_t nlists = parts; // the warps of a group merge their lists before writing
        const int grid = (int)std::min<uint64_t>((uint64_t)groups * parts, sms);
        a.qper = qper;
        a.wgroups = wgroups;
        a.const_queries = const_ok ? 1u : 0u;
        if ((rc = ws->d_cand.ensure((size_t)m * nlists * 32))) return rc;
        a.pq = ws->d_pq.p + stride * q0;
        a.nq = m;
        a.parts = parts;
        static const int adj_env = getenv("SZG_SMALL_ADJ") ? atoi(getenv("SZG_SMALL_ADJ")) : -1;
        a.adjacent = adj_env >= 0 ? (uint32_t)adj_env : (h->C < 48 ? 1u : 0u);
        a.cand = ws->d_cand.p;
        if (timing) CK(cudaEventRecord(ws->t0[tbase + l], main));
        CK(launch_scan_small(h->qt, nd, h->C, grid, stride * wgroups, main, a));
        if (timing) CK(cudaEventRecord(ws->t1[tbase + l], main));
        f.cand = ws->d_cand.p;
        f.nlists = nlists;
        f.queries = d_q + (size_t)q0 * h->dim;
        f.pq = a.pq;
        f.out_ids = d_out_ids + (size_t)q0 * k; f.out_dist = d_out_dist + (size_t)q0 * k;
        f.out_n = d_out_n + q0; f.out_flags = d_out_flags + q0;
        if (f.done_cnt) f.done_cnt = sink->done_cnt + q0;
        CK(launch_finalize(h->qt, 0, m, main, f));
        h->launches += 2;
    }
    if (timing) ws->timed = tbase + nlaunch;
    return SZG_OK;
}

// Enqueues prep + scan + finalize for nq queries on ws->main.  One scan launch serves a whole chunk of
// queries (persistent warps walk query after query); the chunk size is bounded by the candidate
// buffer.  Inputs/outputs are device pointers; `ws` supplies scratch.
int run_topk(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask,
             uint32_t flags, int mode, int nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
             uint32_t *d_out_flags, const PeerSink *sink) {
    const size_t stride = pq_stride(h, nd);
    int rc;
    if ((rc = ws->d_pq.ensure(stride * nq))) return rc;
    ScanPlan plan;
    int grid = 0;
    if ((rc = plan_scan(h, nd, &plan, &grid))) return rc;
    const size_t Kp = 32u << mode;
    // short rows, k <= 24: the kernel of scan_small.cuh (SZG_SCAN_SMALL=0 keeps the general kernel, for comparisons)
    static const bool small_ok = !(getenv("SZG_SCAN_SMALL") && atoi(getenv("SZG_SCAN_SMALL")) == 0);
    static const uint32_t small_maxc = getenv("SZG_SCAN_SMALL_MAXC") ? (uint32_t)atoi(getenv("SZG_SCAN_SMALL_MAXC")) : 48u;
    const bool small = small_ok && mode == 0 && !h->scan_geometry_set && scan_small_supported(h->qt, h->C) && h->C <= small_maxc &&
                       h->nslots >= 32;
    if (small) return run_topk_small(h, ws, d_q, nq, k, mask, flags, nd, d_out_ids, d_out_dist, d_out_n, d_out_flags, sink);
    // one query per launch: the warps of a CTA merge their lists in the (idle) ring memory before writing them
    const bool cta_merge = nq == 1 && plan.smem >= (size_t)plan.warps * Kp * 8 && (plan.warps & (plan.warps - 1)) == 0;
    const size_t nlists = cta_merge ? (size_t)grid : (size_t)grid * plan.warps;
    const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>({(size_t)nq, (size_t)4096, kCandBytes / (nlists * Kp * 8)}));
    if ((rc = ws->d_cand.ensure((size_t)chunk * nlists * Kp))) return rc;

    cudaStream_t main = ws->main;
    PrepArgs pa;
    pa.queries = d_q; pa.pq = ws->d_pq.p; pa.pq_stride = stride;
    pa.dims = (uint32_t)h->dim; pa.C = h->C; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
    pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 0; pa.radius = 0.0;
    CK(launch_prep(nq, main, pa));
    h->launches++;
    const uint32_t nlaunch = (nq + chunk - 1) / chunk;
    uint32_t tbase = 0;
    bool timing = false;
    if ((rc = timing_reserve(h, ws, nlaunch, &tbase, &timing))) return rc;
    ScanArgs a;
    fill_scan_args(h, a, mask);
    a.Ct = plan.Ct; a.stages = plan.stages; a.pq_smem_off = plan.pq_smem_off;
    a.pq_stride = stride;
    a.cand = ws->d_cand.p;
    a.cta_merge = cta_merge ? 1u : 0u;
    FinalizeArgs f;
    fill_finalize_args(h, f, stride, k, flags, sink);
    f.cand = ws->d_cand.p;
    f.nlists = (uint32_t)nlists;
    for (uint32_t l = 0, q0 = 0; q0 < nq; q0 += chunk, ++l) {
        const uint32_t m = std::min(chunk, nq - q0);
        a.pq = ws->d_pq.p + stride * q0;
        a.nq = m;
        if (timing) CK(cudaEventRecord(ws->t0[tbase + l], main));
        CK(launch_scan(h->qt, mode, nd, grid, (int)plan.warps * 32, plan.smem, main, a));
        if (timing) CK(cudaEventRecord(ws->t1[tbase + l], main));
        // merge of the per-warp lists, fp64 re-score, ordered output: one CTA per query
        f.queries = d_q + (size_t)q0 * h->dim;
        f.pq = a.pq;
        f.out_ids = d_out_ids + (size_t)q0 * k; f.out_dist = d_out_dist + (size_t)q0 * k;
        f.out_n = d_out_n + q0; f.out_flags = d_out_flags + q0;
        if (f.done_cnt) f.done_cnt = sink->done_cnt + q0;
        CK(launch_finalize(h->qt, mode, m, main, f));
        h->launches += 2;
    }
    if (timing) ws->timed = tbase + nlaunch;
    return SZG_OK;
}

// ---- batched queries on the tensor cores (batch_q8.cu)
struct BatchPlan {
    uint32_t slice, stages, keep, nranges, gpl, ngroups, Cb;
    uint32_t mth, gm_sp, gm_stride; // published bound keys: mth-best row per range; words per group / per query (batch_q8.cu)
    int mode;
    bool p16, p4;
};

// true when the tensor-core path can serve (collection, k, candidate mode): 4/8/16-bit rows, an even number of 16-byte
// chunks that fits the TMEM columns reserved for the query digits, candidate lists of at most 128 keys, 2-digit queries
static bool plan_batch(const szg_index *h, uint32_t nq, uint32_t k, int mode, BatchPlan *p) {
    if (h->qt > Q16 || h->digits == 3 || k < 1 || nq < 1 || h->batch_disabled || h->live_rows == 0) return false;
    // chunks of the contraction operand (16 dimensions each): the 8-bit row itself, one byte plane of a 16-bit row,
    // or the one-byte-per-code copy of a 4-bit row
    p->p16 = h->qt == Q16;
    p->p4 = h->qt == Q4;
    p->Cb = h->qt == Q8 ? h->C : (uint32_t)(h->dim + 15) / 16;
    if ((p->Cb % 2) != 0 || p->Cb > batch_max_chunks()) return false;
    p->mode = mode; // candidates per list = 32 << mode, as in the streaming scan
    if (p->mode > 2) return false;
    p->keep = 32u << p->mode;
    const size_t ring_limit = batch_dynamic_limit();
    if (ring_limit <= batch_list_bytes(p->keep)) return false;
    const size_t stage_limit = ring_limit - batch_list_bytes(p->keep);
    static const uint32_t want_slice = getenv("SZG_BATCH_SLICE") ? (uint32_t)atoi(getenv("SZG_BATCH_SLICE")) : 0u;
    p->slice = batch_slice_chunks(p->Cb, want_slice, stage_limit);
    p->stages = batch_stages(p->slice, stage_limit);
    if (p->stages < 2) return false;
    p->ngroups = (nq + 63) / 64;
    p->gpl = std::min<uint32_t>(p->ngroups, 16); // query groups per launch (they share the L2 copy of a row range)
    const uint32_t nblk = (h->nslots + 31) / 32;
    p->nranges = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)h->sm_count / p->gpl, (nblk + 3) / 4));
    // ranges are dealt into need = ceil(keep / mth) groups; a group's published keys are read with at most two 16-byte loads
    for (;; --p->nranges) {
        p->mth = (p->keep + p->nranges - 1) / p->nranges;
        const uint32_t need = (p->keep + p->mth - 1) / p->mth;
        p->gm_sp = ((p->nranges + need - 1) / need + 3) & ~3u;
        p->gm_stride = need * p->gm_sp;
        if (p->gm_sp <= 8) break;
    }
    return true;
}

// prep -> batch_kernel (one launch per 16 query groups) -> finalize, all on ws->main; outputs on the device
static int run_batch(szg_index *h, Workspace *ws, const BatchPlan &p, const double *d_q, uint32_t nq, uint32_t k,
                     const uint32_t *mask, uint32_t flags, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                     uint32_t *d_out_flags, const PeerSink *sink) {
    int rc;
    cudaStream_t st = ws->main;
    const int nd = 2; // 2 digit planes x 64 queries = the M dimension
    // the query digits are laid out like an 8-bit row of Cb chunks in both cases
    const size_t stride = sizeof(PQHeader) + (size_t)p.Cb * nd * 16;
    if ((rc = ws->d_pq.ensure(stride * nq)) || (rc = ws->d_cand.ensure((size_t)nq * p.nranges * p.keep)) ||
        (rc = ws->d_gmth.ensure((size_t)nq * p.gm_stride + p.ngroups)))
        return rc;
    const uint32_t nblk_now = (h->nslots + 31) / 32;
    if (p.p16 || p.p4) {
        // (re)build the byte copy after a mutation.  16-bit: [high-byte plane | low-byte plane], each nblk x Cb x 32 uint4;
        // 4-bit: one plane, one byte per code.  Searches may run concurrently (RLock): the first one in rebuilds and
        // waits, the others wait on the mutex.
        std::lock_guard<std::mutex> lk(h->mu);
        if (h->planar_dirty || h->planar_nblk != nblk_now) {
            if (ws->capturing) return fail(SZG_EINTERNAL, "byte-planar copy is stale inside a captured launch sequence");
            const size_t plane = (size_t)nblk_now * p.Cb * 32;
            if ((rc = h->planar.ensure((p.p16 ? 2 : 1) * plane))) return rc;
            CK(cudaStreamSynchronize(h->mut_stream));
            if (p.p16) CK(launch_planar16(h->codes.p, h->C, h->planar.p, h->planar.p + plane, p.Cb, nblk_now, st));
            else CK(launch_expand4(h->codes.p, h->C, h->planar.p, p.Cb, nblk_now, st));
            CK(cudaStreamSynchronize(st));
            h->launches++;
            h->planar_dirty = false;
            h->planar_nblk = nblk_now;
        }
    }
    // The groups' next-tile counters sit behind the bound keys: one memset presets both (a counter's first grab is old + 1 = 0).
    // Before prep, not between prep and the scan: batch_kernel is a programmatic dependent of prep_kernel and starts while it
    // runs, so everything else the scan reads must be complete before prep starts.
    CK(cudaMemsetAsync(ws->d_gmth.p, 0xFF, ((size_t)nq * p.gm_stride + p.ngroups) * sizeof(unsigned int), st));
    PrepArgs pa;
    pa.queries = d_q; pa.pq = ws->d_pq.p; pa.pq_stride = stride;
    pa.dims = (uint32_t)h->dim; pa.C = p.Cb; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
    pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 0; pa.radius = 0.0;
    pa.planar16 = (p.p16 || p.p4) ? 1 : 0; // digits laid out like an 8-bit row of Cb chunks
    CK(launch_prep(nq, st, pa));
    h->launches++;
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.codes = (p.p16 || p.p4) ? h->planar.p : h->codes.p;
    b.codes_lo = p.p16 ? h->planar.p + (size_t)nblk_now * p.Cb * 32 : nullptr;
    b.aux = h->aux.p; b.live = h->live.p; b.mask = mask;
    b.pq = ws->d_pq.p; b.pq_stride = stride; b.cand = ws->d_cand.p; b.keep = p.keep;
    b.C = p.Cb; b.nblk = nblk_now; b.metric = (uint32_t)h->metric; b.nq = nq; b.dims = (uint32_t)h->dim;
    b.nranges = p.nranges; b.nlists = p.nranges; b.stages = p.stages; b.slice = p.slice;
    b.gmth = ws->d_gmth.p; b.mth = p.mth; b.gm_sp = p.gm_sp; b.gm_stride = p.gm_stride;
    static const bool fixed_ranges = getenv("SZG_BATCH_FIXED_RANGES") && atoi(getenv("SZG_BATCH_FIXED_RANGES")) != 0;
    b.tile_ctr = (fixed_ranges || p.nranges < 2) ? nullptr : ws->d_gmth.p + (size_t)nq * p.gm_stride;
    static const uint32_t dbg = getenv("SZG_BATCH_DEBUG") ? (uint32_t)atoi(getenv("SZG_BATCH_DEBUG")) : 0u;
    b.debug = dbg;
    static const uint32_t pns = getenv("SZG_BATCH_POLL_NS") ? (uint32_t)atoi(getenv("SZG_BATCH_POLL_NS")) : 0u;
    b.poll_ns = pns;
    b.trace = h->trace ? h->trace + 8 : nullptr; // words 8..15 of the trace buffer
    const uint32_t nlaunch = (p.ngroups + p.gpl - 1) / p.gpl;
    uint32_t tbase = 0;
    bool timing = false;
    if ((rc = timing_reserve(h, ws, nlaunch, &tbase, &timing))) return rc;
    for (uint32_t l = 0, g0 = 0; g0 < p.ngroups; g0 += p.gpl, ++l) {
        b.group0 = g0;
        b.ngroups = std::min(p.gpl, p.ngroups - g0);
        if (timing) CK(cudaEventRecord(ws->t0[tbase + l], st));
        CK(launch_batch(b, st));
        if (timing) CK(cudaEventRecord(ws->t1[tbase + l], st));
        h->launches++;
    }
    if (timing) ws->timed = tbase + nlaunch;
    FinalizeArgs f;
    fill_finalize_args(h, f, stride, k, flags, sink);
    f.cand = ws->d_cand.p;
    f.pq = ws->d_pq.p; f.queries = d_q;
    f.nlists = p.nranges; // one sorted list of `keep` keys per row range
    f.out_ids = d_out_ids; f.out_dist = d_out_dist; f.out_n = d_out_n; f.out_flags = d_out_flags;
    CK(launch_finalize(h->qt, p.mode, nq, st, f));
    h->launches++;
    h->batch_queries += nq;
    return SZG_OK;
}

// One top-k pass for nq queries: the tensor-core contraction when the call is a batch (prefer_batch, or at least
// batch_min queries, by default a threshold that depends on the size of the mirror, see below)
// and its geometry fits, else the streaming scan.
int enqueue_topk(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags,
                 bool prefer_batch, int min_mode, int force_nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                 uint32_t *d_out_flags, const PeerSink *sink, int *mode_out, int *nd_out) {
    int mode = mode_for_k(h, k);
    if (min_mode > mode) mode = min_mode;
    const int nd = force_nd ? force_nd : first_digits(h);
    if (mode_out) *mode_out = mode;
    if (nd_out) *nd_out = nd;
    if (h->live_rows == 0) { // empty mirror: zero results, nothing to scan (collection.go:706-709)
        CK(cudaMemsetAsync(d_out_n, 0, (size_t)nq * 4, ws->main));
        CK(cudaMemsetAsync(d_out_flags, 0, (size_t)nq * 4, ws->main));
        if (sink && sink->done_cnt) CK(launch_bump(sink->done_cnt, nq, ws->main));
        return SZG_OK;
    }
    BatchPlan p;
    // From how many queries on a call is a batch.  Default (batch_min == 0): by the size of the mirror -- a collection that
    // streams from HBM is read once per call by the contraction against once per query by the scans, so two queries are
    // enough (10 M x 768 8-bit: 1.14 vs 1.40 ms, three: 1.30 vs 1.91 ms); a cache-resident one pays the contraction's fixed
    // ~90-150 us against 3-13 us per extra scan (100 k x 384 8-bit, eight queries: 101 vs 78 us; 1 M x 128 4-bit: 157 vs 129 us).
    // profiles/r02_dispatch_crossover.log
    uint32_t bmin = (uint32_t)h->batch_min;
    if (bmin == 0) {
        const size_t bytes = (size_t)h->nslots * h->rowbytes;
        bmin = bytes >= ((size_t)2 << 30) ? 2u : bytes >= ((size_t)512 << 20) ? 3u : 12u;
    }
    if (nd == 2 && (prefer_batch || nq >= bmin) && plan_batch(h, nq, k, mode, &p))
        return run_batch(h, ws, p, d_q, nq, k, mask, flags, d_out_ids, d_out_dist, d_out_n, d_out_flags, sink);
    return run_topk(h, ws, d_q, nq, k, mask, flags, mode, nd, d_out_ids, d_out_dist, d_out_n, d_out_flags, sink);
}

int check_search(szg_index *h, const void *q, uint32_t nq) {
    if (!q && nq) return fail(SZG_EINVAL, "null query");
    (void)h;
    return SZG_OK;
}

// layout of the packed first-pass outputs of a host-buffer call: [ids on*8 | dist on*8 | n nq*4 | flags nq*4]: one D2H copy
struct OutPack {
    unsigned char *d = nullptr, *h = nullptr;
    size_t on = 0, nq = 0;
    size_t bytes() const { return on * 16 + nq * 8; }
    unsigned long long *d_ids() const { return reinterpret_cast<unsigned long long *>(d); }
    double *d_dist() const { return reinterpret_cast<double *>(d + on * 8); }
    uint32_t *d_n() const { return reinterpret_cast<uint32_t *>(d + on * 16); }
    uint32_t *d_flags() const { return reinterpret_cast<uint32_t *>(d + on * 16 + nq * 4); }
};

// The first pass and the copy of its packed outputs to pack.h are already enqueued on ws->main (directly or as a captured
// launch sequence).  Waits for them, hands the results to the caller and re-runs, together, the queries whose candidate set
// could not be certified: first with the 3-digit (precise) surrogate, then with larger candidate sets (32 -> 64 -> 128 -> 256).
int collect_and_escalate(szg_index *h, Workspace *ws, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags,
                         int nd0, int mode0, uint64_t *out_ids, double *out_dist, uint32_t *out_n, const OutPack &pack) {
    int rc;
    cudaStream_t st = ws->main;
    const size_t on = (size_t)nq * k;
    CK(cudaStreamSynchronize(st));
    drain_timing(h, ws);
    memcpy(out_ids, pack.h, on * 8);
    memcpy(out_dist, pack.h + on * 8, on * 8);
    memcpy(out_n, pack.h + on * 16, nq * 4);
    const uint32_t *fl = reinterpret_cast<const uint32_t *>(pack.h + on * 16 + nq * 4);
    if (flags & SZG_F_NO_FP64_VERIFY) return SZG_OK;
    std::vector<uint32_t> pending;
    for (uint32_t i = 0; i < nq; ++i)
        if (fl[i] & 1u) pending.push_back(i);
    int nd = nd0, mode = mode0;
    while (!pending.empty()) {
        if (nd == 2) nd = 3;
        else if (mode < 3) ++mode;
        else break;
        const uint32_t m = (uint32_t)pending.size();
        h->escalations += m;
        if ((rc = ws->d_q2.ensure((size_t)m * h->dim)) || (rc = ws->d_out_ids.ensure((size_t)m * k)) ||
            (rc = ws->d_out_dist.ensure((size_t)m * k)) || (rc = ws->d_out_n.ensure(m)) || (rc = ws->d_out_flags.ensure(m)) ||
            (rc = ws->h_out_ids.ensure((size_t)m * k)) || (rc = ws->h_out_dist.ensure((size_t)m * k)) ||
            (rc = ws->h_out_n.ensure(m)) || (rc = ws->h_out_flags.ensure(m)))
            return rc;
        for (uint32_t j = 0; j < m; ++j)
            CK(cudaMemcpyAsync(ws->d_q2.p + (size_t)j * h->dim, ws->d_q.p + (size_t)pending[j] * h->dim,
                               (size_t)h->dim * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if ((rc = run_topk(h, ws, ws->d_q2.p, m, k, mask, flags, mode, nd, ws->d_out_ids.p, ws->d_out_dist.p,
                           ws->d_out_n.p, ws->d_out_flags.p, nullptr)))
            return rc;
        CK(cudaMemcpyAsync(ws->h_out_ids.p, ws->d_out_ids.p, (size_t)m * k * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_dist.p, ws->d_out_dist.p, (size_t)m * k * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_n.p, ws->d_out_n.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_flags.p, ws->d_out_flags.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        drain_timing(h, ws);
        std::vector<uint32_t> still;
        for (uint32_t j = 0; j < m; ++j) {
            const uint32_t i = pending[j];
            memcpy(out_ids + (size_t)i * k, ws->h_out_ids.p + (size_t)j * k, (size_t)k * 8);
            memcpy(out_dist + (size_t)i * k, ws->h_out_dist.p + (size_t)j * k, (size_t)k * 8);
            out_n[i] = ws->h_out_n.p[j];
            if (ws->h_out_flags.p[j] & 1u) still.push_back(i);
        }
        pending.swap(still);
    }
    h->uncertain += pending.size();
    return SZG_OK;
}

// the workspace of device-resident calls on `stream` (the caller serialises calls per stream)
int ws_for_stream(szg_index *h, void *stream, Workspace **out) {
    std::lock_guard<std::mutex> lk(h->mu);
    auto it = h->dev_ws.find(stream);
    if (it != h->dev_ws.end()) { *out = it->second; return SZG_OK; }
    Workspace *ws = new Workspace();
    int rc = ws->init(false);
    if (rc) { ws->destroy(); delete ws; return rc; }
    ws->main = (cudaStream_t)stream;
    h->dev_ws[stream] = ws;
    *out = ws;
    return SZG_OK;
}

// Host-buffer top-k of one device: H2D of the queries, the pass, D2H of the packed results, escalation.  Call shapes that
// repeat are replayed as a captured launch sequence (CUDA graph): the first call of a shape runs launch by launch (and
// sizes every buffer), the second one captures, later ones are ONE graph launch -- what makes a single-query call cost
// the scan plus a few microseconds instead of five launch gaps (SZG_OPT_GRAPHS, on unless per-launch timing is requested).
int search_host(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags, uint64_t *out_ids,
                double *out_dist, uint32_t *out_n, uint64_t *scanned, bool prefer_batch) {
    GUARD(h);
    int rc;
    if ((rc = check_search(h, queries, nq))) return rc;
    if (k < 1 || k > SZG_MAX_K) return fail(SZG_EINVAL, "k=%u outside [1, %u]", k, SZG_MAX_K);
    if (nq && (!out_ids || !out_dist || !out_n)) return fail(SZG_EINVAL, "null output");
    if (scanned) *scanned = h->live_rows;
    if (!nq) return SZG_OK;
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    if (h->live_rows == 0) { // empty collection: zero results, nothing to launch (collection.go:706-709)
        for (uint32_t i = 0; i < nq; ++i) out_n[i] = 0;
        return SZG_OK;
    }
    Workspace *ws;
    if ((rc = acquire_ws(h, &ws))) return rc;
    struct Rel { szg_index *h; Workspace *w; ~Rel() { release_ws(h, w); } } rel{h, ws};
    const size_t qn = (size_t)nq * h->dim, on = (size_t)nq * k;
    OutPack pack;
    pack.on = on; pack.nq = nq;
    if ((rc = ws->h_q.ensure(qn)) || (rc = ws->d_q.ensure(qn)) || (rc = ws->d_out_pack.ensure(pack.bytes())) ||
        (rc = ws->h_out_pack.ensure(pack.bytes())))
        return rc;
    pack.d = ws->d_out_pack.p; pack.h = ws->h_out_pack.p;
    memcpy(ws->h_q.p, queries, qn * sizeof(double));
    cudaStream_t st = ws->main;
    int mode0 = 0, nd0 = 0;
    auto enqueue_all = [&]() -> int {
        CK(cudaMemcpyAsync(ws->d_q.p, ws->h_q.p, qn * sizeof(double), cudaMemcpyHostToDevice, st));
        int r = enqueue_topk(h, ws, ws->d_q.p, nq, k, mask, flags, prefer_batch, 0, 0, pack.d_ids(), pack.d_dist(), pack.d_n(),
                             pack.d_flags(), nullptr, &mode0, &nd0);
        if (r) return r;
        CK(cudaMemcpyAsync(pack.h, pack.d, pack.bytes(), cudaMemcpyDeviceToHost, st));
        return SZG_OK;
    };
    bool done = false;
    if (h->use_graphs && h->timing == 0) {
        const uint64_t key = ((uint64_t)nq << 40) ^ ((uint64_t)k << 28) ^ ((uint64_t)(uint32_t)(mask_id + 1) << 4) ^ ((uint64_t)(flags & 3u) << 1) ^
                             (prefer_batch ? 1u : 0u);
        GraphEntry &ge = ws->graphs[key];
        const uint64_t gen = h->generation;
        // the captured nodes hold the addresses of this workspace's buffers: another call shape that made one of them grow
        // (and move) since the capture makes the sequence stale
        auto fingerprint = [&](const void *(&fp)[7]) {
            fp[0] = ws->d_q.p; fp[1] = ws->h_q.p; fp[2] = ws->d_pq.p; fp[3] = ws->d_cand.p; fp[4] = ws->d_gmth.p;
            fp[5] = ws->d_out_pack.p; fp[6] = ws->h_out_pack.p;
        };
        const void *now[7];
        fingerprint(now);
        if (ge.exec && (ge.generation != gen || memcmp(now, ge.buffers, sizeof now) != 0)) {
            cudaGraphExecDestroy(ge.exec);
            ge.exec = nullptr;
            if (ge.generation == gen) ge.generation = 0; // moved buffers: this call runs launch by launch, the next one captures
        }
        if (ge.exec && ge.generation == gen) {
            mode0 = ge.mode0; nd0 = ge.nd0;
            CK(cudaGraphLaunch(ge.exec, st));
            h->graph_launches++;
            h->launches += ge.kernels;
            done = true;
        } else if (!ge.exec && ge.generation == gen && !ge.failed) {
            // second call of this shape on this generation of the mirror: capture it
            const uint64_t l0 = h->launches.load();
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                ws->capturing = true;
                const int r = enqueue_all();
                ws->capturing = false;
                e = cudaStreamEndCapture(st, &g);
                if (r == SZG_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&ge.exec, g, 0);
                else if (e == cudaSuccess) e = cudaErrorUnknown;
                if (g) cudaGraphDestroy(g);
            }
            if (e != cudaSuccess || !ge.exec) {
                cudaGetLastError(); // a failed capture is not an error of the call: this shape stays on plain launches
                ge.exec = nullptr;
                ge.failed = true;
            } else {
                ge.mode0 = mode0; ge.nd0 = nd0;
                ge.kernels = (uint32_t)(h->launches.load() - l0);
                fingerprint(ge.buffers);
                CK(cudaGraphLaunch(ge.exec, st));
                h->graph_launches++;
                done = true;
            }
        } else if (ge.generation != gen) {
            ge.generation = gen;
            ge.failed = false;
            if (ws->graphs.size() > 64) { // a caller that varies its shapes: keep the table small
                for (auto it = ws->graphs.begin(); it != ws->graphs.end();) {
                    if (it->first != key) { if (it->second.exec) cudaGraphExecDestroy(it->second.exec); it = ws->graphs.erase(it); }
                    else ++it;
                }
            }
        }
    }
    if (!done && (rc = enqueue_all())) return rc;
    return collect_and_escalate(h, ws, nq, k, mask, flags, nd0, mode0, out_ids, out_dist, out_n, pack);
}

} // namespace szg

extern "C" {

// Concurrent callers.  The reference answers one query per Search call and lets calls overlap (RLock only,
// collection.go:570); a scan launch, however, owns the whole GPU, so overlapping calls would queue up one launch each and
// every one of them would stream the collection from HBM alone.  Instead the calls combine: a caller that finds no launch in
// flight becomes the leader and runs whatever is queued with its own k / mask / flags as ONE call; callers arriving meanwhile
// wait and are answered together by the next leader.  Nobody waits for company: a lone caller runs at once, exactly as before.
// The leader runs the combined batch through search_host -- never through a public entry point, which would queue it behind
// itself -- and search_host picks the tensor-core contraction or the scan by batch size and geometry.
struct szg_index::PendingSearch {
    const double *q; uint32_t nq, k; int mask_id; uint32_t flags;
    uint64_t *out_ids; double *out_dist; uint32_t *out_n;
    int rc = 0; std::string err; bool done = false;
    std::condition_variable cv; // the caller sleeps on its own variable: a finished launch wakes its callers and one new leader only
};
constexpr uint32_t kCombineMaxCall = 16;   // calls with more queries than this are not combined
constexpr uint32_t kCombineMaxBatch = 128; // queries of one combined launch

static int search_topk_combined(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                                uint64_t *out_ids, double *out_dist, uint32_t *out_n) {
    using P = szg_index::PendingSearch;
    P me;
    me.q = queries; me.nq = nq; me.k = k; me.mask_id = mask_id; me.flags = flags;
    me.out_ids = out_ids; me.out_dist = out_dist; me.out_n = out_n;
    std::unique_lock<std::mutex> lk(h->comb_mu);
    try { h->comb_queue.push_back(&me); } catch (...) { return fail(SZG_ENOMEM, "out of host memory"); }
    while (!me.done) {
        if (h->comb_leader) { me.cv.wait(lk); continue; }
        h->comb_leader = true;
        // the batch: the oldest request and every queued one with the same parameters, in arrival order
        std::vector<P *> batch;
        uint32_t total = 0;
        int rc = SZG_OK;
        try { // nothing may leave this block by exception: the other callers wait for the leader (and the ABI never throws)
            P *first = h->comb_queue.front();
            batch.reserve(h->comb_queue.size());
            for (auto it = h->comb_queue.begin(); it != h->comb_queue.end();) {
                P *p = *it;
                if (p->k == first->k && p->mask_id == first->mask_id && p->flags == first->flags &&
                    (batch.empty() || total + p->nq <= kCombineMaxBatch)) {
                    batch.push_back(p);
                    total += p->nq;
                    it = h->comb_queue.erase(it);
                } else ++it;
            }
            lk.unlock();
            if (batch.size() == 1) {
                P *p = batch[0];
                rc = search_host(h, p->q, p->nq, p->k, p->mask_id, p->flags, p->out_ids, p->out_dist, p->out_n, nullptr, false);
            } else {
                const size_t d = (size_t)h->dim, kk = first->k;
                std::vector<double> q(total * d);
                std::vector<uint64_t> ids(total * kk);
                std::vector<double> dist(total * kk);
                std::vector<uint32_t> n(total);
                size_t off = 0;
                for (P *p : batch) { memcpy(q.data() + off * d, p->q, (size_t)p->nq * d * sizeof(double)); off += p->nq; }
                rc = search_host(h, q.data(), total, first->k, first->mask_id, first->flags, ids.data(), dist.data(), n.data(), nullptr, false);
                off = 0;
                if (!rc)
                    for (P *p : batch) {
                        memcpy(p->out_ids, ids.data() + off * kk, (size_t)p->nq * kk * 8);
                        memcpy(p->out_dist, dist.data() + off * kk, (size_t)p->nq * kk * 8);
                        memcpy(p->out_n, n.data() + off, (size_t)p->nq * 4);
                        off += p->nq;
                    }
            }
        } catch (const std::bad_alloc &) {
            rc = fail(SZG_ENOMEM, "out of host memory while combining %zu concurrent searches", batch.size());
        } catch (...) {
            rc = fail(SZG_EINTERNAL, "unexpected exception while combining concurrent searches");
        }
        std::string err;
        try { if (rc) err = last_error_string(); } catch (...) {}
        if (!lk.owns_lock()) lk.lock();
        if (batch.size() > 1 && !rc) h->combined_queries += total;
        for (P *p : batch) {
            p->rc = rc; p->err = err; p->done = true;
            if (p != &me) p->cv.notify_one();
        }
        h->comb_leader = false;
        if (!h->comb_queue.empty() && h->comb_queue.front() != &me) h->comb_queue.front()->cv.notify_one(); // the next leader
    }
    if (me.rc) set_last_error_string(me.err);
    return me.rc;
}

int szg_search_topk(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                    uint64_t *out_ids, double *out_dist, uint32_t *out_n, uint64_t *scanned) {
    if (h && h->sh) return sharded_search_topk(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n, scanned, false);
    if (h && h->combine && queries && nq >= 1 && nq <= kCombineMaxCall && out_ids && out_dist && out_n && k >= 1 && k <= SZG_MAX_K) {
        if (scanned) *scanned = h->live_rows;
        return search_topk_combined(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n);
    }
    return search_host(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n, scanned, false);
}

// Batched search: same results as szg_search_topk for every query, computed by the tensor-core
// contraction kernel (batch_q8.cu) when the collection is 4/8/16-bit and the geometry fits; every other
// case is served by the streaming scan (still on the GPU).
int szg_search_batch(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                     uint64_t *out_ids, double *out_dist, uint32_t *out_n, uint64_t *scanned) {
    if (h && h->sh) return sharded_search_topk(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n, scanned, true);
    return search_host(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n, scanned, true);
}

static int search_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                      uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, void *stream,
                      bool prefer_batch) {
    if (h && h->sh)
        return sharded_search_topk_dev(h, d_queries, nq, k, mask_id, flags, d_out_ids, d_out_dist, d_out_n, d_out_flags, stream, prefer_batch);
    GUARD(h);
    int rc;
    if ((rc = check_search(h, d_queries, nq))) return rc;
    if (k < 1 || k > SZG_MAX_K) return fail(SZG_EINVAL, "k=%u outside [1, %u]", k, SZG_MAX_K);
    if (!nq) return SZG_OK;
    if (!d_out_ids || !d_out_dist || !d_out_n) return fail(SZG_EINVAL, "null output");
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    Workspace *ws;
    if ((rc = ws_for_stream(h, stream, &ws))) return rc;
    if (!d_out_flags) {
        if ((rc = ws->d_out_flags.ensure(nq))) return rc;
        d_out_flags = ws->d_out_flags.p;
    }
    // no host synchronisation here, hence no escalation: SZG_OPT_DIGITS = 2 (or automatic) runs the fast
    // surrogate and reports uncertified queries in d_out_flags; the caller re-runs those with szg_search_topk
    return enqueue_topk(h, ws, d_queries, nq, k, mask, flags, prefer_batch, 0, 0, (unsigned long long *)d_out_ids, d_out_dist,
                        d_out_n, d_out_flags, nullptr, nullptr, nullptr);
}

// Device-resident forms (queries and outputs in HBM, everything enqueued on `stream`, no host synchronisation): what a
// multi-process row-sharded deployment calls before its all-gather + szg_merge_topk_dev.
int szg_search_topk_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                        uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, void *stream) {
    return search_dev(h, d_queries, nq, k, mask_id, flags, d_out_ids, d_out_dist, d_out_n, d_out_flags, stream, false);
}
int szg_search_batch_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                         uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, void *stream) {
    return search_dev(h, d_queries, nq, k, mask_id, flags, d_out_ids, d_out_dist, d_out_n, d_out_flags, stream, true);
}

int szg_merge_topk_dev(szg_index *h, const uint64_t *d_gathered_ids, const double *d_gathered_dist,
                       const uint32_t *d_gathered_n, const uint32_t *d_gathered_flags, uint64_t rank_stride_bytes,
                       uint32_t nranks, uint32_t nq, uint32_t k, uint64_t *d_out_ids, double *d_out_dist,
                       uint32_t *d_out_n, uint32_t *d_out_flags, void *stream) {
    if (h && h->sh) h = sharded_root(h);
    GUARD(h);
    if (!nq) return SZG_OK;
    if (!d_gathered_ids || !d_gathered_dist || !d_gathered_n || !d_out_ids || !d_out_dist || !d_out_n)
        return fail(SZG_EINVAL, "null argument");
    if (k < 1 || k > SZG_MAX_K || nranks < 1 || (size_t)nranks * k * 16 > 200 * 1024)
        return fail(SZG_EINVAL, "merge of %u lists of k=%u is not supported", nranks, k);
    MergeArgs a;
    memset(&a, 0, sizeof a);
    a.g_ids = (const unsigned long long *)d_gathered_ids; a.g_dist = d_gathered_dist; a.g_n = d_gathered_n;
    a.g_flags = d_gathered_flags; a.out_flags = d_out_flags;
    a.rank_stride = (size_t)rank_stride_bytes;
    a.G = nranks; a.nq = nq; a.k = k;
    a.out_ids = (unsigned long long *)d_out_ids; a.out_dist = d_out_dist; a.out_n = d_out_n;
    CK(launch_merge(a, (cudaStream_t)stream));
    h->launches++;
    return SZG_OK;
}

} // extern "C"

namespace szg {

// m1 of angularDistance (collection.go:821-827: m1 += query[i] * query[i], sequential) -- the same IEEE operations as on the
// device (this file is compiled with -ffp-contract=off): the gather kernels take it from the host instead of re-adding it for
// every candidate
static double query_m1(const double *q, uint32_t dims) {
    volatile double m1 = 0.0; // volatile: every product and every sum is rounded to double, in order
    for (uint32_t i = 0; i < dims; ++i) {
        const volatile double p = q[i] * q[i];
        m1 = m1 + p;
    }
    return m1;
}

// ---- radius search of one device (collection.go:598-605): surrogate scan with warp-ballot compaction, exact fp64
// distances of the compacted rows, the inclusive test and the (distance, lexicographic id) order ON THE DEVICE; the hits
// come back with one copy.  nq queries share the launches (every query still streams the mirror on its own).
// out[q] receives the hits of query q, ascending.
int radius_device(szg_index *h, const double *queries, uint32_t nq, const double *radii, int mask_id, szg_result **out) {
    GUARD(h);
    int rc;
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    for (uint32_t q = 0; q < nq; ++q) out[q] = nullptr;
    std::vector<std::unique_ptr<szg_result>> res(nq);
    for (auto &r : res) r.reset(new szg_result());
    auto give = [&]() { for (uint32_t q = 0; q < nq; ++q) out[q] = res[q].release(); return SZG_OK; };
    if (h->live_rows == 0 || !nq) return give();
    Workspace *ws;
    if ((rc = acquire_ws(h, &ws))) return rc;
    struct Rel { szg_index *h; Workspace *w; ~Rel() { release_ws(h, w); } } rel{h, ws};
    const int nd = first_digits(h); // the radius threshold carries the surrogate error bound: no re-run needed
    const size_t stride = pq_stride(h, nd);
    const size_t qn = (size_t)nq * h->dim;
    if ((rc = ws->h_q.ensure(qn)) || (rc = ws->d_q.ensure(qn)) || (rc = ws->d_pq.ensure(stride * nq)) ||
        (rc = ws->h_out_n.ensure(2 * (size_t)nq)) || (rc = ws->d_ticket.ensure(2 * (size_t)nq + 64)))
        return rc;
    memcpy(ws->h_q.p, queries, qn * sizeof(double));
    cudaStream_t st = ws->main;
    CK(cudaMemcpyAsync(ws->d_q.p, ws->h_q.p, qn * sizeof(double), cudaMemcpyHostToDevice, st));
    ScanPlan plan;
    int grid = 0;
    if ((rc = plan_scan(h, nd, &plan, &grid))) return rc;
    // prep: one launch per distinct radius would be wasteful; the radius lives in the prepared-query header, so prep runs
    // per query with its own radius (launch_prep takes one radius: group equal radii)
    for (uint32_t q0 = 0; q0 < nq;) {
        uint32_t q1 = q0 + 1;
        while (q1 < nq && radii[q1] == radii[q0]) ++q1;
        PrepArgs pa;
        pa.queries = ws->d_q.p + (size_t)q0 * h->dim; pa.pq = ws->d_pq.p + stride * q0; pa.pq_stride = stride;
        pa.dims = (uint32_t)h->dim; pa.C = h->C; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
        pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 1; pa.radius = radii[q0];
        CK(launch_prep(q1 - q0, st, pa));
        h->launches++;
        q0 = q1;
    }
    // compaction buffer: one region per query.  The size is a running estimate (the largest hit count this workspace has
    // seen, with headroom) so that a steady workload never rescans; an overflow sizes the buffer exactly and rescans once.
    unsigned int *d_count = ws->d_ticket.p; // [nq] surrogate hits, [nq .. 2nq) exact hits
    size_t cap = std::max<size_t>({(size_t)4096, (size_t)h->nslots / 64, (size_t)ws->radius_cap_hint});
    uint32_t tbase = 0;
    bool timing = false;
    std::vector<RadiusFinishArgs> fin(nq);
    for (int attempt = 0; attempt < 3; ++attempt) {
        // the hits are ordered by bitonic networks over a power of two of (distance, id) pairs per query
        size_t kcap = 2 * kRadiusSortSmall;
        while (kcap < cap) kcap <<= 1;
        if ((rc = ws->d_slots.ensure(cap * nq))) return rc;
        CK(cudaMemsetAsync(d_count, 0, 2 * (size_t)nq * 4, st));
        if ((rc = timing_reserve(h, ws, nq, &tbase, &timing))) return rc;
        for (uint32_t q = 0; q < nq; ++q) {
            ScanArgs a;
            fill_scan_args(h, a, mask);
            a.pq = ws->d_pq.p + stride * q; a.pq_stride = stride; a.nq = 1;
            a.rad_count = d_count + q; a.rad_slots = ws->d_slots.p + cap * q; a.rad_cap = (uint32_t)cap;
            a.Ct = plan.Ct; a.stages = plan.stages; a.pq_smem_off = plan.pq_smem_off;
            if (timing) CK(cudaEventRecord(ws->t0[tbase + q], st));
            CK(launch_scan(h->qt, MODE_RADIUS, nd, grid, (int)plan.warps * 32, plan.smem, st, a));
            if (timing) CK(cudaEventRecord(ws->t1[tbase + q], st));
            h->launches++;
        }
        if (timing) ws->timed = tbase + nq;
        // exact pass over whatever fitted (an overflowing query is redone anyway): distances + sortable keys
        if ((rc = ws->d_out_ids.ensure(cap * nq)) || (rc = ws->d_out_dist.ensure(cap * nq)) || (rc = ws->d_keys.ensure(2 * kcap * nq)))
            return rc;
        for (uint32_t q = 0; q < nq; ++q) {
            RadiusFinishArgs &fa = fin[q];
            fa.codes = h->codes.p; fa.ids = h->ids.p; fa.lut = h->lut.p; fa.q = ws->d_q.p + (size_t)q * h->dim;
            fa.slots = ws->d_slots.p + cap * q; fa.count_ptr = d_count + q; fa.cap = (uint32_t)cap;
            fa.radius = radii[q];
            fa.m1 = query_m1(queries + (size_t)q * h->dim, (uint32_t)h->dim);
            fa.out_dist = ws->d_out_dist.p + cap * q; fa.out_ids = ws->d_out_ids.p + cap * q;
            fa.keys = ws->d_keys.p + 2 * kcap * q; fa.out_count = d_count + nq + q;
            fa.C = h->C; fa.dims = (uint32_t)h->dim; fa.metric = (uint32_t)h->metric; fa.qt = h->qt;
            CK(launch_radius_finish(fa, st));
            h->launches += 2;
        }
        CK(cudaMemcpyAsync(ws->h_out_n.p, d_count, 2 * (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        drain_timing(h, ws);
        size_t worst = 0;
        for (uint32_t q = 0; q < nq; ++q) worst = std::max<size_t>(worst, ws->h_out_n.p[q]);
        ws->radius_cap_hint = (uint32_t)std::min<size_t>(0xFFFFFFF0u, worst + worst / 4 + 1024);
        if (worst <= cap) break;
        if (attempt == 2) return fail(SZG_EINTERNAL, "radius compaction buffer could not be sized");
        cap = worst; // too small: size it exactly and rescan
    }
    // results above kRadiusSortSmall hits were left unordered by launch_radius_finish: order them now
    size_t total = 0;
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t m = ws->h_out_n.p[nq + q];
        total += m;
        if (m > kRadiusSortSmall) {
            CK(launch_radius_sort_large(fin[q], m, st));
            h->launches++;
        }
    }
    // the exact hits of every query, filtered and ordered: one copy each
    if (total) {
        if ((rc = ws->h_out_ids.ensure(total)) || (rc = ws->h_out_dist.ensure(total))) return rc;
        size_t off = 0;
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t m = ws->h_out_n.p[nq + q];
            if (!m) continue;
            CK(cudaMemcpyAsync(ws->h_out_ids.p + off, ws->d_out_ids.p + cap * q, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ws->h_out_dist.p + off, ws->d_out_dist.p + cap * q, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
            off += m;
        }
        CK(cudaStreamSynchronize(st));
        off = 0;
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t m = ws->h_out_n.p[nq + q];
            res[q]->ids.assign(ws->h_out_ids.p + off, ws->h_out_ids.p + off + m);
            res[q]->dist.assign(ws->h_out_dist.p + off, ws->h_out_dist.p + off + m);
            off += m;
        }
    }
    return give();
}

// fp64 distances of candidate lists in visit order (lshtree.go:316-335 -> collection.go:584-596): nl lists, list l =
// ids[off[l] .. off[l+1]) scored against query l.  One H2D of queries + slots, one launch, one D2H.
int rescore_device(szg_index *h, const double *queries, uint32_t nl, const uint64_t *ids, const uint64_t *off, double *out_dist) {
    GUARD(h);
    int rc;
    const uint64_t m = off[nl];
    if (!m) return SZG_OK;
    if (m > 0xFFFFFFF0ull) return fail(SZG_EINVAL, "too many ids");
    Workspace *ws;
    if ((rc = acquire_ws(h, &ws))) return rc;
    struct Rel { szg_index *h; Workspace *w; ~Rel() { release_ws(h, w); } } rel{h, ws};
    // one pinned staging block: [queries nl*dim f64, m1 nl f64 | list offsets (nl+1) u32, padded | slots m u32]
    const size_t qonly = (size_t)nl * h->dim * 8;
    const size_t qbytes = qonly + (size_t)nl * 8, obytes = ((size_t)(nl + 1) * 4 + 7) / 8 * 8, sbytes = (size_t)m * 4;
    const size_t total = qbytes + obytes + sbytes;
    if ((rc = ws->h_out_pack.ensure(total)) || (rc = ws->d_out_pack.ensure(total)) || (rc = ws->d_out_dist.ensure(m)) ||
        (rc = ws->h_out_dist.ensure(m)))
        return rc;
    unsigned char *hp = ws->h_out_pack.p;
    memcpy(hp, queries, qonly);
    double *hm1 = reinterpret_cast<double *>(hp + qonly);
    for (uint32_t l = 0; l < nl; ++l) hm1[l] = query_m1(queries + (size_t)l * h->dim, (uint32_t)h->dim);
    uint32_t *ho = reinterpret_cast<uint32_t *>(hp + qbytes);
    for (uint32_t l = 0; l <= nl; ++l) ho[l] = (uint32_t)off[l];
    uint32_t *hs = reinterpret_cast<uint32_t *>(hp + qbytes + obytes);
    for (uint64_t i = 0; i < m; ++i) {
        uint32_t s;
        hs[i] = h->lookup(ids[i], &s) ? s : 0xFFFFFFFFu;
    }
    cudaStream_t st = ws->main;
    CK(cudaMemcpyAsync(ws->d_out_pack.p, hp, total, cudaMemcpyHostToDevice, st));
    RescoreArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.codes = h->codes.p; ra.ids = h->ids.p; ra.lut = h->lut.p;
    ra.q = reinterpret_cast<const double *>(ws->d_out_pack.p);
    ra.m1 = reinterpret_cast<const double *>(ws->d_out_pack.p + qonly);
    ra.list_off = reinterpret_cast<const uint32_t *>(ws->d_out_pack.p + qbytes); ra.nlists = nl;
    ra.slots = reinterpret_cast<const uint32_t *>(ws->d_out_pack.p + qbytes + obytes);
    // one LSH batch (a few hundred distances): the kernel stores them straight into the pinned host buffer (unified addressing:
    // the same pointer is valid on the device) -- one asynchronous operation less on a call that is all latency
    const bool direct = m <= 4096;
    ra.out_dist = direct ? ws->h_out_dist.p : ws->d_out_dist.p; ra.out_ids = nullptr;
    ra.C = h->C; ra.dims = (uint32_t)h->dim; ra.metric = (uint32_t)h->metric; ra.m = (uint32_t)m; ra.qt = h->qt;
    CK(launch_rescore(ra, st));
    h->launches++;
    if (!direct) CK(cudaMemcpyAsync(ws->h_out_dist.p, ws->d_out_dist.p, m * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out_dist, ws->h_out_dist.p, m * 8);
    return SZG_OK;
}

} // namespace szg

extern "C" {

int szg_search_radius_batch(szg_index *h, const double *queries, uint32_t nq, const double *radii, int mask_id, uint32_t flags,
                            szg_result **out, uint64_t *scanned) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (!out) return fail(SZG_EINVAL, "null out pointer");
    for (uint32_t q = 0; q < nq; ++q) out[q] = nullptr;
    if (nq && (!queries || !radii)) return fail(SZG_EINVAL, "null query");
    for (uint32_t q = 0; q < nq; ++q)
        if (!(radii[q] > 0)) return fail(SZG_EINVAL, "radius must be > 0 (collection.go:598)");
    (void)flags;
    if (h->sh) return sharded_search_radius(h, queries, nq, radii, mask_id, out, scanned);
    if (scanned) *scanned = h->live_rows;
    return radius_device(h, queries, nq, radii, mask_id, out);
}

int szg_search_radius(szg_index *h, const double *query, double radius, int mask_id, uint32_t flags,
                      szg_result **out, uint64_t *scanned) {
    if (!out) return fail(SZG_EINVAL, "null out pointer");
    *out = nullptr;
    return szg_search_radius_batch(h, query, 1, &radius, mask_id, flags, out, scanned);
}

int szg_result_count(const szg_result *r, uint64_t *n) {
    if (!r || !n) return fail(SZG_EINVAL, "null argument");
    *n = r->ids.size();
    return SZG_OK;
}
int szg_result_fetch(const szg_result *r, uint64_t offset, uint64_t n, uint64_t *out_ids, double *out_dist) {
    if (!r) return fail(SZG_EINVAL, "null result");
    if (offset > r->ids.size() || n > r->ids.size() - offset) return fail(SZG_EINVAL, "range outside the result");
    if (out_ids) memcpy(out_ids, r->ids.data() + offset, n * 8);
    if (out_dist) memcpy(out_dist, r->dist.data() + offset, n * 8);
    return SZG_OK;
}
void szg_result_free(szg_result *r) { delete r; }

int szg_rescore_batch(szg_index *h, const double *queries, uint32_t nlists, const uint64_t *ids, const uint64_t *list_offsets,
                      double *out_dist) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (!nlists) return SZG_OK;
    if (!queries || !list_offsets) return fail(SZG_EINVAL, "null argument");
    for (uint32_t l = 0; l < nlists; ++l)
        if (list_offsets[l] > list_offsets[l + 1]) return fail(SZG_EINVAL, "list offsets must not decrease");
    if (list_offsets[0] != 0) return fail(SZG_EINVAL, "list offsets must start at 0");
    if (!list_offsets[nlists]) return SZG_OK;
    if (!ids || !out_dist) return fail(SZG_EINVAL, "null argument");
    if (h->sh) return sharded_rescore(h, queries, nlists, ids, list_offsets, out_dist);
    return rescore_device(h, queries, nlists, ids, list_offsets, out_dist);
}

int szg_rescore(szg_index *h, const double *query, const uint64_t *ids, uint64_t m, double *out_dist) {
    const uint64_t off[2] = {0, m};
    if (h && !query) return fail(SZG_EINVAL, "null query");
    return szg_rescore_batch(h, query, 1, ids, off, out_dist);
}

int szg_last_scan_times_ms(szg_index *h, float *out_ms, uint32_t cap, uint32_t *n) {
    if (h && h->sh) h = sharded_timing_shard(h);
    GUARD(h);
    if (!n) return fail(SZG_EINVAL, "null n");
    *n = 0;
    std::vector<Workspace *> dev;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        for (auto &kv : h->dev_ws) dev.push_back(kv.second);
    }
    for (Workspace *ws : dev) drain_timing(h, ws); // device-resident calls never synchronise: their events are collected here
    std::lock_guard<std::mutex> lk(h->mu);
    const uint32_t m = (uint32_t)std::min<size_t>(cap, h->last_times.size());
    if (out_ms) memcpy(out_ms, h->last_times.data(), (size_t)m * sizeof(float));
    *n = m;
    h->last_times.clear(); // drained (matters for the accumulating mode)
    return SZG_OK;
}

} // extern "C"
