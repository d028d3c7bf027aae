// index.cu -- the GPU mirror of one collection (or one row shard) and the C ABI of
// include/syzgy_b200.h.  Host logic only: slot allocation, id -> slot map, workspaces,
// streams, launches.  There is deliberately no CPU implementation of any search step.
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/syzgy_b200.h"
#include "kernels.h"
#include "scan_small.cuh"

using namespace szg;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

} // namespace

// error channel and geometry for the other translation units of the library (spanfile.cu)
namespace szg {
int set_error(int code, const char *msg) { return fail(code, "%s", msg); }
} // namespace szg

namespace {

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? SZG_ENOMEM : SZG_ECUDA, "%s failed: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

constexpr int kMaxStreams = 4;
constexpr size_t kStageBytes = 64u << 20;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int ensure(size_t want, bool keep = false, cudaStream_t st = 0) {
        if (want <= n) return SZG_OK;
        T *np = nullptr;
        CK(cudaMalloc(&np, want * sizeof(T)));
        if (keep && p && n) {
            cudaError_t e = cudaMemcpyAsync(np, p, n * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(np); return fail(SZG_ECUDA, "device copy failed: %s", cudaGetErrorString(e)); }
        }
        if (p) cudaFree(p);
        p = np;
        n = want;
        return SZG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t n = 0;
    int ensure(size_t want) {
        if (want <= n) return SZG_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        CK(cudaMallocHost(&p, want * sizeof(T)));
        n = want;
        return SZG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

struct Workspace {
    cudaStream_t main = nullptr; // owned for host calls; the caller's stream for *_dev calls
    bool owns_main = false;
    DevBuf<double> d_q, d_q2;
    DevBuf<unsigned char> d_pq;
    DevBuf<unsigned long long> d_cand; // [nq][scan CTAs][32*E] candidate keys, scan -> finalize
    DevBuf<unsigned int> d_gmth;       // batched path: per (query, row range) shared bound keys
    DevBuf<unsigned int> d_ticket;     // radius hit counter
    DevBuf<unsigned long long> d_out_ids;
    DevBuf<double> d_out_dist;
    DevBuf<uint32_t> d_out_n, d_out_flags;
    DevBuf<unsigned char> d_out_pack; // first-pass outputs of a host-buffer top-k call, packed
    PinBuf<unsigned char> h_out_pack;
    DevBuf<uint32_t> d_slots;
    PinBuf<double> h_q;
    PinBuf<unsigned long long> h_out_ids;
    PinBuf<double> h_out_dist;
    PinBuf<uint32_t> h_out_n, h_out_flags, h_slots;
    std::vector<cudaEvent_t> t0, t1; // per-scan timing events
    uint32_t timed = 0;

    int init(bool own) {
        owns_main = own;
        if (own) CK(cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking));
        int rc = d_ticket.ensure(4);
        if (rc) return rc;
        CK(cudaMemset(d_ticket.p, 0, 4 * sizeof(unsigned int)));
        return SZG_OK;
    }
    void destroy() {
        d_q.release(); d_q2.release(); d_pq.release(); d_ticket.release(); d_out_ids.release(); d_out_dist.release();
        d_out_n.release(); d_out_flags.release(); d_slots.release(); d_out_pack.release(); h_out_pack.release();
        d_cand.release(); d_gmth.release();
        h_q.release(); h_out_ids.release(); h_out_dist.release(); h_out_n.release(); h_out_flags.release();
        h_slots.release();
        for (auto e : t0) cudaEventDestroy(e);
        for (auto e : t1) cudaEventDestroy(e);
        if (owns_main && main) cudaStreamDestroy(main);
    }
};

struct IdRange {
    uint64_t id0;
    uint32_t slot0, n;
};

} // namespace

struct szg_result {
    std::vector<uint64_t> ids;
    std::vector<double> dist;
};

struct szg_index {
    int dim = 0, quant = 0, metric = 0, device = 0, qt = 0;
    uint32_t rowbytes = 0, C = 0, maxint = 0;
    int sm_count = 0;
    // storage
    DevBuf<uint4> codes;
    DevBuf<unsigned long long> ids;
    DevBuf<unsigned long long> aux; // 8 bytes per slot reserved; typed per (quant, metric)
    DevBuf<uint32_t> live;
    DevBuf<double> lut;
    uint32_t capacity = 0; // slots allocated (multiple of 64)
    uint32_t nslots = 0;   // high-water mark
    uint64_t live_rows = 0;
    std::unordered_map<uint64_t, uint32_t> map;
    std::vector<IdRange> ranges;
    std::unordered_set<uint64_t> range_dead;
    std::vector<uint32_t> free_slots;
    std::map<int, uint32_t *> masks;
    int next_mask = 1;
    // staging for mutations
    PinBuf<unsigned char> h_stage;
    DevBuf<unsigned char> d_stage;
    DevBuf<double> d_vec; // float64 vectors of an szg_encode batch
    PinBuf<uint32_t> h_slots;
    DevBuf<uint32_t> d_slots;
    PinBuf<unsigned long long> h_ids;
    DevBuf<unsigned long long> d_ids_in;
    cudaStream_t mut_stream = nullptr;
    // workspaces
    std::mutex mu;
    std::vector<Workspace *> free_ws;
    std::map<void *, Workspace *> dev_ws;
    // options / stats
    int nstreams = 2;
    int timing = 1;
    int force_mode = -1;
    uint64_t launches = 0, escalations = 0, uncertain = 0;
    std::vector<float> last_times;
    Workspace *last_timed_ws = nullptr;
    int scan_warps = 16, scan_stages = 2, scan_tile_chunks = 8;
    bool scan_geometry_set = false; // SZG_OPT_SCAN_* given: no automatic choice
    int batch_disabled = 0; // SZG_OPT_BATCH_TENSOR = 0 routes szg_search_batch to the streaming scan
    // 16-bit collections: byte-planar copy of the codes, the operand of the batched path (rebuilt lazily after mutations)
    DevBuf<uint4> planar;
    bool planar_dirty = true;
    uint32_t planar_nblk = 0;
    uint64_t batch_queries = 0;
    // metadata columns (filter.cu): allocated on first use, sized to `capacity`
    DevBuf<unsigned char> doc_kind;
    DevBuf<unsigned char> col_kind[kFilterMaxCols];
    DevBuf<unsigned long long> col_val[kFilterMaxCols];
    bool meta_used = false;
    // combining of concurrent single-query calls (szg_search_topk)
    struct PendingSearch;
    std::mutex comb_mu;
    std::vector<PendingSearch *> comb_queue;
    bool comb_leader = false;
    int combine = 1;
    uint64_t combined_queries = 0;
    DevBuf<unsigned char> d_filter_blob; // program + tables + ranks of the filter being evaluated (kept between calls)
    std::unordered_map<std::string, uint32_t> dict;
    std::vector<std::string> dict_strs;
    int digits = 0; // 0 = automatic (2-digit fast pass, 3-digit re-run when uncertain), 2 or 3 = forced // streaming geometry (SZG_OPT_SCAN_*)

    bool lookup(uint64_t id, uint32_t *slot) const {
        auto it = map.find(id);
        if (it != map.end()) { *slot = it->second; return true; }
        for (const auto &r : ranges)
            if (id >= r.id0 && id - r.id0 < r.n) {
                if (!range_dead.empty() && range_dead.count(id)) return false;
                *slot = r.slot0 + (uint32_t)(id - r.id0);
                return true;
            }
        return false;
    }
    RowsArgs rows_args() {
        RowsArgs a;
        a.codes = codes.p; a.aux = aux.p; a.live = live.p; a.ids = ids.p;
        a.C = C; a.dims = (uint32_t)dim; a.metric = (uint32_t)metric; a.maxint = maxint; a.rowbytes = rowbytes;
        a.qt = qt;
        return a;
    }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(h)                                                                 \
    if (!(h)) return fail(SZG_EINVAL, "null handle");                            \
    DeviceGuard guard_((h)->device);                                             \
    if (!guard_.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", (h)->device)

int grow(szg_index *h, uint64_t want_slots) {
    if (want_slots <= h->capacity) return SZG_OK;
    if (want_slots > 0xFFFFFF00ull) return fail(SZG_EINVAL, "more than 2^32 rows per mirror are not supported");
    uint64_t cap = std::max<uint64_t>(want_slots, (uint64_t)h->capacity * 2);
    cap = std::max<uint64_t>(cap, 1024);
    cap = (cap + 63) / 64 * 64;
    if (cap > 0xFFFFFF00ull) cap = 0xFFFFFF00ull / 64 * 64;
    cudaStream_t st = h->mut_stream;
    int rc;
    if ((rc = h->codes.ensure((size_t)cap * h->C, true, st))) return rc;
    if ((rc = h->ids.ensure(cap, true, st))) return rc;
    if ((rc = h->aux.ensure(cap, true, st))) return rc;
    size_t old_words = h->live.n, words = cap / 32;
    if ((rc = h->live.ensure(words, true, st))) return rc;
    CK(cudaMemsetAsync(h->live.p + old_words, 0, (words - old_words) * 4, st));
    for (auto &m : h->masks) {
        uint32_t *np = nullptr;
        CK(cudaMalloc(&np, words * 4));
        CK(cudaMemsetAsync(np, 0, words * 4, st));
        if (old_words) CK(cudaMemcpyAsync(np, m.second, old_words * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        cudaFree(m.second);
        m.second = np;
    }
    // metadata columns follow the capacity; new rows read as "no metadata"
    auto grow_bytes = [&](DevBuf<unsigned char> &b) -> int {
        const size_t old_n = b.n;
        if (!old_n) return SZG_OK; // never used: allocated at first use
        int r = b.ensure(cap, true, st);
        if (r) return r;
        CK(cudaMemsetAsync(b.p + old_n, 0, cap - old_n, st));
        return SZG_OK;
    };
    if ((rc = grow_bytes(h->doc_kind))) return rc;
    for (uint32_t c = 0; c < kFilterMaxCols; ++c) {
        if ((rc = grow_bytes(h->col_kind[c]))) return rc;
        if (h->col_val[c].n && (rc = h->col_val[c].ensure(cap, true, st))) return rc;
    }
    CK(cudaStreamSynchronize(st));
    h->capacity = (uint32_t)cap;
    return SZG_OK;
}

int acquire_ws(szg_index *h, Workspace **out) {
    {
        std::lock_guard<std::mutex> lk(h->mu);
        if (!h->free_ws.empty()) {
            *out = h->free_ws.back();
            h->free_ws.pop_back();
            return SZG_OK;
        }
    }
    Workspace *ws = new Workspace();
    int rc = ws->init(true);
    if (rc) { ws->destroy(); delete ws; return rc; }
    *out = ws;
    return SZG_OK;
}
void release_ws(szg_index *h, Workspace *ws) {
    std::lock_guard<std::mutex> lk(h->mu);
    h->free_ws.push_back(ws);
}

int mode_for_k(const szg_index *h, uint32_t k) {
    int mode = 0;
    while (mode < 3 && (32u << mode) < k + std::max<uint32_t>(8, k / 8)) ++mode;
    if (h->force_mode >= 0 && h->force_mode <= 3 && h->force_mode > mode) mode = h->force_mode;
    return mode;
}

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
    if (!scan_plan(h->C, warps, (uint32_t)h->scan_stages, tile_chunks, pq_stride(h, nd), kScanSmemLimit, p))
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

// run_topk for short rows: the launch is cut in (query, part) items handled by one CTA each (scan_small.cuh)
int run_topk_small(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags,
                   int nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags) {
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
    // measured (profiles/r01b_scan_small_vs_general.log): +5..7 % at 48 chunks, -5..10 % on rows of <= 8 chunks, nothing at 24
    static const int const_env = getenv("SZG_SMALL_CONST") ? atoi(getenv("SZG_SMALL_CONST")) : -1;
    const bool const_ok = const_env >= 0 ? const_env != 0 : h->C >= 32;
    const size_t window_q = std::max<size_t>(1, (size_t)kConstSlots * 16 / stride);
    const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>({(size_t)nq, (size_t)4096, kCandBytes / ((size_t)sms * kSmallWarps * 32 * 8),
                                                                             const_ok ? window_q : (size_t)4096}));
    const bool timing = h->timing != 0;
    const uint32_t nlaunch = (nq + chunk - 1) / chunk;
    uint32_t tbase = 0;
    if (timing && h->timing == 2 && h->last_timed_ws == ws && ws->timed + nlaunch <= 65536) tbase = ws->timed;
    if (timing) {
        while (ws->t0.size() < tbase + nlaunch) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            ws->t0.push_back(a);
            ws->t1.push_back(b);
        }
    }
    ScanArgs a;
    fill_scan_args(h, a, mask);
    a.pq_stride = stride;
    FinalizeArgs f;
    f.codes = h->codes.p; f.ids = h->ids.p; f.lut = h->lut.p;
    f.pq_stride = stride;
    f.C = h->C; f.dims = (uint32_t)h->dim; f.metric = (uint32_t)h->metric; f.k = k;
    f.flags = flags & SZG_F_NO_FP64_VERIFY;
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
        CK(launch_finalize(h->qt, 0, m, main, f));
        h->launches += 2;
    }
    if (timing) { ws->timed = tbase + nlaunch; h->last_timed_ws = ws; }
    return SZG_OK;
}

// Enqueues prep + scan + finalize for nq queries on ws->main.  One scan launch serves a whole chunk of
// queries (persistent warps walk query after query); the chunk size is bounded by the candidate
// buffer.  Inputs/outputs are device pointers; `ws` supplies scratch.
int run_topk(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask,
             uint32_t flags, int mode, int nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
             uint32_t *d_out_flags) {
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
    if (small) return run_topk_small(h, ws, d_q, nq, k, mask, flags, nd, d_out_ids, d_out_dist, d_out_n, d_out_flags);
    const size_t nlists = (size_t)grid * plan.warps;
    const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>({(size_t)nq, (size_t)4096, kCandBytes / (nlists * Kp * 8)}));
    if ((rc = ws->d_cand.ensure((size_t)chunk * nlists * Kp))) return rc;

    cudaStream_t main = ws->main;
    PrepArgs pa;
    pa.queries = d_q; pa.pq = ws->d_pq.p; pa.pq_stride = stride;
    pa.dims = (uint32_t)h->dim; pa.C = h->C; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
    pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 0; pa.radius = 0.0;
    CK(launch_prep(nq, main, pa));
    h->launches++;
    const bool timing = h->timing != 0;
    const uint32_t nlaunch = (nq + chunk - 1) / chunk;
    // timing == 2 accumulates events over calls (bounded) until szg_last_scan_times_ms drains them
    uint32_t tbase = 0;
    if (timing && h->timing == 2 && h->last_timed_ws == ws && ws->timed + nlaunch <= 65536) tbase = ws->timed;
    if (timing) {
        while (ws->t0.size() < tbase + nlaunch) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            ws->t0.push_back(a);
            ws->t1.push_back(b);
        }
    }
    ScanArgs a;
    fill_scan_args(h, a, mask);
    a.Ct = plan.Ct; a.stages = plan.stages; a.pq_smem_off = plan.pq_smem_off;
    a.pq_stride = stride;
    a.cand = ws->d_cand.p;
    FinalizeArgs f;
    f.codes = h->codes.p; f.ids = h->ids.p; f.lut = h->lut.p; f.cand = ws->d_cand.p;
    f.pq_stride = stride;
    f.C = h->C; f.dims = (uint32_t)h->dim; f.metric = (uint32_t)h->metric; f.k = k;
    f.flags = flags & SZG_F_NO_FP64_VERIFY; f.nlists = (uint32_t)nlists;
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
        CK(launch_finalize(h->qt, mode, m, main, f));
        h->launches += 2;
    }
    if (timing) { ws->timed = tbase + nlaunch; h->last_timed_ws = ws; }
    return SZG_OK;
}

int get_mask(szg_index *h, int mask_id, const uint32_t **out) {
    *out = nullptr;
    if (mask_id < 0) return SZG_OK;
    auto it = h->masks.find(mask_id);
    if (it == h->masks.end()) return fail(SZG_ENOTFOUND, "unknown mask id %d", mask_id);
    *out = it->second;
    return SZG_OK;
}

int check_search(szg_index *h, const void *q, uint32_t nq) {
    if (!q && nq) return fail(SZG_EINVAL, "null query");
    (void)h;
    return SZG_OK;
}

// Copies the results of the first pass (already enqueued on ws->main into ws->d_out_*) to the host and
// re-runs, together, the queries whose candidate set could not be certified: first with the 3-digit
// (precise) surrogate, then with larger candidate sets.
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

int collect_and_escalate(szg_index *h, Workspace *ws, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags,
                         int nd0, int mode0, uint64_t *out_ids, double *out_dist, uint32_t *out_n, const OutPack *pack = nullptr) {
    int rc;
    cudaStream_t st = ws->main;
    const size_t on = (size_t)nq * k;
    if (pack) {
        CK(cudaMemcpyAsync(pack->h, pack->d, pack->bytes(), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(out_ids, pack->h, on * 8);
        memcpy(out_dist, pack->h + on * 8, on * 8);
        memcpy(out_n, pack->h + on * 16, nq * 4);
        memcpy(ws->h_out_flags.p, pack->h + on * 16 + nq * 4, nq * 4);
    } else {
        CK(cudaMemcpyAsync(ws->h_out_ids.p, ws->d_out_ids.p, on * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_dist.p, ws->d_out_dist.p, on * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_n.p, ws->d_out_n.p, nq * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_flags.p, ws->d_out_flags.p, nq * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(out_ids, ws->h_out_ids.p, on * 8);
        memcpy(out_dist, ws->h_out_dist.p, on * 8);
        memcpy(out_n, ws->h_out_n.p, nq * 4);
    }
    // Queries whose candidate set could not be certified are re-run together, first with the
    // 3-digit (precise) surrogate, then with larger candidate sets.
    if (!(flags & SZG_F_NO_FP64_VERIFY)) {
        std::vector<uint32_t> pending;
        for (uint32_t i = 0; i < nq; ++i)
            if (ws->h_out_flags.p[i] & 1u) pending.push_back(i);
        int nd = nd0, mode = mode0;
        while (!pending.empty()) {
            if (nd == 2) nd = 3;
            else if (mode < 3) ++mode;
            else break;
            const uint32_t m = (uint32_t)pending.size();
            h->escalations += m;
            if ((rc = ws->d_q2.ensure((size_t)m * h->dim))) return rc;
            for (uint32_t j = 0; j < m; ++j)
                CK(cudaMemcpyAsync(ws->d_q2.p + (size_t)j * h->dim, ws->d_q.p + (size_t)pending[j] * h->dim,
                                   (size_t)h->dim * sizeof(double), cudaMemcpyDeviceToDevice, st));
            if ((rc = run_topk(h, ws, ws->d_q2.p, m, k, mask, flags, mode, nd, ws->d_out_ids.p, ws->d_out_dist.p,
                               ws->d_out_n.p, ws->d_out_flags.p)))
                return rc;
            CK(cudaMemcpyAsync(ws->h_out_ids.p, ws->d_out_ids.p, (size_t)m * k * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ws->h_out_dist.p, ws->d_out_dist.p, (size_t)m * k * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ws->h_out_n.p, ws->d_out_n.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ws->h_out_flags.p, ws->d_out_flags.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
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
    }
    return SZG_OK;
}

} // namespace

// geometry of a handle, for spanfile.cu
namespace szg {
void index_geometry(const szg_index *h, int *dim, int *quant, int *metric, uint32_t *rowbytes) {
    *dim = h->dim; *quant = h->quant; *metric = h->metric; *rowbytes = h->rowbytes;
}
} // namespace szg

// ====================================================================== C ABI
extern "C" {

const char *szg_last_error(void) { return g_err.c_str(); }

int szg_create(int dim, int quantization, int metric, int device, szg_index **out) {
    if (!out) return fail(SZG_EINVAL, "null out pointer");
    *out = nullptr;
    if (quantization == 0) quantization = 64; // collection.go:254-256
    int qt;
    switch (quantization) {
    case 4: qt = Q4; break;
    case 8: qt = Q8; break;
    case 16: qt = Q16; break;
    case 32: qt = F32; break;
    case 64: qt = F64; break;
    default: return fail(SZG_EINVAL, "unsupported quantization %d (collection.go:796-811 panics)", quantization);
    }
    if (metric != SZG_EUCLIDEAN && metric != SZG_COSINE)
        return fail(SZG_EINVAL, "unsupported distance method %d (collection.go:275-283)", metric);
    if (dim < 1 || dim > SZG_MAX_DIM) return fail(SZG_EINVAL, "dimension %d outside [1, %d]", dim, SZG_MAX_DIM);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SZG_ECUDA, "no CUDA device: %s (this library has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SZG_EINVAL, "device %d out of range (have %d)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SZG_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    std::unique_ptr<szg_index> h(new szg_index());
    h->dim = dim; h->quant = quantization; h->metric = metric; h->device = device; h->qt = qt;
    h->maxint = quantization <= 16 ? (1u << quantization) - 1u : 0u;
    h->rowbytes = quantization == 4 ? (uint32_t)(dim + 1) / 2 : (uint32_t)dim * (quantization / 8);
    h->C = (h->rowbytes + 15) / 16;
    h->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->mut_stream, cudaStreamNonBlocking));
    CK(scan_configure(qt, kScanSmemLimit));
    if (quantization <= 16) {
        const size_t n = (size_t)1 << quantization;
        std::vector<double> lut(n);
        const double maxInt = (double)h->maxint;
        for (size_t v = 0; v < n; ++v) lut[v] = ((double)v / maxInt) * 2 - 1; // quantization.go:34-35
        int rc = h->lut.ensure(n);
        if (rc) return rc;
        CK(cudaMemcpy(h->lut.p, lut.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    }
    int rc = grow(h.get(), 1024);
    if (rc) return rc;
    *out = h.release();
    return SZG_OK;
}

int szg_destroy(szg_index *h) {
    if (!h) return SZG_OK;
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    for (auto ws : h->free_ws) { ws->destroy(); delete ws; }
    for (auto &kv : h->dev_ws) { kv.second->destroy(); delete kv.second; }
    for (auto &m : h->masks) cudaFree(m.second);
    h->codes.release(); h->ids.release(); h->aux.release(); h->live.release(); h->lut.release(); h->planar.release();
    h->h_stage.release(); h->d_stage.release(); h->d_vec.release();
    h->doc_kind.release(); h->d_filter_blob.release();
    for (uint32_t c = 0; c < kFilterMaxCols; ++c) { h->col_kind[c].release(); h->col_val[c].release(); } h->h_slots.release(); h->d_slots.release();
    h->h_ids.release(); h->d_ids_in.release();
    if (h->mut_stream) cudaStreamDestroy(h->mut_stream);
    delete h;
    return SZG_OK;
}

int szg_set_option(szg_index *h, int option, int64_t value) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    switch (option) {
    case SZG_OPT_STREAMS:
        if (value < 1 || value > kMaxStreams) return fail(SZG_EINVAL, "streams must be in [1, %d]", kMaxStreams);
        h->nstreams = (int)value;
        return SZG_OK;
    case SZG_OPT_TIMING:
        if (value < 0 || value > 2) return fail(SZG_EINVAL, "timing must be 0, 1 or 2");
        h->timing = (int)value;
        return SZG_OK;
    case SZG_OPT_SCAN_WARPS:
        if (value != 8 && value != 16) return fail(SZG_EINVAL, "scan warps must be 8 or 16");
        h->scan_warps = (int)value;
        h->scan_geometry_set = true;
        return SZG_OK;
    case SZG_OPT_SCAN_STAGES:
        if (value < 2 || value > kMaxStages) return fail(SZG_EINVAL, "scan stages must be in [2, %d]", kMaxStages);
        h->scan_stages = (int)value;
        h->scan_geometry_set = true;
        return SZG_OK;
    case SZG_OPT_SCAN_TILE_CHUNKS:
        if (value < 1 || value > kMaxTileChunks) return fail(SZG_EINVAL, "tile chunks must be in [1, %d]", kMaxTileChunks);
        h->scan_tile_chunks = (int)value;
        h->scan_geometry_set = true;
        return SZG_OK;
    case SZG_OPT_BATCH_TENSOR: h->batch_disabled = value == 0; return SZG_OK;
    case SZG_OPT_COMBINE: h->combine = value != 0; return SZG_OK;
    case SZG_OPT_DIGITS:
        if (value != 0 && value != 2 && value != 3) return fail(SZG_EINVAL, "digits must be 0 (auto), 2 or 3");
        h->digits = (int)value;
        return SZG_OK;
    case SZG_OPT_MIN_CANDIDATE_MODE:
        if (value < -1 || value > 3) return fail(SZG_EINVAL, "candidate mode must be in [-1, 3]");
        h->force_mode = (int)value;
        return SZG_OK;
    }
    return fail(SZG_EINVAL, "unknown option %d", option);
}

int szg_reserve(szg_index *h, uint64_t nrows) {
    GUARD(h);
    return grow(h, nrows);
}

int szg_count(szg_index *h, uint64_t *n) {
    if (!h || !n) return fail(SZG_EINVAL, "null argument");
    *n = h->live_rows;
    return SZG_OK;
}

// One staged batch of rows into the mirror: slot assignment on the host, then scatter + aux on the device.  The
// stream-1 bytes come either from the caller (codes) or from encode_kernel over the caller's float64 vectors, in
// which case they are also handed back (out_codes) for the span file.
static int upsert_rows(szg_index *h, const uint64_t *ids, const uint8_t *codes, const double *vectors, uint8_t *out_codes,
                       uint64_t n, bool into_mirror) {
    if (into_mirror) h->planar_dirty = true;
    uint64_t per = std::max<uint64_t>(1, kStageBytes / h->rowbytes);
    if (vectors) per = std::max<uint64_t>(1, std::min<uint64_t>(per, kStageBytes / ((uint64_t)h->dim * sizeof(double))));
    int rc;
    for (uint64_t off = 0; off < n; off += per) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(per, n - off);
        if ((rc = h->h_slots.ensure(m)) || (rc = h->d_slots.ensure(m)) || (rc = h->h_ids.ensure(m)) ||
            (rc = h->d_ids_in.ensure(m)) || (rc = h->h_stage.ensure((size_t)m * h->rowbytes)) ||
            (rc = h->d_stage.ensure((size_t)m * h->rowbytes)))
            return rc;
        if (vectors && (rc = h->d_vec.ensure((size_t)m * h->dim))) return rc;
        cudaStream_t st = h->mut_stream;
        RowsArgs ra = h->rows_args();
        if (vectors) {
            CK(cudaMemcpyAsync(h->d_vec.p, vectors + off * (uint64_t)h->dim, (size_t)m * h->dim * sizeof(double),
                               cudaMemcpyHostToDevice, st));
            CK(launch_encode(ra, h->d_vec.p, h->d_stage.p, m, st));
            h->launches += 1;
            if (out_codes)
                CK(cudaMemcpyAsync(out_codes + off * h->rowbytes, h->d_stage.p, (size_t)m * h->rowbytes, cudaMemcpyDeviceToHost, st));
        } else {
            memcpy(h->h_stage.p, codes + off * h->rowbytes, (size_t)m * h->rowbytes);
            CK(cudaMemcpyAsync(h->d_stage.p, h->h_stage.p, (size_t)m * h->rowbytes, cudaMemcpyHostToDevice, st));
        }
        if (into_mirror) {
            // slot assignment.  A batch may name an id twice: the last one wins, like two
            // AddDocument calls in a row; earlier duplicates are skipped (slot 0xFFFFFFFF).
            if ((rc = grow(h, (uint64_t)h->nslots + m))) return rc;
            std::unordered_map<uint64_t, uint32_t> last_in_batch;
            last_in_batch.reserve(m);
            for (uint32_t i = 0; i < m; ++i) last_in_batch[ids[off + i]] = i;
            for (uint32_t i = 0; i < m; ++i) {
                const uint64_t id = ids[off + i];
                uint32_t slot = 0xFFFFFFFFu;
                if (last_in_batch[id] == i && !h->lookup(id, &slot)) {
                    if (!h->free_slots.empty()) {
                        slot = h->free_slots.back();
                        h->free_slots.pop_back();
                    } else {
                        slot = h->nslots++;
                    }
                    h->map[id] = slot; // shadows a (dead) synthetic-range entry of the same id
                    h->live_rows++;
                }
                h->h_slots.p[i] = slot;
                h->h_ids.p[i] = id;
            }
            ra = h->rows_args(); // grow() may have moved the arrays
            CK(cudaMemcpyAsync(h->d_slots.p, h->h_slots.p, (size_t)m * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(h->d_ids_in.p, h->h_ids.p, (size_t)m * 8, cudaMemcpyHostToDevice, st));
            CK(launch_scatter(ra, h->d_stage.p, h->d_slots.p, h->d_ids_in.p, m, st));
            CK(launch_aux(ra, h->d_slots.p, 0, m, st));
            h->launches += 2;
        }
        CK(cudaStreamSynchronize(st));
    }
    return SZG_OK;
}

int szg_upsert(szg_index *h, const uint64_t *ids, const uint8_t *codes, uint64_t n) {
    GUARD(h);
    if (n && (!ids || !codes)) return fail(SZG_EINVAL, "null ids/codes");
    return upsert_rows(h, ids, codes, nullptr, nullptr, n, true);
}

int szg_encode(szg_index *h, const uint64_t *ids, const double *vectors, uint64_t n, uint8_t *out_codes, int upsert) {
    GUARD(h);
    if (n && !vectors) return fail(SZG_EINVAL, "null vectors");
    if (n && upsert && !ids) return fail(SZG_EINVAL, "null ids");
    if (!upsert && !out_codes) return fail(SZG_EINVAL, "nothing to do: no output buffer and no upsert");
    return upsert_rows(h, ids, nullptr, vectors, out_codes, n, upsert != 0);
}

int szg_remove(szg_index *h, const uint64_t *ids, uint64_t n, uint64_t *n_removed) {
    GUARD(h);
    if (n && !ids) return fail(SZG_EINVAL, "null ids");
    std::vector<uint32_t> slots;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t slot;
        if (!h->lookup(ids[i], &slot)) continue;
        auto it = h->map.find(ids[i]);
        if (it != h->map.end()) h->map.erase(it);
        else h->range_dead.insert(ids[i]);
        h->free_slots.push_back(slot);
        slots.push_back(slot);
        h->live_rows--;
    }
    if (n_removed) *n_removed = slots.size();
    if (slots.empty()) return SZG_OK;
    int rc;
    if ((rc = h->d_slots.ensure(slots.size()))) return rc;
    cudaStream_t st = h->mut_stream;
    CK(cudaMemcpyAsync(h->d_slots.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice, st));
    CK(launch_kill(h->live.p, h->d_slots.p, (uint32_t)slots.size(), st));
    h->launches++;
    if (h->meta_used) { // a slot that is handed to another document later must not show this one's metadata
        MetaPtrs mp;
        mp.doc_kind = h->doc_kind.p;
        for (uint32_t c = 0; c < kFilterMaxCols; ++c) mp.col_kind[c] = h->col_kind[c].p;
        CK(launch_meta_clear(h->d_slots.p, (uint32_t)slots.size(), mp, st));
        h->launches++;
    }
    CK(cudaStreamSynchronize(st));
    return SZG_OK;
}

int szg_fill_synthetic(szg_index *h, uint64_t seed, uint64_t row0, uint64_t nrows) {
    GUARD(h);
    if (!nrows) return SZG_OK;
    h->planar_dirty = true;
    if ((uint64_t)h->nslots + nrows > 0xFFFFFF00ull) return fail(SZG_EINVAL, "too many rows");
    for (const auto &r : h->ranges)
        if (row0 < r.id0 + r.n && r.id0 < row0 + nrows) return fail(SZG_EINVAL, "synthetic range overlaps an existing one");
    int rc = grow(h, (uint64_t)h->nslots + nrows);
    if (rc) return rc;
    RowsArgs ra = h->rows_args();
    cudaStream_t st = h->mut_stream;
    const uint32_t slot0 = h->nslots;
    const uint64_t step = 1u << 22;
    for (uint64_t off = 0; off < nrows; off += step) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(step, nrows - off);
        CK(launch_synth(ra, seed, row0 + off, slot0 + (uint32_t)off, m, st));
        CK(launch_aux(ra, nullptr, slot0 + (uint32_t)off, m, st));
        h->launches += 2;
    }
    CK(cudaStreamSynchronize(st));
    h->ranges.push_back(IdRange{row0, slot0, (uint32_t)nrows});
    h->nslots += (uint32_t)nrows;
    h->live_rows += nrows;
    return SZG_OK;
}

static int map_ids(szg_index *h, const uint64_t *ids, uint64_t n, uint32_t *slots, bool require) {
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s;
        if (h->lookup(ids[i], &s)) slots[i] = s;
        else if (require) return fail(SZG_ENOTFOUND, "id %llu is not in the mirror", (unsigned long long)ids[i]);
        else slots[i] = 0xFFFFFFFFu;
    }
    return SZG_OK;
}

int szg_fetch_codes(szg_index *h, const uint64_t *ids, uint64_t n, uint8_t *out_codes) {
    GUARD(h);
    if (n && (!ids || !out_codes)) return fail(SZG_EINVAL, "null argument");
    const uint64_t per = std::max<uint64_t>(1, kStageBytes / h->rowbytes);
    int rc;
    for (uint64_t off = 0; off < n; off += per) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(per, n - off);
        if ((rc = h->h_slots.ensure(m)) || (rc = h->d_slots.ensure(m)) ||
            (rc = h->h_stage.ensure((size_t)m * h->rowbytes)) || (rc = h->d_stage.ensure((size_t)m * h->rowbytes)))
            return rc;
        if ((rc = map_ids(h, ids + off, m, h->h_slots.p, true))) return rc;
        cudaStream_t st = h->mut_stream;
        CK(cudaMemcpyAsync(h->d_slots.p, h->h_slots.p, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(launch_fetch(h->rows_args(), h->d_slots.p, m, h->d_stage.p, st));
        h->launches++;
        CK(cudaMemcpyAsync(h->h_stage.p, h->d_stage.p, (size_t)m * h->rowbytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(out_codes + off * h->rowbytes, h->h_stage.p, (size_t)m * h->rowbytes);
    }
    return SZG_OK;
}

int szg_mask_create(szg_index *h, const uint64_t *ids, const uint8_t *pass, uint64_t n, int *mask_id) {
    GUARD(h);
    if (!mask_id || (n && (!ids || !pass))) return fail(SZG_EINVAL, "null argument");
    if (n > 0xFFFFFFFFull) return fail(SZG_EINVAL, "too many ids");
    const size_t words = h->capacity / 32;
    uint32_t *mask = nullptr;
    CK(cudaMalloc(&mask, words * 4));
    cudaStream_t st = h->mut_stream;
    cudaError_t e = cudaMemsetAsync(mask, 0, words * 4, st);
    int rc = SZG_OK;
    if (e != cudaSuccess) rc = fail(SZG_ECUDA, "memset failed: %s", cudaGetErrorString(e));
    if (!rc && n) {
        std::vector<uint32_t> slots(n);
        map_ids(h, ids, n, slots.data(), false);
        DevBuf<unsigned char> d_pass;
        if (!(rc = h->d_slots.ensure(n)) && !(rc = d_pass.ensure(n))) {
            e = cudaMemcpyAsync(h->d_slots.p, slots.data(), n * 4, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_pass.p, pass, n, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = launch_mask_set(mask, h->d_slots.p, d_pass.p, (uint32_t)n, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = fail(SZG_ECUDA, "mask upload failed: %s", cudaGetErrorString(e));
            h->launches++;
        }
        d_pass.release();
    } else if (!rc) {
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(SZG_ECUDA, "sync failed: %s", cudaGetErrorString(e));
    }
    if (rc) { cudaFree(mask); return rc; }
    *mask_id = h->next_mask++;
    h->masks[*mask_id] = mask;
    return SZG_OK;
}

// ---------------------------------------------------------------- metadata columns and device-side filters
static int meta_column_ready(szg_index *h, DevBuf<unsigned char> &kind, DevBuf<unsigned long long> *val) {
    int rc;
    cudaStream_t st = h->mut_stream;
    if (!kind.n) {
        if ((rc = kind.ensure(std::max<size_t>(h->capacity, 64)))) return rc;
        CK(cudaMemsetAsync(kind.p, 0, kind.n, st));
    }
    if (val && !val->n) {
        if ((rc = val->ensure(std::max<size_t>(h->capacity, 64)))) return rc;
        CK(cudaMemsetAsync(val->p, 0, val->n * 8, st));
    }
    return SZG_OK;
}

static uint32_t dict_code(szg_index *h, const char *sp, uint32_t len, bool insert) {
    std::string key(sp ? sp : "", sp ? len : 0);
    auto it = h->dict.find(key);
    if (it != h->dict.end()) return it->second;
    if (!insert) return 0xFFFFFFFFu;
    const uint32_t code = (uint32_t)h->dict_strs.size();
    h->dict_strs.push_back(key);
    h->dict.emplace(std::move(key), code);
    return code;
}

int szg_meta_upsert(szg_index *h, const uint64_t *ids, uint64_t n, const uint8_t *doc_kind, const uint32_t *cols,
                    uint32_t ncols, const szg_meta_value *values) {
    GUARD(h);
    if (n && (!ids || !doc_kind || (ncols && (!cols || !values)))) return fail(SZG_EINVAL, "null argument");
    if (n > 0xFFFFFFFFull) return fail(SZG_EINVAL, "too many ids");
    for (uint32_t j = 0; j < ncols; ++j)
        if (cols[j] >= kFilterMaxCols) return fail(SZG_EINVAL, "metadata column %u out of range (max %u)", cols[j], kFilterMaxCols - 1);
    if (!n) return SZG_OK;
    int rc;
    std::vector<uint32_t> slots(n);
    if ((rc = map_ids(h, ids, n, slots.data(), true))) return rc;
    h->meta_used = true;
    cudaStream_t st = h->mut_stream;
    if ((rc = meta_column_ready(h, h->doc_kind, nullptr))) return rc;
    DevBuf<unsigned char> d_kinds;
    DevBuf<unsigned long long> d_vals;
    if ((rc = h->d_slots.ensure(n)) || (rc = d_kinds.ensure(n)) || (rc = d_vals.ensure(n))) return rc;
    std::vector<unsigned char> kinds(n);
    std::vector<unsigned long long> vals(n);
    auto finish = [&](int code) { d_kinds.release(); d_vals.release(); return code; };
#define CKF(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return finish(fail(SZG_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); \
    } while (0)
    CKF(cudaMemcpyAsync(h->d_slots.p, slots.data(), n * 4, cudaMemcpyHostToDevice, st));
    for (uint64_t i = 0; i < n; ++i) {
        if (doc_kind[i] > SZG_DOC_OTHER) return finish(fail(SZG_EINVAL, "bad document kind %u", doc_kind[i]));
        kinds[i] = doc_kind[i];
    }
    CKF(cudaMemcpyAsync(d_kinds.p, kinds.data(), n, cudaMemcpyHostToDevice, st));
    CKF(launch_meta_scatter(h->d_slots.p, d_kinds.p, nullptr, h->doc_kind.p, nullptr, (uint32_t)n, st));
    CKF(cudaStreamSynchronize(st)); // the staging vectors are reused per column
    h->launches++;
    for (uint32_t j = 0; j < ncols; ++j) {
        const uint32_t c = cols[j];
        if ((rc = meta_column_ready(h, h->col_kind[c], &h->col_val[c]))) return finish(rc);
        for (uint64_t i = 0; i < n; ++i) {
            const szg_meta_value &v = values[i * ncols + j];
            if (v.kind > SZG_MV_ERROR) return finish(fail(SZG_EINVAL, "bad value kind %u", v.kind));
            kinds[i] = (unsigned char)v.kind;
            unsigned long long bits = 0;
            if (v.kind == SZG_MV_NUMBER) memcpy(&bits, &v.num, 8);
            else if (v.kind == SZG_MV_BOOL) bits = v.num != 0.0;
            else if (v.kind == SZG_MV_STRING) bits = dict_code(h, v.str, v.str_len, true);
            vals[i] = bits;
        }
        CKF(cudaMemcpyAsync(d_kinds.p, kinds.data(), n, cudaMemcpyHostToDevice, st));
        CKF(cudaMemcpyAsync(d_vals.p, vals.data(), n * 8, cudaMemcpyHostToDevice, st));
        CKF(launch_meta_scatter(h->d_slots.p, d_kinds.p, d_vals.p, h->col_kind[c].p, h->col_val[c].p, (uint32_t)n, st));
        CKF(cudaStreamSynchronize(st));
        h->launches++;
    }
    return finish(SZG_OK);
}

int szg_meta_dictionary_size(szg_index *h, uint32_t *size) {
    GUARD(h);
    if (!size) return fail(SZG_EINVAL, "null argument");
    *size = (uint32_t)h->dict_strs.size();
    return SZG_OK;
}

int szg_meta_dictionary_get(szg_index *h, uint32_t code, const char **str, uint32_t *len) {
    GUARD(h);
    if (!str || !len) return fail(SZG_EINVAL, "null argument");
    if (code >= h->dict_strs.size()) return fail(SZG_ENOTFOUND, "no string with code %u", code);
    *str = h->dict_strs[code].data();
    *len = (uint32_t)h->dict_strs[code].size();
    return SZG_OK;
}

int szg_filter_mask(szg_index *h, const szg_filter_op *ops, uint32_t nops, int *mask_id) {
    GUARD(h);
    if (!ops || !nops || !mask_id) return fail(SZG_EINVAL, "null argument");
    if (nops > 4096) return fail(SZG_EINVAL, "filter program too long");
    int rc;
    cudaStream_t st = h->mut_stream;
    // ---- validate the stack discipline and lower the literals
    const uint32_t D = (uint32_t)h->dict_strs.size();
    std::vector<std::string> lits; // literal strings that are not in the dictionary get virtual codes D, D + 1, ...
    auto literal_code = [&](const szg_filter_op &o) -> uint32_t {
        uint32_t c = dict_code(h, o.str, o.str_len, false);
        if (c != 0xFFFFFFFFu) return c;
        std::string key(o.str ? o.str : "", o.str ? o.str_len : 0);
        for (size_t i = 0; i < lits.size(); ++i)
            if (lits[i] == key) return D + (uint32_t)i;
        lits.push_back(key);
        return D + (uint32_t)lits.size() - 1;
    };
    std::vector<FilterOp> prog(nops);
    std::vector<std::vector<unsigned char>> tables(nops);
    bool need_rank = false;
    int sp = 0;
    for (uint32_t i = 0; i < nops; ++i) {
        const szg_filter_op &o = ops[i];
        FilterOp &f = prog[i];
        f.op = 0; f.arg = 0; f.bits = 0; f.table = nullptr;
        int pops = 0, pushes = 1;
        switch (o.op) {
        case SZG_FOP_COL:
        case SZG_FOP_EXISTS:
        case SZG_FOP_NOT_EXISTS:
            if (o.arg >= kFilterMaxCols) return fail(SZG_EINVAL, "op %u: column %u out of range", i, o.arg);
            f.op = o.op == SZG_FOP_COL ? FOP_COL : (o.op == SZG_FOP_EXISTS ? FOP_EXISTS : FOP_NOT_EXISTS);
            f.arg = o.arg;
            if ((rc = meta_column_ready(h, h->col_kind[o.arg], &h->col_val[o.arg]))) return rc; // never set: all missing
            break;
        case SZG_FOP_NUM: f.op = FOP_NUM; memcpy(&f.bits, &o.num, 8); break;
        case SZG_FOP_STR: f.op = FOP_STR; f.bits = literal_code(o); break;
        case SZG_FOP_BOOL: f.op = FOP_BOOL; f.bits = o.num != 0.0; break;
        case SZG_FOP_NULL: f.op = FOP_NULL; break;
        case SZG_FOP_EQ: f.op = FOP_EQ; pops = 2; break;
        case SZG_FOP_NE: f.op = FOP_NE; pops = 2; break;
        case SZG_FOP_LT: f.op = FOP_LT; pops = 2; need_rank = true; break;
        case SZG_FOP_LE: f.op = FOP_LE; pops = 2; need_rank = true; break;
        case SZG_FOP_GT: f.op = FOP_GT; pops = 2; need_rank = true; break;
        case SZG_FOP_GE: f.op = FOP_GE; pops = 2; need_rank = true; break;
        case SZG_FOP_AND: f.op = FOP_AND; pops = 2; break;
        case SZG_FOP_OR: f.op = FOP_OR; pops = 2; break;
        case SZG_FOP_NOT: f.op = FOP_NOT; pops = 1; break;
        case SZG_FOP_IN:
        case SZG_FOP_NOT_IN:
            f.op = o.op == SZG_FOP_IN ? FOP_IN : FOP_NOT_IN;
            f.arg = o.arg;
            if (o.arg >= (uint32_t)kFilterMaxStack) return fail(SZG_EINVAL, "op %u: list of %u elements is too long", i, o.arg);
            pops = (int)o.arg + 1;
            break;
        case SZG_FOP_CONTAINS:
        case SZG_FOP_STARTS_WITH:
        case SZG_FOP_ENDS_WITH: {
            // strings.Contains / HasPrefix / HasSuffix (compiler.go:395-420) over the dictionary, once per program
            const std::string lit(o.str ? o.str : "", o.str ? o.str_len : 0);
            std::vector<unsigned char> &t = tables[i];
            t.resize(std::max<uint32_t>(D, 1));
            for (uint32_t c = 0; c < D; ++c) {
                const std::string &x = h->dict_strs[c];
                bool r;
                if (o.op == SZG_FOP_CONTAINS) r = x.find(lit) != std::string::npos;
                else if (o.op == SZG_FOP_STARTS_WITH) r = x.size() >= lit.size() && x.compare(0, lit.size(), lit) == 0;
                else r = x.size() >= lit.size() && x.compare(x.size() - lit.size(), lit.size(), lit) == 0;
                t[c] = r;
            }
            f.op = FOP_STR_TABLE; f.arg = D; pops = 1;
            break;
        }
        case SZG_FOP_STR_TABLE:
            if (o.table_len && !o.table) return fail(SZG_EINVAL, "op %u: null table", i);
            tables[i].assign(o.table, o.table + o.table_len);
            if (tables[i].empty()) tables[i].push_back(0);
            f.op = FOP_STR_TABLE; f.arg = o.table_len; pops = 1;
            break;
        default: return fail(SZG_EINVAL, "op %u: unknown opcode %u", i, o.op);
        }
        if (sp < pops) return fail(SZG_EINVAL, "op %u: value stack underflow", i);
        sp += pushes - pops;
        if (sp > kFilterMaxStack) return fail(SZG_EINVAL, "op %u: expression nests deeper than %d values", i, kFilterMaxStack);
    }
    if (sp != 1) return fail(SZG_EINVAL, "the program leaves %d values, not one", sp);
    // ---- bytewise order of all string codes (Go compares strings bytewise, compiler.go:306-320)
    std::vector<uint32_t> rank;
    if (need_rank) {
        const uint32_t total = D + (uint32_t)lits.size();
        std::vector<uint32_t> order(total);
        for (uint32_t i = 0; i < total; ++i) order[i] = i;
        auto str_of = [&](uint32_t c) -> const std::string & { return c < D ? h->dict_strs[c] : lits[c - D]; };
        std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return str_of(x) < str_of(y); });
        rank.resize(total);
        for (uint32_t i = 0; i < total; ++i) rank[order[i]] = i;
    }
    if (rank.empty()) rank.push_back(0);
    // ---- upload and run
    if ((rc = meta_column_ready(h, h->doc_kind, nullptr))) return rc;
    size_t table_bytes = 0;
    for (auto &t : tables) table_bytes += (t.size() + 15) / 16 * 16;
    DevBuf<unsigned char> &d_blob = h->d_filter_blob; // [tables][program][ranks]
    const size_t prog_off = table_bytes, rank_off = (prog_off + nops * sizeof(FilterOp) + 15) / 16 * 16;
    const size_t blob_bytes = rank_off + rank.size() * 4;
    if ((rc = d_blob.ensure(blob_bytes))) return rc;
    std::vector<unsigned char> blob(blob_bytes, 0);
    size_t off = 0;
    for (uint32_t i = 0; i < nops; ++i) {
        if (tables[i].empty()) continue;
        memcpy(blob.data() + off, tables[i].data(), tables[i].size());
        prog[i].table = d_blob.p + off;
        off += (tables[i].size() + 15) / 16 * 16;
    }
    memcpy(blob.data() + prog_off, prog.data(), nops * sizeof(FilterOp));
    memcpy(blob.data() + rank_off, rank.data(), rank.size() * 4);
    const size_t words = h->capacity / 32;
    uint32_t *mask = nullptr;
    cudaError_t e = cudaMalloc(&mask, std::max<size_t>(words, 1) * 4);
    if (e != cudaSuccess) return fail(SZG_ENOMEM, "mask allocation failed: %s", cudaGetErrorString(e));
    FilterArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.prog = reinterpret_cast<const FilterOp *>(d_blob.p + prog_off);
    fa.nops = nops;
    fa.doc_kind = h->doc_kind.p;
    for (uint32_t c = 0; c < kFilterMaxCols; ++c) { fa.col_kind[c] = h->col_kind[c].p; fa.col_val[c] = h->col_val[c].p; }
    fa.rank = reinterpret_cast<const uint32_t *>(d_blob.p + rank_off);
    fa.nrank = (uint32_t)rank.size();
    fa.mask = mask;
    fa.nwords = (uint32_t)words;
    fa.nslots = h->nslots;
    e = cudaMemsetAsync(mask, 0, std::max<size_t>(words, 1) * 4, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_blob.p, blob.data(), blob_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = launch_filter(fa, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    h->launches++;
    if (e != cudaSuccess) { cudaFree(mask); return fail(SZG_ECUDA, "filter evaluation failed: %s", cudaGetErrorString(e)); }
    *mask_id = h->next_mask++;
    h->masks[*mask_id] = mask;
    return SZG_OK;
}

int szg_mask_destroy(szg_index *h, int mask_id) {
    GUARD(h);
    auto it = h->masks.find(mask_id);
    if (it == h->masks.end()) return fail(SZG_ENOTFOUND, "unknown mask id %d", mask_id);
    cudaFree(it->second);
    h->masks.erase(it);
    return SZG_OK;
}

static int search_topk_impl(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                            uint64_t *out_ids, double *out_dist, uint32_t *out_n, uint64_t *scanned);

// Concurrent callers.  The reference answers one query per Search call and lets calls overlap (RLock only,
// collection.go:570); a scan launch, however, owns the whole GPU, so overlapping calls would queue up one launch each and
// every one of them would stream the collection from HBM alone.  Instead the calls combine: a caller that finds no launch in
// flight becomes the leader and runs whatever is queued with its own k / mask / flags as ONE call (scan_small_kernel then
// deals the queries to CTA groups and they share rows in L2); callers arriving meanwhile wait and are answered together by
// the next leader.  Nobody waits for company: a lone caller runs at once, exactly as before.
struct szg_index::PendingSearch {
    const double *q; uint32_t nq, k; int mask_id; uint32_t flags;
    uint64_t *out_ids; double *out_dist; uint32_t *out_n;
    int rc = 0; std::string err; bool done = false;
    std::condition_variable cv; // the caller sleeps on its own variable: a finished launch wakes its callers and one new leader only
};
constexpr uint32_t kCombineMaxCall = 16;   // calls with more queries than this are not combined
constexpr uint32_t kCombineMaxBatch = 128; // queries of one combined launch
constexpr uint32_t kCombineTensorMin = 4;  // combined batches from this size on go to szg_search_batch

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
                rc = search_topk_impl(h, p->q, p->nq, p->k, p->mask_id, p->flags, p->out_ids, p->out_dist, p->out_n, nullptr);
            } else {
                const size_t d = (size_t)h->dim, kk = first->k;
                std::vector<double> q(total * d);
                std::vector<uint64_t> ids(total * kk);
                std::vector<double> dist(total * kk);
                std::vector<uint32_t> n(total);
                size_t off = 0;
                for (P *p : batch) { memcpy(q.data() + off * d, p->q, (size_t)p->nq * d * sizeof(double)); off += p->nq; }
                // a combined batch is a batch: from a few queries on, the tensor-core contraction (identical results, it falls
                // back to the scan by itself where it does not apply) answers it in about the time of one or two scans
                if (total >= kCombineTensorMin && !h->batch_disabled)
                    rc = szg_search_batch(h, q.data(), total, first->k, first->mask_id, first->flags, ids.data(), dist.data(), n.data(), nullptr);
                else
                    rc = search_topk_impl(h, q.data(), total, first->k, first->mask_id, first->flags, ids.data(), dist.data(), n.data(), nullptr);
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
        try { if (rc) err = g_err; } catch (...) {}
        if (!lk.owns_lock()) lk.lock();
        if (batch.size() > 1 && !rc) h->combined_queries += total;
        for (P *p : batch) {
            p->rc = rc; p->err = err; p->done = true;
            if (p != &me) p->cv.notify_one();
        }
        h->comb_leader = false;
        if (!h->comb_queue.empty() && h->comb_queue.front() != &me) h->comb_queue.front()->cv.notify_one(); // the next leader
    }
    if (me.rc) g_err = me.err;
    return me.rc;
}

int szg_search_topk(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                    uint64_t *out_ids, double *out_dist, uint32_t *out_n, uint64_t *scanned) {
    if (h && h->combine && queries && nq >= 1 && nq <= kCombineMaxCall && out_ids && out_dist && out_n && k >= 1 && k <= SZG_MAX_K) {
        if (scanned) *scanned = h->live_rows;
        return search_topk_combined(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n);
    }
    return search_topk_impl(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n, scanned);
}

static int search_topk_impl(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                            uint64_t *out_ids, double *out_dist, uint32_t *out_n, uint64_t *scanned) {
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
    if ((rc = ws->h_q.ensure(qn)) || (rc = ws->d_q.ensure(qn)) || (rc = ws->d_out_ids.ensure(on)) ||
        (rc = ws->d_out_dist.ensure(on)) || (rc = ws->d_out_n.ensure(nq)) || (rc = ws->d_out_flags.ensure(nq)) ||
        (rc = ws->h_out_ids.ensure(on)) || (rc = ws->h_out_dist.ensure(on)) || (rc = ws->h_out_n.ensure(nq)) ||
        (rc = ws->h_out_flags.ensure(nq)))
        return rc;
    memcpy(ws->h_q.p, queries, qn * sizeof(double));
    cudaStream_t st = ws->main;
    CK(cudaMemcpyAsync(ws->d_q.p, ws->h_q.p, qn * sizeof(double), cudaMemcpyHostToDevice, st));
    const int mode0 = mode_for_k(h, k), nd0 = first_digits(h);
    // the first pass writes its four outputs into one buffer: one copy back instead of four
    OutPack pack;
    pack.on = on; pack.nq = nq;
    if ((rc = ws->d_out_pack.ensure(pack.bytes())) || (rc = ws->h_out_pack.ensure(pack.bytes()))) return rc;
    pack.d = ws->d_out_pack.p; pack.h = ws->h_out_pack.p;
    if ((rc = run_topk(h, ws, ws->d_q.p, nq, k, mask, flags, mode0, nd0, pack.d_ids(), pack.d_dist(), pack.d_n(), pack.d_flags())))
        return rc;
    return collect_and_escalate(h, ws, nq, k, mask, flags, nd0, mode0, out_ids, out_dist, out_n, &pack);
}

// ---- batched queries on the tensor cores (batch_q8.cu)
struct BatchPlan { uint32_t slice, stages, keep, nranges, gpl, ngroups, Cb; int mode; bool p16, p4; };

// true when the tensor-core path can serve (collection, k): 8-bit rows, an even number of 16-byte chunks that
// fits the TMEM columns reserved for the query digits, candidate lists of at most 128 keys, 2-digit queries
static bool plan_batch(const szg_index *h, uint32_t nq, uint32_t k, BatchPlan *p) {
    if (h->qt > Q16 || h->digits == 3 || k < 1 || nq < 1 || h->batch_disabled || h->live_rows == 0) return false;
    // chunks of the contraction operand (16 dimensions each): the 8-bit row itself, one byte plane of a 16-bit row,
    // or the one-byte-per-code copy of a 4-bit row
    p->p16 = h->qt == Q16;
    p->p4 = h->qt == Q4;
    p->Cb = h->qt == Q8 ? h->C : (uint32_t)(h->dim + 15) / 16;
    if ((p->Cb % 2) != 0 || p->Cb > batch_max_chunks()) return false;
    p->mode = mode_for_k(h, k); // candidates per list = 32 << mode, as in the streaming scan
    if (p->mode > 2) return false;
    p->keep = 32u << p->mode;
    const size_t ring_limit = batch_dynamic_limit();
    if (ring_limit <= batch_list_bytes(p->keep)) return false;
    const size_t stage_limit = ring_limit - batch_list_bytes(p->keep);
    uint32_t want_slice = 0;
    if (const char *e = getenv("SZG_BATCH_SLICE")) want_slice = (uint32_t)atoi(e);
    p->slice = batch_slice_chunks(p->Cb, want_slice, stage_limit);
    p->stages = batch_stages(p->slice, stage_limit);
    if (p->stages < 2) return false;
    p->ngroups = (nq + 63) / 64;
    p->gpl = std::min<uint32_t>(p->ngroups, 16); // query groups per launch (they share the L2 copy of a row range)
    const uint32_t nblk = (h->nslots + 31) / 32;
    p->nranges = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)h->sm_count / p->gpl, (nblk + 3) / 4));
    return true;
}

// prep -> batch_kernel (one launch per 16 query groups) -> finalize, all on ws->main; outputs on the device
static int run_batch(szg_index *h, Workspace *ws, const BatchPlan &p, const double *d_q, uint32_t nq, uint32_t k,
                     const uint32_t *mask, uint32_t flags, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                     uint32_t *d_out_flags) {
    int rc;
    cudaStream_t st = ws->main;
    const int nd = 2; // 2 digit planes x 64 queries = the M dimension
    // the query digits are laid out like an 8-bit row of Cb chunks in both cases
    const size_t stride = sizeof(PQHeader) + (size_t)p.Cb * nd * 16;
    if ((rc = ws->d_pq.ensure(stride * nq)) || (rc = ws->d_cand.ensure((size_t)nq * p.nranges * p.keep)) ||
        (rc = ws->d_gmth.ensure((size_t)nq * p.nranges)))
        return rc;
    const uint32_t nblk_now = (h->nslots + 31) / 32;
    if (p.p16 || p.p4) {
        // (re)build the byte copy after a mutation.  16-bit: [high-byte plane | low-byte plane], each nblk x Cb x 32 uint4;
        // 4-bit: one plane, one byte per code.  Searches may run concurrently (RLock): the first one in rebuilds and
        // waits, the others wait on the mutex.
        std::lock_guard<std::mutex> lk(h->mu);
        if (h->planar_dirty || h->planar_nblk != nblk_now) {
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
    PrepArgs pa;
    pa.queries = d_q; pa.pq = ws->d_pq.p; pa.pq_stride = stride;
    pa.dims = (uint32_t)h->dim; pa.C = p.Cb; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
    pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 0; pa.radius = 0.0;
    pa.planar16 = (p.p16 || p.p4) ? 1 : 0; // digits laid out like an 8-bit row of Cb chunks
    CK(launch_prep(nq, st, pa));
    h->launches++;
    CK(batch_configure(batch_dynamic_limit()));
    BatchArgs b;
    memset(&b, 0, sizeof b);
    b.codes = (p.p16 || p.p4) ? h->planar.p : h->codes.p;
    b.codes_lo = p.p16 ? h->planar.p + (size_t)nblk_now * p.Cb * 32 : nullptr;
    b.aux = h->aux.p; b.live = h->live.p; b.mask = mask;
    b.pq = ws->d_pq.p; b.pq_stride = stride; b.cand = ws->d_cand.p; b.keep = p.keep;
    b.C = p.Cb; b.nblk = nblk_now; b.metric = (uint32_t)h->metric; b.nq = nq; b.dims = (uint32_t)h->dim;
    b.nranges = p.nranges; b.nlists = p.nranges; b.stages = p.stages; b.slice = p.slice;
    b.gmth = ws->d_gmth.p; b.mth = (p.keep + p.nranges - 1) / p.nranges;
    CK(cudaMemsetAsync(ws->d_gmth.p, 0xFF, (size_t)nq * p.nranges * sizeof(unsigned int), st));
    if (const char *dbg = getenv("SZG_BATCH_DEBUG")) b.debug = (uint32_t)atoi(dbg);
    const bool timing = h->timing != 0;
    const uint32_t nlaunch = (p.ngroups + p.gpl - 1) / p.gpl;
    uint32_t tbase = 0;
    if (timing && h->timing == 2 && h->last_timed_ws == ws && ws->timed + nlaunch <= 65536) tbase = ws->timed;
    if (timing) {
        while (ws->t0.size() < tbase + nlaunch) {
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            ws->t0.push_back(e0);
            ws->t1.push_back(e1);
        }
    }
    for (uint32_t l = 0, g0 = 0; g0 < p.ngroups; g0 += p.gpl, ++l) {
        b.group0 = g0;
        b.ngroups = std::min(p.gpl, p.ngroups - g0);
        if (timing) CK(cudaEventRecord(ws->t0[tbase + l], st));
        CK(launch_batch(b, st));
        if (timing) CK(cudaEventRecord(ws->t1[tbase + l], st));
        h->launches++;
    }
    if (timing) { ws->timed = tbase + nlaunch; h->last_timed_ws = ws; }
    FinalizeArgs f;
    f.codes = h->codes.p; f.ids = h->ids.p; f.lut = h->lut.p; f.cand = ws->d_cand.p;
    f.pq = ws->d_pq.p; f.pq_stride = stride; f.queries = d_q;
    f.C = h->C; f.dims = (uint32_t)h->dim; f.metric = (uint32_t)h->metric; f.k = k;
    f.flags = flags & SZG_F_NO_FP64_VERIFY; f.nlists = p.nranges; // one sorted list of `keep` keys per row range
    f.out_ids = d_out_ids; f.out_dist = d_out_dist; f.out_n = d_out_n; f.out_flags = d_out_flags;
    CK(launch_finalize(h->qt, p.mode, nq, st, f));
    h->launches++;
    h->batch_queries += nq;
    return SZG_OK;
}

// Batched search: same results as szg_search_topk for every query, computed by the tensor-core
// contraction kernel (batch_q8.cu) when the collection is 8-bit and the geometry fits; every other
// case is routed to the streaming scan (still on the GPU).
int szg_search_batch(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                     uint64_t *out_ids, double *out_dist, uint32_t *out_n, uint64_t *scanned) {
    GUARD(h);
    int rc;
    BatchPlan p;
    if (!plan_batch(h, nq, k, &p)) return szg_search_topk(h, queries, nq, k, mask_id, flags, out_ids, out_dist, out_n, scanned);
    if ((rc = check_search(h, queries, nq))) return rc;
    if (!out_ids || !out_dist || !out_n) return fail(SZG_EINVAL, "null output");
    if (scanned) *scanned = h->live_rows;
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    Workspace *ws;
    if ((rc = acquire_ws(h, &ws))) return rc;
    struct Rel { szg_index *h; Workspace *w; ~Rel() { release_ws(h, w); } } rel{h, ws};
    const size_t qn = (size_t)nq * h->dim, on = (size_t)nq * k;
    if ((rc = ws->h_q.ensure(qn)) || (rc = ws->d_q.ensure(qn)) || (rc = ws->d_out_ids.ensure(on)) ||
        (rc = ws->d_out_dist.ensure(on)) || (rc = ws->d_out_n.ensure(nq)) || (rc = ws->d_out_flags.ensure(nq)) ||
        (rc = ws->h_out_ids.ensure(on)) || (rc = ws->h_out_dist.ensure(on)) || (rc = ws->h_out_n.ensure(nq)) ||
        (rc = ws->h_out_flags.ensure(nq)))
        return rc;
    memcpy(ws->h_q.p, queries, qn * sizeof(double));
    CK(cudaMemcpyAsync(ws->d_q.p, ws->h_q.p, qn * sizeof(double), cudaMemcpyHostToDevice, ws->main));
    if ((rc = run_batch(h, ws, p, ws->d_q.p, nq, k, mask, flags, ws->d_out_ids.p, ws->d_out_dist.p, ws->d_out_n.p,
                        ws->d_out_flags.p)))
        return rc;
    return collect_and_escalate(h, ws, nq, k, mask, flags, 2, p.mode, out_ids, out_dist, out_n);
}

// Device-resident form of szg_search_batch (queries and outputs in HBM, everything enqueued on `stream`, no host
// synchronisation): what a row-sharded deployment calls before its all-gather + szg_merge_topk_dev.
int szg_search_batch_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                         uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, void *stream) {
    GUARD(h);
    BatchPlan p;
    if (!plan_batch(h, nq, k, &p))
        return szg_search_topk_dev(h, d_queries, nq, k, mask_id, flags, d_out_ids, d_out_dist, d_out_n, d_out_flags, stream);
    int rc;
    if ((rc = check_search(h, d_queries, nq))) return rc;
    if (k > SZG_MAX_K) return fail(SZG_EINVAL, "k=%u outside [1, %u]", k, SZG_MAX_K);
    if (!d_out_ids || !d_out_dist || !d_out_n) return fail(SZG_EINVAL, "null output");
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    Workspace *ws;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        auto it = h->dev_ws.find(stream);
        if (it == h->dev_ws.end()) {
            ws = new Workspace();
            if ((rc = ws->init(false))) { ws->destroy(); delete ws; return rc; }
            ws->main = (cudaStream_t)stream;
            h->dev_ws[stream] = ws;
        } else ws = it->second;
    }
    if (!d_out_flags) {
        if ((rc = ws->d_out_flags.ensure(nq))) return rc;
        d_out_flags = ws->d_out_flags.p;
    }
    return run_batch(h, ws, p, d_queries, nq, k, mask, flags, (unsigned long long *)d_out_ids, d_out_dist, d_out_n, d_out_flags);
}

int szg_search_topk_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                        uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags,
                        void *stream) {
    GUARD(h);
    int rc;
    if ((rc = check_search(h, d_queries, nq))) return rc;
    if (k < 1 || k > SZG_MAX_K) return fail(SZG_EINVAL, "k=%u outside [1, %u]", k, SZG_MAX_K);
    if (!nq) return SZG_OK;
    if (!d_out_ids || !d_out_dist || !d_out_n) return fail(SZG_EINVAL, "null output");
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    Workspace *ws;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        auto it = h->dev_ws.find(stream);
        if (it == h->dev_ws.end()) {
            ws = new Workspace();
            if ((rc = ws->init(false))) { ws->destroy(); delete ws; return rc; }
            ws->main = (cudaStream_t)stream;
            h->dev_ws[stream] = ws;
        } else ws = it->second;
    }
    if (!d_out_flags) {
        if ((rc = ws->d_out_flags.ensure(nq))) return rc;
        d_out_flags = ws->d_out_flags.p;
    }
    if (h->live_rows == 0) {
        CK(cudaMemsetAsync(d_out_n, 0, nq * 4, ws->main));
        CK(cudaMemsetAsync(d_out_flags, 0, nq * 4, ws->main));
        return SZG_OK;
    }
    // no host synchronisation here, hence no escalation: SZG_OPT_DIGITS = 2 (or automatic) runs the fast
    // surrogate and reports uncertified queries in d_out_flags; the caller re-runs those with szg_search_topk
    return run_topk(h, ws, d_queries, nq, k, mask, flags, mode_for_k(h, k), first_digits(h),
                    (unsigned long long *)d_out_ids, d_out_dist, d_out_n, d_out_flags);
}

int szg_merge_topk_dev(szg_index *h, const uint64_t *d_gathered_ids, const double *d_gathered_dist,
                       const uint32_t *d_gathered_n, const uint32_t *d_gathered_flags, uint64_t rank_stride_bytes,
                       uint32_t nranks, uint32_t nq, uint32_t k, uint64_t *d_out_ids, double *d_out_dist,
                       uint32_t *d_out_n, uint32_t *d_out_flags, void *stream) {
    GUARD(h);
    if (!nq) return SZG_OK;
    if (!d_gathered_ids || !d_gathered_dist || !d_gathered_n || !d_out_ids || !d_out_dist || !d_out_n)
        return fail(SZG_EINVAL, "null argument");
    if (k < 1 || k > SZG_MAX_K || nranks < 1 || (size_t)nranks * k * 16 > 200 * 1024)
        return fail(SZG_EINVAL, "merge of %u lists of k=%u is not supported", nranks, k);
    MergeArgs a;
    a.g_ids = (const unsigned long long *)d_gathered_ids; a.g_dist = d_gathered_dist; a.g_n = d_gathered_n;
    a.g_flags = d_gathered_flags; a.out_flags = d_out_flags;
    a.rank_stride = (size_t)rank_stride_bytes;
    a.G = nranks; a.nq = nq; a.k = k;
    a.out_ids = (unsigned long long *)d_out_ids; a.out_dist = d_out_dist; a.out_n = d_out_n;
    CK(launch_merge(a, (cudaStream_t)stream));
    h->launches++;
    return SZG_OK;
}

int szg_search_radius(szg_index *h, const double *query, double radius, int mask_id, uint32_t flags,
                      szg_result **out, uint64_t *scanned) {
    GUARD(h);
    int rc;
    if (!out) return fail(SZG_EINVAL, "null out pointer");
    *out = nullptr;
    if ((rc = check_search(h, query, 1))) return rc;
    if (!(radius > 0)) return fail(SZG_EINVAL, "radius must be > 0 (collection.go:598)");
    if (scanned) *scanned = h->live_rows;
    const uint32_t *mask;
    if ((rc = get_mask(h, mask_id, &mask))) return rc;
    std::unique_ptr<szg_result> res(new szg_result());
    if (h->live_rows == 0) { *out = res.release(); return SZG_OK; }
    Workspace *ws;
    if ((rc = acquire_ws(h, &ws))) return rc;
    struct Rel { szg_index *h; Workspace *w; ~Rel() { release_ws(h, w); } } rel{h, ws};
    const int nd = first_digits(h); // the radius threshold carries the surrogate error bound: no re-run needed
    const size_t stride = pq_stride(h, nd);
    if ((rc = ws->h_q.ensure(h->dim)) || (rc = ws->d_q.ensure(h->dim)) || (rc = ws->d_pq.ensure(stride)) ||
        (rc = ws->h_out_n.ensure(1)))
        return rc;
    memcpy(ws->h_q.p, query, (size_t)h->dim * sizeof(double));
    cudaStream_t st = ws->main;
    CK(cudaMemcpyAsync(ws->d_q.p, ws->h_q.p, (size_t)h->dim * sizeof(double), cudaMemcpyHostToDevice, st));
    PrepArgs pa;
    pa.queries = ws->d_q.p; pa.pq = ws->d_pq.p; pa.pq_stride = stride;
    pa.dims = (uint32_t)h->dim; pa.C = h->C; pa.metric = (uint32_t)h->metric; pa.maxint = h->maxint;
    pa.qt = h->qt; pa.nd = nd; pa.radius_mode = 1; pa.radius = radius;
    CK(launch_prep(1, st, pa));
    h->launches++;
    ScanPlan plan;
    int grid = 0;
    if ((rc = plan_scan(h, nd, &plan, &grid))) return rc;
    unsigned int *d_count = ws->d_ticket.p;
    uint32_t count = 0;
    size_t cap = std::max<size_t>(4096, h->nslots / 64);
    for (int attempt = 0; attempt < 3; ++attempt) {
        if ((rc = ws->d_slots.ensure(cap))) return rc;
        CK(cudaMemsetAsync(d_count, 0, 4, st));
        ScanArgs a;
        fill_scan_args(h, a, mask);
        a.pq = ws->d_pq.p; a.pq_stride = stride; a.nq = 1;
        a.rad_count = d_count; a.rad_slots = ws->d_slots.p; a.rad_cap = (uint32_t)ws->d_slots.n;
        a.Ct = plan.Ct; a.stages = plan.stages; a.pq_smem_off = plan.pq_smem_off;
        CK(launch_scan(h->qt, MODE_RADIUS, nd, grid, (int)plan.warps * 32, plan.smem, st, a));
        h->launches++;
        CK(cudaMemcpyAsync(ws->h_out_n.p, d_count, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        count = ws->h_out_n.p[0];
        if (count <= ws->d_slots.n) break;
        cap = count; // the compaction buffer was too small: size it exactly and rescan
    }
    if (count > ws->d_slots.n) return fail(SZG_EINTERNAL, "radius compaction buffer could not be sized");
    if (count) {
        if ((rc = ws->d_out_ids.ensure(count)) || (rc = ws->d_out_dist.ensure(count)) ||
            (rc = ws->h_out_ids.ensure(count)) || (rc = ws->h_out_dist.ensure(count)))
            return rc;
        RescoreArgs ra;
        ra.codes = h->codes.p; ra.ids = h->ids.p; ra.lut = h->lut.p; ra.q = ws->d_q.p; ra.slots = ws->d_slots.p;
        ra.count_ptr = nullptr; ra.out_dist = ws->d_out_dist.p; ra.out_ids = ws->d_out_ids.p;
        ra.C = h->C; ra.dims = (uint32_t)h->dim; ra.metric = (uint32_t)h->metric; ra.m = count; ra.qt = h->qt;
        CK(launch_rescore(ra, st));
        h->launches++;
        CK(cudaMemcpyAsync(ws->h_out_ids.p, ws->d_out_ids.p, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ws->h_out_dist.p, ws->d_out_dist.p, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        // exact inclusive test (collection.go:598) on the fp64 distances, then ascending order
        // distances are >= 0 (or NaN, which fails the test): their bit patterns order like the values, so the sort runs
        // on packed integers and only equal distances fall back to the lexicographic id comparison
        const double *hd = ws->h_out_dist.p;
        const unsigned long long *hi = ws->h_out_ids.p;
        struct Hit { unsigned long long bits; uint32_t i; };
        std::vector<Hit> keep;
        keep.reserve(count);
        for (uint32_t i = 0; i < count; ++i)
            if (hd[i] <= radius) {
                const double dv = hd[i] == 0.0 ? 0.0 : hd[i]; // -0.0 orders with +0.0
                unsigned long long b;
                memcpy(&b, &dv, 8);
                keep.push_back(Hit{b, i});
            }
        std::sort(keep.begin(), keep.end(), [&](const Hit &x, const Hit &y) {
            if (x.bits != y.bits) return x.bits < y.bits;
            return lex_less_u64(hi[x.i], hi[y.i]);
        });
        res->ids.resize(keep.size());
        res->dist.resize(keep.size());
        for (size_t i = 0; i < keep.size(); ++i) { res->ids[i] = hi[keep[i].i]; res->dist[i] = hd[keep[i].i]; }
    }
    (void)flags;
    *out = res.release();
    return SZG_OK;
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

int szg_rescore(szg_index *h, const double *query, const uint64_t *ids, uint64_t m, double *out_dist) {
    GUARD(h);
    int rc;
    if ((rc = check_search(h, query, 1))) return rc;
    if (!m) return SZG_OK;
    if (!ids || !out_dist) return fail(SZG_EINVAL, "null argument");
    if (m > 0xFFFFFFF0ull) return fail(SZG_EINVAL, "too many ids");
    Workspace *ws;
    if ((rc = acquire_ws(h, &ws))) return rc;
    struct Rel { szg_index *h; Workspace *w; ~Rel() { release_ws(h, w); } } rel{h, ws};
    if ((rc = ws->h_q.ensure(h->dim)) || (rc = ws->d_q.ensure(h->dim)) || (rc = ws->h_slots.ensure(m)) ||
        (rc = ws->d_slots.ensure(m)) || (rc = ws->d_out_dist.ensure(m)) || (rc = ws->h_out_dist.ensure(m)))
        return rc;
    memcpy(ws->h_q.p, query, (size_t)h->dim * sizeof(double));
    map_ids(h, ids, m, ws->h_slots.p, false);
    cudaStream_t st = ws->main;
    CK(cudaMemcpyAsync(ws->d_q.p, ws->h_q.p, (size_t)h->dim * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ws->d_slots.p, ws->h_slots.p, m * 4, cudaMemcpyHostToDevice, st));
    RescoreArgs ra;
    ra.codes = h->codes.p; ra.ids = h->ids.p; ra.lut = h->lut.p; ra.q = ws->d_q.p; ra.slots = ws->d_slots.p;
    ra.count_ptr = nullptr; ra.out_dist = ws->d_out_dist.p; ra.out_ids = nullptr;
    ra.C = h->C; ra.dims = (uint32_t)h->dim; ra.metric = (uint32_t)h->metric; ra.m = (uint32_t)m; ra.qt = h->qt;
    CK(launch_rescore(ra, st));
    h->launches++;
    CK(cudaMemcpyAsync(ws->h_out_dist.p, ws->d_out_dist.p, m * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out_dist, ws->h_out_dist.p, m * 8);
    return SZG_OK;
}

int szg_get_stats(szg_index *h, szg_stats *out) {
    GUARD(h);
    if (!out) return fail(SZG_EINVAL, "null out");
    memset(out, 0, sizeof *out);
    out->kernel_launches = h->launches;
    out->escalations = h->escalations;
    out->uncertain_results = h->uncertain;
    out->batch_queries = h->batch_queries;
    out->reserved0 = 0;
    out->combined_queries = h->combined_queries;
    out->device_bytes = h->codes.n * sizeof(uint4) + h->planar.n * sizeof(uint4) + h->ids.n * 8 + h->aux.n * 8 + h->live.n * 4 +
                        h->lut.n * 8 + h->masks.size() * (h->capacity / 32) * 4;
    out->live_rows = h->live_rows;
    out->slots = h->nslots;
    out->rowbytes = h->rowbytes;
    out->pitch = h->C * 16;
    out->sm_count = (uint32_t)h->sm_count;
    int grid = 0;
    ScanPlan plan;
    int rc = plan_scan(h, first_digits(h), &plan, &grid);
    if (rc) return rc;
    out->scan_grid = (uint32_t)grid;
    out->scan_block = plan.warps * 32;
    out->scan_stages = plan.stages;
    out->scan_tile_bytes = plan.Ct * 512;
    out->scan_smem_bytes = (uint32_t)plan.smem;
    return SZG_OK;
}

int szg_last_scan_times_ms(szg_index *h, float *out_ms, uint32_t cap, uint32_t *n) {
    GUARD(h);
    if (!n) return fail(SZG_EINVAL, "null n");
    *n = 0;
    Workspace *ws = h->last_timed_ws;
    if (!ws || !ws->timed) return SZG_OK;
    const uint32_t m = std::min(cap, ws->timed);
    for (uint32_t i = 0; i < m; ++i) {
        CK(cudaEventSynchronize(ws->t1[i]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ws->t0[i], ws->t1[i]));
        if (out_ms) out_ms[i] = ms;
    }
    *n = m;
    ws->timed = 0; // drained (matters for the accumulating mode)
    return SZG_OK;
}

} // extern "C"
