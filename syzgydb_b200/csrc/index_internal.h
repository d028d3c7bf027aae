// index_internal.h -- host-side state of a GPU mirror (struct szg_index), shared by the translation units that implement
// the C ABI: index.cu (one device), sharded.cu (one handle over several devices), spanfile.cu.  Host logic only.
#pragma once
#include <algorithm>
#include <atomic>
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

namespace szg {

// error channel (thread-local message behind szg_last_error)
int fail(int code, const char *fmt, ...);
const std::string &last_error_string();
void set_last_error_string(const std::string &s);

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return ::szg::fail(e_ == cudaErrorMemoryAllocation ? SZG_ENOMEM : SZG_ECUDA, "%s failed: %s (%s:%d)", \
                               #call, cudaGetErrorString(e_), __FILE__, __LINE__);                       \
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

// A captured launch sequence of one search shape (CUDA graph): H2D of the queries, prep, scan / batch, finalize, D2H.
// Valid for one workspace (the node parameters are its buffers) and one mirror generation.
struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    uint64_t generation = 0; // mirror generation the shape was last seen / captured on
    bool failed = false;     // capture did not work for this shape: plain launches
    int mode0 = 0, nd0 = 0;  // what the captured first pass runs with (the escalation ladder continues from there)
    uint32_t kernels = 0;    // kernel nodes (statistics)
    const void *buffers[7] = {}; // addresses of the workspace buffers the nodes refer to
};

struct Workspace {
    cudaStream_t main = nullptr; // owned for host calls; the caller's stream for *_dev calls
    bool owns_main = false;
    DevBuf<double> d_q, d_q2;
    DevBuf<unsigned char> d_pq;
    DevBuf<unsigned long long> d_cand; // [nq][lists][32*E] candidate keys, scan -> finalize
    DevBuf<unsigned int> d_gmth;       // batched path: per (query, row range) shared bound keys
    DevBuf<unsigned int> d_ticket;     // radius hit counters
    DevBuf<unsigned long long> d_out_ids;
    DevBuf<double> d_out_dist;
    DevBuf<uint32_t> d_out_n, d_out_flags;
    DevBuf<unsigned char> d_out_pack; // first-pass outputs of a host-buffer top-k call, packed
    PinBuf<unsigned char> h_out_pack;
    DevBuf<uint32_t> d_slots;
    DevBuf<unsigned long long> d_keys; // radius: sortable (distance bits, lexicographic rank) keys
    PinBuf<double> h_q;
    PinBuf<unsigned long long> h_out_ids;
    PinBuf<double> h_out_dist;
    PinBuf<uint32_t> h_out_n, h_out_flags, h_slots;
    std::vector<cudaEvent_t> t0, t1; // per-scan timing events
    uint32_t timed = 0;
    bool capturing = false;          // the launches being enqueued are recorded into a CUDA graph: no allocation, no events
    uint32_t radius_cap_hint = 0;    // running estimate of the radius compaction buffer (largest hit count seen + headroom)
    std::map<uint64_t, GraphEntry> graphs; // key: (nq, k, mask, flags) of a host-buffer top-k call

    int init(bool own) {
        owns_main = own;
        if (own) CK(cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking));
        int rc = d_ticket.ensure(64);
        if (rc) return rc;
        CK(cudaMemset(d_ticket.p, 0, 64 * sizeof(unsigned int)));
        return SZG_OK;
    }
    void destroy() {
        for (auto &g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        graphs.clear();
        d_q.release(); d_q2.release(); d_pq.release(); d_ticket.release(); d_out_ids.release(); d_out_dist.release();
        d_out_n.release(); d_out_flags.release(); d_slots.release(); d_out_pack.release(); h_out_pack.release();
        d_cand.release(); d_gmth.release(); d_keys.release();
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

// the string dictionary of the metadata columns: one per collection, shared by the shards of a sharded handle so that
// string codes mean the same on every device
struct MetaDict {
    std::unordered_map<std::string, uint32_t> codes;
    std::vector<std::string> strs;
};

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
    if (!(h)) return ::szg::fail(SZG_EINVAL, "null handle");                     \
    ::szg::DeviceGuard guard_((h)->device);                                      \
    if (!guard_.ok) return ::szg::fail(SZG_ECUDA, "cannot select CUDA device %d", (h)->device)

struct Sharded; // sharded.cu: the router state of a handle created by szg_create_sharded

} // namespace szg

struct szg_result {
    std::vector<uint64_t> ids;
    std::vector<double> dist;
};

struct szg_index {
    int dim = 0, quant = 0, metric = 0, device = 0, qt = 0;
    uint32_t rowbytes = 0, C = 0, maxint = 0;
    int sm_count = 0;
    // storage
    szg::DevBuf<uint4> codes;
    szg::DevBuf<unsigned long long> ids;
    szg::DevBuf<unsigned long long> aux; // 8 bytes per slot reserved; typed per (quant, metric)
    szg::DevBuf<uint32_t> live;
    szg::DevBuf<double> lut;
    uint32_t capacity = 0; // slots allocated (multiple of 64)
    uint32_t nslots = 0;   // high-water mark
    uint64_t live_rows = 0;
    uint64_t generation = 1; // bumped by every mutation: captured launch sequences of older generations are stale
    std::unordered_map<uint64_t, uint32_t> map;
    std::vector<szg::IdRange> ranges;
    std::unordered_set<uint64_t> range_dead;
    std::vector<uint32_t> free_slots;
    // filter bitmaps: creation, lookup and destruction are safe next to running searches (mask_mu); a mask must not be
    // destroyed while a search that names it is in flight
    std::mutex mask_mu;       // the table below
    std::mutex mask_build_mu; // builders (szg_mask_create / szg_filter_mask): one at a time on the staging buffers
    std::map<int, uint32_t *> masks;
    int next_mask = 1;
    // staging for mutations
    szg::PinBuf<unsigned char> h_stage;
    szg::DevBuf<unsigned char> d_stage;
    szg::DevBuf<double> d_vec; // float64 vectors of an szg_encode batch
    szg::PinBuf<uint32_t> h_slots;
    szg::DevBuf<uint32_t> d_slots;
    szg::PinBuf<unsigned long long> h_ids;
    szg::DevBuf<unsigned long long> d_ids_in;
    cudaStream_t mut_stream = nullptr;
    // workspaces
    std::mutex mu;
    std::vector<szg::Workspace *> free_ws;
    std::map<void *, szg::Workspace *> dev_ws;
    // options / stats (the counters are bumped by concurrent searches)
    int nstreams = 2;
    int timing = 0;
    int force_mode = -1;
    std::atomic<uint64_t> launches{0}, escalations{0}, uncertain{0}, batch_queries{0}, combined_queries{0}, graph_launches{0};
    std::vector<float> last_times; // scan-launch durations drained from the workspaces (guarded by mu)
    int scan_warps = 16, scan_stages = 2, scan_tile_chunks = 8;
    bool scan_geometry_set = false; // SZG_OPT_SCAN_* given: no automatic choice
    int batch_disabled = 0; // SZG_OPT_BATCH_TENSOR = 0 routes batches to the streaming scan
    int batch_min = 0;      // SZG_OPT_BATCH_MIN_QUERIES: calls with at least this many queries take the tensor-core path (0: by size)
    int use_graphs = 1;     // SZG_OPT_GRAPHS
    long long *trace = nullptr; // SZG_OPT_TRACE_BUFFER: device buffer of 8 clock64 stamps written by finalize_kernel (profiling)
    // 16-bit / 4-bit collections: byte copy of the codes, the operand of the batched path (rebuilt lazily after mutations)
    szg::DevBuf<uint4> planar;
    bool planar_dirty = true;
    uint32_t planar_nblk = 0;
    // metadata columns (filter.cu): allocated on first use, sized to `capacity`
    szg::DevBuf<unsigned char> doc_kind;
    szg::DevBuf<unsigned char> col_kind[szg::kFilterMaxCols];
    szg::DevBuf<unsigned long long> col_val[szg::kFilterMaxCols];
    bool meta_used = false;
    // combining of concurrent single-query calls (szg_search_topk)
    struct PendingSearch;
    std::mutex comb_mu;
    std::vector<PendingSearch *> comb_queue;
    bool comb_leader = false;
    int combine = 1;
    szg::DevBuf<unsigned char> d_filter_blob; // program + tables + ranks of the filter being evaluated (guarded by mask_mu)
    std::shared_ptr<szg::MetaDict> dict;
    int digits = 0; // 0 = automatic (2-digit fast pass, 3-digit re-run when uncertain), 2 or 3 = forced
    // sharded handle (szg_create_sharded): this object is then only the router; the mirrors are sh->shards
    szg::Sharded *sh = nullptr;
    szg_index *parent = nullptr; // a shard's router (NULL for a stand-alone handle)

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
    szg::RowsArgs rows_args() {
        szg::RowsArgs a;
        a.codes = codes.p; a.aux = aux.p; a.live = live.p; a.ids = ids.p;
        a.C = C; a.dims = (uint32_t)dim; a.metric = (uint32_t)metric; a.maxint = maxint; a.rowbytes = rowbytes;
        a.qt = qt;
        return a;
    }
};

namespace szg {

// ---- single-device internals (index.cu) used by the router (sharded.cu)
int acquire_ws(szg_index *h, Workspace **out);
void release_ws(szg_index *h, Workspace *ws);
int get_mask(szg_index *h, int mask_id, const uint32_t **out);
int mode_for_k(const szg_index *h, uint32_t k);
int first_digits(const szg_index *h);
int plan_scan(szg_index *h, int nd, ScanPlan *p, int *grid);

// Where a finished per-shard result goes when the handle is one shard of a sharded search: finalize_kernel then writes into
// the root device's gather buffer (peer stores) and bumps the query's arrival counter there (system-scope release).
struct PeerSink {
    uint32_t *done_cnt = nullptr; // [nq] on the root device, NULL = stand-alone
};

// Enqueues prep + (scan | tensor-core batch) + finalize for nq queries on ws->main; all pointers are device pointers.
// prefer_batch: take the tensor-core path whenever its geometry fits (szg_search_batch); otherwise from batch_min queries on.
// *mode_out / *nd_out: what the pass ran with (the escalation ladder continues from there).
int enqueue_topk(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags,
                 bool prefer_batch, int min_mode, int force_nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                 uint32_t *d_out_flags, const PeerSink *sink, int *mode_out, int *nd_out);

// moves the completed timing events of ws into h->last_times (the stream must be idle)
void drain_timing(szg_index *h, Workspace *ws);

} // namespace szg
