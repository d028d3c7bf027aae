// sharded.cu -- ONE handle over the GPUs of one box (szg_create_sharded): what a single Go process serving a collection
// (rest.go:20-23 keeps its collections in one process; Search is a method call, collection.go:569) binds when the mirror
// does not fit, or should not be scanned by, one device.  SURVEY.md section 8e: rows are dealt to the devices, every device
// scans its rows for the whole query batch, the per-device lists are merged on devices[0].
//
// No NCCL and no host hop in the exchange step: the shards' finalize kernels write their k x 16-byte lists STRAIGHT INTO
// devices[0]'s gather buffer (peer stores over NVLink / NVSwitch; every device maps the root's memory) and then bump the
// query's arrival counter there with a system-scope release; the root's merge kernel is already launched behind the root's
// own scan and waits on the counters (acquire).  The only cross-device stream dependency is the start event (queries
// uploaded, counters cleared).  The multi-process / NCCL all-gather variant (syzgydb_b200/sharded.py) remains as a cross-check.
//
// Row ownership: a record lives on shard id % G, except ids inside a synthetic range (szg_fill_synthetic), which is cut in
// G contiguous pieces.  Results do not depend on G: the final order is (distance, lexicographic id).
#include <thread>

#include "sharded.h"

using namespace szg;

namespace szg {

namespace {

struct ShardWs { // one in-flight sharded search
    std::vector<Workspace *> ws;     // per shard; ws[0] is the root's (or the caller's stream for device-resident calls)
    DevBuf<unsigned char> d_gather;  // root: G records [ids nq*k | dist nq*k | n nq | flags nq]
    DevBuf<uint32_t> d_done;         // root: [nq] arrival counters + one error word
    DevBuf<unsigned char> d_out;     // root: merged [ids | dist | n | flags | err]
    PinBuf<unsigned char> h_out;
    DevBuf<double> d_q2;             // root: queries of an escalation pass
    cudaEvent_t ev_start = nullptr;
    std::vector<cudaEvent_t> ev_done; // per shard: its lists are written (used when the merge does not spin, see sharded_pass)
    bool bound = false;              // ws[0] belongs to a caller's stream
    bool capturing = false;          // the pass being enqueued is recorded into a CUDA graph (all streams of the workspace)
    std::map<uint64_t, GraphEntry> graphs; // captured host-buffer calls, as in search_host (search.cu)
};

struct SynthRange {
    uint64_t id0, n;
    std::vector<uint64_t> cut; // G + 1 offsets: shard g holds ids [id0 + cut[g], id0 + cut[g + 1])
};

size_t record_bytes(uint32_t nq, uint32_t k) { return (size_t)nq * k * 16 + (size_t)nq * 8; }

} // namespace

struct Sharded {
    std::vector<szg_index *> shards;
    std::vector<int> devices;
    std::vector<SynthRange> ranges;
    std::mutex mu;
    std::vector<ShardWs *> free_ws;
    std::map<void *, ShardWs *> dev_ws;
    std::mutex mask_mu;
    std::map<int, std::vector<int>> masks;
    int next_mask = 1;
    unsigned long long wait_timeout_ns = 20ull * 1000 * 1000 * 1000;
    // The merge kernel may wait in-kernel for the other shards' arrival counters only when every one of them runs on
    // ANOTHER GPU: kernels that wait on one another as separate launches on ONE GPU are not guaranteed to run at the same
    // time (B200_PROFILING.md).  With a device listed twice (tests on single-GPU boxes) the root stream waits for the
    // shards' events instead and the merge launches afterwards.
    bool spin_merge = true;

    uint32_t G() const { return (uint32_t)shards.size(); }
    uint32_t owner(uint64_t id) const {
        for (const auto &r : ranges)
            if (id >= r.id0 && id - r.id0 < r.n) {
                const uint64_t o = id - r.id0;
                uint32_t g = 0;
                while (g + 1 < G() && o >= r.cut[g + 1]) ++g;
                return g;
            }
        return (uint32_t)(id % G());
    }
    int shard_masks(int mask_id, std::vector<int> *out) {
        out->assign(G(), -1);
        if (mask_id < 0) return SZG_OK;
        std::lock_guard<std::mutex> lk(mask_mu);
        auto it = masks.find(mask_id);
        if (it == masks.end()) return fail(SZG_ENOTFOUND, "unknown mask id %d", mask_id);
        *out = it->second;
        return SZG_OK;
    }
};

namespace {

// runs fn(g) for every shard, shards 1 .. G-1 on their own host threads (each call synchronises its own device)
template <typename F>
int for_shards(Sharded *S, F fn) {
    const uint32_t G = S->G();
    std::vector<int> rc(G, SZG_OK);
    std::vector<std::string> err(G);
    std::vector<std::thread> th;
    try {
        for (uint32_t g = 1; g < G; ++g)
            th.emplace_back([&, g]() {
                rc[g] = fn(g);
                if (rc[g]) err[g] = last_error_string();
            });
    } catch (...) {
        for (auto &t : th) t.join();
        return fail(SZG_ENOMEM, "cannot start a host thread per shard");
    }
    rc[0] = fn(0);
    if (rc[0]) err[0] = last_error_string();
    for (auto &t : th) t.join();
    for (uint32_t g = 0; g < G; ++g)
        if (rc[g]) { set_last_error_string(err[g]); return rc[g]; }
    return SZG_OK;
}

void free_shard_ws(Sharded *S, ShardWs *W) {
    for (uint32_t g = 0; g < W->ws.size(); ++g) {
        if (!W->ws[g] || (g == 0 && W->bound)) continue; // a bound root workspace belongs to the root's dev_ws table
        release_ws(S->shards[g], W->ws[g]);
    }
    for (uint32_t g = 0; g < W->ev_done.size(); ++g)
        if (W->ev_done[g]) { DeviceGuard gd(S->devices[g]); cudaEventDestroy(W->ev_done[g]); }
    for (auto &ge : W->graphs)
        if (ge.second.exec) cudaGraphExecDestroy(ge.second.exec);
    {
        DeviceGuard gd(S->devices[0]);
        W->d_gather.release(); W->d_done.release(); W->d_out.release(); W->h_out.release(); W->d_q2.release();
        if (W->ev_start) cudaEventDestroy(W->ev_start);
    }
    delete W;
}

int new_shard_ws(Sharded *S, void *bound_stream, ShardWs **out) {
    std::unique_ptr<ShardWs> W(new ShardWs());
    W->ws.assign(S->G(), nullptr);
    int rc = SZG_OK;
    for (uint32_t g = 0; g < S->G() && !rc; ++g) {
        DeviceGuard gd(S->devices[g]);
        if (!gd.ok) { rc = fail(SZG_ECUDA, "cannot select CUDA device %d", S->devices[g]); break; }
        if (g == 0 && bound_stream != (void *)-1) {
            rc = ws_for_stream(S->shards[0], bound_stream, &W->ws[0]);
            W->bound = true;
        } else rc = acquire_ws(S->shards[g], &W->ws[g]);
        if (!rc) {
            W->ev_done.resize(S->G(), nullptr);
            cudaError_t e = cudaEventCreateWithFlags(&W->ev_done[g], cudaEventDisableTiming);
            if (e != cudaSuccess) rc = fail(SZG_ECUDA, "event creation failed: %s", cudaGetErrorString(e));
        }
    }
    if (!rc) {
        DeviceGuard gd(S->devices[0]);
        cudaError_t e = cudaEventCreateWithFlags(&W->ev_start, cudaEventDisableTiming);
        if (e != cudaSuccess) rc = fail(SZG_ECUDA, "event creation failed: %s", cudaGetErrorString(e));
    }
    if (rc) { free_shard_ws(S, W.release()); return rc; }
    *out = W.release();
    return SZG_OK;
}

int acquire_shard_ws(Sharded *S, ShardWs **out) {
    {
        std::lock_guard<std::mutex> lk(S->mu);
        if (!S->free_ws.empty()) { *out = S->free_ws.back(); S->free_ws.pop_back(); return SZG_OK; }
    }
    return new_shard_ws(S, (void *)-1, out);
}
void release_shard_ws(Sharded *S, ShardWs *W) {
    std::lock_guard<std::mutex> lk(S->mu);
    S->free_ws.push_back(W);
}

// One pass over all shards for nq queries that sit in the ROOT device's memory (every shard reads them through its peer
// mapping): start event -> per shard prep + scan | batch + finalize (outputs and arrival counters land on the root) -> merge
// on the root stream.  Outputs (root memory): d_out_ids/dist [nq*k], d_out_n/flags [nq]; *err_word gets bit0 on a timed-out wait.
int sharded_pass(szg_index *h, ShardWs *W, const double *d_q_root, uint32_t nq, uint32_t k, const std::vector<int> &mask_ids,
                 uint32_t flags, bool prefer_batch, int min_mode, int force_nd, unsigned long long *d_out_ids, double *d_out_dist,
                 uint32_t *d_out_n, uint32_t *d_out_flags, int *mode_out, int *nd_out) {
    Sharded *S = h->sh;
    const uint32_t G = S->G();
    int rc;
    const size_t rec = record_bytes(nq, k);
    cudaStream_t st0 = W->ws[0]->main;
    {
        DeviceGuard gd(S->devices[0]);
        if (!gd.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", S->devices[0]);
        if ((rc = W->d_gather.ensure(rec * G)) || (rc = W->d_done.ensure((size_t)nq + 1))) return rc;
        CK(cudaMemsetAsync(W->d_done.p, 0, ((size_t)nq + 1) * 4, st0));
        CK(cudaEventRecord(W->ev_start, st0));
    }
    // in-kernel wait of the merge for the other devices' counters -- or, when shards share a GPU or the pass is being
    // captured into a graph (whose completion joins every stream anyway), plain event dependencies
    const bool spin = S->spin_merge && !W->capturing;
    PeerSink sink;
    sink.done_cnt = spin ? W->d_done.p : nullptr;
    for (uint32_t g = 0; g < G; ++g) {
        szg_index *sh = S->shards[g];
        DeviceGuard gd(sh->device);
        if (!gd.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", sh->device);
        Workspace *ws = W->ws[g];
        if (g) CK(cudaStreamWaitEvent(ws->main, W->ev_start, 0));
        const uint32_t *mask;
        if ((rc = get_mask(sh, mask_ids[g], &mask))) return rc;
        unsigned char *r = W->d_gather.p + rec * g;
        int mode = 0, nd = 0;
        if ((rc = enqueue_topk(sh, ws, d_q_root, nq, k, mask, flags, prefer_batch, min_mode, force_nd,
                               reinterpret_cast<unsigned long long *>(r), reinterpret_cast<double *>(r + (size_t)nq * k * 8),
                               reinterpret_cast<uint32_t *>(r + (size_t)nq * k * 16),
                               reinterpret_cast<uint32_t *>(r + (size_t)nq * k * 16 + (size_t)nq * 4), &sink, &mode, &nd)))
            return rc;
        if (g == 0) { if (mode_out) *mode_out = mode; if (nd_out) *nd_out = nd; }
        if (!spin && g) CK(cudaEventRecord(W->ev_done[g], ws->main));
    }
    DeviceGuard gd(S->devices[0]);
    if (!spin)
        for (uint32_t g = 1; g < G; ++g) CK(cudaStreamWaitEvent(st0, W->ev_done[g], 0));
    MergeArgs a;
    memset(&a, 0, sizeof a);
    unsigned char *r0 = W->d_gather.p;
    a.g_ids = reinterpret_cast<const unsigned long long *>(r0);
    a.g_dist = reinterpret_cast<const double *>(r0 + (size_t)nq * k * 8);
    a.g_n = reinterpret_cast<const uint32_t *>(r0 + (size_t)nq * k * 16);
    a.g_flags = reinterpret_cast<const uint32_t *>(r0 + (size_t)nq * k * 16 + (size_t)nq * 4);
    a.rank_stride = rec;
    a.G = G; a.nq = nq; a.k = k;
    a.out_ids = d_out_ids; a.out_dist = d_out_dist; a.out_n = d_out_n; a.out_flags = d_out_flags;
    if (spin) { a.wait_cnt = W->d_done.p; a.wait_target = G; a.wait_timeout_ns = S->wait_timeout_ns; a.err = W->d_done.p + nq; }
    CK(launch_merge(a, st0));
    h->launches++;
    return SZG_OK;
}

} // namespace

// --------------------------------------------------------------------------------------------------- lifecycle
int sharded_create(int dim, int quantization, int metric, const int *devices, int ndev, szg_index **out) {
    if (!out) return fail(SZG_EINVAL, "null out pointer");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > 64) return fail(SZG_EINVAL, "need 1..64 devices");
    if ((size_t)ndev * SZG_MAX_K * 16 > 200 * 1024) return fail(SZG_EINVAL, "too many devices");
    std::unique_ptr<szg_index> h(new szg_index());
    std::unique_ptr<Sharded> S(new Sharded());
    auto dict = std::make_shared<MetaDict>();
    int rc = SZG_OK;
    for (int g = 0; g < ndev && !rc; ++g) {
        szg_index *sh = nullptr;
        rc = create_single(dim, quantization, metric, devices[g], dict, &sh);
        if (!rc) {
            sh->parent = h.get();
            sh->combine = 0;
            S->shards.push_back(sh);
            S->devices.push_back(devices[g]);
        }
    }
    // every device writes its lists into, and reads the queries from, devices[0]'s memory
    for (int g = 1; g < ndev && !rc; ++g) {
        if (devices[g] == devices[0]) continue;
        int can = 0;
        cudaError_t e = cudaDeviceCanAccessPeer(&can, devices[g], devices[0]);
        if (e != cudaSuccess || !can) {
            rc = fail(SZG_ECUDA, "device %d cannot map the memory of device %d (peer access is required for a sharded mirror)",
                      devices[g], devices[0]);
            break;
        }
        DeviceGuard gd(devices[g]);
        e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) rc = fail(SZG_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", devices[g], devices[0], cudaGetErrorString(e));
    }
    if (rc) {
        for (auto sh : S->shards) destroy_single(sh);
        return rc;
    }
    for (int g = 0; g < ndev; ++g)
        for (int j = g + 1; j < ndev; ++j)
            if (devices[g] == devices[j]) S->spin_merge = false;
    if (const char *e = getenv("SZG_SHARD_SPIN")) S->spin_merge = S->spin_merge && atoi(e) != 0; // 0: event joins (A/B measurements)
    szg_index *s0 = S->shards[0];
    h->dim = s0->dim; h->quant = s0->quant; h->metric = s0->metric; h->device = s0->device; h->qt = s0->qt;
    h->rowbytes = s0->rowbytes; h->C = s0->C; h->maxint = s0->maxint; h->sm_count = s0->sm_count;
    h->dict = dict;
    h->sh = S.release();
    *out = h.release();
    return SZG_OK;
}

int sharded_destroy(szg_index *h) {
    Sharded *S = h->sh;
    for (uint32_t g = 0; g < S->G(); ++g) {
        DeviceGuard gd(S->devices[g]);
        cudaDeviceSynchronize();
    }
    for (auto W : S->free_ws) free_shard_ws(S, W);
    for (auto &kv : S->dev_ws) free_shard_ws(S, kv.second);
    for (auto sh : S->shards) destroy_single(sh);
    delete S;
    delete h;
    return SZG_OK;
}

szg_index *sharded_root(szg_index *h) { return h->sh->shards[0]; }
szg_index *sharded_timing_shard(szg_index *h) { return h->sh->shards[0]; }

uint64_t sharded_count(szg_index *h) {
    uint64_t n = 0;
    for (auto sh : h->sh->shards) n += sh->live_rows;
    return n;
}

// --------------------------------------------------------------------------------------------------- mutations
int sharded_reserve(szg_index *h, uint64_t nrows) {
    Sharded *S = h->sh;
    const uint64_t per = (nrows + S->G() - 1) / S->G();
    for (auto sh : S->shards) {
        DeviceGuard gd(sh->device);
        int rc = grow(sh, per + per / 16 + 64); // id % G is even only on average
        if (rc) return rc;
    }
    return SZG_OK;
}

int sharded_upsert(szg_index *h, const uint64_t *ids, const uint8_t *codes, const double *vectors, uint8_t *out_codes, uint64_t n,
                   bool into_mirror) {
    Sharded *S = h->sh;
    const uint32_t G = S->G();
    const size_t rb = h->rowbytes, d = (size_t)h->dim;
    std::vector<std::vector<uint64_t>> where(G); // positions of each shard's records in the caller's arrays
    for (uint64_t i = 0; i < n; ++i) where[ids ? S->owner(ids[i]) : (uint32_t)(i % G)].push_back(i);
    return for_shards(S, [&](uint32_t g) -> int {
        const auto &w = where[g];
        if (w.empty()) return SZG_OK;
        szg_index *sh = S->shards[g];
        DeviceGuard gd(sh->device);
        if (!gd.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", sh->device);
        std::vector<uint64_t> sid(w.size());
        for (size_t j = 0; j < w.size(); ++j) sid[j] = ids ? ids[w[j]] : 0;
        std::vector<uint8_t> sc, so;
        std::vector<double> sv;
        if (codes) {
            sc.resize(w.size() * rb);
            for (size_t j = 0; j < w.size(); ++j) memcpy(sc.data() + j * rb, codes + w[j] * rb, rb);
        }
        if (vectors) {
            sv.resize(w.size() * d);
            for (size_t j = 0; j < w.size(); ++j) memcpy(sv.data() + j * d, vectors + w[j] * d, d * sizeof(double));
        }
        if (out_codes) so.resize(w.size() * rb);
        int rc = upsert_rows(sh, sid.data(), codes ? sc.data() : nullptr, vectors ? sv.data() : nullptr, out_codes ? so.data() : nullptr,
                             w.size(), into_mirror);
        if (rc) return rc;
        if (out_codes)
            for (size_t j = 0; j < w.size(); ++j) memcpy(out_codes + w[j] * rb, so.data() + j * rb, rb);
        return SZG_OK;
    });
}

int sharded_remove(szg_index *h, const uint64_t *ids, uint64_t n, uint64_t *n_removed) {
    Sharded *S = h->sh;
    std::vector<std::vector<uint64_t>> per(S->G());
    for (uint64_t i = 0; i < n; ++i) per[S->owner(ids[i])].push_back(ids[i]);
    uint64_t total = 0;
    for (uint32_t g = 0; g < S->G(); ++g) {
        if (per[g].empty()) continue;
        uint64_t r = 0;
        int rc = szg_remove(S->shards[g], per[g].data(), per[g].size(), &r);
        if (rc) return rc;
        total += r;
    }
    if (n_removed) *n_removed = total;
    return SZG_OK;
}

int sharded_fill_synthetic(szg_index *h, uint64_t seed, uint64_t row0, uint64_t nrows) {
    Sharded *S = h->sh;
    const uint32_t G = S->G();
    for (const auto &r : S->ranges)
        if (row0 < r.id0 + r.n && r.id0 < row0 + nrows) return fail(SZG_EINVAL, "synthetic range overlaps an existing one");
    SynthRange r;
    r.id0 = row0; r.n = nrows;
    r.cut.resize(G + 1);
    const uint64_t base = nrows / G, extra = nrows % G;
    r.cut[0] = 0;
    for (uint32_t g = 0; g < G; ++g) r.cut[g + 1] = r.cut[g] + base + (g < extra ? 1 : 0);
    int rc = for_shards(S, [&](uint32_t g) -> int {
        return szg_fill_synthetic(S->shards[g], seed, row0 + r.cut[g], r.cut[g + 1] - r.cut[g]);
    });
    if (rc) return rc;
    S->ranges.push_back(r);
    return SZG_OK;
}

int sharded_fetch_codes(szg_index *h, const uint64_t *ids, uint64_t n, uint8_t *out_codes) {
    Sharded *S = h->sh;
    const size_t rb = h->rowbytes;
    std::vector<std::vector<uint64_t>> where(S->G());
    for (uint64_t i = 0; i < n; ++i) where[S->owner(ids[i])].push_back(i);
    for (uint32_t g = 0; g < S->G(); ++g) {
        const auto &w = where[g];
        if (w.empty()) continue;
        std::vector<uint64_t> sid(w.size());
        for (size_t j = 0; j < w.size(); ++j) sid[j] = ids[w[j]];
        std::vector<uint8_t> buf(w.size() * rb);
        int rc = szg_fetch_codes(S->shards[g], sid.data(), sid.size(), buf.data());
        if (rc) return rc;
        for (size_t j = 0; j < w.size(); ++j) memcpy(out_codes + w[j] * rb, buf.data() + j * rb, rb);
    }
    return SZG_OK;
}

int sharded_mask_create(szg_index *h, const uint64_t *ids, const uint8_t *pass, uint64_t n, int *mask_id) {
    Sharded *S = h->sh;
    std::vector<std::vector<uint64_t>> sid(S->G());
    std::vector<std::vector<uint8_t>> sp(S->G());
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t g = S->owner(ids[i]);
        sid[g].push_back(ids[i]);
        sp[g].push_back(pass[i]);
    }
    std::vector<int> mids(S->G(), -1);
    for (uint32_t g = 0; g < S->G(); ++g) {
        int rc = szg_mask_create(S->shards[g], sid[g].data(), sp[g].data(), sid[g].size(), &mids[g]);
        if (rc) {
            for (uint32_t j = 0; j < g; ++j) szg_mask_destroy(S->shards[j], mids[j]);
            return rc;
        }
    }
    std::lock_guard<std::mutex> lk(S->mask_mu);
    *mask_id = S->next_mask++;
    S->masks[*mask_id] = mids;
    return SZG_OK;
}

int sharded_filter_mask(szg_index *h, const szg_filter_op *ops, uint32_t nops, int *mask_id) {
    Sharded *S = h->sh;
    std::vector<int> mids(S->G(), -1);
    for (uint32_t g = 0; g < S->G(); ++g) { // one dictionary for all shards: the program means the same everywhere
        int rc = szg_filter_mask(S->shards[g], ops, nops, &mids[g]);
        if (rc) {
            for (uint32_t j = 0; j < g; ++j) szg_mask_destroy(S->shards[j], mids[j]);
            return rc;
        }
    }
    std::lock_guard<std::mutex> lk(S->mask_mu);
    *mask_id = S->next_mask++;
    S->masks[*mask_id] = mids;
    return SZG_OK;
}

int sharded_mask_destroy(szg_index *h, int mask_id) {
    Sharded *S = h->sh;
    std::vector<int> mids;
    {
        std::lock_guard<std::mutex> lk(S->mask_mu);
        auto it = S->masks.find(mask_id);
        if (it == S->masks.end()) return fail(SZG_ENOTFOUND, "unknown mask id %d", mask_id);
        mids = it->second;
        S->masks.erase(it);
    }
    for (uint32_t g = 0; g < S->G(); ++g) szg_mask_destroy(S->shards[g], mids[g]);
    return SZG_OK;
}

int sharded_meta_upsert(szg_index *h, const uint64_t *ids, uint64_t n, const uint8_t *doc_kind, const uint32_t *cols, uint32_t ncols,
                        const szg_meta_value *values) {
    Sharded *S = h->sh;
    std::vector<std::vector<uint64_t>> where(S->G());
    for (uint64_t i = 0; i < n; ++i) where[S->owner(ids[i])].push_back(i);
    for (uint32_t g = 0; g < S->G(); ++g) {
        const auto &w = where[g];
        if (w.empty()) continue;
        std::vector<uint64_t> sid(w.size());
        std::vector<uint8_t> sk(w.size());
        std::vector<szg_meta_value> sv(w.size() * (size_t)ncols);
        for (size_t j = 0; j < w.size(); ++j) {
            sid[j] = ids[w[j]];
            sk[j] = doc_kind[w[j]];
            for (uint32_t c = 0; c < ncols; ++c) sv[j * ncols + c] = values[w[j] * ncols + c];
        }
        int rc = szg_meta_upsert(S->shards[g], sid.data(), sid.size(), sk.data(), cols, ncols, sv.data());
        if (rc) return rc;
    }
    return SZG_OK;
}

int sharded_set_option(szg_index *h, int option, int64_t value) {
    for (auto sh : h->sh->shards) {
        int rc = szg_set_option(sh, option, value);
        if (rc) return rc;
    }
    return SZG_OK;
}

int sharded_get_stats(szg_index *h, szg_stats *out) {
    Sharded *S = h->sh;
    szg_stats acc;
    memset(&acc, 0, sizeof acc);
    for (uint32_t g = 0; g < S->G(); ++g) {
        szg_stats s;
        int rc = szg_get_stats(S->shards[g], &s);
        if (rc) return rc;
        if (g == 0) acc = s;
        else {
            acc.kernel_launches += s.kernel_launches; acc.escalations += s.escalations; acc.uncertain_results += s.uncertain_results;
            acc.batch_queries += s.batch_queries; acc.device_bytes += s.device_bytes; acc.live_rows += s.live_rows; acc.slots += s.slots;
            acc.combined_queries += s.combined_queries; acc.graph_launches += s.graph_launches;
        }
    }
    acc.kernel_launches += h->launches.load();
    acc.graph_launches += h->graph_launches.load();
    acc.escalations += h->escalations.load();
    acc.uncertain_results += h->uncertain.load();
    acc.batch_queries /= S->G(); // every shard serves every query of a batch
    acc.shards = S->G();
    *out = acc;
    return SZG_OK;
}

// --------------------------------------------------------------------------------------------------- searches
int sharded_search_topk(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags, uint64_t *out_ids,
                        double *out_dist, uint32_t *out_n, uint64_t *scanned, bool prefer_batch) {
    Sharded *S = h->sh;
    int rc;
    if (!queries && nq) return fail(SZG_EINVAL, "null query");
    if (k < 1 || k > SZG_MAX_K) return fail(SZG_EINVAL, "k=%u outside [1, %u]", k, SZG_MAX_K);
    if (nq && (!out_ids || !out_dist || !out_n)) return fail(SZG_EINVAL, "null output");
    const uint64_t live = sharded_count(h);
    if (scanned) *scanned = live;
    if (!nq) return SZG_OK;
    std::vector<int> mids;
    if ((rc = S->shard_masks(mask_id, &mids))) return rc;
    if (live == 0) {
        for (uint32_t i = 0; i < nq; ++i) out_n[i] = 0;
        return SZG_OK;
    }
    ShardWs *W;
    if ((rc = acquire_shard_ws(S, &W))) return rc;
    struct Rel { Sharded *S; ShardWs *W; ~Rel() { release_shard_ws(S, W); } } rel{S, W};
    szg_index *root = S->shards[0];
    DeviceGuard gd(root->device);
    if (!gd.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", root->device);
    Workspace *w0 = W->ws[0];
    cudaStream_t st0 = w0->main;
    const size_t qn = (size_t)nq * h->dim, on = (size_t)nq * k;
    const size_t pack = on * 16 + (size_t)nq * 8 + 8; // [ids | dist | n | flags | err word]
    if ((rc = w0->h_q.ensure(qn)) || (rc = w0->d_q.ensure(qn)) || (rc = W->d_out.ensure(pack)) || (rc = W->h_out.ensure(pack))) return rc;
    memcpy(w0->h_q.p, queries, qn * sizeof(double));
    // enqueue: one pass over all shards + the copy of the merged results (and the error word) to the host
    auto enqueue = [&](const double *dq, uint32_t m, int min_mode, int force_nd, bool batch, int *mode, int *nd) -> int {
        unsigned char *o = W->d_out.p;
        int r = sharded_pass(h, W, dq, m, k, mids, flags, batch, min_mode, force_nd, reinterpret_cast<unsigned long long *>(o),
                             reinterpret_cast<double *>(o + (size_t)m * k * 8), reinterpret_cast<uint32_t *>(o + (size_t)m * k * 16),
                             reinterpret_cast<uint32_t *>(o + (size_t)m * k * 16 + (size_t)m * 4), mode, nd);
        if (r) return r;
        const size_t bytes = (size_t)m * k * 16 + (size_t)m * 8;
        CK(cudaMemcpyAsync(W->h_out.p, W->d_out.p, bytes, cudaMemcpyDeviceToHost, st0));
        CK(cudaMemcpyAsync(W->h_out.p + bytes, W->d_done.p + m, 4, cudaMemcpyDeviceToHost, st0));
        return SZG_OK;
    };
    auto finish = [&](uint32_t m) -> int {
        CK(cudaStreamSynchronize(st0));
        const size_t bytes = (size_t)m * k * 16 + (size_t)m * 8;
        uint32_t err;
        memcpy(&err, W->h_out.p + bytes, 4);
        if (err) return fail(SZG_EINTERNAL, "a shard did not deliver its results within %.0f s", (double)S->wait_timeout_ns / 1e9);
        return SZG_OK;
    };
    auto run = [&](const double *dq, uint32_t m, int min_mode, int force_nd, bool batch, int *mode, int *nd) -> int {
        int r = enqueue(dq, m, min_mode, force_nd, batch, mode, nd);
        return r ? r : finish(m);
    };
    int mode0 = 0, nd0 = 0;
    // Repeated call shapes are replayed as ONE captured launch sequence over all devices (the copy of the queries, every
    // shard's prep + scan | batch + finalize, the merge, the copy back): a call then costs one graph launch instead of a
    // few dozen stream operations issued device by device.  Same scheme as search_host: first call of a shape launch by
    // launch (sizes the buffers), second one captures.
    bool done = false;
    const bool graphs_ok = root->use_graphs && root->timing == 0;
    if (graphs_ok) {
        const uint64_t key = ((uint64_t)nq << 40) ^ ((uint64_t)k << 28) ^ ((uint64_t)(uint32_t)(mask_id + 1) << 4) ^ ((uint64_t)(flags & 3u) << 1) ^
                             (prefer_batch ? 1u : 0u);
        GraphEntry &ge = W->graphs[key];
        uint64_t gen = 0;
        for (auto sh : S->shards) gen += sh->generation;
        auto fingerprint = [&]() -> const void * {
            uint64_t f = 1469598103934665603ULL;
            auto mix = [&](const void *p) { f = (f ^ (uint64_t)(uintptr_t)p) * 1099511628211ULL; };
            mix(w0->d_q.p); mix(w0->h_q.p); mix(W->d_gather.p); mix(W->d_done.p); mix(W->d_out.p); mix(W->h_out.p);
            for (auto ws : W->ws) { mix(ws->d_pq.p); mix(ws->d_cand.p); mix(ws->d_gmth.p); }
            return reinterpret_cast<const void *>((uintptr_t)f);
        };
        if (ge.exec && (ge.generation != gen || ge.buffers[0] != fingerprint())) {
            cudaGraphExecDestroy(ge.exec);
            ge.exec = nullptr;
            if (ge.generation == gen) ge.generation = 0;
        }
        if (ge.exec && ge.generation == gen) {
            mode0 = ge.mode0; nd0 = ge.nd0;
            CK(cudaGraphLaunch(ge.exec, st0));
            h->graph_launches++;
            h->launches += ge.kernels;
            done = true;
        } else if (!ge.exec && ge.generation == gen && !ge.failed) {
            uint64_t l0 = h->launches.load();
            for (auto sh : S->shards) l0 += sh->launches.load();
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamBeginCapture(st0, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                W->capturing = true;
                for (auto ws : W->ws) ws->capturing = true;
                int r = SZG_OK;
                cudaError_t ce = cudaMemcpyAsync(w0->d_q.p, w0->h_q.p, qn * sizeof(double), cudaMemcpyHostToDevice, st0);
                if (ce != cudaSuccess) r = SZG_ECUDA;
                if (!r) r = enqueue(w0->d_q.p, nq, 0, 0, prefer_batch, &mode0, &nd0);
                W->capturing = false;
                for (auto ws : W->ws) ws->capturing = false;
                e = cudaStreamEndCapture(st0, &g);
                if (r == SZG_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&ge.exec, g, 0);
                else if (e == cudaSuccess) e = cudaErrorUnknown;
                if (g) cudaGraphDestroy(g);
            }
            if (e != cudaSuccess || !ge.exec) {
                cudaGetLastError();
                ge.exec = nullptr;
                ge.failed = true;
            } else {
                uint64_t l1 = h->launches.load();
                for (auto sh : S->shards) l1 += sh->launches.load();
                ge.mode0 = mode0; ge.nd0 = nd0;
                ge.kernels = (uint32_t)(l1 - l0);
                ge.buffers[0] = fingerprint();
                CK(cudaGraphLaunch(ge.exec, st0));
                h->graph_launches++;
                done = true;
            }
        } else if (ge.generation != gen) {
            ge.generation = gen;
            ge.failed = false;
        }
    }
    if (done) {
        if ((rc = finish(nq))) return rc;
    } else {
        CK(cudaMemcpyAsync(w0->d_q.p, w0->h_q.p, qn * sizeof(double), cudaMemcpyHostToDevice, st0));
        if ((rc = run(w0->d_q.p, nq, 0, 0, prefer_batch, &mode0, &nd0))) return rc;
    }
    for (uint32_t g = 0; g < S->G(); ++g) drain_timing(S->shards[g], W->ws[g]);
    memcpy(out_ids, W->h_out.p, on * 8);
    memcpy(out_dist, W->h_out.p + on * 8, on * 8);
    memcpy(out_n, W->h_out.p + on * 16, (size_t)nq * 4);
    if (flags & SZG_F_NO_FP64_VERIFY) return SZG_OK;
    // queries some shard could not certify: re-run on every shard, first with the precise surrogate, then with larger
    // candidate sets -- the ladder of the single-device path (collect_and_escalate)
    std::vector<uint32_t> pending;
    {
        const uint32_t *fl = reinterpret_cast<const uint32_t *>(W->h_out.p + on * 16 + (size_t)nq * 4);
        for (uint32_t i = 0; i < nq; ++i)
            if (fl[i] & 1u) pending.push_back(i);
    }
    int nd = nd0, mode = mode0;
    while (!pending.empty()) {
        if (nd == 2) nd = 3;
        else if (mode < 3) ++mode;
        else break;
        const uint32_t m = (uint32_t)pending.size();
        h->escalations += m;
        if ((rc = W->d_q2.ensure((size_t)m * h->dim))) return rc;
        for (uint32_t j = 0; j < m; ++j)
            CK(cudaMemcpyAsync(W->d_q2.p + (size_t)j * h->dim, w0->d_q.p + (size_t)pending[j] * h->dim, (size_t)h->dim * sizeof(double),
                               cudaMemcpyDeviceToDevice, st0));
        if ((rc = run(W->d_q2.p, m, mode, nd, false, nullptr, nullptr))) return rc;
        const size_t om = (size_t)m * k;
        const uint64_t *ri = reinterpret_cast<const uint64_t *>(W->h_out.p);
        const double *rd = reinterpret_cast<const double *>(W->h_out.p + om * 8);
        const uint32_t *rn = reinterpret_cast<const uint32_t *>(W->h_out.p + om * 16);
        const uint32_t *rf = rn + m;
        std::vector<uint32_t> still;
        for (uint32_t j = 0; j < m; ++j) {
            const uint32_t i = pending[j];
            memcpy(out_ids + (size_t)i * k, ri + (size_t)j * k, (size_t)k * 8);
            memcpy(out_dist + (size_t)i * k, rd + (size_t)j * k, (size_t)k * 8);
            out_n[i] = rn[j];
            if (rf[j] & 1u) still.push_back(i);
        }
        pending.swap(still);
    }
    h->uncertain += pending.size();
    return SZG_OK;
}

int sharded_search_topk_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                            uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, void *stream,
                            bool prefer_batch) {
    Sharded *S = h->sh;
    int rc;
    if (!d_queries && nq) return fail(SZG_EINVAL, "null query");
    if (k < 1 || k > SZG_MAX_K) return fail(SZG_EINVAL, "k=%u outside [1, %u]", k, SZG_MAX_K);
    if (!nq) return SZG_OK;
    if (!d_out_ids || !d_out_dist || !d_out_n) return fail(SZG_EINVAL, "null output");
    std::vector<int> mids;
    if ((rc = S->shard_masks(mask_id, &mids))) return rc;
    ShardWs *W = nullptr;
    {
        std::lock_guard<std::mutex> lk(S->mu);
        auto it = S->dev_ws.find(stream);
        if (it != S->dev_ws.end()) W = it->second;
    }
    if (!W) {
        if ((rc = new_shard_ws(S, stream, &W))) return rc;
        std::lock_guard<std::mutex> lk(S->mu);
        S->dev_ws[stream] = W;
    }
    DeviceGuard gd(S->devices[0]);
    if (!gd.ok) return fail(SZG_ECUDA, "cannot select CUDA device %d", S->devices[0]);
    if (!d_out_flags) {
        if ((rc = W->ws[0]->d_out_flags.ensure(nq))) return rc;
        d_out_flags = W->ws[0]->d_out_flags.p;
    }
    // queries and outputs live on devices[0]; nothing is synchronised with the host (no escalation: see szg_search_topk_dev)
    return sharded_pass(h, W, d_queries, nq, k, mids, flags, prefer_batch, 0, 0, (unsigned long long *)d_out_ids, d_out_dist, d_out_n,
                        d_out_flags, nullptr, nullptr);
}

// Radius search (collection.go:598-605): every shard returns its hits ordered on the device (radius_device); the variable-
// length lists are merged here by (distance, lexicographic id).  SURVEY.md 8e: "allgather counts, then variable-length
// gather" -- inside one process that is one copy per shard into this result object.
int sharded_search_radius(szg_index *h, const double *queries, uint32_t nq, const double *radii, int mask_id, szg_result **out,
                          uint64_t *scanned) {
    Sharded *S = h->sh;
    const uint32_t G = S->G();
    int rc;
    if (scanned) *scanned = sharded_count(h);
    std::vector<int> mids;
    if ((rc = S->shard_masks(mask_id, &mids))) return rc;
    std::vector<std::vector<szg_result *>> part(G, std::vector<szg_result *>(nq, nullptr));
    rc = for_shards(S, [&](uint32_t g) -> int { return radius_device(S->shards[g], queries, nq, radii, mids[g], part[g].data()); });
    auto cleanup = [&]() {
        for (auto &v : part)
            for (auto r : v) delete r;
    };
    if (rc) { cleanup(); return rc; }
    try {
        for (uint32_t q = 0; q < nq; ++q) {
            std::unique_ptr<szg_result> res(new szg_result());
            size_t total = 0;
            for (uint32_t g = 0; g < G; ++g) total += part[g][q]->ids.size();
            res->ids.reserve(total);
            res->dist.reserve(total);
            std::vector<size_t> pos(G, 0);
            for (size_t i = 0; i < total; ++i) { // G-way merge of ascending lists
                int best = -1;
                for (uint32_t g = 0; g < G; ++g) {
                    const szg_result *p = part[g][q];
                    if (pos[g] >= p->ids.size()) continue;
                    if (best < 0) { best = (int)g; continue; }
                    const szg_result *b = part[best][q];
                    const double dg = p->dist[pos[g]], db = b->dist[pos[best]];
                    if (dg < db || (dg == db && lex_less_u64(p->ids[pos[g]], b->ids[pos[best]]))) best = (int)g;
                }
                res->ids.push_back(part[best][q]->ids[pos[best]]);
                res->dist.push_back(part[best][q]->dist[pos[best]]);
                ++pos[best];
            }
            out[q] = res.release();
        }
    } catch (...) {
        for (uint32_t q = 0; q < nq; ++q) { delete out[q]; out[q] = nullptr; }
        cleanup();
        return fail(SZG_ENOMEM, "out of host memory");
    }
    cleanup();
    return SZG_OK;
}

// candidate re-scoring: every id goes to the device that owns its row; distances return in visit order
int sharded_rescore(szg_index *h, const double *queries, uint32_t nlists, const uint64_t *ids, const uint64_t *list_offsets,
                    double *out_dist) {
    Sharded *S = h->sh;
    const uint32_t G = S->G();
    const uint64_t m = list_offsets[nlists];
    std::vector<std::vector<uint64_t>> sid(G), where(G), off(G, std::vector<uint64_t>(nlists + 1, 0));
    for (uint32_t l = 0; l < nlists; ++l) {
        for (uint64_t i = list_offsets[l]; i < list_offsets[l + 1]; ++i) {
            const uint32_t g = S->owner(ids[i]);
            sid[g].push_back(ids[i]);
            where[g].push_back(i);
        }
        for (uint32_t g = 0; g < G; ++g) off[g][l + 1] = sid[g].size();
    }
    (void)m;
    return for_shards(S, [&](uint32_t g) -> int {
        if (sid[g].empty()) return SZG_OK;
        std::vector<double> d(sid[g].size());
        int rc = rescore_device(S->shards[g], queries, nlists, sid[g].data(), off[g].data(), d.data());
        if (rc) return rc;
        for (size_t j = 0; j < d.size(); ++j) out_dist[where[g][j]] = d[j];
        return SZG_OK;
    });
}

} // namespace szg

extern "C" int szg_create_sharded(int dim, int quantization, int metric, const int *devices, int ndev, szg_index **out) {
    return szg::sharded_create(dim, quantization, metric, devices, ndev, out);
}
