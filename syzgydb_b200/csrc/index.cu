// index.cu -- the GPU mirror of one collection (or one row shard) and the C ABI of
// include/syzgy_b200.h.  Host logic only: slot allocation, id -> slot map, workspaces,
// streams, launches.  There is deliberately no CPU implementation of any search step.
#include <cmath>

#include "index_internal.h"
#include "sharded.h"

using namespace szg;

namespace szg {

static thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
const std::string &last_error_string() { return g_err; }
void set_last_error_string(const std::string &s) { g_err = s; }

// error channel for the other translation units of the library (spanfile.cu)
int set_error(int code, const char *msg) { return fail(code, "%s", msg); }

} // namespace szg

namespace szg {

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

int get_mask(szg_index *h, int mask_id, const uint32_t **out) {
    *out = nullptr;
    if (mask_id < 0) return SZG_OK;
    std::lock_guard<std::mutex> lk(h->mask_mu);
    auto it = h->masks.find(mask_id);
    if (it == h->masks.end()) return fail(SZG_ENOTFOUND, "unknown mask id %d", mask_id);
    *out = it->second;
    return SZG_OK;
}

// geometry of a handle, for spanfile.cu
void index_geometry(const szg_index *h, int *dim, int *quant, int *metric, uint32_t *rowbytes) {
    *dim = h->dim; *quant = h->quant; *metric = h->metric; *rowbytes = h->rowbytes;
}

int create_single(int dim, int quantization, int metric, int device, std::shared_ptr<MetaDict> dict, szg_index **out) {
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
    if (grouped_layout(qt)) h->C = (h->C + kGroupChunks - 1) / kGroupChunks * kGroupChunks; // float rows: chunk groups of 8 (common.cuh)
    h->sm_count = prop.multiProcessorCount;
    h->dict = dict ? dict : std::make_shared<MetaDict>();
    CK(cudaStreamCreateWithFlags(&h->mut_stream, cudaStreamNonBlocking));
    CK(scan_configure(qt, 224 * 1024));
    if (qt <= Q16) CK(batch_configure(batch_dynamic_limit()));
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

int destroy_single(szg_index *h) {
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

} // namespace szg

// ====================================================================== C ABI
extern "C" {

const char *szg_last_error(void) { return szg::last_error_string().c_str(); }

int szg_create(int dim, int quantization, int metric, int device, szg_index **out) {
    return create_single(dim, quantization, metric, device, nullptr, out);
}

int szg_destroy(szg_index *h) {
    if (h && h->sh) return sharded_destroy(h);
    return destroy_single(h);
}

int szg_set_option(szg_index *h, int option, int64_t value) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (h->sh) return sharded_set_option(h, option, value);
    h->generation++; // captured launch sequences were recorded under the old options
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
    case SZG_OPT_BATCH_MIN_QUERIES:
        if (value < 0 || value > 4096) return fail(SZG_EINVAL, "batch threshold must be in [0, 4096]");
        h->batch_min = (int)value;
        return SZG_OK;
    case SZG_OPT_GRAPHS: h->use_graphs = value != 0; return SZG_OK;
    case SZG_OPT_TRACE_BUFFER: h->trace = reinterpret_cast<long long *>((uintptr_t)value); return SZG_OK;
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
    if (h && h->sh) return sharded_reserve(h, nrows);
    GUARD(h);
    return grow(h, nrows);
}

int szg_count(szg_index *h, uint64_t *n) {
    if (!h || !n) return fail(SZG_EINVAL, "null argument");
    *n = h->sh ? sharded_count(h) : h->live_rows;
    return SZG_OK;
}

} // extern "C"

namespace szg {

// One staged batch of rows into the mirror: slot assignment on the host, then scatter + aux on the device.  The
// stream-1 bytes come either from the caller (codes) or from encode_kernel over the caller's float64 vectors, in
// which case they are also handed back (out_codes) for the span file.
int upsert_rows(szg_index *h, const uint64_t *ids, const uint8_t *codes, const double *vectors, uint8_t *out_codes,
                uint64_t n, bool into_mirror) {
    if (into_mirror) { h->planar_dirty = true; h->generation++; }
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

} // namespace szg

extern "C" {

int szg_upsert(szg_index *h, const uint64_t *ids, const uint8_t *codes, uint64_t n) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (n && (!ids || !codes)) return fail(SZG_EINVAL, "null ids/codes");
    if (h->sh) return sharded_upsert(h, ids, codes, nullptr, nullptr, n, true);
    GUARD(h);
    return upsert_rows(h, ids, codes, nullptr, nullptr, n, true);
}

int szg_encode(szg_index *h, const uint64_t *ids, const double *vectors, uint64_t n, uint8_t *out_codes, int upsert) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (n && !vectors) return fail(SZG_EINVAL, "null vectors");
    if (n && upsert && !ids) return fail(SZG_EINVAL, "null ids");
    if (!upsert && !out_codes) return fail(SZG_EINVAL, "nothing to do: no output buffer and no upsert");
    if (h->sh) return sharded_upsert(h, upsert ? ids : nullptr, nullptr, vectors, out_codes, n, upsert != 0);
    GUARD(h);
    return upsert_rows(h, ids, nullptr, vectors, out_codes, n, upsert != 0);
}

int szg_remove(szg_index *h, const uint64_t *ids, uint64_t n, uint64_t *n_removed) {
    if (h && h->sh) return n && !ids ? fail(SZG_EINVAL, "null ids") : sharded_remove(h, ids, n, n_removed);
    GUARD(h);
    if (n && !ids) return fail(SZG_EINVAL, "null ids");
    h->generation++;
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
    if (h && h->sh) return nrows ? sharded_fill_synthetic(h, seed, row0, nrows) : SZG_OK;
    GUARD(h);
    if (!nrows) return SZG_OK;
    h->planar_dirty = true;
    h->generation++;
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
    if (h && h->sh) return n && (!ids || !out_codes) ? fail(SZG_EINVAL, "null argument") : sharded_fetch_codes(h, ids, n, out_codes);
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
    if (h && h->sh) return !mask_id || (n && (!ids || !pass)) ? fail(SZG_EINVAL, "null argument") : sharded_mask_create(h, ids, pass, n, mask_id);
    GUARD(h);
    if (!mask_id || (n && (!ids || !pass))) return fail(SZG_EINVAL, "null argument");
    if (n > 0xFFFFFFFFull) return fail(SZG_EINVAL, "too many ids");
    // Search builds masks while it holds only the RLock (collection.go:570, 592-594): builders take turns on the staging
    // buffers, running searches are not disturbed (they only look masks up)
    std::lock_guard<std::mutex> build(h->mask_build_mu);
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
    std::lock_guard<std::mutex> lk(h->mask_mu);
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
    auto it = h->dict->codes.find(key);
    if (it != h->dict->codes.end()) return it->second;
    if (!insert) return 0xFFFFFFFFu;
    const uint32_t code = (uint32_t)h->dict->strs.size();
    h->dict->strs.push_back(key);
    h->dict->codes.emplace(std::move(key), code);
    return code;
}

int szg_meta_upsert(szg_index *h, const uint64_t *ids, uint64_t n, const uint8_t *doc_kind, const uint32_t *cols,
                    uint32_t ncols, const szg_meta_value *values) {
    if (h && h->sh) {
        if (n && (!ids || !doc_kind || (ncols && (!cols || !values)))) return fail(SZG_EINVAL, "null argument");
        for (uint32_t j = 0; j < ncols; ++j)
            if (cols[j] >= kFilterMaxCols) return fail(SZG_EINVAL, "metadata column %u out of range (max %u)", cols[j], kFilterMaxCols - 1);
        return sharded_meta_upsert(h, ids, n, doc_kind, cols, ncols, values);
    }
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
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (!size) return fail(SZG_EINVAL, "null argument");
    *size = (uint32_t)h->dict->strs.size();
    return SZG_OK;
}

int szg_meta_dictionary_get(szg_index *h, uint32_t code, const char **str, uint32_t *len) {
    if (!h) return fail(SZG_EINVAL, "null handle");
    if (!str || !len) return fail(SZG_EINVAL, "null argument");
    if (code >= h->dict->strs.size()) return fail(SZG_ENOTFOUND, "no string with code %u", code);
    *str = h->dict->strs[code].data();
    *len = (uint32_t)h->dict->strs[code].size();
    return SZG_OK;
}

int szg_filter_mask(szg_index *h, const szg_filter_op *ops, uint32_t nops, int *mask_id) {
    if (h && h->sh) return !ops || !nops || !mask_id ? fail(SZG_EINVAL, "null argument") : sharded_filter_mask(h, ops, nops, mask_id);
    GUARD(h);
    if (!ops || !nops || !mask_id) return fail(SZG_EINVAL, "null argument");
    if (nops > 4096) return fail(SZG_EINVAL, "filter program too long");
    std::lock_guard<std::mutex> build(h->mask_build_mu);
    int rc;
    cudaStream_t st = h->mut_stream;
    // ---- validate the stack discipline and lower the literals
    const uint32_t D = (uint32_t)h->dict->strs.size();
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
                const std::string &x = h->dict->strs[c];
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
        auto str_of = [&](uint32_t c) -> const std::string & { return c < D ? h->dict->strs[c] : lits[c - D]; };
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
    std::lock_guard<std::mutex> lk(h->mask_mu);
    *mask_id = h->next_mask++;
    h->masks[*mask_id] = mask;
    return SZG_OK;
}

int szg_mask_destroy(szg_index *h, int mask_id) {
    if (h && h->sh) return sharded_mask_destroy(h, mask_id);
    GUARD(h);
    uint32_t *p = nullptr;
    {
        std::lock_guard<std::mutex> lk(h->mask_mu);
        auto it = h->masks.find(mask_id);
        if (it == h->masks.end()) return fail(SZG_ENOTFOUND, "unknown mask id %d", mask_id);
        p = it->second;
        h->masks.erase(it);
    }
    h->generation++; // a captured launch sequence may hold this mask's address
    cudaFree(p);
    return SZG_OK;
}

int szg_get_stats(szg_index *h, szg_stats *out) {
    if (h && h->sh) return out ? sharded_get_stats(h, out) : fail(SZG_EINVAL, "null out");
    GUARD(h);
    if (!out) return fail(SZG_EINVAL, "null out");
    memset(out, 0, sizeof *out);
    out->kernel_launches = h->launches;
    out->escalations = h->escalations;
    out->uncertain_results = h->uncertain;
    out->batch_queries = h->batch_queries;
    out->shards = 1;
    out->combined_queries = h->combined_queries;
    out->graph_launches = h->graph_launches;
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


} // extern "C"
