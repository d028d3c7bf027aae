// gather.cu -- K3 and the exact half of K2: fp64 distances of GATHERED rows, in the reference's operation order.
//
//   rescore_kernel        candidate lists of the LSH index (lshtree.go:316-335 calling `consider`, collection.go:584-596:
//                         getDocument + decodeVector + c.distance per candidate), visit order preserved; many lists
//                         (queries) per launch
//   radius_exact_kernel   the rows a radius scan compacted (collection.go:598-605): exact distance, inclusive test
//   radius_sort_*         ascending (distance, lexicographic decimal id) order of the hits (collection.go:693-697 pops
//                         the heap back to front; ties: scan order spanfile.go:540-560), bitonic networks on the device
//
// HBM gather-bound.  A CTA takes a batch of candidates of one list, fetches their rows slab by slab with all its threads
// (exact_staged, scan_impl.cuh) and lets thread r run candidate r's sequential fp64 chain out of shared memory.  Float rows
// keep their 16-byte chunks in groups of 8 (common.cuh), so the fetch reads whole 128-byte lines: algorithmic bytes =
// m x rowbytes, and that is what comes from DRAM.
#include <stdlib.h>

#include "kernels.h"

namespace szg {

namespace {

constexpr size_t kGatherStageBytes = 64 * 1024;

constexpr int kGatherThreads = 256; // all of them fetch

// KP = 32: few candidates (one speculative batch of the LSH walk) -- the latency form (exact_staged: products in parallel,
// one thread per running sum); KP = 128: many -- the throughput form (exact_stream: a thread per candidate, fp64-pipe bound)
template <int QT, int KP>
__device__ __forceinline__ void score_batch(const uint4 *codes, const double *lut, uint32_t C, uint32_t dims, uint32_t metric,
                                            const double *q, double m1, const uint32_t *s_slot, unsigned char *stage, double *s_out,
                                            int tid) {
    if (KP <= 32) {
        if (metric == COSINE) exact_staged<QT, COSINE, kGatherThreads, 16>(codes, lut, C, dims, q, s_slot, KP, stage, kGatherStageBytes, s_out, tid);
        else exact_staged<QT, EUCLID, kGatherThreads, 16>(codes, lut, C, dims, q, s_slot, KP, stage, kGatherStageBytes, s_out, tid);
    } else {
        if (metric == COSINE) exact_stream<QT, COSINE, kGatherThreads>(codes, lut, C, dims, q, m1, s_slot, KP, stage, kGatherStageBytes, s_out, tid);
        else exact_stream<QT, EUCLID, kGatherThreads>(codes, lut, C, dims, q, m1, s_slot, KP, stage, kGatherStageBytes, s_out, tid);
    }
}

// CTA b scores candidates [b KP, b KP + KP) of the flat candidate array; where that range crosses a list boundary the
// pieces are scored one after the other, each against its own query.
template <int QT, int KP>
__global__ void __launch_bounds__(kGatherThreads, 2) rescore_kernel(const RescoreArgs a) {
    extern __shared__ __align__(16) unsigned char stage[];
    __shared__ uint32_t s_slot[KP];
    __shared__ double s_out[KP];
    const int tid = threadIdx.x;
    uint32_t base = blockIdx.x * KP;
    const uint32_t end = min(base + (uint32_t)KP, a.m);
    // the list that holds candidate `base`: the last l with list_off[l] <= base
    uint32_t l = 0;
    if (a.list_off) {
        uint32_t lo = 0, hi = a.nlists; // invariant: list_off[lo] <= base < list_off[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.list_off[mid] <= base) lo = mid; else hi = mid;
        }
        l = lo;
    }
    while (base < end) {
        const uint32_t lend = a.list_off ? min(end, a.list_off[l + 1]) : end;
        if (lend > base) {
            const uint32_t n = lend - base;
            const uint32_t slot = (uint32_t)tid < n ? a.slots[base + tid] : 0xFFFFFFFFu;
            __syncthreads(); // the previous piece's s_out / s_slot readers are done
            if (tid < KP) s_slot[tid] = slot;
            __syncthreads();
            score_batch<QT, KP>(a.codes, a.lut, a.C, a.dims, a.metric, a.q + (size_t)l * a.dims, a.m1 ? a.m1[l] : 0.0, s_slot, stage, s_out, tid);
            if ((uint32_t)tid < n) {
                a.out_dist[base + tid] = slot == 0xFFFFFFFFu ? -1.0 /* SZG_MISSING_DISTANCE */ : s_out[tid];
                if (a.out_ids) a.out_ids[base + tid] = slot == 0xFFFFFFFFu ? 0ull : a.ids[slot];
            }
            base = lend;
        }
        ++l;
    }
}

template <int QT, int KP>
__global__ void __launch_bounds__(kGatherThreads, 2) radius_exact_kernel(const RadiusFinishArgs a) {
    extern __shared__ __align__(16) unsigned char stage[];
    __shared__ uint32_t s_slot[KP];
    __shared__ double s_out[KP];
    const int tid = threadIdx.x;
    const uint32_t m = min(*a.count_ptr, a.cap);
    const uint32_t base = blockIdx.x * KP;
    if (base >= m) return;
    const uint32_t n = min((uint32_t)KP, m - base);
    if (tid < KP) s_slot[tid] = (uint32_t)tid < n ? a.slots[base + tid] : 0xFFFFFFFFu;
    __syncthreads();
    score_batch<QT, KP>(a.codes, a.lut, a.C, a.dims, a.metric, a.q, a.m1, s_slot, stage, s_out, tid);
    if ((uint32_t)tid < n) {
        const double d = s_out[tid];
        if (d <= a.radius) { // inclusive (collection.go:598); NaN fails
            const uint32_t pos = atomicAdd(a.out_count, 1u);
            a.keys[2 * (size_t)pos] = (unsigned long long)__double_as_longlong(d); // d >= 0: the bits order like the values
            a.keys[2 * (size_t)pos + 1] = a.ids[s_slot[tid]];
        }
    }
}

// ---- ordering of the hits: bitonic networks over (distance bits, id) pairs
struct Hit { unsigned long long d, id; };
__device__ __forceinline__ bool hit_less(const Hit &x, const Hit &y) {
    return x.d < y.d || (x.d == y.d && lex_less_u64(x.id, y.id));
}
constexpr uint32_t kSortTile = kRadiusSortSmall; // pairs sorted per CTA in shared memory (32 KB)
constexpr int kSortThreads = 1024;

// steps j = jmax .. 1 of stage k on the tile in shared memory; i0 = global index of the tile's first element
__device__ __forceinline__ void tile_steps(Hit *s, uint32_t i0, uint32_t k, uint32_t jmax, int tid) {
    for (uint32_t j = jmax; j > 0; j >>= 1) {
        for (uint32_t t = tid; t < kSortTile / 2; t += kSortThreads) {
            const uint32_t i = 2 * t - (t & (j - 1)); // the lower index of the pair
            const uint32_t p = i + j;
            const bool up = ((i0 + i) & k) == 0;
            const Hit x = s[i], y = s[p];
            if (hit_less(y, x) == up) { s[i] = y; s[p] = x; }
        }
        __syncthreads();
    }
}

// FULL: sorts the tile from scratch (stages 2 .. tile) -- else only the in-tile steps of stage k.  Elements at global index
// >= load_m read as +infinity; with out_dist set (last pass) the first out_m elements are written out unzipped.  count_ptr
// != NULL is the speculative launch of the small case: load_m = out_m = *count_ptr, and nothing happens above only_if_le.
template <bool FULL>
__global__ void __launch_bounds__(kSortThreads) radius_sort_tile_kernel(unsigned long long *keys, const uint32_t *count_ptr, uint32_t only_if_le,
                                                                        uint32_t load_m, uint32_t out_m, uint32_t k, double *out_dist,
                                                                        unsigned long long *out_ids) {
    __shared__ Hit s[kSortTile];
    if (count_ptr) {
        const uint32_t m = *count_ptr;
        if (m > only_if_le) return; // larger results are sorted by later launches (launch_radius_sort_large)
        load_m = out_m = m;
    }
    const int tid = threadIdx.x;
    const uint32_t i0 = blockIdx.x * kSortTile;
    for (uint32_t t = tid; t < kSortTile; t += kSortThreads) {
        const uint32_t i = i0 + t;
        Hit h;
        h.d = i < load_m ? keys[2 * (size_t)i] : ~0ull;
        h.id = i < load_m ? keys[2 * (size_t)i + 1] : ~0ull;
        s[t] = h;
    }
    __syncthreads();
    if (FULL) {
        for (uint32_t kk = 2; kk <= kSortTile; kk <<= 1) tile_steps(s, i0, kk, kk >> 1, tid);
    } else {
        tile_steps(s, i0, k, kSortTile >> 1, tid);
    }
    for (uint32_t t = tid; t < kSortTile; t += kSortThreads) {
        const uint32_t i = i0 + t;
        if (out_dist) {
            if (i < out_m) { out_dist[i] = __longlong_as_double((long long)s[t].d); out_ids[i] = s[t].id; }
        } else {
            keys[2 * (size_t)i] = s[t].d;
            keys[2 * (size_t)i + 1] = s[t].id;
        }
    }
}

// one step (k, j) with j >= the tile size, in global memory over n2 (a power of two) padded elements
__global__ void radius_sort_step_kernel(unsigned long long *keys, uint32_t n2, uint32_t k, uint32_t j) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n2 / 2) return;
    const uint32_t i = 2 * t - (t & (j - 1)), p = i + j;
    const bool up = (i & k) == 0;
    Hit x, y;
    x.d = keys[2 * (size_t)i]; x.id = keys[2 * (size_t)i + 1];
    y.d = keys[2 * (size_t)p]; y.id = keys[2 * (size_t)p + 1];
    if (hit_less(y, x) == up) {
        keys[2 * (size_t)i] = y.d; keys[2 * (size_t)i + 1] = y.id;
        keys[2 * (size_t)p] = x.d; keys[2 * (size_t)p + 1] = x.id;
    }
}

__global__ void radius_pad_kernel(unsigned long long *keys, uint32_t m, uint32_t n2) {
    const uint32_t i = m + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) { keys[2 * (size_t)i] = ~0ull; keys[2 * (size_t)i + 1] = ~0ull; }
}

template <int NT>
cudaError_t rescore_nt(const RescoreArgs &a, cudaStream_t st) { // NT = candidates per CTA
    const unsigned grid = (a.m + NT - 1) / NT;
    cudaError_t e = cudaSuccess;
#define SZG_RS(QT)                                                                                                          \
    do {                                                                                                                    \
        e = cudaFuncSetAttribute(rescore_kernel<QT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGatherStageBytes); \
        if (e == cudaSuccess) rescore_kernel<QT, NT><<<grid, kGatherThreads, kGatherStageBytes, st>>>(a);                               \
    } while (0)
    switch (a.qt) {
    case Q4: SZG_RS(Q4); break;
    case Q8: SZG_RS(Q8); break;
    case Q16: SZG_RS(Q16); break;
    case F32: SZG_RS(F32); break;
    default: SZG_RS(F64); break;
    }
#undef SZG_RS
    return e != cudaSuccess ? e : cudaGetLastError();
}

} // namespace

cudaError_t launch_rescore(const RescoreArgs &a, cudaStream_t st) {
    if (!a.m) return cudaSuccess;
    // few candidates (one speculative batch of the LSH walk: ~200): small CTAs spread them over the SMs and the slabs get
    // long; many (radius hits, bulk re-scoring): 128 candidates per CTA keep the fetches wide
    // (8 per CTA for one LSH batch: a CTA's rows then fit one slab -- one round of fetch latency instead of four -- and the
    // products of 8 candidates leave the fp64 pipe to the chains; SZG_RESCORE_KP=32 for the A/B)
    static const unsigned small_kp = getenv("SZG_RESCORE_KP") ? (unsigned)atoi(getenv("SZG_RESCORE_KP")) : 8u;
    if (small_kp == 8u && a.m <= 8u * 296u) return rescore_nt<8>(a, st);
    return a.m <= 32u * 296u ? rescore_nt<32>(a, st) : rescore_nt<128>(a, st);
}

cudaError_t launch_radius_finish(const RadiusFinishArgs &a, cudaStream_t st) {
    constexpr int NT = 128;
    const unsigned grid = (a.cap + NT - 1) / NT;
    cudaError_t e = cudaSuccess;
#define SZG_RX(QT)                                                                                                           \
    do {                                                                                                                     \
        e = cudaFuncSetAttribute(radius_exact_kernel<QT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGatherStageBytes); \
        if (e == cudaSuccess) radius_exact_kernel<QT, NT><<<grid, kGatherThreads, kGatherStageBytes, st>>>(a);                           \
    } while (0)
    switch (a.qt) {
    case Q4: SZG_RX(Q4); break;
    case Q8: SZG_RX(Q8); break;
    case Q16: SZG_RX(Q16); break;
    case F32: SZG_RX(F32); break;
    default: SZG_RX(F64); break;
    }
#undef SZG_RX
    if (e != cudaSuccess) return e;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // the common case -- a few hundred hits -- is ordered right away by one CTA; it does nothing for larger results
    radius_sort_tile_kernel<true><<<1, kSortThreads, 0, st>>>(a.keys, a.out_count, kRadiusSortSmall, 0, 0, 0, a.out_dist, a.out_ids);
    return cudaGetLastError();
}

// m > kRadiusSortSmall hits in a.keys (capacity: the power of two >= m): pad, sort every tile, merge stage by stage
cudaError_t launch_radius_sort_large(const RadiusFinishArgs &a, uint32_t m, cudaStream_t st) {
    uint32_t n2 = 2 * kSortTile;
    while (n2 < m) n2 <<= 1;
    if (n2 > m) radius_pad_kernel<<<(n2 - m + 255) / 256, 256, 0, st>>>(a.keys, m, n2);
    const unsigned tiles = n2 / kSortTile;
    radius_sort_tile_kernel<true><<<tiles, kSortThreads, 0, st>>>(a.keys, nullptr, 0, n2, 0, 0, nullptr, nullptr);
    for (uint32_t k = 2 * kSortTile; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j >= kSortTile; j >>= 1)
            radius_sort_step_kernel<<<(n2 / 2 + 255) / 256, 256, 0, st>>>(a.keys, n2, k, j);
        const bool last = k == n2;
        radius_sort_tile_kernel<false><<<tiles, kSortThreads, 0, st>>>(a.keys, nullptr, 0, n2, m, k, last ? a.out_dist : nullptr,
                                                                         last ? a.out_ids : nullptr);
    }
    return cudaGetLastError();
}

} // namespace szg
