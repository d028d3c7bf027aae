// scan_impl.cuh -- K1/K2: single-query streaming scan of the column-blocked mirror with
// fused decode, distance surrogate, top-k (per-warp register lists merged with shuffles,
// block merge in shared memory, last-CTA final merge + fp64 re-score) or radius
// compaction (warp ballot).  One launch per query.
//
// Replaces, per record: getDocument/decodeVector (collection.go:470-484, 768-794), the
// distance call (596) and the heap logic of `consider` (598-628) -- N calls become one
// kernel.  HBM-bound: every code byte is read exactly once with coalesced 128-bit loads.
//
// Arithmetic: for 4/8/16-bit codes the row-dependent part of both
// metrics is I = sum_i u_i * W_i with W_i = round(w_i * 2^F) a 21-bit fixed-point copy of
// the query coefficient, split in three signed base-128 digits so that I is three exact
// integer dot products (IDP.4A / IDP.2A).  All cancellation happens on exact integers;
// only the final key is rounded to fp32.  32/64-bit rows use fp32/fp64 FMAs directly.
#pragma once
#include "common.cuh"
#include "exact.cuh"

namespace szg {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kMaxListE = 8;            // candidates per lane; candidate set = 32 * E
constexpr int MODE_RADIUS = 4;          // MODE 0..3: top-k with E = 1 << MODE

struct ScanArgs {
    const uint4 *codes;
    const void *aux;
    const uint32_t *live; // bit r of live[b]: row r of block b holds a live record
    const uint32_t *mask; // optional filter bitmask, same indexing (NULL = none)
    const unsigned long long *ids;
    const double *lut;          // dequantize table (4/8/16-bit)
    const unsigned char *pq;    // PQHeader + payload of this query
    const double *q;            // raw float64 query (re-score)
    uint32_t C, nblk, dims, metric;
    uint32_t k, flags;
    // top-k workspace / outputs
    unsigned long long *cand;   // [gridDim.x][32*E]
    unsigned int *ticket;
    unsigned long long *out_ids;
    double *out_dist;
    uint32_t *out_n;
    uint32_t *out_flags;        // bit0: candidate margin below tolerance ("uncertain")
    // radius outputs
    uint32_t *rad_count;
    uint32_t *rad_slots;
    uint32_t rad_cap;
};

// ------------------------------------------------------------------ per-warp sorted list
template <int E>
struct WarpList {
    unsigned long long v[E]; // lane-major: lane i holds ranks [i*E, i*E+E)
    unsigned long long thr;  // warp-uniform copy of the worst kept key

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = kNoKey;
        thr = kNoKey;
    }
    __device__ __forceinline__ void insert(unsigned long long nk, int lane) {
        unsigned gt = __ballot_sync(0xffffffffu, v[E - 1] > nk);
        int p = __ffs(gt) - 1;
        unsigned long long carry = __shfl_up_sync(0xffffffffu, v[E - 1], 1);
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) cnt += (v[e] < nk);
#pragma unroll
        for (int e = E - 1; e >= 1; --e) {
            unsigned long long prev = v[e - 1];
            if (lane > p) v[e] = prev;
            else if (lane == p) {
                if (e > cnt) v[e] = prev;
                else if (e == cnt) v[e] = nk;
            }
        }
        if (lane > p) v[0] = carry;
        else if (lane == p && cnt == 0) v[0] = nk;
        thr = __shfl_sync(0xffffffffu, v[E - 1], 31);
    }
    // every lane offers one key (kNoKey = nothing)
    __device__ __forceinline__ void offer(unsigned long long ck, int lane) {
        unsigned m = __ballot_sync(0xffffffffu, ck < thr);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            unsigned long long nk = __shfl_sync(0xffffffffu, ck, src);
            if (nk < thr) insert(nk, lane);
        }
    }
};

// ascending bitonic sort of n (power of two) keys in shared memory by the whole CTA
__device__ __forceinline__ void block_bitonic_sort(unsigned long long *s, int n, int tid) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += kScanThreads) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = s[i], b = s[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { s[i] = b; s[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

template <int E>
__device__ __forceinline__ void block_merge(const WarpList<E> &L, unsigned long long *pool, int tid, int lane,
                                            int warp) {
#pragma unroll
    for (int e = 0; e < E; ++e) pool[(warp * 32 + lane) * E + e] = L.v[e];
    __syncthreads();
    block_bitonic_sort(pool, kScanWarps * 32 * E, tid);
}

// ------------------------------------------------------------------------- row scoring
__device__ __forceinline__ double digits_total(const int (&a)[ND]) {
    double t = (double)a[0];
#pragma unroll
    for (int j = 1; j < ND; ++j) t = t * 128.0 + (double)a[j];
    return t;
}

template <int QT>
__device__ __forceinline__ float finish_quant(const ScanArgs &a, const PQHeader &h, double I, uint32_t slot) {
    if (a.metric == COSINE) {
        float rn = reinterpret_cast<const float *>(a.aux)[slot];
        double num = 2.0 * I + h.numc;
        float c = (float)(num * h.c_key) * rn;
        return (h.zero_query || rn == 0.f) ? 1.0f : -c;
    }
    double s2 = (QT == Q16) ? (double)reinterpret_cast<const unsigned long long *>(a.aux)[slot]
                            : (double)reinterpret_cast<const uint32_t *>(a.aux)[slot];
    double Ev = fma(-h.pow2F1, I, fma(s2, h.pow2F2, h.base));
    return (float)(Ev * h.c_key);
}

template <int QT>
struct Scorer;

// 8-bit codes: 16 dims per chunk, payload = ND uint4 of digits per chunk
template <>
struct Scorer<Q8> {
    static __device__ __forceinline__ void step(const uint4 &v, const uint4 *dg, int (&acc)[ND]) {
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            uint4 d = dg[j];
            acc[j] = dp4a_us(v.x, (int)d.x, acc[j]);
            acc[j] = dp4a_us(v.y, (int)d.y, acc[j]);
            acc[j] = dp4a_us(v.z, (int)d.z, acc[j]);
            acc[j] = dp4a_us(v.w, (int)d.w, acc[j]);
        }
    }
    static __device__ __forceinline__ void run(const ScanArgs &a, const unsigned char *spq, const PQHeader &h,
                                               const uint4 *p0, const uint4 *p1, uint32_t slot0, uint32_t slot1,
                                               float (&key)[2]) {
        const uint4 *dg = reinterpret_cast<const uint4 *>(spq);
        const uint32_t C = a.C;
        int acc0[ND], acc1[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) acc0[j] = acc1[j] = 0;
        constexpr int U = 4;
        uint32_t c = 0;
        for (; c + U <= C; c += U) {
            uint4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v0[u] = ldg_stream(p0 + (size_t)(c + u) * 32);
                v1[u] = ldg_stream(p1 + (size_t)(c + u) * 32);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                step(v0[u], dg + (c + u) * ND, acc0);
                step(v1[u], dg + (c + u) * ND, acc1);
            }
        }
        for (; c < C; ++c) {
            uint4 v0 = ldg_stream(p0 + (size_t)c * 32), v1 = ldg_stream(p1 + (size_t)c * 32);
            step(v0, dg + c * ND, acc0);
            step(v1, dg + c * ND, acc1);
        }
        key[0] = finish_quant<Q8>(a, h, digits_total(acc0), slot0);
        key[1] = finish_quant<Q8>(a, h, digits_total(acc1), slot1);
    }
};

// 4-bit codes: 32 dims per chunk; byte = (even dim << 4) | odd dim.  Payload per chunk =
// ND uint4 for the even dims (applied to w & 0xF0F0F0F0, i.e. 16*u) + ND uint4 for the odd.
template <>
struct Scorer<Q4> {
    static __device__ __forceinline__ void step(const uint4 &v, const uint4 *dg, int (&hi)[ND], int (&lo)[ND]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            uint4 da = dg[j], db = dg[ND + j];
            const uint32_t a4[4] = {da.x, da.y, da.z, da.w}, b4[4] = {db.x, db.y, db.z, db.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                hi[j] = dp4a_us(w[k] & 0xF0F0F0F0u, (int)a4[k], hi[j]);
                lo[j] = dp4a_us(w[k] & 0x0F0F0F0Fu, (int)b4[k], lo[j]);
            }
        }
    }
    static __device__ __forceinline__ void run(const ScanArgs &a, const unsigned char *spq, const PQHeader &h,
                                               const uint4 *p0, const uint4 *p1, uint32_t slot0, uint32_t slot1,
                                               float (&key)[2]) {
        const uint4 *dg = reinterpret_cast<const uint4 *>(spq);
        const uint32_t C = a.C;
        int hi0[ND], lo0[ND], hi1[ND], lo1[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) hi0[j] = lo0[j] = hi1[j] = lo1[j] = 0;
        constexpr int U = 4;
        uint32_t c = 0;
        for (; c + U <= C; c += U) {
            uint4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v0[u] = ldg_stream(p0 + (size_t)(c + u) * 32);
                v1[u] = ldg_stream(p1 + (size_t)(c + u) * 32);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                step(v0[u], dg + (c + u) * 2 * ND, hi0, lo0);
                step(v1[u], dg + (c + u) * 2 * ND, hi1, lo1);
            }
        }
        for (; c < C; ++c) {
            uint4 v0 = ldg_stream(p0 + (size_t)c * 32), v1 = ldg_stream(p1 + (size_t)c * 32);
            step(v0, dg + c * 2 * ND, hi0, lo0);
            step(v1, dg + c * 2 * ND, hi1, lo1);
        }
        // hi accumulates 16 * u_even * W: I = hi/16 + lo (exact in double)
        double I0 = digits_total(hi0) * 0.0625 + digits_total(lo0);
        double I1 = digits_total(hi1) * 0.0625 + digits_total(lo1);
        key[0] = finish_quant<Q4>(a, h, I0, slot0);
        key[1] = finish_quant<Q4>(a, h, I1, slot1);
    }
};

// 16-bit codes, stored as little-endian int16 of (u - 32768): 8 dims per chunk, payload =
// ND uint2 of digits per chunk.  |s16 * s8| <= 2^22, so int32 partials are flushed to
// double every 32 chunks (256 dims).
template <>
struct Scorer<Q16> {
    static __device__ __forceinline__ void step(const uint4 &v, const uint2 *dg, int (&acc)[ND]) {
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            uint2 d = dg[j];
            acc[j] = dp2a_lo_ss((int)v.x, (int)d.x, acc[j]);
            acc[j] = dp2a_hi_ss((int)v.y, (int)d.x, acc[j]);
            acc[j] = dp2a_lo_ss((int)v.z, (int)d.y, acc[j]);
            acc[j] = dp2a_hi_ss((int)v.w, (int)d.y, acc[j]);
        }
    }
    static __device__ __forceinline__ void run(const ScanArgs &a, const unsigned char *spq, const PQHeader &h,
                                               const uint4 *p0, const uint4 *p1, uint32_t slot0, uint32_t slot1,
                                               float (&key)[2]) {
        const uint2 *dg = reinterpret_cast<const uint2 *>(spq);
        const uint32_t C = a.C;
        double I0 = 0.0, I1 = 0.0;
        constexpr uint32_t W = 32;
        for (uint32_t cw = 0; cw < C; cw += W) {
            const uint32_t cend = min(C, cw + W);
            int acc0[ND], acc1[ND];
#pragma unroll
            for (int j = 0; j < ND; ++j) acc0[j] = acc1[j] = 0;
            constexpr int U = 4;
            uint32_t c = cw;
            for (; c + U <= cend; c += U) {
                uint4 v0[U], v1[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    v0[u] = ldg_stream(p0 + (size_t)(c + u) * 32);
                    v1[u] = ldg_stream(p1 + (size_t)(c + u) * 32);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    step(v0[u], dg + (c + u) * ND, acc0);
                    step(v1[u], dg + (c + u) * ND, acc1);
                }
            }
            for (; c < cend; ++c) {
                uint4 v0 = ldg_stream(p0 + (size_t)c * 32), v1 = ldg_stream(p1 + (size_t)c * 32);
                step(v0, dg + c * ND, acc0);
                step(v1, dg + c * ND, acc1);
            }
            I0 += digits_total(acc0);
            I1 += digits_total(acc1);
        }
        key[0] = finish_quant<Q16>(a, h, I0, slot0);
        key[1] = finish_quant<Q16>(a, h, I1, slot1);
    }
};

// 32-bit float rows: payload = the query as float4 per chunk
template <>
struct Scorer<F32> {
    template <int METRIC>
    static __device__ __forceinline__ void step(const uint4 &v, const float4 &q, float (&acc)[4]) {
        const float x[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
        const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (METRIC == COSINE) acc[k] = fmaf(x[k], qq[k], acc[k]);
            else { float d = qq[k] - x[k]; acc[k] = fmaf(d, d, acc[k]); }
        }
    }
    template <int METRIC>
    static __device__ __forceinline__ void loop(const ScanArgs &a, const float4 *sq, const uint4 *p0, const uint4 *p1,
                                                float &r0, float &r1) {
        const uint32_t C = a.C;
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
        constexpr int U = 4;
        uint32_t c = 0;
        for (; c + U <= C; c += U) {
            uint4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v0[u] = ldg_stream(p0 + (size_t)(c + u) * 32);
                v1[u] = ldg_stream(p1 + (size_t)(c + u) * 32);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float4 q = sq[c + u];
                step<METRIC>(v0[u], q, acc0);
                step<METRIC>(v1[u], q, acc1);
            }
        }
        for (; c < C; ++c) {
            float4 q = sq[c];
            step<METRIC>(ldg_stream(p0 + (size_t)c * 32), q, acc0);
            step<METRIC>(ldg_stream(p1 + (size_t)c * 32), q, acc1);
        }
        r0 = (acc0[0] + acc0[1]) + (acc0[2] + acc0[3]);
        r1 = (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]);
    }
    static __device__ __forceinline__ void run(const ScanArgs &a, const unsigned char *spq, const PQHeader &h,
                                               const uint4 *p0, const uint4 *p1, uint32_t slot0, uint32_t slot1,
                                               float (&key)[2]) {
        const float4 *sq = reinterpret_cast<const float4 *>(spq);
        float r0, r1;
        if (a.metric == COSINE) {
            loop<COSINE>(a, sq, p0, p1, r0, r1);
            const float *rn = reinterpret_cast<const float *>(a.aux);
            float rn0 = rn[slot0], rn1 = rn[slot1], ck = (float)h.c_key;
            key[0] = (h.zero_query || rn0 == 0.f) ? 1.0f : -(r0 * ck) * rn0;
            key[1] = (h.zero_query || rn1 == 0.f) ? 1.0f : -(r1 * ck) * rn1;
        } else {
            loop<EUCLID>(a, sq, p0, p1, r0, r1);
            key[0] = r0;
            key[1] = r1;
        }
    }
};

// 64-bit float rows: payload = the query as double2 per chunk
template <>
struct Scorer<F64> {
    template <int METRIC>
    static __device__ __forceinline__ void step(const uint4 &v, const double2 &q, double (&acc)[2]) {
        const double x0 = __hiloint2double((int)v.y, (int)v.x), x1 = __hiloint2double((int)v.w, (int)v.z);
        if (METRIC == COSINE) {
            acc[0] = fma(x0, q.x, acc[0]);
            acc[1] = fma(x1, q.y, acc[1]);
        } else {
            double d0 = q.x - x0, d1 = q.y - x1;
            acc[0] = fma(d0, d0, acc[0]);
            acc[1] = fma(d1, d1, acc[1]);
        }
    }
    template <int METRIC>
    static __device__ __forceinline__ void loop(const ScanArgs &a, const double2 *sq, const uint4 *p0, const uint4 *p1,
                                                double &r0, double &r1) {
        const uint32_t C = a.C;
        double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
        constexpr int U = 4;
        uint32_t c = 0;
        for (; c + U <= C; c += U) {
            uint4 v0[U], v1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v0[u] = ldg_stream(p0 + (size_t)(c + u) * 32);
                v1[u] = ldg_stream(p1 + (size_t)(c + u) * 32);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double2 q = sq[c + u];
                step<METRIC>(v0[u], q, acc0);
                step<METRIC>(v1[u], q, acc1);
            }
        }
        for (; c < C; ++c) {
            double2 q = sq[c];
            step<METRIC>(ldg_stream(p0 + (size_t)c * 32), q, acc0);
            step<METRIC>(ldg_stream(p1 + (size_t)c * 32), q, acc1);
        }
        r0 = acc0[0] + acc0[1];
        r1 = acc1[0] + acc1[1];
    }
    static __device__ __forceinline__ void run(const ScanArgs &a, const unsigned char *spq, const PQHeader &h,
                                               const uint4 *p0, const uint4 *p1, uint32_t slot0, uint32_t slot1,
                                               float (&key)[2]) {
        const double2 *sq = reinterpret_cast<const double2 *>(spq);
        double r0, r1;
        if (a.metric == COSINE) {
            loop<COSINE>(a, sq, p0, p1, r0, r1);
            const float *rn = reinterpret_cast<const float *>(a.aux);
            float rn0 = rn[slot0], rn1 = rn[slot1];
            key[0] = (h.zero_query || rn0 == 0.f) ? 1.0f : -((float)(r0 * h.c_key)) * rn0;
            key[1] = (h.zero_query || rn1 == 0.f) ? 1.0f : -((float)(r1 * h.c_key)) * rn1;
        } else {
            loop<EUCLID>(a, sq, p0, p1, r0, r1);
            key[0] = (float)r0;
            key[1] = (float)r1;
        }
    }
};

// ------------------------------------------------------- last-CTA finalisation (top-k)
// pool[0 .. K') holds the K' best (surrogate, slot) keys, ascending.  Re-scores them in
// fp64, orders by (distance, lexicographic id) and writes min(k, #) results.
template <int QT>
__device__ void finalize_topk(const ScanArgs &a, const unsigned long long *pool, int Kp, double *s_ex,
                              unsigned long long *s_id, int tid) {
    __shared__ double s_dk, s_dmax;
    if (tid == 0) { s_dk = 0.0; s_dmax = 0.0; }
    bool valid = false;
    double d = 0.0;
    unsigned long long id = 0;
    if (tid < Kp) {
        unsigned long long key = pool[tid];
        if (key != kNoKey) {
            uint32_t slot = (uint32_t)key;
            id = a.ids[slot];
            if (a.flags & 1u) d = key_to_distance(a.metric, key_to_float((uint32_t)(key >> 32)));
            else d = exact_distance<QT>(a.codes, a.C, a.dims, a.metric, a.lut, a.q, slot);
            valid = (d == d); // NaN is never returned
        }
        s_ex[tid] = valid ? d : __longlong_as_double(0x7ff8000000000000ll);
        s_id[tid] = id;
    }
    int nfull = __syncthreads_count(tid < Kp && pool[tid < Kp ? tid : 0] != kNoKey);
    int cnt = __syncthreads_count(valid);
    int rank = -1;
    if (valid) {
        rank = 0;
        for (int j = 0; j < Kp; ++j) {
            double dj = s_ex[j];
            if (j != tid && dj == dj && (dj < d || (dj == d && lex_less_u64(s_id[j], id)))) ++rank;
        }
        if ((uint32_t)rank < a.k) {
            a.out_ids[rank] = id;
            a.out_dist[rank] = d;
        }
        if ((uint32_t)rank + 1 == a.k) s_dk = d;
        if (rank == cnt - 1) s_dmax = d;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t n = (uint32_t)cnt < a.k ? (uint32_t)cnt : a.k;
        *a.out_n = n;
        // all non-candidates have a surrogate no better than the worst candidate; the result is
        // certain when that candidate is clearly (1e-4 relative) farther than the k-th result
        bool uncertain = (nfull == Kp) && ((uint32_t)cnt >= a.k ? !(s_dmax > s_dk * (1.0 + 1e-4)) : true);
        *a.out_flags = uncertain ? 1u : 0u;
    }
}

// ------------------------------------------------------------------------ the scan kernel
template <int QT, int MODE>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const ScanArgs a) {
    constexpr int E = (MODE == MODE_RADIUS) ? 1 : (1 << MODE);
    constexpr int Kp = 32 * E;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const PQHeader h = *reinterpret_cast<const PQHeader *>(a.pq);
    {
        const uint32_t n16 = (a.C * (uint32_t)pq_bytes_per_chunk(QT) + 15) / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(a.pq + sizeof(PQHeader));
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (uint32_t i = tid; i < n16; i += kScanThreads) dst[i] = src[i];
    }
    __syncthreads();

    WarpList<E> list;
    list.init();
    const float radius_key = (float)h.radius_key;

    const uint32_t npairs = (a.nblk + 1) >> 1;
    const uint32_t total_warps = gridDim.x * kScanWarps;
    for (uint32_t pair = blockIdx.x * kScanWarps + warp; pair < npairs; pair += total_warps) {
        const uint32_t blk0 = pair * 2;
        const bool has1 = blk0 + 1 < a.nblk;
        uint32_t live0 = a.live[blk0], live1 = has1 ? a.live[blk0 + 1] : 0u;
        if (a.mask) {
            live0 &= a.mask[blk0];
            if (has1) live1 &= a.mask[blk0 + 1];
        }
        if ((live0 | live1) == 0u) continue; // warp-uniform
        const uint4 *p0 = a.codes + ((size_t)blk0 * a.C) * 32 + lane;
        const uint4 *p1 = has1 ? p0 + (size_t)a.C * 32 : p0;
        const uint32_t slot0 = blk0 * 32 + lane, slot1 = has1 ? slot0 + 32 : slot0;
        float key[2];
        Scorer<QT>::run(a, smem, h, p0, p1, slot0, slot1, key);
        const bool ok0 = (live0 >> lane) & 1u, ok1 = (live1 >> lane) & 1u;
        if (MODE == MODE_RADIUS) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const bool pass = (r ? ok1 : ok0) && key[r] <= radius_key;
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(a.rad_count, (uint32_t)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (pass) {
                        uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < a.rad_cap) a.rad_slots[pos] = r ? slot1 : slot0;
                    }
                }
            }
        } else {
            list.offer(ok0 ? make_key64(key[0], slot0) : kNoKey, lane);
            list.offer(ok1 ? make_key64(key[1], slot1) : kNoKey, lane);
        }
    }
    if (MODE == MODE_RADIUS) return;

    // ---- block merge in shared memory (the digit payload is dead from here on)
    unsigned long long *pool = reinterpret_cast<unsigned long long *>(smem);
    __syncthreads();
    block_merge<E>(list, pool, tid, lane, warp);
    if (tid < Kp) a.cand[(size_t)blockIdx.x * Kp + tid] = pool[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;

    // ---- last CTA: merge all CTA lists, re-score, emit
    __threadfence();
    list.init();
    const uint32_t total = gridDim.x * Kp;
    for (uint32_t base = warp * 32; base < total; base += kScanThreads) {
        const uint32_t i = base + lane;
        list.offer(i < total ? ld_cg_u64(a.cand + i) : kNoKey, lane);
    }
    __syncthreads();
    block_merge<E>(list, pool, tid, lane, warp);
    double *s_ex = reinterpret_cast<double *>(pool + kScanWarps * Kp);
    unsigned long long *s_id = reinterpret_cast<unsigned long long *>(s_ex + Kp);
    finalize_topk<QT>(a, pool, Kp, s_ex, s_id, tid);
    if (tid == 0) *a.ticket = 0u; // self-cleaning for the next launch on this workspace
}

// dynamic shared memory a launch needs
inline size_t scan_smem_bytes(int qt, uint32_t C, int mode) {
    size_t payload = ((size_t)C * pq_bytes_per_chunk(qt) + 15) / 16 * 16;
    if (mode == MODE_RADIUS) return payload;
    size_t Kp = 32u << mode;
    size_t pool = (size_t)kScanWarps * Kp * 8 + Kp * 16;
    return payload > pool ? payload : pool;
}

// host-side launcher, instantiated per quantization in scan_<qt>.cu
template <int QT>
cudaError_t launch_scan_t(int mode, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    switch (mode) {
    case 0: scan_kernel<QT, 0><<<grid, kScanThreads, smem, st>>>(a); break;
    case 1: scan_kernel<QT, 1><<<grid, kScanThreads, smem, st>>>(a); break;
    case 2: scan_kernel<QT, 2><<<grid, kScanThreads, smem, st>>>(a); break;
    case 3: scan_kernel<QT, 3><<<grid, kScanThreads, smem, st>>>(a); break;
    case MODE_RADIUS: scan_kernel<QT, MODE_RADIUS><<<grid, kScanThreads, smem, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int QT>
cudaError_t scan_configure_t(size_t max_smem, int *blocks_per_sm) {
    cudaError_t e;
    int best = 0;
#define SZG_CFG(M)                                                                                                \
    e = cudaFuncSetAttribute(scan_kernel<QT, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);       \
    if (e != cudaSuccess) return e;                                                                                \
    if (M == 0) {                                                                                                  \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&best, scan_kernel<QT, M>, kScanThreads, 8192);           \
        if (e != cudaSuccess) return e;                                                                            \
    }
    SZG_CFG(0) SZG_CFG(1) SZG_CFG(2) SZG_CFG(3) SZG_CFG(MODE_RADIUS)
#undef SZG_CFG
    *blocks_per_sm = best;
    return cudaSuccess;
}

} // namespace szg
