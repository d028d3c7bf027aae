// scan_impl.cuh -- K1/K2: single-query streaming scan of the column-blocked mirror with
// fused decode, distance surrogate, top-k (per-warp register lists merged with shuffles,
// block merge in shared memory, last-CTA final merge + fp64 re-score) or radius
// compaction (warp ballot).  One launch per query.
//
// Replaces, per record: getDocument/decodeVector (collection.go:470-484, 768-794), the
// distance call (596) and the heap logic of `consider` (598-628) -- N calls become one
// kernel.  HBM-bound: every code byte is read exactly once.
//
// Data movement: persistent CTAs, one per SM.  A 32-row block of the
// column-blocked layout is one contiguous span of C*512 bytes, cut into tiles of Ct chunks.
// Every warp owns a private ring of S shared-memory stages and streams its blocks tile by
// tile with cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on an mbarrier per
// stage; one elected lane issues the copies S tiles ahead, all 32 lanes (one row each)
// consume a stage with conflict-free LDS.128.  Bytes in flight per SM = warps * S * tile,
// independent of the register file, which is what a latency-bound stream needs.
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

constexpr int kMaxScanWarps = 16;       // warps per CTA is a launch parameter (8 or 16)
constexpr int kMaxStages = 8;           // ring stages per warp (launch parameter, 2..8)
constexpr int kMaxTileChunks = 32;      // chunks per tile (Q16 flushes int32 partials per tile)
constexpr int kMaxListE = 8;            // candidates per lane; candidate set = 32 * E
constexpr int MODE_RADIUS = 4;          // MODE 0..3: top-k with E = 1 << MODE

struct ScanArgs {
    const uint4 *codes;
    const void *aux;
    const uint32_t *live; // bit r of live[b]: row r of block b holds a live record
    const uint32_t *mask; // optional filter bitmask, same indexing (NULL = none)
    const unsigned long long *ids;
    const double *lut;          // dequantize table (4/8/16-bit)
    const unsigned char *pq;    // prepared queries: nq x (PQHeader + payload), pq_stride bytes apart
    size_t pq_stride;
    uint32_t nq;
    uint32_t C, nblk, dims, metric;
    // streaming geometry (host-chosen, see scan_plan)
    uint32_t Ct;                // chunks per tile
    uint32_t stages;            // ring stages per warp
    uint32_t pq_smem_off;       // != 0: every warp keeps a private copy of its current prepared query at this
                                // offset of dynamic shared memory (+ warp * pq_stride); 0: read it through L1
    uint32_t parts;             // scan_small_kernel only: work items (row parts) per query
    uint32_t adjacent;          // scan_small_kernel only: the blocks of a step are adjacent (else a warp stride apart)
    uint32_t qper;              // scan_small_kernel only: queries per work item (1 or 2)
    uint32_t const_queries;     // scan_small_kernel only: the launch's prepared queries go through constant memory
    uint32_t wgroups;           // scan_small_kernel only: warp groups per CTA, each on another query of the same row part (1, 2, 4)
    uint32_t cta_merge;         // scan_kernel, one query per launch: the warps of a CTA merge their lists before writing
                                // (one list per CTA instead of one per warp: finalize_kernel reads 16x fewer keys)
    // top-k output: per-warp candidate lists, consumed by finalize_kernel
    unsigned long long *cand;   // [nq][grid warps][32*E]
    // radius outputs
    uint32_t *rad_count;
    uint32_t *rad_slots;
    uint32_t rad_cap;
};

// One finalize launch serves every query of a call: CTA q merges the per-CTA candidate lists of
// query q's scan, re-scores the survivors in fp64 and writes the ordered result.
struct FinalizeArgs {
    const uint4 *codes;
    const unsigned long long *ids;
    const double *lut;
    const double *queries;            // [nq][dims] raw float64 queries
    const unsigned long long *cand;   // [nq][nlists][32*E]: one sorted list per scan warp
    const unsigned char *pq;          // prepared queries (error bounds live in the header)
    size_t pq_stride;
    uint32_t C, dims, metric, k, flags, nlists;
    unsigned long long *out_ids;      // [nq][k]
    double *out_dist;                 // [nq][k]
    uint32_t *out_n;                  // [nq]
    uint32_t *out_flags;              // [nq] bit0: candidate margin below tolerance ("uncertain")
    long long *trace;                 // optional (SZG_OPT_TRACE_BUFFER): clock64 of CTA 0 at the phase boundaries, 8 words
    uint32_t *done_cnt;               // optional [nq], possibly in a PEER device's memory (sharded search): bumped, system
                                      // scope, once query q's results are written -- the merge kernel on the root device
                                      // waits for it (sharded.cu)
};

// ------------------------------------------------------------------ mbarrier / bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

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
    // E == 1 only: merge a batch of 32 keys (one per lane) with two bitonic networks
    // (sort the batch, then merge with the sorted list): bounded cost when many keys pass.
    __device__ __forceinline__ void merge_batch(unsigned long long ck, int lane) {
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                unsigned long long o = __shfl_xor_sync(0xffffffffu, ck, j);
                bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
                ck = keep_min ? (ck < o ? ck : o) : (ck > o ? ck : o);
            }
        }
        unsigned long long rev = __shfl_sync(0xffffffffu, ck, 31 - lane); // descending batch
        unsigned long long m = v[0] < rev ? v[0] : rev;                   // bitonic: the 32 smallest of the union
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, m, j);
            bool keep_min = (lane & j) == 0;
            m = keep_min ? (m < o ? m : o) : (m > o ? m : o);
        }
        v[0] = m;
        thr = __shfl_sync(0xffffffffu, m, 31);
    }
    // E == 1 only: `other` is another SORTED list (lane i = rank i); keeps the 32 smallest of the union, sorted
    __device__ __forceinline__ void merge_sorted(unsigned long long other, int lane) {
        unsigned long long rev = __shfl_sync(0xffffffffu, other, 31 - lane);
        unsigned long long m = v[0] < rev ? v[0] : rev; // bitonic
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            unsigned long long o = __shfl_xor_sync(0xffffffffu, m, j);
            bool keep_min = (lane & j) == 0;
            m = keep_min ? (m < o ? m : o) : (m > o ? m : o);
        }
        v[0] = m;
        thr = __shfl_sync(0xffffffffu, m, 31);
    }
    // every lane offers one key (kNoKey = nothing)
    __device__ __forceinline__ void offer(unsigned long long ck, int lane) {
        unsigned m = __ballot_sync(0xffffffffu, ck < thr);
        if (E == 1 && __popc(m) > 5) {
            merge_batch(ck, lane);
            return;
        }
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            unsigned long long nk = __shfl_sync(0xffffffffu, ck, src);
            if (nk < thr) insert(nk, lane);
        }
    }
};

// ascending bitonic sort of n (power of two) keys in shared memory by the whole CTA
__device__ __forceinline__ void block_bitonic_sort(unsigned long long *s, int n, int tid, int nthreads) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += nthreads) {
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
__device__ __forceinline__ void block_merge(WarpList<E> &L, unsigned long long *pool, int tid, int lane,
                                            int warp, int nwarps) {
#pragma unroll
    for (int e = 0; e < E; ++e) pool[(warp * 32 + lane) * E + e] = L.v[e];
    if (E == 1 && (nwarps & (nwarps - 1)) == 0) {
        // 32-key lists: a tree of pairwise merges of sorted lists (log2(warps) barriers) instead of sorting warps * 32 keys
        for (int half = nwarps >> 1; half >= 1; half >>= 1) {
            __syncthreads();
            if (warp < half) {
                L.merge_sorted(pool[(warp + half) * 32 + lane], lane);
                pool[warp * 32 + lane] = L.v[0];
            }
        }
        __syncthreads();
        return;
    }
    __syncthreads();
    block_bitonic_sort(pool, nwarps * 32 * E, tid, nwarps * 32);
}

// ------------------------------------------------------------------------- row scoring
template <int ND>
__device__ __forceinline__ double digits_total(const int (&a)[ND]) {
    double t = (double)a[0];
#pragma unroll
    for (int j = 1; j < ND; ++j) t = t * 128.0 + (double)a[j];
    return t;
}

// aux values of one row {1/||x||, ||x||^2}, fetched when its block starts so that the latency hides
// behind the tiles
__device__ __forceinline__ float2 load_aux(const ScanArgs &a, uint32_t slot) {
    return __ldg(reinterpret_cast<const float2 *>(a.aux) + slot);
}

__device__ __forceinline__ float finish_quant(const ScanArgs &a, const PQHeader &h, double I, const float2 &x) {
    const double num = 2.0 * I + h.numc;
    if (a.metric == COSINE) {
        float c = (float)(num * h.c_key) * x.x;
        return (h.zero_query || x.x == 0.f) ? 1.0f : -c;
    }
    return (float)(((double)x.y + h.qn2) - 2.0 * (num * h.c_dot));
}

// A Scorer accumulates one row (one lane) over the tiles of its block:
//   reset(acc); tile(acc, stage + lane, n, payload + c0 * bytes_per_chunk, metric) per tile; finish(...) -> key
template <int QT, int ND>
struct Scorer;

// 8-bit codes: 16 dims per chunk, payload = ND uint4 of digits per chunk
template <int ND>
struct Scorer<Q8, ND> {
    struct Acc { int a[ND]; };
    static __device__ __forceinline__ void reset(Acc &s) {
#pragma unroll
        for (int j = 0; j < ND; ++j) s.a[j] = 0;
    }
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
    static __device__ __forceinline__ void tile(Acc &s, const uint4 *stage, int lane, uint32_t n, const unsigned char *spq, int) {
        const uint4 *sd = stage + lane;
        const uint4 *dg = reinterpret_cast<const uint4 *>(spq);
        int b[ND]; // second accumulator set: shortens the dependent IDP chains
#pragma unroll
        for (int j = 0; j < ND; ++j) b[j] = 0;
        uint32_t c = 0;
        for (; c + 4 <= n; c += 4) {
            uint4 v0 = sd[(c + 0) * 32], v1 = sd[(c + 1) * 32], v2 = sd[(c + 2) * 32], v3 = sd[(c + 3) * 32];
            step(v0, dg + (c + 0) * ND, s.a);
            step(v1, dg + (c + 1) * ND, b);
            step(v2, dg + (c + 2) * ND, s.a);
            step(v3, dg + (c + 3) * ND, b);
        }
        for (; c < n; ++c) step(sd[c * 32], dg + c * ND, s.a);
#pragma unroll
        for (int j = 0; j < ND; ++j) s.a[j] += b[j];
    }
    static __device__ __forceinline__ float finish(const Acc &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        return finish_quant(a, h, digits_total<ND>(s.a), x);
    }
};

// 4-bit codes: 32 dims per chunk; byte = (even dim << 4) | odd dim.  Payload per chunk =
// ND uint4 for the even dims (applied to w & 0xF0F0F0F0, i.e. 16*u) + ND uint4 for the odd.
template <int ND>
struct Scorer<Q4, ND> {
    struct Acc { int hi[ND], lo[ND]; };
    static __device__ __forceinline__ void reset(Acc &s) {
#pragma unroll
        for (int j = 0; j < ND; ++j) s.hi[j] = s.lo[j] = 0;
    }
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
    static __device__ __forceinline__ void tile(Acc &s, const uint4 *stage, int lane, uint32_t n, const unsigned char *spq, int) {
        const uint4 *sd = stage + lane;
        const uint4 *dg = reinterpret_cast<const uint4 *>(spq);
        uint32_t c = 0;
        for (; c + 2 <= n; c += 2) {
            uint4 v0 = sd[(c + 0) * 32], v1 = sd[(c + 1) * 32];
            step(v0, dg + (c + 0) * 2 * ND, s.hi, s.lo);
            step(v1, dg + (c + 1) * 2 * ND, s.hi, s.lo);
        }
        for (; c < n; ++c) step(sd[c * 32], dg + c * 2 * ND, s.hi, s.lo);
    }
    static __device__ __forceinline__ float finish(const Acc &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        // hi accumulates 16 * u_even * W: I = hi/16 + lo (exact in double)
        return finish_quant(a, h, digits_total<ND>(s.hi) * 0.0625 + digits_total<ND>(s.lo), x);
    }
};

// 16-bit codes, stored as little-endian int16 of (u - 32768): 8 dims per chunk, payload =
// ND uint2 of digits per chunk.  |s16 * s8| <= 2^22, so int32 partials are flushed to
// double at the end of every tile (<= 32 chunks = 256 dims).
template <int ND>
struct Scorer<Q16, ND> {
    struct Acc { double I; };
    static __device__ __forceinline__ void reset(Acc &s) { s.I = 0.0; }
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
    static __device__ __forceinline__ void tile(Acc &s, const uint4 *stage, int lane, uint32_t n, const unsigned char *spq, int) {
        const uint4 *sd = stage + lane;
        const uint2 *dg = reinterpret_cast<const uint2 *>(spq);
        int acc[ND];
#pragma unroll
        for (int j = 0; j < ND; ++j) acc[j] = 0;
        uint32_t c = 0;
        for (; c + 4 <= n; c += 4) {
            uint4 v0 = sd[(c + 0) * 32], v1 = sd[(c + 1) * 32], v2 = sd[(c + 2) * 32], v3 = sd[(c + 3) * 32];
            step(v0, dg + (c + 0) * ND, acc);
            step(v1, dg + (c + 1) * ND, acc);
            step(v2, dg + (c + 2) * ND, acc);
            step(v3, dg + (c + 3) * ND, acc);
        }
        for (; c < n; ++c) step(sd[c * 32], dg + c * ND, acc);
        s.I += digits_total<ND>(acc);
    }
    static __device__ __forceinline__ float finish(const Acc &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        return finish_quant(a, h, s.I, x);
    }
};

// 32-bit float rows: payload = the query as float4 per chunk
template <int ND>
struct Scorer<F32, ND> {
    struct Acc { float a[4]; };
    static __device__ __forceinline__ void reset(Acc &s) { s.a[0] = s.a[1] = s.a[2] = s.a[3] = 0.f; }
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
    // the staged tile holds whole chunk groups: [group][lane][8 chunks, XOR-swizzled by lane % 8] (common.cuh); n % 8 == 0
    template <int METRIC>
    static __device__ __forceinline__ void loop(Acc &s, const uint4 *stage, int lane, uint32_t n, const float4 *sq) {
        const uint32_t sw = (uint32_t)lane & 7u;
        for (uint32_t c = 0; c < n; c += 8) {
            const uint4 *g = stage + ((size_t)(c >> 3) * 32 + lane) * 8;
            uint4 v[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) v[j] = g[j ^ sw];
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) step<METRIC>(v[j], sq[c + j], s.a);
        }
    }
    static __device__ __forceinline__ void tile(Acc &s, const uint4 *stage, int lane, uint32_t n, const unsigned char *spq, int metric) {
        const float4 *sq = reinterpret_cast<const float4 *>(spq);
        if (metric == COSINE) loop<COSINE>(s, stage, lane, n, sq);
        else loop<EUCLID>(s, stage, lane, n, sq);
    }
    static __device__ __forceinline__ float finish(const Acc &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        const float r = (s.a[0] + s.a[1]) + (s.a[2] + s.a[3]);
        if (a.metric == COSINE) return (h.zero_query || x.x == 0.f) ? 1.0f : -(r * (float)h.c_key) * x.x;
        return r;
    }
};

// 64-bit float rows: payload = the query as double2 per chunk
template <int ND>
struct Scorer<F64, ND> {
    struct Acc { double a[2]; };
    static __device__ __forceinline__ void reset(Acc &s) { s.a[0] = s.a[1] = 0.0; }
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
    static __device__ __forceinline__ void loop(Acc &s, const uint4 *stage, int lane, uint32_t n, const double2 *sq) {
        const uint32_t sw = (uint32_t)lane & 7u;
        for (uint32_t c = 0; c < n; c += 8) {
            const uint4 *g = stage + ((size_t)(c >> 3) * 32 + lane) * 8;
            uint4 v[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) v[j] = g[j ^ sw];
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) step<METRIC>(v[j], sq[c + j], s.a);
        }
    }
    static __device__ __forceinline__ void tile(Acc &s, const uint4 *stage, int lane, uint32_t n, const unsigned char *spq, int metric) {
        const double2 *sq = reinterpret_cast<const double2 *>(spq);
        if (metric == COSINE) loop<COSINE>(s, stage, lane, n, sq);
        else loop<EUCLID>(s, stage, lane, n, sq);
    }
    static __device__ __forceinline__ float finish(const Acc &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        const double r = s.a[0] + s.a[1];
        if (a.metric == COSINE) return (h.zero_query || x.x == 0.f) ? 1.0f : -((float)(r * h.c_key)) * x.x;
        return (float)r;
    }
};

// ------------------------------------------------------------------- finalisation (top-k)
constexpr int kFinalizeThreads = 512;
constexpr size_t kFinalizeStageBytes = 64 * 1024; // staging area of the fp64 re-score

// Fetches chunks [c0, c0 + nc) of the Kp candidates' rows into s_codes (row r at r * (SC + 1) uint4).  Gathered rows: every
// load is a DRAM round trip of its own, so a thread keeps 8 of them in flight.
template <int QT, int NT>
__device__ __forceinline__ void gather_slab(const uint4 *__restrict__ codes, uint32_t C, uint32_t c0, uint32_t nc, uint32_t SC,
                                            const uint32_t *s_slot, int Kp, uint4 *s_codes, int tid) {
    constexpr int PF = 8;
    const uint32_t total = (uint32_t)Kp * nc;
    for (uint32_t idx0 = tid; idx0 < total; idx0 += NT * PF) {
        uint4 v[PF];
        uint32_t at[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const uint32_t idx = idx0 + (uint32_t)u * NT;
            at[u] = 0xFFFFFFFFu;
            if (idx < total) {
                const uint32_t r = idx / nc, c = idx - r * nc;
                const uint32_t slot = s_slot[r];
                if (slot != 0xFFFFFFFFu) {
                    v[u] = __ldg(codes + chunk_at<QT>(slot, C, c0 + c));
                    at[u] = r * (SC + 1) + c;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (at[u] != 0xFFFFFFFFu) s_codes[at[u]] = v[u];
    }
}

// fp64 distances of up to NT / 2 candidates (s_slot[0 .. Kp), 0xFFFFFFFF = none) to the query q, bit for bit what the
// reference computes: decodeVector + dequantize (collection.go:768-794, quantization.go:25-36), then euclideanDistance
// (812-819) or angularDistance (821-832) -- sequential over the dimensions, un-fused multiply and add.  Organised so that
// only what MUST be sequential is:
//   fetch     the NT threads gather the candidates' chunks slab by slab into shared memory, 8 loads in flight per thread
//             (float rows keep their chunks in groups of 8: whole 128-byte lines per row);
//   products  every rounded product of the reference's loop body -- q_i x_i and x_i x_i (cosine), (q_i - x_i)^2 (euclid),
//             q_i q_i once per slab -- is independent of the running sums, so all threads compute them in parallel
//             (__dmul_rn / __dsub_rn: the same IEEE operations, never contracted);
//   chains    one thread per running sum (dot and m2 of a candidate; m1 = sum q_i^2 rides with candidate 0's) adds the
//             products in dimension order with __dadd_rn: a pure dependent-add chain out of shared memory, which is the
//             irreducible critical path (d adds).
// dequantize: 4/8-bit through a shared copy of the host-built table, 16-bit with the same three IEEE operations
// ((v / maxInt) * 2 - 1, quantization.go:34-35).  Used by finalize_kernel (the survivors of a top-k scan), rescore_kernel
// (candidate lists of the LSH index) and radius_exact_kernel.  s_out[r] = distance (NaN possible: math.Acos of a ratio > 1).
template <int QT>
__device__ __forceinline__ double staged_element(const uint4 *row, uint32_t e, const double *s_lut) {
    if (QT == Q4) {
        const uint32_t byte = reinterpret_cast<const unsigned char *>(row)[e >> 1];
        return s_lut[(e & 1u) ? (byte & 0x0Fu) : (byte >> 4)]; // even index: high nibble (collection.go:774-779)
    } else if (QT == Q8) {
        return s_lut[reinterpret_cast<const unsigned char *>(row)[e]];
    } else if (QT == Q16) {
        const uint32_t u = (uint32_t)reinterpret_cast<const unsigned short *>(row)[e] ^ 0x8000u; // stored centred
        return __dsub_rn(__dmul_rn(__ddiv_rn((double)u, 65535.0), 2.0), 1.0);
    } else if (QT == F32) {
        return (double)reinterpret_cast<const float *>(row)[e]; // widened, quantization.go:27-28
    } else {
        return reinterpret_cast<const double *>(row)[e];
    }
}


template <int QT, int METRIC, int NT, int ESMAX = 32>
__device__ void exact_staged(const uint4 *__restrict__ codes, const double *__restrict__ lut, uint32_t C, uint32_t dims,
                             const double *__restrict__ q, const uint32_t *s_slot, int Kp, unsigned char *stage,
                             size_t stage_bytes, double *s_out, int tid, long long *trace = nullptr) {
    long long t_chain = 0, t_bar = 0;
    constexpr int EPC = QT == Q4 ? 32 : QT == Q8 ? 16 : QT == Q16 ? 8 : QT == F32 ? 4 : 2;
    constexpr int LUTN = QT == Q4 ? 16 : QT == Q8 ? 256 : 0;
    constexpr uint32_t GR = QT >= F32 ? (uint32_t)kGroupChunks : 1u; // slabs of float rows hold whole chunk groups
    constexpr int NA = METRIC == COSINE ? 2 : 1;                       // running sums per candidate
    double *s_lut = reinterpret_cast<double *>(stage);
    // products: [Kp * NA][ESP] + (cosine) q_i^2 [ES]; ES dimensions per round, rows ESP = ES | 1 doubles apart so that the
    // chain threads (one row each) spread over the banks.  Half of a 64 KB stage, a quarter for the largest candidate sets.
    // When the chains occupy at most half of the threads (a 32-candidate set: 2 warps of 16), the other warps compute the
    // products of round r + 1 into a second buffer WHILE the chain warps add round r: the dependent fp64 adds (~40 cycles
    // each on this part) are then the only thing on the critical path.
    // chains: thread t < Kp * NA -> candidate t / NA, running sum t % NA; cosine: thread Kp * NA carries m1 = sum q_i^2 (a
    // thread of its own, in a warp of its own when Kp * NA fills whole warps: no divergence inside the chain warps)
    const bool m1_own = METRIC == COSINE && Kp * NA < NT; // else (the largest candidate sets) thread 0 carries m1 as a second sum
    const int nchains = Kp * NA + (m1_own ? 1 : 0);
    const int chainT = (nchains + 31) / 32 * 32; // threads of the warps that run chains
    const bool overlap = chainT * 2 <= NT;
    double *s_prod = reinterpret_cast<double *>(stage + LUTN * sizeof(double));
    const size_t prod_bytes = Kp > 128 ? 16 * 1024 : 32 * 1024;
    // ES = dimensions per round: a power of two <= 32, so that a warp's lanes map to (candidate, dimension) pairs with shifts
    uint32_t ES = (uint32_t)(prod_bytes / (overlap ? 2 : 1) / 8 / ((size_t)Kp * NA + 1));
    ES = ES > 1 ? ES - 1 : 1;
    uint32_t ES_log = 0;
    while ((2u << ES_log) <= (uint32_t)ESMAX && (2u << ES_log) <= ES) ++ES_log;
    ES = 1u << ES_log;
    const uint32_t ESP = ES | 1u;
    const size_t buf_doubles = (size_t)Kp * NA * ESP + ((ES + 1) & ~1u); // one buffer: the products, then q_i^2
    unsigned char *body = reinterpret_cast<unsigned char *>(s_prod + buf_doubles * (overlap ? 2 : 1));
    const size_t budget = stage_bytes - (size_t)(body - stage);
    // per chunk: Kp uint4 of codes + EPC doubles of the query; rows padded by one uint4 against bank conflicts
    uint32_t SC = (uint32_t)((budget - (size_t)Kp * 16) / ((size_t)Kp * 16 + EPC * 8));
    if (SC > C) SC = C;
    SC = SC / GR * GR;
    if (SC < GR) SC = GR;
    uint4 *s_codes = reinterpret_cast<uint4 *>(body);
    double *s_q = reinterpret_cast<double *>(body + (size_t)Kp * (SC + 1) * 16);
    for (int i = tid; i < LUTN; i += NT) s_lut[i] = lut[i];

    const bool chain = tid < Kp * NA;
    const int cr = tid / NA;
    const bool has_m1 = METRIC == COSINE && tid == (m1_own ? Kp * NA : 0);
    const bool live = chain && s_slot[chain ? cr : 0] != 0xFFFFFFFFu;
    double acc = 0.0, m1 = 0.0;
    for (uint32_t c0 = 0; c0 < C; c0 += SC) {
        const uint32_t nc = min(SC, C - c0);
        __syncthreads(); // previous slab fully consumed (and the table written)
        gather_slab<QT, NT>(codes, C, c0, nc, SC, s_slot, Kp, s_codes, tid);
        for (uint32_t e = tid; e < nc * EPC; e += NT) {
            const uint32_t i = c0 * EPC + e;
            s_q[e] = i < dims ? q[i] : 0.0;
        }
        __syncthreads();
        const uint32_t i_slab = c0 * EPC;
        if (trace && tid == 0 && c0 == 0) trace[6] = clock64();
        if (i_slab >= dims) break; // padding chunks only (uniform)
        const uint32_t ne_slab = min(nc * (uint32_t)EPC, dims - i_slab); // real dimensions in this slab
        // ---- rounds of ES dimensions: products (parallel), then the chains (one thread per running sum, dimension order)
        auto produce = [&](uint32_t e0, uint32_t ne, double *buf, uint32_t pt, uint32_t pn) {
            // pt / pn: this thread's index among the pn producing threads (whole warps).  A warp covers 32 / ES candidates
            // per pass: lane = (candidate offset, dimension)
            double *bq = buf + (size_t)Kp * NA * ESP;
            const uint32_t e = pt & (ES - 1);
            const uint32_t cpp = pn >> ES_log; // candidates per pass of all producing threads
            for (uint32_t r = pt >> ES_log; r < (uint32_t)Kp; r += cpp) {
                if (e >= ne || s_slot[r] == 0xFFFFFFFFu) continue;
                const double x = staged_element<QT>(s_codes + (size_t)r * (SC + 1), e0 + e, s_lut);
                const double qi = s_q[e0 + e];
                if (METRIC == COSINE) {
                    buf[((size_t)r * 2 + 0) * ESP + e] = __dmul_rn(qi, x); // dot += query[i] * vec[i]   (collection.go:824)
                    buf[((size_t)r * 2 + 1) * ESP + e] = __dmul_rn(x, x);  // m2 += vec[i] * vec[i]     (826)
                } else {
                    const double diff = __dsub_rn(qi, x);                   // diff := query[i] - vec[i]  (815)
                    buf[(size_t)r * ESP + e] = __dmul_rn(diff, diff);       // sum += diff * diff         (816)
                }
            }
            if (METRIC == COSINE)
                for (uint32_t e = pt; e < ne; e += pn) bq[e] = __dmul_rn(s_q[e0 + e], s_q[e0 + e]); // m1 += query[i] * query[i] (825)
        };
        auto consume = [&](uint32_t ne, const double *buf) {
            // all products of the round into registers first (independent loads), then the dependent adds: the critical
            // path is one shared-memory latency plus ne fp64 add latencies
            if (live) {
                const double *p = buf + (size_t)tid * ESP;
                double v[ESMAX];
#pragma unroll
                for (uint32_t e = 0; e < (uint32_t)ESMAX; ++e)
                    if (e < ES) v[e] = e < ne ? p[e] : 0.0;
#pragma unroll
                for (uint32_t e = 0; e < (uint32_t)ESMAX; ++e)
                    if (e < ne) acc = __dadd_rn(acc, v[e]);
            }
            if (has_m1) {
                const double *p = buf + (size_t)Kp * NA * ESP;
                double v[ESMAX];
#pragma unroll
                for (uint32_t e = 0; e < (uint32_t)ESMAX; ++e)
                    if (e < ES) v[e] = e < ne ? p[e] : 0.0;
#pragma unroll
                for (uint32_t e = 0; e < (uint32_t)ESMAX; ++e)
                    if (e < ne) m1 = __dadd_rn(m1, v[e]);
            }
        };
        if (!overlap) {
            for (uint32_t e0 = 0; e0 < ne_slab; e0 += ES) {
                const uint32_t ne = min(ES, ne_slab - e0);
                produce(e0, ne, s_prod, (uint32_t)tid, (uint32_t)NT);
                __syncthreads();
                consume(ne, s_prod);
                __syncthreads();
            }
        } else {
            const bool producer = tid >= chainT;
            const uint32_t pt = (uint32_t)(tid - chainT), pn = (uint32_t)(NT - chainT);
            if (producer) produce(0, min(ES, ne_slab), s_prod, pt, pn);
            __syncthreads();
            uint32_t b = 0;
            for (uint32_t e0 = 0; e0 < ne_slab; e0 += ES, b ^= 1u) {
                const uint32_t ne = min(ES, ne_slab - e0);
                const long long ta = trace ? clock64() : 0;
                if (!producer) consume(ne, s_prod + buf_doubles * b);
                else if (e0 + ES < ne_slab) produce(e0 + ES, min(ES, ne_slab - e0 - ES), s_prod + buf_doubles * (b ^ 1u), pt, pn);
                const long long tb = trace ? clock64() : 0;
                __syncthreads();
                if (trace) { t_chain += tb - ta; t_bar += clock64() - tb; }
            }
        }
    }
    // ---- the running sums meet: s_prod is free now
    __syncthreads();
    if (trace && tid == 0) trace[7] = (t_chain << 32) | (t_bar & 0xFFFFFFFFll);
    double *s_m1 = s_prod + buf_doubles - 1;
    if (chain) s_prod[tid] = acc;
    if (has_m1) *s_m1 = m1;
    __syncthreads();
    if (tid < Kp && s_slot[tid] != 0xFFFFFFFFu) {
        double d;
        if (METRIC == COSINE) {
            const double dot = s_prod[2 * tid], m2 = s_prod[2 * tid + 1], mm1 = *s_m1;
            if (mm1 == 0.0 || m2 == 0.0) d = 1.0; // collection.go:828-830
            else {
                const double r = __ddiv_rn(dot, __dmul_rn(__dsqrt_rn(mm1), __dsqrt_rn(m2)));
                d = __ddiv_rn(go_acos(r), 3.141592653589793); // math.Acos(r > 1) = NaN
            }
        } else {
            d = __dsqrt_rn(s_prod[tid]);
        }
        s_out[tid] = d;
    }
    __syncthreads();
}

// Throughput form of the same arithmetic, for gathers of many candidates (szg_rescore of long lists, radius hits): thread r
// runs candidate r's whole loop body -- decode, the rounded products, the sequential adds -- out of the staged slab.  With
// a warp-wide fp64 instruction occupying its pipe for ~16 cycles, four instructions per dimension keep the pipe busy from a
// single warp's dependent chain, so nothing is gained by splitting products and sums (exact_staged does that for the
// latency of FEW candidates); two CTAs per SM alternate between fetching and computing.  m1 = sum q_i^2 (cosine) is the
// same for every candidate of a query: the caller supplies it (computed once on the host with the same IEEE operations).
template <int QT, int METRIC, int NT>
__device__ void exact_stream(const uint4 *__restrict__ codes, const double *__restrict__ lut, uint32_t C, uint32_t dims,
                             const double *__restrict__ q, double m1, const uint32_t *s_slot, int Kp, unsigned char *stage,
                             size_t stage_bytes, double *s_out, int tid) {
    constexpr int EPC = QT == Q4 ? 32 : QT == Q8 ? 16 : QT == Q16 ? 8 : QT == F32 ? 4 : 2;
    constexpr int LUTN = QT == Q4 ? 16 : QT == Q8 ? 256 : 0;
    constexpr uint32_t GR = QT >= F32 ? (uint32_t)kGroupChunks : 1u;
    double *s_lut = reinterpret_cast<double *>(stage);
    unsigned char *body = stage + LUTN * sizeof(double);
    const size_t budget = stage_bytes - LUTN * sizeof(double);
    uint32_t SC = (uint32_t)((budget - (size_t)Kp * 16) / ((size_t)Kp * 16 + EPC * 8));
    if (SC > C) SC = C;
    SC = SC / GR * GR;
    if (SC < GR) SC = GR;
    uint4 *s_codes = reinterpret_cast<uint4 *>(body);
    double *s_q = reinterpret_cast<double *>(body + (size_t)Kp * (SC + 1) * 16);
    for (int i = tid; i < LUTN; i += NT) s_lut[i] = lut[i];
    const bool mine = tid < Kp && s_slot[tid < Kp ? tid : 0] != 0xFFFFFFFFu;
    double a0 = 0.0, a1 = 0.0; // cosine: dot, m2; euclid: sum
    for (uint32_t c0 = 0; c0 < C; c0 += SC) {
        const uint32_t nc = min(SC, C - c0);
        __syncthreads();
        gather_slab<QT, NT>(codes, C, c0, nc, SC, s_slot, Kp, s_codes, tid);
        for (uint32_t e = tid; e < nc * EPC; e += NT) {
            const uint32_t i = c0 * EPC + e;
            s_q[e] = i < dims ? q[i] : 0.0;
        }
        __syncthreads();
        const uint32_t i_slab = c0 * EPC;
        if (i_slab >= dims) break;
        const uint32_t ne = min(nc * (uint32_t)EPC, dims - i_slab);
        if (mine) {
            const uint4 *row = s_codes + (size_t)tid * (SC + 1);
#pragma unroll 4
            for (uint32_t e = 0; e < ne; ++e) {
                const double x = staged_element<QT>(row, e, s_lut);
                const double qi = s_q[e];
                if (METRIC == COSINE) {
                    a0 = __dadd_rn(a0, __dmul_rn(qi, x)); // dot += query[i] * vec[i]   (collection.go:824)
                    a1 = __dadd_rn(a1, __dmul_rn(x, x));  // m2 += vec[i] * vec[i]     (826)
                } else {
                    const double diff = __dsub_rn(qi, x);  // (815)
                    a0 = __dadd_rn(a0, __dmul_rn(diff, diff));
                }
            }
        }
    }
    if (mine) {
        double d;
        if (METRIC == COSINE) {
            if (m1 == 0.0 || a1 == 0.0) d = 1.0; // collection.go:828-830
            else {
                const double r = __ddiv_rn(a0, __dmul_rn(__dsqrt_rn(m1), __dsqrt_rn(a1)));
                d = __ddiv_rn(go_acos(r), 3.141592653589793);
            }
        } else {
            d = __dsqrt_rn(a0);
        }
        s_out[tid] = d;
    }
    __syncthreads();
}

// pool[0 .. K') holds the K' best (surrogate, slot) keys, ascending.  Re-scores them in
// fp64, orders by (distance, lexicographic id), writes min(k, #) results and certifies.
template <int QT>
__device__ void finalize_topk(const FinalizeArgs &a, const double *q, const PQHeader *hdr, unsigned long long *out_ids,
                              double *out_dist, uint32_t *out_n, uint32_t *out_flags, uint32_t *done_cnt,
                              const unsigned long long *pool, int Kp, double *s_ex, unsigned long long *s_id,
                              unsigned char *stage, int tid) {
    __shared__ double s_dk;
    __shared__ uint32_t s_slot[32 * kMaxListE];
    if (tid == 0) s_dk = 0.0;
    unsigned long long id = 0; // fetched now: the load's latency hides behind the exact pass
    if (tid < Kp && pool[tid] != kNoKey) id = __ldg(a.ids + (uint32_t)pool[tid]);
    if (!(a.flags & 1u)) {
        if (tid < Kp) s_slot[tid] = pool[tid] == kNoKey ? 0xFFFFFFFFu : (uint32_t)pool[tid];
        __syncthreads();
        if (a.metric == COSINE)
            exact_staged<QT, COSINE, kFinalizeThreads>(a.codes, a.lut, a.C, a.dims, q, s_slot, Kp, stage, kFinalizeStageBytes, s_ex, tid,
                                                       blockIdx.x == 0 ? a.trace : nullptr);
        else
            exact_staged<QT, EUCLID, kFinalizeThreads>(a.codes, a.lut, a.C, a.dims, q, s_slot, Kp, stage, kFinalizeStageBytes, s_ex, tid,
                                                       blockIdx.x == 0 ? a.trace : nullptr);
    }
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[3] = clock64();
    bool valid = false;
    double d = 0.0;
    if (tid < Kp) {
        unsigned long long key = pool[tid];
        if (key != kNoKey) {
            if (a.flags & 1u) d = key_to_distance(a.metric, key_to_float((uint32_t)(key >> 32)));
            else d = s_ex[tid];
            valid = (d == d); // NaN is never returned
        }
    }
    __syncthreads();
    if (tid < Kp) {
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
            out_ids[rank] = id;
            out_dist[rank] = d;
        }
        if ((uint32_t)rank + 1 == a.k) s_dk = d;
    }
    __syncthreads();
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[4] = clock64();
    if (tid == 0) {
        uint32_t n = (uint32_t)cnt < a.k ? (uint32_t)cnt : a.k;
        *out_n = n;
        // Certification.  Every row outside the candidate set has a surrogate key >= s_last (the
        // worst candidate's), hence a true key >= s_last - e_abs - e_rel |s_last| (PQHeader bound).
        // The result is certain when that lower bound still exceeds the true key of the k-th
        // result.  A candidate set that is not full holds every live row: always certain.
        bool uncertain = false;
        if (nfull == Kp) {
            if ((uint32_t)cnt < a.k) uncertain = true; // NaN candidates displaced real ones
            else {
                const double s_last = (double)key_to_float((uint32_t)(pool[Kp - 1] >> 32));
                const double lower = s_last - hdr->e_abs - hdr->e_rel * fabs(s_last);
                const double t_k = a.metric == COSINE ? -cospi(s_dk) : s_dk * s_dk;
                uncertain = !(lower > t_k);
            }
        }
        *out_flags = uncertain ? 1u : 0u;
    }
    if (done_cnt) {
        // sharded search: the outputs above went to the root device's gather buffer; publish them (every writer fences, the
        // barrier orders the fences before thread 0's release) and tell the root's merge kernel
        __threadfence_system();
        __syncthreads();
        if (tid == 0) atomicAdd_system(done_cnt, 1u);
    }
}

template <int QT, int MODE>
__global__ void __launch_bounds__(kFinalizeThreads) finalize_kernel(const FinalizeArgs a) {
    constexpr int E = 1 << MODE;
    constexpr int Kp = 32 * E;
    constexpr int NW = kFinalizeThreads / 32;
    constexpr int PF = 8; // candidate keys fetched ahead per lane: the merge is load-latency-bound otherwise
    extern __shared__ __align__(16) unsigned char fsm[];
    unsigned long long *pool = reinterpret_cast<unsigned long long *>(fsm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qi = blockIdx.x;
    const unsigned long long *cand = a.cand + (size_t)qi * a.nlists * Kp;
    const bool tr = a.trace && qi == 0 && tid == 0;
    grid_dependency_wait(); // launched as a programmatic dependent of the scan (launch_dependent): its lists are complete from here on
    if (tr) a.trace[0] = clock64();
    WarpList<E> list;
    list.init();
    const uint32_t total = a.nlists * Kp;
    for (uint32_t base = warp * 32; base < total; base += kFinalizeThreads * PF) {
        unsigned long long v[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const uint32_t i = base + u * kFinalizeThreads + lane;
            v[u] = i < total ? __ldg(cand + i) : kNoKey;
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (E == 1) {
                // every 32 keys are one SORTED list of a scan CTA / row range (lane = rank): six merge steps instead of a
                // sort of the batch
                if (__ballot_sync(0xffffffffu, v[u] < list.thr)) list.merge_sorted(v[u], lane);
            } else list.offer(v[u], lane);
        }
    }
    if (tr) a.trace[1] = clock64();
    block_merge<E>(list, pool, tid, lane, warp, NW);
    if (tr) a.trace[2] = clock64();
    double *s_ex = reinterpret_cast<double *>(pool + NW * Kp);
    unsigned long long *s_id = reinterpret_cast<unsigned long long *>(s_ex + Kp);
    unsigned char *stage = reinterpret_cast<unsigned char *>(s_id + Kp);
    finalize_topk<QT>(a, a.queries + (size_t)qi * a.dims,
                      reinterpret_cast<const PQHeader *>(a.pq + (size_t)qi * a.pq_stride), a.out_ids + (size_t)qi * a.k,
                      a.out_dist + (size_t)qi * a.k, a.out_n + qi, a.out_flags + qi, a.done_cnt ? a.done_cnt + qi : nullptr,
                      pool, Kp, s_ex, s_id, stage, tid);
    if (tr) a.trace[5] = clock64();
}
inline size_t finalize_smem_bytes(int mode) {
    const size_t Kp = 32u << mode;
    return (size_t)(kFinalizeThreads / 32) * Kp * 8 + Kp * 16 + kFinalizeStageBytes;
}

// ------------------------------------------------------------------------ the scan kernel
// per-warp streaming state of the tile ring: which (query, block, tile) each stage holds
struct RingMeta {
    uint32_t blk[kMaxStages];   // 0xFFFFFFFF = no more tiles
    uint32_t tile[kMaxStages];  // tile index | query index << 16
    uint32_t live[kMaxStages];  // live & filter word of that block
};

// One launch serves all nq queries of a call: every warp walks the sequence
// (query 0: its blocks), (query 1: its blocks), ... without any CTA-level synchronisation, so
// the copy pipeline never drains between queries and there is no per-query launch gap.  Each
// (query, block) pair is still streamed from HBM once: single-query GEMV semantics.  The
// prepared query (digits) is read through L1 (warp-uniform 128-bit loads): a few KB per
// query, resident next to the streaming traffic, which bypasses L1 via the bulk copies.
template <int QT, int MODE, int ND, bool SHARE = false>
__global__ void __launch_bounds__(kMaxScanWarps * 32, 1) scan_kernel(const ScanArgs a) {
    constexpr int E = (MODE == MODE_RADIUS) ? 1 : (1 << MODE);
    constexpr int Kp = 32 * E;
    grid_launch_dependents(); // a finalize_kernel launched as programmatic dependent may queue behind this grid's CTAs now
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar[kMaxScanWarps][kMaxStages];
    __shared__ RingMeta s_meta[kMaxScanWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const uint32_t S = a.stages, Ct = a.Ct, C = a.C;

    if (lane == 0) {
        for (uint32_t s = 0; s < S; ++s) mbar_init(&s_bar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // CTA-shared bound of short scans (see the use below): (query << 32 | key of the warp's mth-best row), 4 query slots
    __shared__ unsigned long long s_pub[SHARE ? 4 : 1][kMaxScanWarps];
    if (SHARE) {
        if (lane < 4) s_pub[lane][warp] = ~0ull;
        __syncthreads();
    }

    // ---- streaming: this warp's blocks are gw, gw + stride, ...; tiles of Ct chunks
    const uint32_t stride = gridDim.x * nwarps;
    const uint32_t gw = blockIdx.x * nwarps + warp;
    const uint32_t T = (C + Ct - 1) / Ct;
    const uint32_t stage_bytes = Ct * 512u;
    // When a warp sees only a few dozen blocks per query, most of its selection work is the warm-up of its own list
    // (every one of the grid's lists starts empty: measured 35 % of the instructions at 13 blocks per warp, cfg2).
    // The warps of a CTA then share a bound: each publishes the key of its mth-best row so far, mth = ceil(Kp / warps);
    // once all have, warps * mth >= Kp rows of this CTA lie at or below the largest published key, so no row above it
    // can be among the query's best Kp.  Warps move from query to query independently: entries carry the query index.
    const uint32_t mth = ((uint32_t)Kp + (uint32_t)nwarps - 1) / (uint32_t)nwarps;
    constexpr bool share_bound = SHARE; // a separate instantiation: the long-scan kernel keeps its code and registers
    unsigned char *ring = smem + (size_t)warp * S * stage_bytes;
    RingMeta &meta = s_meta[warp];
    uint64_t *bars = s_bar[warp];

    // issue iterator (warp-uniform): next (query, block, tile) to fetch
    uint32_t iq = 0, iblk = gw, itile = 0, ilive = 0;
    auto seek_live = [&]() { // advance (iq, iblk) to the next block with a live, unfiltered row
        while (iq < a.nq) {
            while (iblk < a.nblk) {
                uint32_t w = __ldg(a.live + iblk);
                if (a.mask) w &= __ldg(a.mask + iblk);
                if (w) { ilive = w; return; }
                iblk += stride;
            }
            ++iq;
            iblk = gw;
        }
    };
    seek_live();
    auto issue = [&](uint32_t s) { // all lanes run it (uniform control), lane 0 talks to the hardware
        if (iq >= a.nq) {
            if (lane == 0) meta.blk[s] = 0xFFFFFFFFu;
            return;
        }
        const uint32_t c0 = itile * Ct, n = min(Ct, C - c0);
        if (lane == 0) {
            meta.blk[s] = iblk; meta.tile[s] = itile | (iq << 16); meta.live[s] = ilive;
            mbar_expect_tx(&bars[s], n * 512u);
            bulk_g2s(ring + (size_t)s * stage_bytes, a.codes + ((size_t)iblk * C + c0) * 32, n * 512u, &bars[s]);
        }
        if (++itile == T) { itile = 0; iblk += stride; seek_live(); }
    };
    for (uint32_t s = 0; s < S; ++s) issue(s);
    __syncwarp();

    WarpList<E> list;
    list.init();
    using Sc = Scorer<QT, ND>;
    typename Sc::Acc acc;
    Sc::reset(acc);
    float2 aux = make_float2(0.f, 0.f);
    uint32_t phases = 0, cs = 0, cq = 0;
    const uint32_t bpc = (uint32_t)pq_bytes_per_chunk(QT, ND);
    // prepared query cq (header + payload).  Normally a per-warp private copy in shared memory
    // (digit loads are then warp-uniform LDS.128 broadcasts, 1 wavefront each; as global loads
    // they cost ~4x the L1 data-pipe time and made that pipe the limiter); warps change query
    // independently, so the copy needs no CTA-level synchronisation.
    const unsigned char *gpq = a.pq;
    unsigned char *spq = a.pq_smem_off ? smem + a.pq_smem_off + (size_t)warp * a.pq_stride : nullptr;
    auto load_pq = [&]() {
        if (!spq) return;
        const uint4 *src = reinterpret_cast<const uint4 *>(gpq);
        uint4 *dst = reinterpret_cast<uint4 *>(spq);
        const uint32_t n16 = (uint32_t)(a.pq_stride / 16);
        for (uint32_t i = lane; i < n16; i += 32) dst[i] = __ldg(src + i);
        __syncwarp();
    };
    load_pq();
    const unsigned char *pq = spq ? spq : gpq;
    auto flush = [&]() {                                  // this warp's list of query cq -> finalize_kernel
        if (MODE != MODE_RADIUS) {
            unsigned long long *dst = a.cand + ((size_t)cq * stride + gw) * Kp + (size_t)lane * E;
#pragma unroll
            for (int e = 0; e < E; ++e) dst[e] = list.v[e];
            list.init();
        }
    };
    while (true) {
        const uint32_t blk = meta.blk[cs];
        if (blk == 0xFFFFFFFFu) break;
        const uint32_t tq = meta.tile[cs], lv = meta.live[cs];
        const uint32_t tile = tq & 0xFFFFu, qn = tq >> 16;
        if (cq < qn) {
            while (cq < qn) { flush(); ++cq; gpq += a.pq_stride; }
            load_pq();
            pq = spq ? spq : gpq;
        }
        const uint32_t slot = blk * 32 + lane;
        if (tile == 0) {
            Sc::reset(acc);
            aux = load_aux(a, slot);
        }
        mbar_wait(&bars[cs], (phases >> cs) & 1u);
        phases ^= 1u << cs;
        const uint32_t c0 = tile * Ct, n = min(Ct, C - c0);
        Sc::tile(acc, reinterpret_cast<const uint4 *>(ring + (size_t)cs * stage_bytes), lane, n,
                         pq + sizeof(PQHeader) + (size_t)c0 * bpc, (int)a.metric);
        if (tile == T - 1) {
            const PQHeader &h = *reinterpret_cast<const PQHeader *>(pq);
            const float key = Sc::finish(acc, a, h, aux);
            const bool ok = (lv >> lane) & 1u;
            if (MODE == MODE_RADIUS) {
                const bool pass = ok && key <= (float)h.radius_key;
                const unsigned m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(a.rad_count, (uint32_t)__popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (pass) {
                        uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < a.rad_cap) a.rad_slots[pos] = slot;
                    }
                }
            } else {
                if (share_bound) {
                    const unsigned long long e = lane < nwarps ? s_pub[cq & 3u][lane] : ((unsigned long long)cq << 32);
                    if (__all_sync(0xffffffffu, (uint32_t)(e >> 32) == cq)) {
                        const unsigned long long b = ((unsigned long long)__reduce_max_sync(0xffffffffu, (uint32_t)e) << 32) | 0xFFFFFFFFull;
                        if (b < list.thr) list.thr = b; // keys equal to the bound still pass
                    }
                }
                list.offer(ok ? make_key64(key, slot) : kNoKey, lane);
                if (share_bound && (uint32_t)lane == (mth - 1) / E) {
                    unsigned long long mv = list.v[0];
#pragma unroll
                    for (int e2 = 1; e2 < E; ++e2)
                        if ((mth - 1) % E == (uint32_t)e2) mv = list.v[e2];
                    if (mv != kNoKey) s_pub[cq & 3u][warp] = ((unsigned long long)cq << 32) | (mv >> 32);
                }
            }
        }
        __syncwarp(); // every lane is done reading stage cs (and its meta) before it is refilled
        issue(cs);
        __syncwarp(); // meta written by lane 0 is visible to the warp
        cs = (cs + 1 == S) ? 0 : cs + 1;
    }
    if (MODE != MODE_RADIUS && a.cta_merge) {
        // a.nq == 1: every warp of the CTA ends here.  The rings are idle now; their memory holds the merge pool.
        __syncthreads();
        unsigned long long *pool = reinterpret_cast<unsigned long long *>(smem);
        block_merge<E>(list, pool, tid, lane, warp, nwarps);
        unsigned long long *dst = a.cand + (size_t)blockIdx.x * Kp;
        for (int i = tid; i < Kp; i += (int)blockDim.x) dst[i] = pool[i];
        return;
    }
    while (cq < a.nq) { flush(); ++cq; } // remaining queries (also the ones this warp had no block for)
}

// ---- host-side launch plan: shared memory carve-up for (quantization, C, mode, warps, stages)
struct ScanPlan {
    uint32_t Ct, stages, warps, pq_smem_off;
    size_t smem;
};
inline bool scan_plan(uint32_t C, uint32_t warps, uint32_t stages, uint32_t max_tile_chunks, size_t pq_stride,
                      size_t smem_limit, ScanPlan *p, uint32_t granule = 1) {
    // granule: tiles hold whole multiples of it (8 for float rows, whose chunks are grouped by 8; C is a multiple of it)
    if ((size_t)warps * stages * 512 * granule > smem_limit) return false;
    // per-warp private copies of the prepared query, unless they would squeeze the rings below 2 KB tiles
    const size_t pq_all = (size_t)warps * pq_stride;
    const uint32_t want = max_tile_chunks < C ? max_tile_chunks : C;
    uint32_t floor_ct = want < 4 ? want : 4;
    if (floor_ct < granule) floor_ct = granule;
    const bool pq_in_smem = pq_all + (size_t)warps * stages * 512 * floor_ct <= smem_limit;
    const size_t ring_budget = smem_limit - (pq_in_smem ? pq_all : 0);
    size_t per_stage = ring_budget / ((size_t)warps * stages) / 512;
    uint32_t Ct = (uint32_t)(per_stage < max_tile_chunks ? per_stage : max_tile_chunks);
    if (Ct > C) Ct = C;
    if (Ct > (uint32_t)kMaxTileChunks) Ct = kMaxTileChunks;
    Ct = Ct / granule * granule;
    if (Ct < granule) {
        if (per_stage < granule) return false;
        Ct = granule;
    }
    // equalise tiles: the smallest Ct that keeps the same number of tiles per block
    const uint32_t T = (C + Ct - 1) / Ct;
    Ct = ((C + T - 1) / T + granule - 1) / granule * granule;
    p->Ct = Ct; p->stages = stages; p->warps = warps;
    const size_t rings = (size_t)warps * stages * Ct * 512;
    p->pq_smem_off = pq_in_smem ? (uint32_t)rings : 0; // rings start at 0, so a non-zero offset doubles as the flag
    p->smem = rings + (pq_in_smem ? pq_all : 0);
    return true;
}

// host-side launcher, instantiated per quantization in scan_<qt>.cu.  Float rows have no digits:
// only the ND = 3 instantiation exists for them.
template <int QT, int ND>
cudaError_t launch_scan_nd(int mode, int grid, int threads, size_t smem, cudaStream_t st, const ScanArgs &a) {
    switch (mode) {
    case 0:
        // short per-warp streams (< 96 blocks per warp and query): the variant with the CTA-shared bound
        if (a.nblk / ((uint32_t)grid * (uint32_t)(threads / 32)) < 96) scan_kernel<QT, 0, ND, true><<<grid, threads, smem, st>>>(a);
        else scan_kernel<QT, 0, ND><<<grid, threads, smem, st>>>(a);
        break;
    case 1: scan_kernel<QT, 1, ND><<<grid, threads, smem, st>>>(a); break;
    case 2: scan_kernel<QT, 2, ND><<<grid, threads, smem, st>>>(a); break;
    case 3: scan_kernel<QT, 3, ND><<<grid, threads, smem, st>>>(a); break;
    case MODE_RADIUS: scan_kernel<QT, MODE_RADIUS, ND><<<grid, threads, smem, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
template <int QT>
cudaError_t launch_scan_t(int mode, int nd, int grid, int threads, size_t smem, cudaStream_t st, const ScanArgs &a) {
    if (QT <= Q16 && nd == 2) return launch_scan_nd<QT, (QT <= Q16 ? 2 : 3)>(mode, grid, threads, smem, st, a);
    return launch_scan_nd<QT, 3>(mode, grid, threads, smem, st, a);
}

template <int QT>
cudaError_t launch_finalize_t(int mode, uint32_t nq, cudaStream_t st, const FinalizeArgs &a) {
    const size_t smem = finalize_smem_bytes(mode);
    switch (mode) {
    case 0: return launch_dependent(finalize_kernel<QT, 0>, nq, kFinalizeThreads, smem, st, a);
    case 1: return launch_dependent(finalize_kernel<QT, 1>, nq, kFinalizeThreads, smem, st, a);
    case 2: return launch_dependent(finalize_kernel<QT, 2>, nq, kFinalizeThreads, smem, st, a);
    case 3: return launch_dependent(finalize_kernel<QT, 3>, nq, kFinalizeThreads, smem, st, a);
    default: return cudaErrorInvalidValue;
    }
}

template <int QT>
cudaError_t scan_attr_t(size_t max_smem) {
    cudaError_t e;
#define SZG_FATTR(M)                                                                                             \
    e = cudaFuncSetAttribute(finalize_kernel<QT, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                             (int)finalize_smem_bytes(M));                                                        \
    if (e != cudaSuccess) return e;
    SZG_FATTR(0) SZG_FATTR(1) SZG_FATTR(2) SZG_FATTR(3)
#undef SZG_FATTR
#define SZG_ATTR(M)                                                                                              \
    e = cudaFuncSetAttribute(scan_kernel<QT, M, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem); \
    if (e != cudaSuccess) return e;                                                                              \
    if (M == 0) {                                                                                                \
        e = cudaFuncSetAttribute(scan_kernel<QT, 0, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem); \
        if (e != cudaSuccess) return e;                                                                          \
        if (QT <= Q16) {                                                                                         \
            e = cudaFuncSetAttribute(scan_kernel<QT, 0, (QT <= Q16 ? 2 : 3), true>,                              \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);                \
            if (e != cudaSuccess) return e;                                                                      \
        }                                                                                                        \
    }                                                                                                            \
    if (QT <= Q16) {                                                                                             \
        e = cudaFuncSetAttribute(scan_kernel<QT, M, (QT <= Q16 ? 2 : 3)>,                                        \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);                    \
        if (e != cudaSuccess) return e;                                                                          \
    }
    SZG_ATTR(0) SZG_ATTR(1) SZG_ATTR(2) SZG_ATTR(3) SZG_ATTR(MODE_RADIUS)
#undef SZG_ATTR
    return cudaSuccess;
}

} // namespace szg
