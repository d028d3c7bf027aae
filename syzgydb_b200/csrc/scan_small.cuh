// scan_small.cuh -- the top-k scan (k <= 24) for quantized rows of 2 .. 48 chunks of 16 bytes (at most 768 bytes): BASELINE
// configs[0] 100k x 384 8-bit, configs[1] 1M x 128 4-bit (64-byte rows), configs[3] 10M x 768 8-bit.
//
// Why a second kernel: the general kernel (scan_impl.cuh) was built for one query streaming gigabytes.  On short rows a
// 32-row block is only C * 512 bytes and it spends ~420 instructions per block of which the arithmetic is ~150 (ring
// stages, barriers, per-tile bookkeeping, run-time chunk loops, per-chunk digit loads); with every warp of the grid walking
// every query a query ends after a dozen blocks per warp (2368 candidate lists per query, each warmed up from empty); and
// when a call brings several queries each of them streams the collection from HBM alone.
//
// What is different here:
//  * chunk count, digit count and rows per lane are compile-time: everything is unrolled, no ring, no barriers in the
//    loop -- a lane loads its row's chunks with plain coalesced 128-bit loads, 16 uint4 in flight per lane;
//  * the query's digits are read once per chunk and applied to the U rows a lane holds; from 32 chunks per row on they
//    come from constant memory (warp-uniform LDC through the constant cache instead of the L1 data pipe);
//  * queries are dealt to CTAs: a launch of nq queries is cut in work items (query, row part), an item is processed by
//    ONE CTA whose warps stride over the part's blocks, so a query ends in P x warps lists (16 when nq >= SM count)
//    and a warp sees hundreds of blocks per query; the warps on a query share one bound (the argument of the general
//    kernel's short-scan variant).  P = SM count for a single query: the decomposition the general kernel has;
//  * CTAs of different queries sweep the collection in the same order at the same pace: a row fetched from HBM for one
//    query is an L2 hit for the others; with two warp groups per CTA (a.wgroups), each on another query of the same
//    part, also an L1 hit (cfg4, 32 queries per launch: 7.76 GB of DRAM traffic for 245.8 GB of scans).
// Keys are computed by the same Scorer<QT, ND>::step / finish as the general kernel: identical surrogate keys,
// identical candidate semantics, finalize_kernel unchanged.  Measurements and the variants that were tried and dropped:
// DESIGN.md section 5a, profiles/r01b_scan_small_vs_general.log.
#pragma once
#include <mutex>

#include "scan_impl.cuh"

namespace szg {

#ifndef SZG_SMALL_WARPS
#define SZG_SMALL_WARPS 16 // warps per CTA (tuning builds: make EXTRA=-DSZG_SMALL_WARPS=24)
#endif
#ifndef SZG_SMALL_PCBIG
#define SZG_SMALL_PCBIG 8  // chunks per piece for rows above 8 chunks
#endif
constexpr int kSmallWarps = SZG_SMALL_WARPS;
// Prepared queries of one launch in constant memory (bank 3): the digits are warp-uniform, so read from there they arrive
// in uniform registers through the constant cache (SASS: LDCU.64 + IDP.4A with a UR operand) and leave the L1 data pipe --
// the limiter of this kernel -- to the row loads.  One window per translation unit (quantization) and device; launches
// that use it are chained by an event (scan_small launches fill the GPU and would not overlap anyway).
constexpr uint32_t kConstSlots = 3840; // uint4 slots: 60 KB
__constant__ uint4 c_pq[kConstSlots];

template <int QT, int ND>
struct SmallOps;
template <int ND>
struct SmallOps<Q8, ND> {
    using Dig = uint4;
    static constexpr int DPC = ND;
    struct A { int a[ND]; };
    static __device__ __forceinline__ void reset(A &s) {
#pragma unroll
        for (int j = 0; j < ND; ++j) s.a[j] = 0;
    }
    static __device__ __forceinline__ void apply(const uint4 &v, const Dig *d, A &s) { Scorer<Q8, ND>::step(v, d, s.a); }
    static __device__ __forceinline__ float finish(const A &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        typename Scorer<Q8, ND>::Acc t;
#pragma unroll
        for (int j = 0; j < ND; ++j) t.a[j] = s.a[j];
        return Scorer<Q8, ND>::finish(t, a, h, x);
    }
};
template <int ND>
struct SmallOps<Q4, ND> {
    using Dig = uint4;
    static constexpr int DPC = 2 * ND;
    using A = typename Scorer<Q4, ND>::Acc;
    static __device__ __forceinline__ void reset(A &s) { Scorer<Q4, ND>::reset(s); }
    static __device__ __forceinline__ void apply(const uint4 &v, const Dig *d, A &s) { Scorer<Q4, ND>::step(v, d, s.hi, s.lo); }
    static __device__ __forceinline__ float finish(const A &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        return Scorer<Q4, ND>::finish(s, a, h, x);
    }
};
template <int ND>
struct SmallOps<Q16, ND> {
    using Dig = uint2;
    static constexpr int DPC = ND;
    struct A { int a[ND]; }; // |s16 * s8| <= 2^22 and at most 384 dims (C <= 48): int32 partials hold a whole row
    static __device__ __forceinline__ void reset(A &s) {
#pragma unroll
        for (int j = 0; j < ND; ++j) s.a[j] = 0;
    }
    static __device__ __forceinline__ void apply(const uint4 &v, const Dig *d, A &s) { Scorer<Q16, ND>::step(v, d, s.a); }
    static __device__ __forceinline__ float finish(const A &s, const ScanArgs &a, const PQHeader &h, const float2 &x) {
        typename Scorer<Q16, ND>::Acc t;
        t.I = digits_total<ND>(s.a); // the general kernel flushes per tile: one tile here
        return Scorer<Q16, ND>::finish(t, a, h, x);
    }
};

template <int QT, int ND, int C, int Q, bool CQ>
__global__ void __launch_bounds__(kSmallWarps * 32, 1) scan_small_kernel(const ScanArgs a) {
    grid_launch_dependents(); // see launch_dependent (common.cuh): finalize_kernel queues behind this grid's CTAs
    // a lane holds 16 uint4 of row data at a time: U = 16 / C whole rows (of U different blocks) when C <= 8, else
    // pieces of 8 chunks of two rows (C = 12: 4 chunks of four rows).  The digits of a chunk are read once for the U
    // rows: the L1 data pipe, which carries the row loads and (through shared memory) the digit broadcasts, limited the
    // first version of this kernel (89 % with one row per lane; cfg4: 2084 -> 2575 QPS with two).
    // Q (queries per WARP) stays 1: see launch_scan_small_t.
    constexpr int PC = C <= 8 ? C : (C % SZG_SMALL_PCBIG == 0 ? SZG_SMALL_PCBIG : 4); // chunks per piece (measured: 8 beats 4 and 16 at C = 24 .. 48)
    constexpr int U = 16 / PC;                             // blocks per step
    constexpr int NP = (C + PC - 1) / PC;
    using Ops = SmallOps<QT, ND>;
    using Dig = typename Ops::Dig;
    constexpr int DPC = Ops::DPC;
    extern __shared__ __align__(16) unsigned char s_pq[]; // the item's prepared queries: G x (header + digits)
    __shared__ uint32_t s_pub[kSmallWarps];               // key of each warp's mth-best row so far (0xFFFFFFFF: none yet)
    __shared__ unsigned long long s_merge[kSmallWarps * 32]; // the warps' lists at the end of an item, merged per group
    static_assert(Q == 1, "queries per warp: only 1 is instantiated");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    // a.wgroups = G: the CTA's warps form G groups, each answering ANOTHER query over the SAME row part: the groups walk the
    // part's blocks in the same order at the same pace, so a row brought into L1 for one group is a hit for the others
    const uint32_t G = a.wgroups, gw = (uint32_t)nw / G;  // warps per group
    const uint32_t wg = (uint32_t)warp / gw, wl = (uint32_t)warp - wg * gw;
    const uint32_t P = a.parts, ngroups = (a.nq + G - 1) / G, nitems = ngroups * P, nlists = P; // one list per (query, part)
    const uint32_t mth = (32u + gw - 1) / gw;
    const uint32_t stride = P * gw;
    const uint32_t n16 = (uint32_t)(a.pq_stride / 16);
    for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const uint32_t qg = item / P, p = item - qg * P;
        const uint32_t q0 = qg * G;
        const uint32_t nqv = min(G, a.nq - q0); // queries of this item (the last one may be short: its extra groups idle)
        const uint32_t q = q0 + wg;
        __syncthreads(); // the previous item's readers of s_pq / s_pub are done
        {
            if (!CQ)
                for (uint32_t i = tid; i < n16 * nqv; i += blockDim.x) {
                    const uint32_t qq = i / n16, kk = i - qq * n16;
                    reinterpret_cast<uint4 *>(s_pq + (size_t)qq * a.pq_stride)[kk] =
                        __ldg(reinterpret_cast<const uint4 *>(a.pq + (size_t)(q0 + qq) * a.pq_stride) + kk);
                }
            if (tid < nw) s_pub[tid] = 0xFFFFFFFFu;
        }
        __syncthreads();
        const bool active = wg < nqv; // warp-uniform: the groups beyond the item's queries only keep the barriers company
        const uint32_t cbase = q * n16; // CQ: this group's query in the constant window (uint4 slots; warp-uniform)
        const unsigned char *my_pq = s_pq + (size_t)wg * a.pq_stride;
        WarpList<1> list;
        list.init();
        // a step takes U blocks: adjacent ones (contiguous in the mirror; a.adjacent) or `stride` apart
        const uint32_t ustep = a.adjacent ? 1u : stride;
        for (uint32_t g = p * gw + wl; active && (a.adjacent ? g * U : g) < a.nblk; g += a.adjacent ? stride : stride * U) {
            const uint32_t b0 = a.adjacent ? g * U : g;
            float2 ax[U];
            uint32_t lv[U];
            const uint4 *src[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t blk = b0 + (uint32_t)u * ustep;
                const bool in = blk < a.nblk;                    // warp-uniform
                const uint32_t bc = in ? blk : a.nblk - 1;       // clamped: loads stay unconditional and in flight together
                uint32_t w = __ldg(a.live + bc);
                if (a.mask) w &= __ldg(a.mask + bc);
                lv[u] = in ? w : 0u;
                src[u] = a.codes + (size_t)bc * C * 32 + lane;
                ax[u] = load_aux(a, bc * 32 + lane);
            }
            typename Ops::A acc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) Ops::reset(acc[u]);
#pragma unroll
            for (int pz = 0; pz < NP; ++pz) {
                uint4 data[U][PC];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int c = 0; c < PC; ++c)
                        if (pz * PC + c < C) data[u][c] = __ldg(src[u] + (pz * PC + c) * 32);
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    if (pz * PC + c < C) {
                        // this chunk's digits: warp-uniform loads (constant cache or shared memory), used for the U rows of the lane
                        Dig d[DPC];
                        if (CQ) {
                            constexpr uint32_t per16 = 16 / sizeof(Dig); // digit vectors per uint4 slot
                            const Dig *cd = reinterpret_cast<const Dig *>(c_pq);
                            const uint32_t at = cbase * per16 + (uint32_t)(sizeof(PQHeader) / sizeof(Dig));
#pragma unroll
                            for (int j = 0; j < DPC; ++j) d[j] = cd[at + (pz * PC + c) * DPC + j];
                        } else {
                            const Dig *dig = reinterpret_cast<const Dig *>(my_pq + sizeof(PQHeader));
#pragma unroll
                            for (int j = 0; j < DPC; ++j) d[j] = dig[(pz * PC + c) * DPC + j];
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) Ops::apply(data[u][c], d, acc[u]);
                    }
                }
            }
            const PQHeader &h = CQ ? *reinterpret_cast<const PQHeader *>(&c_pq[cbase]) : *reinterpret_cast<const PQHeader *>(my_pq);
            unsigned long long keys[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float key = Ops::finish(acc[u], a, h, ax[u]);
                const uint32_t blk = b0 + (uint32_t)u * ustep;
                keys[u] = ((lv[u] >> lane) & 1u) ? make_key64(key, blk * 32 + lane) : kNoKey;
            }
            { // the group's shared bound: once every warp of it has published, nothing above the largest published key matters
                const uint32_t e = (uint32_t)lane < gw ? *reinterpret_cast<volatile uint32_t *>(&s_pub[wg * gw + lane]) : 0u;
                const uint32_t mx = __reduce_max_sync(0xffffffffu, e);
                if (mx != 0xFFFFFFFFu) {
                    const unsigned long long b = ((unsigned long long)mx << 32) | 0xFFFFFFFFull;
                    if (b < list.thr) list.thr = b; // keys equal to the bound still pass
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (lv[u]) list.offer(keys[u], lane);
            if ((uint32_t)lane == mth - 1 && list.v[0] != kNoKey) {
                const uint32_t pk = (uint32_t)(list.v[0] >> 32);
                if (pk < s_pub[warp]) *reinterpret_cast<volatile uint32_t *>(&s_pub[warp]) = pk;
            }
        }
        // the group's lists -> one list of (query q, part p) for finalize_kernel: a tree of pairwise merges of sorted lists
        s_merge[warp * 32 + lane] = list.v[0];
        for (uint32_t half = gw >> 1; half >= 1; half >>= 1) {
            __syncthreads();
            if (active && wl < half) {
                list.merge_sorted(s_merge[(warp + half) * 32 + lane], lane);
                s_merge[warp * 32 + lane] = list.v[0];
            }
        }
        if (active && wl == 0) a.cand[((size_t)q * nlists + p) * 32 + lane] = list.v[0];
    }
}

// parts per query that minimise the makespan of nq x P items dealt round-robin to `grid` CTAs (ties: fewer parts =
// fewer, longer lists); P <= 32 unless a single round needs more
inline uint32_t small_parts(uint32_t nq, uint32_t grid) {
    if (nq == 0) return 1;
    if (nq * 32u <= grid) return grid / nq; // few queries: one round, every CTA busy
    uint32_t best = 1;
    double best_t = 1e30;
    for (uint32_t P = 1; P <= 32; ++P) {
        const uint32_t rounds = (nq * P + grid - 1) / grid;
        const double t = (double)rounds / P + 0.002 * rounds; // + a small per-item cost
        if (t < best_t - 1e-12) { best_t = t; best = P; }
    }
    return best;
}

template <int QT, int ND, int Q, bool CQ>
cudaError_t launch_scan_small_nd(uint32_t C, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    switch (C) {
#define SZG_SMALL_CASE(CC) case CC: scan_small_kernel<QT, ND, CC, Q, CQ><<<grid, kSmallWarps * 32, smem, st>>>(a); break;
    SZG_SMALL_CASE(2) SZG_SMALL_CASE(4) SZG_SMALL_CASE(8) SZG_SMALL_CASE(12) SZG_SMALL_CASE(16) SZG_SMALL_CASE(24)
    SZG_SMALL_CASE(32) SZG_SMALL_CASE(48)
#undef SZG_SMALL_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
// Q (queries per item) = 2 was measured and is not instantiated: scoring every loaded chunk against a pair of queries halves the
// row loads per query, but the pair's accumulators, lists and digits cost the registers that keep 16 loads in flight per lane --
// cfg4 2374 -> 1217 QPS, 24-chunk rows 20 -> 37 us/query; only rows of <= 4 chunks gained (~10 %): profiles/r01b_scan_small_vs_general.log
template <int QT>
cudaError_t launch_scan_small_t(int nd, uint32_t C, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    if (a.qper != 1) return cudaErrorInvalidValue;
    const size_t bytes = (size_t)a.nq * a.pq_stride;
    if (a.const_queries && bytes <= (size_t)kConstSlots * 16) {
        // the launch's prepared queries -> this translation unit's constant window; users of the window are chained
        static std::mutex mu;
        static cudaEvent_t ev[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        std::lock_guard<std::mutex> lk(mu);
        if (!ev[dev]) {
            if ((e = cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming)) != cudaSuccess) return e;
        } else if ((e = cudaStreamWaitEvent(st, ev[dev], 0)) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbolAsync(c_pq, a.pq, bytes, 0, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        e = nd == 2 ? launch_scan_small_nd<QT, 2, 1, true>(C, grid, 16, st, a) : launch_scan_small_nd<QT, 3, 1, true>(C, grid, 16, st, a);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(ev[dev], st);
    }
    if (nd == 2) return launch_scan_small_nd<QT, 2, 1, false>(C, grid, smem, st, a);
    return launch_scan_small_nd<QT, 3, 1, false>(C, grid, smem, st, a);
}

inline bool scan_small_supported(int qt, uint32_t C) {
    return qt <= Q16 && (C == 2 || C == 4 || C == 8 || C == 12 || C == 16 || C == 24 || C == 32 || C == 48);
}

} // namespace szg
