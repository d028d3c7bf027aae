// batch_q8.cu -- K4: batched-query contraction on the 5th-generation tensor cores (tcgen05 / TMEM) with a
// fused threshold top-k epilogue, for 8-bit collections.
//
// Replaces B independent Search calls (collection.go:569-711) on the same collection: the distance
// surrogates of a block of queries against every row are one dense integer contraction
//     D[query-digit, row] = sum_i W_digit[query][i] * u[row][i]          (s8 x u8 -> s32, exact)
// which is what tensor cores are for (a single query is a memory-bound GEMV and stays on the streaming
// scan kernel).  Per CTA:
//   A (M = 128) : 64 queries x 2 base-128 digit planes of the fixed-point query, resident in shared memory
//                 in the canonical K-major no-swizzle core-matrix layout (prepared by batch_pack_kernel);
//   B (N = 32)  : one 32-row block of the column-blocked mirror; its HBM image IS the canonical K-major
//                 layout (8 rows x 16 B core matrices), so a block lands in shared memory with one
//                 cp.async.bulk and is consumed by tcgen05.mma without any reshuffle;
//   D           : 128 x 32 s32 accumulators in TMEM, double buffered (2 x 32 columns).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-5 = epilogue
// (tcgen05.ld, combine the two digit planes, key, per-query threshold, append to the query's candidate
// buffer, occasional warp-level compaction by bitonic sort).  Pipelines: block ring (full/empty
// mbarriers), accumulator ring (tcgen05.commit -> epilogue -> release).
// CTAs are (query group g, row range r): the 16 groups of a 1024-query batch walk the same row range at
// the same time, so HBM is read once per range and the other 15 reads hit L2.
// The candidate buffers feed the same finalize_kernel as the scan path (fp64 re-score, ordering,
// certification against the surrogate error bound), so results are identical to single queries.
#include <cuda.h>

#include "kernels.h"

namespace szg {

namespace {

constexpr int kBatchQueries = 64;          // queries per CTA (x 2 digit planes = M 128)
constexpr int kBatchThreads = 192;         // producer, MMA, 4 epilogue warps
constexpr int kBatchCap = 512;             // candidate buffer entries per (CTA, query)
constexpr int kBatchKp = 128;              // survivors of a compaction (= finalize MODE 2)
constexpr uint32_t kNoBlock = 0xFFFFFFFFu;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B; SBO between 8-row groups, LBO between the two
    // 16-byte K halves of one MMA (K = 32 bytes); bit 46 = descriptor version of sm_100
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }


} // namespace

// ---- query digits -> the A operand image of each 64-query group:
//      img[g][c][16 row groups][8 rows][16 B], row m = plane * 64 + (query % 64), plane 0 = most significant digit
__global__ void batch_pack_kernel(const unsigned char *__restrict__ pq, size_t pq_stride, uint32_t nq, uint32_t C,
                                  unsigned char *__restrict__ img) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ngroups = (nq + kBatchQueries - 1) / kBatchQueries;
    const size_t total = (size_t)ngroups * 128 * C;
    if (t >= total) return;
    const uint32_t c = (uint32_t)(t % C), m = (uint32_t)((t / C) % 128), g = (uint32_t)(t / ((size_t)C * 128));
    const uint32_t q = g * kBatchQueries + (m & 63), plane = m >> 6;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q < nq) v = *reinterpret_cast<const uint4 *>(pq + (size_t)q * pq_stride + sizeof(PQHeader) + ((size_t)c * 2 + plane) * 16);
    *reinterpret_cast<uint4 *>(img + (size_t)g * C * 2048 + ((size_t)c * 16 + (m >> 3)) * 128 + (m & 7) * 16) = v;
}

// One "super tile" = kNB = 4 consecutive 32-row blocks = N 128: a tcgen05.mma costs ~64 cycles whether N is 32
// or 128 (measured, tools/umma_test.cu), so the rows operand must be 128 wide.  The four blocks are laid out in
// shared memory chunk-major, [chunk][block][8-row group][128 B], by 512-byte bulk copies (one per block and
// chunk, issued by all 32 producer lanes), which gives the 16 row groups the uniform 128-byte stride the
// K-major no-swizzle descriptor needs.  The query digit planes (A, M = 128) live in TMEM (192 columns,
// written once per CTA with tcgen05.st: lane = row, 4 K-bytes per column), so no shared memory or bandwidth
// is spent on A and both accumulator buffers (2 x 128 columns) fit beside it.
constexpr uint32_t kNB = 4;                     // blocks per super tile (N = 4 x 32)
constexpr uint32_t kAccCols = kNB * 32;         // columns of one accumulator buffer
constexpr uint32_t kTmemCols = 512;             // 2 accumulator buffers + up to 256 columns of A

struct SuperMeta { uint32_t blk, live[kNB]; };

__global__ void __launch_bounds__(kBatchThreads, 1) batch_kernel(const BatchArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t b_full[kMaxStages], b_empty[kMaxStages], d_full[2], d_empty[2];
    __shared__ SuperMeta s_meta[kMaxStages], s_dmeta[2];
    __shared__ unsigned long long s_thr[kBatchQueries];
    __shared__ uint32_t s_cnt[kBatchQueries];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t C = a.C, S = a.stages;
    const uint32_t stage_bytes = kNB * C * 512u;
    unsigned char *sB = smem;
    int32_t *s_x = reinterpret_cast<int32_t *>(sB + (size_t)S * stage_bytes); // [2][128][17] plane exchange

    const uint32_t g = blockIdx.x % a.ngroups, r = blockIdx.x / a.ngroups; // (query group, row range)
    const uint32_t nsup = (a.nblk + kNB - 1) / kNB;
    const uint32_t per = (nsup + a.nranges - 1) / a.nranges;
    const uint32_t sup0 = min(nsup, r * per), sup1 = min(nsup, sup0 + per);
    const uint32_t q0 = (a.group0 + g) * kBatchQueries;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&d_full[d], 1); mbar_init(&d_empty[d], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kBatchQueries) { s_thr[tid] = (q0 + tid < a.nq) ? kNoKey : 0ull; s_cnt[tid] = 0; }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t tmemA = tmem + 2 * kAccCols;

    // ---- A operand: the epilogue threads own TMEM lanes; lane L = plane * 64 + query holds that digit plane
    if (warp >= 2) {
        const uint32_t lq = (uint32_t)(warp & 3), L = lq * 32 + lane;
        const uint32_t q = q0 + (L & 63), plane = L >> 6;
        const unsigned char *src = a.pq + (size_t)(q < a.nq ? q : 0) * a.pq_stride + sizeof(PQHeader);
        for (uint32_t ks = 0; ks < C / 2; ++ks) { // 2 chunks = 32 K-bytes = 8 columns = one MMA K step
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
            if (q < a.nq) {
                v0 = __ldg(reinterpret_cast<const uint4 *>(src + ((size_t)(2 * ks) * 2 + plane) * 16));
                v1 = __ldg(reinterpret_cast<const uint4 *>(src + ((size_t)(2 * ks + 1) * 2 + plane) * 16));
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                             tmemA + ((lq * 32u) << 16) + ks * 8),
                         "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w)
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == 0) {
        // ================================================================ TMA producer (whole warp)
        uint32_t t = 0; // stage counter
        for (uint32_t sup = sup0; sup <= sup1; ++sup) {
            uint32_t lv = 0;
            if (sup < sup1 && lane < (int)kNB) {
                const uint32_t blk = sup * kNB + lane;
                if (blk < a.nblk) {
                    lv = __ldg(a.live + blk);
                    if (a.mask) lv &= __ldg(a.mask + blk);
                }
            }
            const uint32_t any = __ballot_sync(0xffffffffu, lv != 0);
            if (sup < sup1 && !any) continue; // nothing live in these blocks: never fetched
            const uint32_t s = t % S;
            if (t >= S) mbar_wait(&b_empty[s], ((t / S) - 1) & 1u);
            if (lane < (int)kNB) s_meta[s].live[lane] = lv;
            __syncwarp(); // the meta words are in place before lane 0 arms / arrives on the barrier
            if (sup < sup1) {
                if (lane == 0) {
                    // one TMA tensor copy per stage: box (512 B, 4 blocks, C chunks) of the 3-D view
                    // (512 B | block, stride C*512 | chunk, stride 512) lands chunk-major, [chunk][block][512 B];
                    // blocks past the end of the mirror are zero-filled by the copy engine
                    s_meta[s].blk = sup * kNB;
                    mbar_expect_tx(&b_full[s], stage_bytes);
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                            smem_u32(sB + (size_t)s * stage_bytes)),
                        "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&b_full[s])), "r"(0), "r"((int)(sup * kNB)), "r"(0)
                        : "memory");
                }
            } else { // end of the range: a sentinel travels through both pipelines
                if (lane == 0) { s_meta[s].blk = kNoBlock; mbar_arrive(&b_full[s]); }
            }
            __syncwarp();
            ++t;
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            // D = S32, A = S8 (query digits, TMEM), B = U8 (codes, smem K-major), N = 128, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((kAccCols >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t sb = smem_u32(sB);
            for (uint32_t t = 0;; ++t) {
                const uint32_t s = t % S, d = t & 1u;
                mbar_wait(&b_full[s], (t / S) & 1u);
                const uint32_t blk = *reinterpret_cast<volatile uint32_t *>(&s_meta[s].blk);
                if (t >= 2) mbar_wait(&d_empty[d], ((t >> 1) - 1) & 1u);
                s_dmeta[d].blk = blk;
                for (uint32_t j = 0; j < kNB; ++j) s_dmeta[d].live[j] = *reinterpret_cast<volatile uint32_t *>(&s_meta[s].live[j]);
                __threadfence_block();
                if (blk == kNoBlock) { mbar_arrive(&d_full[d]); break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (!(a.debug & 4u))
                    for (uint32_t ks = 0; ks < C / 2; ++ks) {
                        const uint64_t db = umma_desc(sb + s * stage_bytes + ks * (2 * kNB * 512u), kNB * 512u, 128);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(
                                tmem + d * kAccCols),
                            "r"(tmemA + ks * 8), "l"(db), "r"(idesc), "r"((uint32_t)(ks > 0)), "r"(0u)
                            : "memory");
                    }
                umma_commit(&b_empty[s]); // the stage may be refilled once these MMAs have read it
                umma_commit(&d_full[d]);  // accumulators complete
            }
        }
    } else {
        // ================================================================ epilogue (128 threads)
        const int ew = warp - 2;                        // 0..3, compaction work split
        const uint32_t lq = (uint32_t)(warp & 3);       // TMEM lane quarter this warp may read
        const uint32_t L = lq * 32 + lane;              // TMEM lane = A row: plane * 64 + query
        const uint32_t qi = L & 63, upper = L >> 6;     // upper half holds the least significant digit plane
        const uint32_t q = q0 + qi;
        const bool qvalid = q < a.nq;
        const PQHeader *hdr = reinterpret_cast<const PQHeader *>(a.pq + (size_t)(qvalid ? q : 0) * a.pq_stride);
        // the cancellation num = 2 I + numc happens in exact 64-bit integers; everything after it is fp32
        const long long numc = (long long)hdr->numc;
        // |num| = M 2^F |x.q| <= M 2^F sqrt(d) ||q||: when that fits 31 bits the whole sum can run in wrapping
        // 32-bit arithmetic (exact mod 2^32, and the true value fits), which spares the 64-bit ops and the slow
        // s64 -> f32 conversion
        const bool fits32 = 255.0 * ldexp(1.0, hdr->F) * sqrt((double)a.dims * hdr->qn2) < 2.0e9;
        const int numc32 = (int)numc;
        const float c_key = (float)hdr->c_key, c_dot2 = (float)(2.0 * hdr->c_dot), qn2 = (float)hdr->qn2;
        const bool zero_query = hdr->zero_query != 0;
        const bool cosine = a.metric == COSINE;
        unsigned long long *gbuf = a.cand + ((size_t)(qvalid ? q : 0) * a.nlists + (size_t)r * (kBatchCap / kBatchKp)) * kBatchKp;
        const float2 *aux = reinterpret_cast<const float2 *>(a.aux);
        const uint32_t partner = upper ? L - 64 : L + 64;
        const uint32_t slot_max = a.nblk * 32 - 1; // the last super tile may reach past the mirror: clamp aux reads
        uint32_t xb = 0; // exchange buffer parity

        for (uint32_t tile = 0;; ++tile) {
            const uint32_t d = tile & 1u;
            mbar_wait(&d_full[d], (tile >> 1) & 1u);
            const uint32_t blk_base = *reinterpret_cast<volatile uint32_t *>(&s_dmeta[d].blk);
            if (blk_base == kNoBlock) break;
            uint32_t lives[kNB];
#pragma unroll
            for (uint32_t j = 0; j < kNB; ++j) lives[j] = *reinterpret_cast<volatile uint32_t *>(&s_dmeta[d].live[j]);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned long long thr = s_thr[qi];
#pragma unroll 1
            for (uint32_t j = 0; j < kNB; ++j) {
                uint32_t acc[32];
                tmem_ld32(tmem + ((lq * 32u) << 16) + d * kAccCols + j * 32, acc);
                if (j == kNB - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[d]); // this warp's quarter of the buffer is in registers
                }
                if (a.debug & 1u) continue;
                // exchange the halves: lower threads (digit 1) handle rows 0..15, upper threads (digit 0) rows 16..31
                int32_t *xw = s_x + xb * (128 * 17);
                xb ^= 1u;
#pragma unroll
                for (int i = 0; i < 16; ++i) xw[L * 17 + i] = (int32_t)acc[upper ? i : 16 + i];
                epi_barrier();
                const uint32_t slot0 = (blk_base + j) * 32 + (upper ? 16 : 0);
                const uint32_t live = lives[j] >> (upper ? 16 : 0);
                unsigned long long k64[16];
                unsigned long long best = kNoKey;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int32_t other = xw[partner * 17 + i];
                    const int32_t d1 = upper ? other : (int32_t)acc[i];
                    const int32_t d0 = upper ? (int32_t)acc[16 + i] : other;
                    float numf;
                    if (fits32) numf = (float)(2 * (d1 * 128 + d0) + numc32);
                    else numf = (float)(2 * ((long long)d1 * 128 + (long long)d0) + numc);
                    const float2 ax = (a.debug & 2u) ? make_float2(0.1f, 1.f) : __ldg(aux + min(slot0 + i, slot_max));
                    float key;
                    if (cosine) key = (zero_query || ax.x == 0.f) ? 1.0f : -((numf * c_key) * ax.x);
                    else key = (ax.y + qn2) - numf * c_dot2;
                    const bool ok = (live >> i) & 1u;
                    k64[i] = ok ? make_key64(key, slot0 + i) : kNoKey;
                    best = k64[i] < best ? k64[i] : best;
                }
                if (best < thr) { // rare once the threshold has tightened
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (k64[i] < thr) {
                            const uint32_t pos = atomicAdd(&s_cnt[qi], 1u);
                            if (pos < (uint32_t)kBatchCap) gbuf[pos] = k64[i];
                        }
                }
            }
            if (a.debug & 1u) continue;
            epi_barrier();
            // compaction: a super tile appends at most kNB * 32 keys per query
            for (uint32_t cq = ew; cq < (uint32_t)kBatchQueries; cq += 4) {
                const uint32_t n = s_cnt[cq];
                if (n <= (uint32_t)(kBatchCap - kNB * 32)) continue;
                unsigned long long *gb = a.cand + ((size_t)(q0 + cq) * a.nlists + (size_t)r * (kBatchCap / kBatchKp)) * kBatchKp;
                // 512 keys = 16 per lane in registers (element e = lane * 16 + r): bitonic network, register
                // compare-exchanges for partner distance < 16, shuffles above
                unsigned long long v[16];
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) { const uint32_t e = (uint32_t)lane * 16 + rr; v[rr] = e < n ? gb[e] : kNoKey; }
#pragma unroll
                for (int k = 2; k <= kBatchCap; k <<= 1) {
#pragma unroll
                    for (int jj = k >> 1; jj > 0; jj >>= 1) {
                        if (jj >= 16) {
                            const int lj = jj >> 4;
                            const bool lower = (lane & lj) == 0;
#pragma unroll
                            for (int rr = 0; rr < 16; ++rr) {
                                const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[rr], lj);
                                const bool up = (((lane * 16 + rr) & k) == 0);
                                const bool keep_min = (lower == up);
                                v[rr] = keep_min ? (v[rr] < o ? v[rr] : o) : (v[rr] > o ? v[rr] : o);
                            }
                        } else {
#pragma unroll
                            for (int rr = 0; rr < 16; ++rr) {
                                if ((rr & jj) == 0) {
                                    const int r2 = rr | jj;
                                    const bool up = (((lane * 16 + rr) & k) == 0);
                                    const unsigned long long x = v[rr], y = v[r2];
                                    const bool sw = (x > y) == up;
                                    v[rr] = sw ? y : x;
                                    v[r2] = sw ? x : y;
                                }
                            }
                        }
                    }
                }
                // ranks [lane*16, lane*16+16): the first 128 ranks live in lanes 0..7
                if (lane < kBatchKp / 16) {
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) gb[lane * 16 + rr] = v[rr];
                }
                if (lane == kBatchKp / 16 - 1) { s_cnt[cq] = kBatchKp; s_thr[cq] = v[15]; }
                __syncwarp();
            }
            epi_barrier();
        }
        // pad the unused tail of every candidate buffer so that finalize sees (cap / Kp) well-formed lists
        epi_barrier();
        for (uint32_t cq = ew; cq < (uint32_t)kBatchQueries; cq += 4) {
            if (q0 + cq >= a.nq) continue;
            unsigned long long *gb = a.cand + ((size_t)(q0 + cq) * a.nlists + (size_t)r * (kBatchCap / kBatchKp)) * kBatchKp;
            for (uint32_t i = s_cnt[cq] + lane; i < (uint32_t)kBatchCap; i += 32) gb[i] = kNoKey;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

size_t batch_smem_bytes(uint32_t C, uint32_t stages) { return (size_t)stages * kNB * C * 512 + 2 * 128 * 17 * 4; }
uint32_t batch_max_chunks() { return 256 * 4 / 16; } // A lives in <= 256 TMEM columns: rows of at most 1024 bytes
uint32_t batch_lists_per_range() { return kBatchCap / kBatchKp; }

cudaError_t batch_configure(size_t max_smem) {
    return cudaFuncSetAttribute(batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
}

cudaError_t launch_batch_pack(const unsigned char *pq, size_t pq_stride, uint32_t nq, uint32_t C, unsigned char *img,
                              cudaStream_t st) {
    const uint32_t ngroups = (nq + kBatchQueries - 1) / kBatchQueries;
    const size_t total = (size_t)ngroups * 128 * C;
    batch_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pq, pq_stride, nq, C, img);
    return cudaGetLastError();
}

// 3-D tensor map over the column-blocked mirror: (64 x u64 = one chunk of a block's 32 rows | block | chunk)
static cudaError_t make_tmap(const BatchArgs &a, CUtensorMap *tm) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t gdim[3] = {64, a.nblk, a.C};
    const cuuint64_t gstride[2] = {(cuuint64_t)a.C * 512, 512}; // bytes, dims 1 and 2
    const cuuint32_t box[3] = {64, kNB, a.C};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint4 *>(a.codes), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_batch(const BatchArgs &a, cudaStream_t st) {
    CUtensorMap tm;
    cudaError_t e = make_tmap(a, &tm);
    if (e != cudaSuccess) return e;
    batch_kernel<<<a.ngroups * a.nranges, kBatchThreads, batch_smem_bytes(a.C, a.stages), st>>>(a, tm);
    return cudaGetLastError();
}

} // namespace szg
