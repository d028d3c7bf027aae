// batch_q8.cu -- K4: batched-query contraction on the 5th-generation tensor cores (tcgen05 / TMEM) with a
// fused threshold top-k epilogue, for 8-bit collections.
//
// Replaces B independent Search calls (collection.go:569-711) on the same collection: the distance
// surrogates of a block of queries against every row are one dense integer contraction
//     D[query-digit, row] = sum_i W_digit[query][i] * u[row][i]          (s8 x u8 -> s32, exact)
// which is what tensor cores are for (a single query is a memory-bound GEMV and stays on the streaming
// scan kernel).  Per CTA:
//   A (M = 128) : 64 queries x 2 base-128 digit planes of the fixed-point query, resident in TMEM for the
//                 whole kernel (written once with tcgen05.st, C * 4 columns);
//   B (N = 128) : one "super tile" = 4 consecutive 32-row blocks of the column-blocked mirror.  The HBM
//                 image of a block IS the canonical K-major no-swizzle layout (8 rows x 16 B core matrices),
//                 so TMA lands K slices of a super tile in shared memory ([chunk][block][512 B]) and
//                 tcgen05.mma consumes them without any reshuffle;
//   D           : 128 x 128 s32 accumulators in TMEM, double buffered.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-9 = epilogue (8 queries each), warps 10-11 = bound
// pollers (they keep the bound shared by a query's CTAs current, see poll_bounds).
// Pipelines: a ring of K slices (full/empty mbarriers; the ring never drains between super tiles) and the
// two accumulator buffers (tcgen05.commit -> epilogue -> release).
//
// Epilogue.  TMEM lane L holds (query, plane) with the two planes of a query 8 lanes apart, so that
// tcgen05.ld.16x256b hands thread t BOTH planes of query t/4 for 2 columns (= rows) per 8-column group
// (layout measured with tools/tmem_probe.cu): no exchange between threads is needed.  Per (query, row) the
// thread combines the planes in exact integer arithmetic, converts once, and compares against a per-query
// pre-transformed threshold (cosine: num * (1/||x||) > T;  euclid: num * c - ||x||^2 > T) that is slightly
// conservative; the few survivors recompute the exact surrogate key of the streaming scan and append it to
// the query's candidate buffer.  Every query belongs to exactly one epilogue warp, so appends, the
// occasional compaction (bitonic sort in registers) and threshold updates need warp-level
// synchronisation only.  Liveness / filter bits are checked only for rows that pass the fast test (a removed or
// filtered row passes it as rarely as a live one does).
// CTAs are (query group g, range r).  The 128-row tiles of the collection are dealt to the CTAs of a group on demand (one atomic
// counter per group, next_live_tile), every group in ascending tile order at about the same pace, so HBM is read once and the
// other groups' reads hit L2.  A CTA keeps one sorted candidate list per query for the rows it saw ("range" below = the rows of
// one CTA).  The launch is a programmatic dependent of prep_kernel and finalize_kernel one of this launch (common.cuh).
// The candidate buffers feed the same finalize_kernel as the scan path (fp64 re-score, ordering,
// certification against the surrogate error bound), so results are identical to single queries.
#include <cuda.h>

#include "kernels.h"

namespace szg {

namespace {

constexpr int kBatchQueries = 64;          // queries per CTA (x 2 digit planes = M 128)
constexpr int kBatchThreads = 384;         // producer, MMA, 8 epilogue warps, 2 bound pollers
constexpr int kBatchPollWarp = 10;         // first poller warp (queries 0..31 of the group; the next one takes 32..63)
constexpr uint32_t kNoBlock = 0xFFFFFFFFu;
constexpr uint32_t kNB = 4;                // blocks per super tile (N = 4 x 32 = 128 rows)
constexpr uint32_t kTileRows = kNB * 32;
constexpr uint32_t kAccCols = kTileRows;   // columns of one accumulator buffer
constexpr uint32_t kTmemCols = 512;        // 2 accumulator buffers + up to 256 columns of A
constexpr int kBatchMaxStages = 13;
constexpr uint32_t kAuxSlots = 6;          // ring of per-tile side data (aux pairs, live words)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B; SBO between 8-row groups, LBO between the two
    // 16-byte K halves of one MMA (K = 32 bytes); bit 46 = descriptor version of sm_100
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 TMEM lanes x 64 columns: thread t receives lanes (t/4, t/4 + 8) x columns 8 * rep + 2 * (t%4) + {0, 1}:
//   r[4 rep + 0/1] = lane t/4 (plane 0), r[4 rep + 2/3] = lane t/4 + 8 (plane 1)
__device__ __forceinline__ void tmem_ld_16x64(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// waits for every tcgen05.ld of this thread; the registers are passed through the statement so that no use of
// them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
    asm volatile(""
                 : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// per-query state of the epilogue (shared memory; owned by one epilogue warp)
struct __align__(16) QState {
    unsigned long long thr;   // admission threshold: min(last key of the query's list, gbound)
    unsigned long long gbound; // bound shared by the row ranges of this query (see poll_bounds), kNoKey = none yet
    long long numc;
    float c_key, inv_ckey, c_dot2, qn2; // surrogate constants (PQHeader)
    float T;                  // pre-transformed, conservative threshold of the fast test
    float slack;              // euclid: bound of the rounding differences between fast test and exact key
    uint32_t zero;            // zero query
    uint32_t valid;
    uint32_t pub;             // smallest mth-best key this range has published for the query (0xFFFFFFFF = none)
    uint32_t inserts;         // rows inserted into the list (profiling)
};

// threshold of the fast test "v > T" such that every key <= thr passes (see the file comment); fp32 arithmetic
// with margins far above its own rounding
template <bool COS>
__device__ __forceinline__ float fast_threshold(const QState &s) {
    if (s.thr == kNoKey) return -INFINITY;
    const float thr_f = key_to_float((uint32_t)(s.thr >> 32));
    if (thr_f != thr_f) return -INFINITY;
    if (COS) {
        // every key is 1.0.  Against a list entry (real slot) nothing later can win the tie, rows arrive in ascending
        // slot order; against the shared bound (slot part all ones) every such row is still admissible
        if (s.zero || s.c_key <= 0.f)
            return (thr_f > 1.0f || (thr_f == 1.0f && (uint32_t)s.thr == 0xFFFFFFFFu)) ? -INFINITY : INFINITY;
        const float t = -thr_f * s.inv_ckey;
        return t - fabsf(t) * 3.814697265625e-6f - 1e-30f; // 2^-18
    }
    const float t = (s.qn2 - thr_f) - s.slack - fabsf(thr_f) * 9.5367431640625e-7f;
    return t - fabsf(t) * 4.76837158203125e-7f - 1e-30f;
}

template <int E>
__device__ __forceinline__ unsigned long long pick(const unsigned long long (&v)[E], uint32_t idx) {
    unsigned long long r = v[0];
#pragma unroll
    for (int e = 1; e < E; ++e)
        if (idx == (uint32_t)e) r = v[e];
    return r;
}

// sorted insert of nk (< the key at rank capm1) into a query's candidate list in shared memory; lane i holds ranks
// [i E, i E + E).  Publishes the range's mth-best key when it improved and returns the key at rank capm1
// (warp-uniform): capm1 = Kp - 1 normally, mth - 1 while the range is seeding the shared bound.
template <int E>
__device__ __forceinline__ unsigned long long list_insert(unsigned long long *l, unsigned long long nk, int lane, uint32_t capm1,
                                                          uint32_t mthm1, uint32_t *gm, uint32_t *pub) {
    unsigned long long v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = l[lane * E + e];
    const unsigned gt = __ballot_sync(0xffffffffu, v[E - 1] > nk);
    const int p = __ffs(gt) - 1;
    const unsigned long long carry = __shfl_up_sync(0xffffffffu, v[E - 1], 1);
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) cnt += (v[e] < nk);
#pragma unroll
    for (int e = E - 1; e >= 1; --e) {
        const unsigned long long prev = v[e - 1];
        if (lane > p) v[e] = prev;
        else if (lane == p) {
            if (e > cnt) v[e] = prev;
            else if (e == cnt) v[e] = nk;
        }
    }
    if (lane > p) v[0] = carry;
    else if (lane == p && cnt == 0) v[0] = nk;
    if (lane >= p) {
#pragma unroll
        for (int e = 0; e < E; ++e) l[lane * E + e] = v[e];
    }
    if ((uint32_t)lane == mthm1 / E) {
        const uint32_t pk = (uint32_t)(pick<E>(v, mthm1 % E) >> 32);
        if (pk < *pub) { *pub = pk; *reinterpret_cast<volatile uint32_t *>(gm) = pk; }
    }
    return __shfl_sync(0xffffffffu, pick<E>(v, capm1 % E), capm1 / E);
}

// Rows that passed the fast test for column index `colb + 2 * (lane & 3)` (one bit of m per lane): the whole warp
// handles them one by one -- exact surrogate key of the streaming scan, sorted insert into the query's list,
// new threshold.
template <bool COS, int E>
__device__ __noinline__ void batch_hits(unsigned m, float numf, uint32_t colb, const float2 *xaux /*shared: aux pairs of the tile*/,
                                        const uint32_t *xwords /*shared: live & filter words of the tile's 4 blocks*/,
                                        QState *wq /*the warp's 8 queries*/, unsigned long long *wl /*their lists*/, uint32_t slot0,
                                        uint32_t *gmw /*gmth of the warp's first query, this range*/, uint32_t gstride /*words per query*/, uint32_t mthm1,
                                        uint32_t capm1) {
    const int lane = threadIdx.x & 31;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t j = (uint32_t)src >> 2;
        const float nf = __shfl_sync(0xffffffffu, numf, src);
        const uint32_t col = colb + 2 * ((uint32_t)src & 3u);
        if (!((xwords[col >> 5] >> (col & 31u)) & 1u)) continue; // removed or filtered row (or past the end of the mirror)
        const float2 ax = xaux[col];
        QState &s = wq[j];
        float key;
        if (COS) key = (s.zero || ax.x == 0.f) ? 1.0f : -((nf * s.c_key) * ax.x);
        else key = (ax.y + s.qn2) - nf * s.c_dot2;
        const unsigned long long k64 = make_key64(key, slot0 + col);
        if (k64 < s.thr) {
            if (lane == 0) ++s.inserts;
            const unsigned long long last = list_insert<E>(wl + j * (32 * E), k64, lane, capm1, mthm1, gmw + (size_t)j * gstride, &s.pub);
            __syncwarp();
            if (lane == 0) {
                const unsigned long long gb = *reinterpret_cast<volatile unsigned long long *>(&s.gbound);
                s.thr = last < gb ? last : gb;
                s.T = fast_threshold<COS>(s);
            }
            __syncwarp();
        }
    }
}

// 16 (query, row) pairs of one tcgen05.ld: 8 column groups x 2 rows.  The fast test of all 16 pairs is
// straight-line code (independent chains, a max tree, one warp vote); only when some lane has a pair that
// passes are the pairs looked at, group of four by group of four, by the whole warp.
template <bool COS, bool FITS, int E, bool P16 = false>
__device__ __forceinline__ float score16(const uint32_t (&r)[32], const uint32_t (&rl)[32], const float4 (&ax)[8], int numc32, long long numc64,
                                         float T, float c2,
                                         uint32_t colp, const float2 *xaux, const uint32_t *xwords, QState *wq, unsigned long long *wl,
                                         uint32_t slot0, const float *Tsrc, uint32_t *gmw, uint32_t gstride, uint32_t mthm1, uint32_t capm1,
                                         bool seeding = false) {
    float numf[16], v[16];
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint32_t hi = r[4 * rep + e], lo = r[4 * rep + 2 + e];
            // num = 2 I + numc with I = 128 * D_hi + D_lo: exact in wrapping 32-bit arithmetic when |num| < 2^31
            if (P16) {
                // 16-bit rows: r = contraction with the HIGH bytes, rl = with the LOW bytes of the uncentred codes;
                // I = 256 I_H + I_L, num = 2 I + numc (numc = -65535 sum W), exact in 64-bit integers
                const long long ih = (long long)(int)hi * 128 + (long long)(int)lo;
                const long long il = (long long)(int)rl[4 * rep + e] * 128 + (long long)(int)rl[4 * rep + 2 + e];
                numf[2 * rep + e] = (float)(2 * (256 * ih + il) + numc64);
            } else if (FITS) numf[2 * rep + e] = (float)(int)(hi * 256u + (2u * lo + (uint32_t)numc32));
            else numf[2 * rep + e] = (float)(2 * ((long long)(int)hi * 128 + (long long)(int)lo) + numc64);
            // aux pairs {1/||x||, ||x||^2} of rows col, col + 1: cosine uses .x/.z, euclid .y/.w
            const float a = COS ? (e ? ax[rep].z : ax[rep].x) : (e ? ax[rep].w : ax[rep].y);
            v[2 * rep + e] = COS ? numf[2 * rep + e] * a : fmaf(numf[2 * rep + e], c2, -a);
        }
    }
    float m4[4];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) m4[g4] = fmaxf(fmaxf(v[4 * g4], v[4 * g4 + 1]), fmaxf(v[4 * g4 + 2], v[4 * g4 + 3]));
    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])); // fmaxf ignores NaN (dead rows)
    if (seeding && mthm1 < 8) {
        // Seeding pass with no threshold yet: all that is wanted from this tile is the key of SOME mth-best row to publish (any
        // mth rows of the range bound its mth-best key from above).  Without this, T = -inf sends every one of the 64 rows
        // x 8 queries of the call down the one-pair-at-a-time path (measured: ~100 us per launch, the fixed cost that bent
        // the multi-GPU curve).  Keep the best pair of each of the query's 4 lanes (the 4 lanes of a query hold 16 rows
        // each): threshold = the mth-largest of the 4 lane maxima (the smallest when mth > 4).
        const bool open = T == -INFINITY;
        const float m0 = mx == mx ? mx : -INFINITY;
        const float m1 = __shfl_xor_sync(0xffffffffu, m0, 1);
        const float hi01 = fmaxf(m0, m1), lo01 = fminf(m0, m1);
        const float hi23 = __shfl_xor_sync(0xffffffffu, hi01, 2), lo23 = __shfl_xor_sync(0xffffffffu, lo01, 2);
        // the four values of the query, ordered: s0 >= s1 >= s2 >= s3
        const float s0 = fmaxf(hi01, hi23), s3 = fminf(lo01, lo23);
        const float mid_a = fminf(hi01, hi23), mid_b = fmaxf(lo01, lo23);
        const float s1 = fmaxf(mid_a, mid_b), s2 = fminf(mid_a, mid_b);
        const float pick = mthm1 == 0 ? s0 : mthm1 == 1 ? s1 : mthm1 == 2 ? s2 : s3;
        if (open && pick > -INFINITY) T = pick - fabsf(pick) * 9.5367431640625e-7f - 1e-30f; // just below: the picked pair passes
    }
    if (__any_sync(0xffffffffu, mx > T)) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
            if (__any_sync(0xffffffffu, m4[g4] > T)) {
#pragma unroll
                for (int i = 4 * g4; i < 4 * g4 + 4; ++i) {
                    const unsigned m = __ballot_sync(0xffffffffu, v[i] > T);
                    if (m) batch_hits<COS, E>(m, numf[i], colp + (i >> 1) * 8 + (i & 1), xaux, xwords, wq, wl, slot0, gmw, gstride, mthm1, capm1);
                }
            }
        }
        T = *Tsrc; // thresholds may have tightened
    }
    return T;
}


// Shared bound of a query from the keys its R row ranges (CTAs) have published.  Any value B such that `need` = ceil(Kp / mth)
// different ranges have published a key <= B is valid: each of them holds mth rows at or below its key, so need * mth >= Kp
// rows lie at or below B.  The ranges are dealt into `need` groups (range r -> group r % need); B = the largest of the
// per-group minima.  With R = 148 ranges and Kp = 32 that is ~6x tighter than the plain maximum over all ranges.
// Layout of the published keys: gmth[query][group][SP] (SP = members per group rounded up to 4 or 8 words, padding and
// unpublished entries 0xFFFFFFFF), so a lane reads one group with one or two 16-byte loads and the loads of 8 (query, group)
// units are in flight together: one L2 round trip per 8 / ceil(need / 32) queries.
// Called by a poller warp for `nqv` consecutive valid queries; writes QState::gbound only (the owning epilogue warp folds
// it into its thresholds).  Returns true when every one of them has a bound.
__device__ __noinline__ bool poll_bounds(const uint32_t *gmth, uint32_t QS, uint32_t SP, uint32_t need, uint32_t qfirst, uint32_t nqv,
                                         QState *sq, const uint32_t *epi_done /*shared: finished epilogue warps*/) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lg = need <= 32 ? 0u : need <= 64 ? 1u : 2u; // 32-group slabs per query = 1 << lg
    const uint32_t QB = 8u >> lg;                                // queries per batch of 8 units
    bool all = true;
    for (uint32_t j0 = 0; j0 < nqv; j0 += QB) {
        if (*reinterpret_cast<const volatile uint32_t *>(epi_done) >= 8u) return true; // the CTA is done: do not hold up its exit
        uint4 x0[8], x1[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) {
            const uint32_t j = min(j0 + (u >> lg), nqv - 1);
            const uint32_t g = min(lane + 32u * (u & ((1u << lg) - 1u)), need - 1);
            const uint32_t *p = gmth + (size_t)(qfirst + j) * QS + (size_t)g * SP;
            x0[u] = __ldcg(reinterpret_cast<const uint4 *>(p));
            x1[u] = SP > 4 ? __ldcg(reinterpret_cast<const uint4 *>(p + 4)) : make_uint4(~0u, ~0u, ~0u, ~0u);
        }
        uint32_t acc = 0;
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) {
            const uint32_t gi = u & ((1u << lg) - 1u);
            const bool ok = lane + 32u * gi < need; // lanes past the last group do not vote in the maximum
            const uint32_t m = min(min(min(x0[u].x, x0[u].y), min(x0[u].z, x0[u].w)), min(min(x1[u].x, x1[u].y), min(x1[u].z, x1[u].w)));
            const uint32_t b = __reduce_max_sync(0xffffffffu, ok ? m : 0u);
            acc = gi == 0 ? b : max(acc, b);
            const uint32_t j = j0 + (u >> lg);
            if (gi == (1u << lg) - 1u && j < nqv) { // last slab of query j: acc = its bound (0xFFFFFFFF: some group has not published)
                if (acc == 0xFFFFFFFFu) all = false;
                else if (lane == 0) {
                    const unsigned long long gb = ((unsigned long long)acc << 32) | 0xFFFFFFFFull;
                    volatile unsigned long long *dst = &sq[j].gbound;
                    if (gb < *dst) *dst = gb;
                }
            }
        }
    }
    return all;
}

// next super tile with a live, unfiltered row; lanes 0..3 return the live words of its blocks.  Tiles are dealt to the CTAs
// of a query group on demand (one atomic counter per group, pre-set to 0xFFFFFFFF so that old + 1 is the tile): a CTA that
// reaches its SM late -- the previous call's finalize kernel may still hold it -- simply takes fewer tiles, and the tail
// of the launch is balanced to one tile.  Without a counter (ctr == nullptr) the CTA walks its fixed range [*, sup1).
// Only the producer warp walks; the epilogue warps learn each tile's index through the side-data ring.
// The counter is read one tile ahead (*pend = the grab issued during the previous call, lane 0): its L2 round trip runs
// under the previous tile's ring waits instead of delaying this tile's copy (measured: +4 % on a 10M-row launch otherwise).
__device__ __forceinline__ uint32_t grab_tile(uint32_t *ctr, int lane) {
    uint32_t t = 0;
    if (ctr && lane == 0) t = atomicAdd(ctr, 1u) + 1u;
    return t;
}
__device__ __forceinline__ uint32_t next_live_tile(const BatchArgs &a, uint32_t *ctr, uint32_t *pend, uint32_t sup, uint32_t sup1,
                                                   int lane, uint32_t *words) {
    for (;; ++sup) {
        if (ctr) {
            sup = __shfl_sync(0xffffffffu, *pend, 0);
            *pend = grab_tile(ctr, lane);
        }
        if (sup >= sup1) break;
        uint32_t lv = 0;
        if (lane < (int)kNB) {
            const uint32_t blk = sup * kNB + lane;
            if (blk < a.nblk) {
                lv = __ldg(a.live + blk);
                if (a.mask) lv &= __ldg(a.mask + blk);
            }
        }
        if (__ballot_sync(0xffffffffu, lv != 0)) { *words = lv; return sup; }
    }
    *words = 0;
    return sup1;
}

} // namespace

template <bool COS, int E, bool P16>
__global__ void __launch_bounds__(kBatchThreads, 1) batch_kernel(const BatchArgs a, const __grid_constant__ CUtensorMap tmap,
                                                                 const __grid_constant__ CUtensorMap tmapL) {
    constexpr uint32_t kHalves = P16 ? 2u : 1u; // 16-bit rows: every super tile is contracted twice (high bytes, low bytes)
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t b_full[kBatchMaxStages], b_empty[kBatchMaxStages], d_full[2], d_empty[2];
    __shared__ uint32_t s_first[kBatchMaxStages]; // first block of the super tile a stage belongs to (kNoBlock = end)
    __shared__ QState s_q[kBatchQueries];
    // per-tile side data, producer -> epilogue: the rows' aux pairs (bulk copy), the live words, the tile index
    __shared__ __align__(16) float2 s_xaux[kAuxSlots][kTileRows];
    __shared__ uint32_t s_xwords[kAuxSlots][kNB], s_xsup[kAuxSlots];
    __shared__ __align__(8) uint64_t x_full[kAuxSlots], x_empty[kAuxSlots];
    __shared__ uint32_t s_tmem;
    __shared__ int s_fits;
    __shared__ uint32_t s_epi_done; // epilogue warps that have finished (the poller warp's exit condition)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t C = a.C, S = a.stages;
    const uint32_t slc = a.slice;                        // chunks per K slice (= TMA box)
    const uint32_t nsl = (C + slc - 1) / slc;            // slices per super tile
    const uint32_t stage_bytes = kNB * slc * 512u;
    unsigned char *sB = smem;

    const uint32_t g = blockIdx.x % a.ngroups, r = blockIdx.x / a.ngroups; // (query group, row range)
    const uint32_t nsup = (a.nblk + kNB - 1) / kNB;
    const uint32_t per = (nsup + a.nranges - 1) / a.nranges;
    uint32_t *const ctr = a.tile_ctr ? a.tile_ctr + a.group0 + g : nullptr;
    const uint32_t sup0 = ctr ? 0u : min(nsup, r * per), sup1 = ctr ? nsup : min(nsup, sup0 + per);
    const uint32_t q0 = (a.group0 + g) * kBatchQueries;
    grid_launch_dependents(); // finalize_kernel may take its place on an SM as soon as one has room (it waits for this grid)
    if (a.trace && blockIdx.x == 0 && tid == 128) a.trace[7] = clock64(); // kernel entry, same SM and thread as stamps 0..3

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&d_full[d], 1); mbar_init(&d_empty[d], 8); }
        for (uint32_t x = 0; x < kAuxSlots; ++x) { mbar_init(&x_full[x], 1); mbar_init(&x_empty[x], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_fits = 1;
        s_epi_done = 0;
    }
    __syncthreads();
    // The launch is a programmatic dependent of prep_kernel (launch_batch): everything up to here, the TMEM allocation and the
    // producer warp's first row tiles (rows, live words and tile counters are older than prep) overlap prep's execution.  The
    // warps that read prep's output -- query headers and digits -- wait for it here; warp 0 (TMA producer) never does.
    uint32_t tmem = 0, tmemA = 0;
    if (warp != 0) {
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    grid_dependency_wait();
    if (tid >= 64 && tid < 64 + kBatchQueries) {
        const uint32_t q = q0 + (uint32_t)(tid - 64);
        QState &s = s_q[tid - 64];
        s.valid = q < a.nq;
        const PQHeader *hdr = reinterpret_cast<const PQHeader *>(a.pq + (size_t)(s.valid ? q : 0) * a.pq_stride);
        s.thr = s.valid ? kNoKey : 0ull;
        s.gbound = kNoKey;
        s.pub = 0xFFFFFFFFu;
        s.inserts = 0;
        // 16-bit: the header's numc belongs to centred codes (+sum W); the byte planes hold uncentred ones: -65535 sum W
        s.numc = P16 ? -65535ll * (long long)hdr->numc : (long long)hdr->numc;
        s.c_key = (float)hdr->c_key; s.c_dot2 = (float)(2.0 * hdr->c_dot); s.qn2 = (float)hdr->qn2;
        s.inv_ckey = hdr->c_key > 0.0 ? (float)(1.0 / hdr->c_key) : 0.f;
        s.zero = hdr->zero_query != 0;
        // |s| <= d + ||q||^2, |p| <= 2 sqrt(d) ||q|| (+ fixed-point rounding): 8 eps of their sum
        s.slack = (float)(4.76837158203125e-7 * ((double)a.dims + hdr->qn2 + 2.0 * sqrt((double)a.dims * hdr->qn2) + 1.0));
        s.T = s.valid ? -INFINITY : INFINITY;
        // |num| = M 2^F |x.q| <= M 2^F sqrt(d) ||q||: when that fits 31 bits the whole sum can run in wrapping
        // 32-bit arithmetic (exact mod 2^32, and the true value fits)
        const bool fits = 255.0 * ldexp(1.0, hdr->F) * sqrt((double)a.dims * hdr->qn2) < 2.0e9;
        if (s.valid && (!fits || P16)) atomicAnd(&s_fits, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kBatchThreads - 32) : "memory"); // query state and the TMEM address are in shared memory
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem = s_tmem;
    tmemA = tmem + 2 * kAccCols;

    // ---- A operand: TMEM lane L = 32 lq + 16 h + 8 plane + j holds digit plane `plane` of query 16 lq + 8 h + j
    if (warp >= 2 && warp < 6) {
        const uint32_t lq = (uint32_t)(warp & 3), L = lq * 32 + lane;
        const uint32_t q = q0 + lq * 16 + ((L >> 4) & 1u) * 8 + (L & 7u), plane = (L >> 3) & 1u;
        const unsigned char *src = a.pq + (size_t)(q < a.nq ? q : 0) * a.pq_stride + sizeof(PQHeader);
        // 2 chunks = 32 K-bytes = 8 columns = one MMA K step.  The digit loads of 4 K steps are issued together before their
        // stores (the stores are asm volatile with a memory clobber: one step at a time, every step paid a full L2 round
        // trip -- ~15 us of every launch for 768-dimension rows)
        constexpr uint32_t KB = 12;
        for (uint32_t ks0 = 0; ks0 < C / 2; ks0 += KB) {
            uint4 v0[KB], v1[KB];
#pragma unroll
            for (uint32_t u = 0; u < KB; ++u) {
                const uint32_t ks = ks0 + u;
                v0[u] = make_uint4(0, 0, 0, 0);
                v1[u] = v0[u];
                if (q < a.nq && ks < C / 2) {
                    v0[u] = __ldg(reinterpret_cast<const uint4 *>(src + ((size_t)(2 * ks) * 2 + plane) * 16));
                    v1[u] = __ldg(reinterpret_cast<const uint4 *>(src + ((size_t)(2 * ks + 1) * 2 + plane) * 16));
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < KB; ++u) {
                const uint32_t ks = ks0 + u;
                if (ks < C / 2) // warp-uniform
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                                     tmemA + ((lq * 32u) << 16) + ks * 8),
                                 "r"(v0[u].x), "r"(v0[u].y), "r"(v0[u].z), "r"(v0[u].w), "r"(v1[u].x), "r"(v1[u].y), "r"(v1[u].z), "r"(v1[u].w)
                                 : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // the query operand must be in TMEM before the first MMA -- a matter between the staging warps and the MMA warp.  The
    // TMA producer (warp 0) does not wait: the first row tiles are in flight while the digits are still being staged.
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(kBatchThreads - 32) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    if (warp == 0) {
        // ================================================================ TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            if (P16) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapL)) : "memory");
        }
        uint32_t t = 0, xt = 0, words; // stage counter, tile counter
        bool again = a.nranges > 1; // the first tile of a range is sent twice (seeding pass, see the epilogue)
        uint32_t pend = grab_tile(ctr, lane);
        for (uint32_t sup = next_live_tile(a, ctr, &pend, sup0, sup1, lane, &words);;
             sup = again ? sup : next_live_tile(a, ctr, &pend, sup + 1, sup1, lane, &words), again = false, ++xt) {
            // side data of the tile (or the end marker) for the epilogue warps
            const uint32_t x = xt % kAuxSlots;
            if (xt >= kAuxSlots) mbar_wait(&x_empty[x], ((xt / kAuxSlots) - 1) & 1u);
            if (lane < (int)kNB) s_xwords[x][lane] = words;
            __syncwarp();
            if (lane == 0) {
                if (sup < sup1) {
                    const uint32_t row0 = sup * kTileRows, nrows = min(kTileRows, a.nblk * 32 - row0);
                    s_xsup[x] = sup;
                    mbar_expect_tx(&x_full[x], nrows * 8u);
                    bulk_g2s(&s_xaux[x][0], reinterpret_cast<const float2 *>(a.aux) + row0, nrows * 8u, &x_full[x]);
                } else {
                    s_xsup[x] = kNoBlock;
                    mbar_arrive(&x_full[x]);
                }
            }
            const uint32_t n = sup < sup1 ? nsl * kHalves : 1u; // the end of the range travels through the ring as a sentinel stage
            for (uint32_t i = 0; i < n; ++i, ++t) {
                const uint32_t s = t % S;
                const uint32_t half = i / nsl, sl = i % nsl;
                if (lane == 0) {
                    if (t >= S) mbar_wait(&b_empty[s], ((t / S) - 1) & 1u);
                    if (sup < sup1) {
                        // one TMA tensor copy per stage: box (512 B, 4 blocks, slc chunks) of the 3-D view
                        // (512 B | block, stride C * 512 | chunk, stride 512) lands chunk-major, [chunk][block][512 B];
                        // blocks past the end of the mirror and chunks past the end of a row are zero-filled
                        s_first[s] = sup * kNB;
                        mbar_expect_tx(&b_full[s], stage_bytes);
                        asm volatile(
                            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                                smem_u32(sB + (size_t)s * stage_bytes)),
                            "l"(reinterpret_cast<uint64_t>(half ? &tmapL : &tmap)), "r"(smem_u32(&b_full[s])), "r"(0), "r"((int)(sup * kNB)),
                            "r"((int)(sl * slc))
                            : "memory");
                    } else {
                        s_first[s] = kNoBlock;
                        mbar_arrive(&b_full[s]);
                    }
                }
                __syncwarp();
            }
            if (sup >= sup1) break;
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        // The whole warp walks the pipeline in uniform control flow (operands of tcgen05.mma live in uniform
        // registers; computed by one divergent thread they cost a register-to-uniform "waterfall" per MMA, and
        // that issue loop, not the tensor core, paced the kernel); one elected lane issues.
        // D = S32, A = S8 (query digits, TMEM), B = U8 (codes, smem K-major), N = 128, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((kAccCols >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sb = smem_u32(sB);
        uint32_t t = 0;
        for (uint32_t tile = 0;; ++tile) {
            bool end = false;
            for (uint32_t i = 0; i < nsl * kHalves; ++i, ++t) {
                const uint32_t s = t % S;
                const uint32_t half = i / nsl, sl = i % nsl;
                // 8-bit: the two accumulator buffers alternate between tiles.  16-bit: buffer 0 takes the high-byte
                // contraction of every tile, buffer 1 the low-byte one; the epilogue drains each to registers as soon as
                // it completes, so the next tile's high-byte pass starts while this tile is still being scored
                const uint32_t d = P16 ? half : (tile & 1u);
                mbar_wait(&b_full[s], (t / S) & 1u);
                if (i == 0 && *reinterpret_cast<volatile uint32_t *>(&s_first[s]) == kNoBlock) { end = true; break; }
                if (sl == 0) {
                    if (P16) { if (tile >= 1) mbar_wait(&d_empty[d], (tile - 1) & 1u); }
                    else if (tile >= 2) mbar_wait(&d_empty[d], ((tile >> 1) - 1) & 1u);
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ks0 = sl * slc / 2, ks1 = min(C / 2, ks0 + slc / 2);
                const uint64_t db0 = umma_desc(sb + s * stage_bytes, kNB * 512u, 128);
                const uint32_t dacc = tmem + d * kAccCols;
                if (elect_one()) {
                    if (!(a.debug & 4u)) {
                        // one descriptor per stage, advanced by adding the K-step offset to its address field;
                        // the accumulate flag is an immediate (the first K step of a tile overwrites)
                        uint64_t db = db0;
                        uint32_t ta = tmemA + ks0 * 8;
                        uint32_t ks = ks0;
                        if (ks0 == 0) {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}\n" ::"r"(dacc),
                                "r"(ta), "l"(db), "r"(idesc), "r"(0u)
                                : "memory");
                            db += (2 * kNB * 512u) >> 4;
                            ta += 8;
                            ++ks;
                        }
#pragma unroll 4
                        for (; ks < ks1; ++ks) {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n\t}\n" ::"r"(dacc),
                                "r"(ta), "l"(db), "r"(idesc), "r"(0u)
                                : "memory");
                            db += (2 * kNB * 512u) >> 4;
                            ta += 8;
                        }
                    }
                    umma_commit(&b_empty[s]); // the stage may be refilled once these MMAs have read it
                    if (sl == nsl - 1) umma_commit(&d_full[d]); // this buffer's accumulators are complete
                }
                __syncwarp();
            }
            if (end) break;
        }
    } else if (warp >= kBatchPollWarp) {
        // ================================================================ bound pollers (2 warps, 32 queries each)
        // Read the keys the R row ranges of each query have published and keep QState::gbound current.  Done by the epilogue
        // warps themselves (round 1), a poll -- L2 round trips under full streaming load, ~14 us for 148 ranges -- kept the
        // accumulators undrained and stalled the ring: 2.8 us per tile on a 66-tile range against 2.0 us on a 528-tile one
        // with six polls each.  Here it costs the pipeline nothing and the bound is fresher.
        const uint32_t R = a.nranges;
        const uint32_t qp = (uint32_t)(warp - kBatchPollWarp) * 32u; // first query of this poller within the group
        const uint32_t nqv = q0 + qp < a.nq ? min(32u, a.nq - (q0 + qp)) : 0u;
        if (R > 1 && nqv) {
            constexpr uint32_t Kp = 32 * E;
            const uint32_t need = (Kp + a.mth - 1) / a.mth;
            // pause between rounds once every query has a bound: doubling from 0.5 us (the bound tightens fastest over the first
            // tiles) up to 2 us x (query groups per launch) -- with several groups the row tiles come from L2 and the polls
            // compete with them (measured, 1024 queries = 16 groups on 1.25 M rows: cap 16 / 64 / 256 / 1000 us -> 1.77 / 1.80 / 1.87 / 1.88 ms)
            const uint32_t pause_max = a.poll_ns ? a.poll_ns : (a.ngroups == 1 ? 1000u : 2000u * a.ngroups);
            uint32_t pause = min(500u, pause_max);
            while (*reinterpret_cast<volatile uint32_t *>(&s_epi_done) < 8u) {
                const bool all = poll_bounds(a.gmth, a.gm_stride, a.gm_sp, need, q0 + qp, nqv, &s_q[qp], &s_epi_done);
                // sleep in short slices: the CTA must not outlive its epilogue warps by a pause
                const uint32_t nap = all ? pause : 100u;
                if (all) pause = min(pause * 2u, pause_max);
                for (uint32_t t = 0; t < nap && *reinterpret_cast<volatile uint32_t *>(&s_epi_done) < 8u; t += 250u)
                    __nanosleep(all ? 250u : 100u);
            }
        }
    } else {
        // ================================================================ epilogue (8 warps, 8 queries each)
        // warp = (TMEM lane quarter lq, half hh): TMEM lanes 32 lq + 16 hh + {0..15} = both planes of 8 queries
        const uint32_t lq = (uint32_t)(warp & 3);       // TMEM lane quarter this warp may read
        const uint32_t hh = (uint32_t)(warp - 2) >> 2;  // which 16 lanes of the quarter
        const uint32_t c4 = (uint32_t)lane & 3u;        // column pair within each group of 8 columns
        const uint32_t qw = lq * 16 + hh * 8;           // the warp's first query within the group
        QState *qs = &s_q[qw + (lane >> 2)];
        const bool fits = s_fits != 0;
        const long long numc64 = qs->numc;
        const int numc32 = (int)numc64;
        const float c2 = qs->c_dot2;
        float T = qs->T;
        constexpr uint32_t Kp = 32 * E;
        QState *wq = &s_q[qw];
        // the warp's 8 candidate lists (sorted, Kp keys each) live behind the ring in dynamic shared memory
        unsigned long long *wl = reinterpret_cast<unsigned long long *>(smem + (size_t)S * stage_bytes) + (size_t)qw * Kp;
        for (uint32_t i = lane; i < 8 * Kp; i += 32) wl[i] = kNoKey;
        __syncwarp();
        // Bound shared by the R row ranges of a query: every range publishes the key of its mth-best row so far
        // (mth = ceil(Kp / R)); once all have one, at least R * mth >= Kp rows lie at or below the largest of them,
        // so that value bounds the query's Kp-th best key and no row above it can reach the candidate set.  This
        // keeps the number of rows that pass the fast test near Kp (1 + ln) per QUERY instead of per range.
        // To have the bound from the start, the first tile of a range is processed twice: once "seeding"
        // (admission by the mth-best key, cheap), then -- after every range of the group has published -- for real.
        const uint32_t R = a.nranges, mthm1 = a.mth - 1;
        const uint32_t need = (Kp + a.mth - 1) / a.mth, QS = a.gm_stride;
        // this range's slot in the published keys of the warp's first query: [query][group r % need][member r / need]
        uint32_t *gmw = a.gmth + ((size_t)min(q0 + qw, a.nq - 1) * QS + (size_t)(r % need) * a.gm_sp + r / need);
        // The bound itself is kept current by the poller warp (QState::gbound); this warp folds it into its queries'
        // thresholds at the start of every tile: shared-memory reads only.
        auto fold_bounds = [&]() -> bool {
            bool have = true;
            if (lane < 8 && wq[lane].valid) {
                QState &s = wq[lane];
                const unsigned long long gb = *reinterpret_cast<volatile unsigned long long *>(&s.gbound);
                have = gb != kNoKey;
                if (gb < s.thr) { s.thr = gb; s.T = fast_threshold<COS>(s); }
            }
            have = __all_sync(0xffffffffu, have);
            T = qs->T;
            return have;
        };
        bool seeding = R > 1;
        const bool idle = q0 + qw >= a.nq; // none of this warp's queries exists
        const bool tr = a.trace && blockIdx.x == 0 && warp == 4 && lane == 0; // warp 4 owns queries 0..7 of the group
        if (tr) a.trace[0] = clock64();
        for (uint32_t tile = 0;; ++tile) {
            const uint32_t d = P16 ? 0u : (tile & 1u), x = tile % kAuxSlots;
            // side data of the tile (aux pairs, live words): read in place; removed / filtered rows are rejected only
            // when they pass the fast test (as rare as for live rows), see batch_hits
            mbar_wait(&x_full[x], (tile / kAuxSlots) & 1u);
            const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&s_xsup[x]);
            if (cur == kNoBlock) break;
            if (idle) {
                // a warp whose 8 queries lie past the end of the batch (32 queries per call: half of the epilogue warps) only
                // keeps the pipeline's books: no accumulator reads, no scoring -- that work bought nothing and, at the power
                // cap this kernel runs into under sustained load, cost clock
                mbar_wait(&d_full[d], (P16 ? tile : (tile >> 1)) & 1u);
                if (lane == 0) mbar_arrive(&d_empty[d]);
                if (P16) {
                    mbar_wait(&d_full[1], tile & 1u);
                    if (lane == 0) mbar_arrive(&d_empty[1]);
                }
                if (lane == 0) mbar_arrive(&x_empty[x]);
                seeding = false;
                continue;
            }
            if (R > 1 && !seeding) fold_bounds();
            const uint32_t capm1 = seeding ? mthm1 : Kp - 1;
            mbar_wait(&d_full[d], (P16 ? tile : (tile >> 1)) & 1u); // 16-bit: d = 0, the high-byte buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tbase = tmem + ((lq * 32u + hh * 16u) << 16) + d * kAccCols;
            if (a.debug & 1u) { // profiling aid: drain the accumulators, skip the arithmetic
                uint32_t ra[32];
                for (int i = 0; i < 2; ++i) { tmem_ld_16x64(tbase + i * 64, ra); tmem_ld_wait(ra); }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[d]);
                if (P16) {
                    mbar_wait(&d_full[1], tile & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int i = 0; i < 2; ++i) { tmem_ld_16x64(tbase + kAccCols + i * 64, ra); tmem_ld_wait(ra); }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[1]);
                }
                if (lane == 0) mbar_arrive(&x_empty[x]);
                continue;
            }
            uint32_t ra[32], rb[32];
            float4 ax[8];
            auto load_ax = [&](uint32_t part) { // the aux pairs of this lane's two rows per 8-column group, straight from the side ring
#pragma unroll
                for (int rep = 0; rep < 8; ++rep) ax[rep] = *reinterpret_cast<const float4 *>(&s_xaux[x][part * 64 + rep * 8 + 2 * c4]);
            };
            const uint32_t slot0 = cur * kTileRows;
            auto score = [&](const uint32_t (&rr)[32], uint32_t part) {
                if (fits) T = score16<COS, true, E>(rr, rr, ax, numc32, numc64, T, c2, part * 64, s_xaux[x], s_xwords[x], wq, wl, slot0, &qs->T, gmw, QS, mthm1, capm1, seeding);
                else T = score16<COS, false, E>(rr, rr, ax, numc32, numc64, T, c2, part * 64, s_xaux[x], s_xwords[x], wq, wl, slot0, &qs->T, gmw, QS, mthm1, capm1, seeding);
            };
            if (P16) {
                // high-byte accumulators in buffer 0, low-byte ones in buffer 1: each is drained to registers and released
                // as soon as it completes, the scoring follows
                uint32_t rc[32], rd[32];
                tmem_ld_16x64(tbase, ra);
                tmem_ld_16x64(tbase + 64, rc);
                load_ax(0);
                tmem_ld_wait(ra);
                tmem_ld_wait(rc);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[0]); // high-byte accumulators are in registers
                mbar_wait(&d_full[1], tile & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld_16x64(tbase + kAccCols, rb);
                tmem_ld_16x64(tbase + kAccCols + 64, rd);
                tmem_ld_wait(rb);
                tmem_ld_wait(rd);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[1]); // low-byte accumulators too
                T = score16<COS, false, E, true>(ra, rb, ax, numc32, numc64, T, c2, 0, s_xaux[x], s_xwords[x], wq, wl, slot0, &qs->T, gmw, QS, mthm1, capm1, seeding);
                load_ax(1);
                T = score16<COS, false, E, true>(rc, rd, ax, numc32, numc64, T, c2, 64, s_xaux[x], s_xwords[x], wq, wl, slot0, &qs->T, gmw, QS, mthm1, capm1, seeding);
            } else {
            // both halves of this warp's part of the accumulator go to registers first, so the buffer returns to the
            // MMA warp before any scoring (a warp that has rows to insert would otherwise hold it)
            tmem_ld_16x64(tbase, ra);
            tmem_ld_16x64(tbase + 64, rb);
            load_ax(0);
            tmem_ld_wait(ra);
            tmem_ld_wait(rb);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[d]); // this warp's part of the buffer is in registers
            score(ra, 0);
            load_ax(1);
            score(rb, 1);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[x]); // the tile's aux pairs are no longer needed
            if (seeding) {
                // end of the seeding pass over the first tile: a range that could not seed publishes "no bound"
                seeding = false;
                if (lane < 8 && wq[lane].valid && wq[lane].pub == 0xFFFFFFFFu) {
                    wq[lane].pub = 0xFFFFFFFEu;
                    *reinterpret_cast<volatile uint32_t *>(gmw + (size_t)lane * QS) = 0xFFFFFFFEu;
                }
                __syncwarp();
                for (uint32_t i = lane; i < 8 * Kp; i += 32) wl[i] = kNoKey; // the tile comes again
                if (lane < 8) { wq[lane].thr = wq[lane].valid ? kNoKey : 0ull; wq[lane].T = wq[lane].valid ? -INFINITY : INFINITY; }
                __syncwarp();
                // all ranges of the group are normally co-resident (grid <= SM count) and publish within a tile time; the
                // bound is an optimisation only, so the wait is bounded (~4 ms) and the range simply goes on without it
                if (tr) a.trace[1] = clock64();
                int spin = 0;
                for (; !fold_bounds() && spin < 40000; ++spin) __nanosleep(100);
                if (tr) { a.trace[2] = clock64(); a.trace[5] = spin; }
            }
        }
        __syncwarp();
        if (tr) {
            a.trace[3] = clock64();
            uint32_t ins = 0;
            for (int j = 0; j < 8; ++j) ins += wq[j].inserts;
            a.trace[4] = ins;
            a.trace[6] = 0;
        }
        if (seeding && lane < 8 && wq[lane].valid) // a range without live rows: tell the others not to wait for it
            *reinterpret_cast<volatile uint32_t *>(gmw + (size_t)lane * QS) = 0xFFFFFFFEu;
        // hand the lists to finalize_kernel: cand[query][row range][Kp]
        __syncwarp();
        if (lane == 0) atomicAdd(&s_epi_done, 1u);
        for (uint32_t j = 0; j < 8; ++j) {
            if (!wq[j].valid) continue;
            unsigned long long *gb = a.cand + ((size_t)(q0 + qw + j) * a.nranges + r) * Kp;
            for (uint32_t i = lane; i < Kp; i += 32) gb[i] = wl[j * Kp + i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

// chunks per ring stage: measured on B200, the fewer stage hand-offs per super tile the better (a whole tile per
// stage with 2 stages beats 6 slices with 13 stages by 2x), so: the largest slice that still leaves >= 2 stages
uint32_t batch_slice_chunks(uint32_t C, uint32_t want, size_t smem_limit) {
    uint32_t s = want ? want : C;
    if (s > C) s = C;
    if (!want)
        while (s > 2 && (size_t)2 * kNB * s * 512u > smem_limit) s = ((s + 1) / 2 + 1) & ~1u;
    s &= ~1u;
    return s < 2 ? 2 : s;
}
uint32_t batch_stages(uint32_t slice, size_t smem_limit) {
    uint32_t s = (uint32_t)(smem_limit / (kNB * slice * 512u));
    return s > (uint32_t)kBatchMaxStages ? (uint32_t)kBatchMaxStages : s;
}
size_t batch_list_bytes(uint32_t keep) { return (size_t)kBatchQueries * keep * 8; }
size_t batch_smem_bytes(uint32_t slice, uint32_t stages, uint32_t keep) { return (size_t)stages * kNB * slice * 512u + batch_list_bytes(keep); }
uint32_t batch_max_chunks() { return 256 * 4 / 16; } // A lives in <= 256 TMEM columns: rows of at most 1024 bytes

// dynamic shared memory available to the ring: the CTA limit minus the kernel's static shared memory
size_t batch_dynamic_limit() {
    cudaFuncAttributes fa;
    size_t stat = 16 * 1024;
    if (cudaFuncGetAttributes(&fa, batch_kernel<true, 4, true>) == cudaSuccess) stat = fa.sharedSizeBytes;
    return 227 * 1024 - stat - 1024;
}

template <bool P16>
static cudaError_t configure_all(size_t max_smem) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(batch_kernel<true, 1, P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem))) return e;
    if ((e = cudaFuncSetAttribute(batch_kernel<true, 2, P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem))) return e;
    if ((e = cudaFuncSetAttribute(batch_kernel<true, 4, P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem))) return e;
    if ((e = cudaFuncSetAttribute(batch_kernel<false, 1, P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem))) return e;
    if ((e = cudaFuncSetAttribute(batch_kernel<false, 2, P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem))) return e;
    return cudaFuncSetAttribute(batch_kernel<false, 4, P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
}
cudaError_t batch_configure(size_t max_smem) {
    cudaError_t e = configure_all<false>(max_smem);
    return e ? e : configure_all<true>(max_smem);
}

// 3-D tensor map over the column-blocked mirror: (64 x u64 = one chunk of a block's 32 rows | block | chunk)
static cudaError_t make_tmap(const BatchArgs &a, const uint4 *codes, CUtensorMap *tm) {
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
    const cuuint32_t box[3] = {64, kNB, a.slice};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint4 *>(codes), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <bool P16>
static cudaError_t launch_variant(const BatchArgs &a, const CUtensorMap &tm, const CUtensorMap &tl, dim3 grid, size_t smem, cudaStream_t st) {
    const bool cos = a.metric == COSINE;
    // programmatic dependent of the prep_kernel before it in the stream (see the kernel's set-up)
    if (a.keep == 32) return cos ? launch_dependent(batch_kernel<true, 1, P16>, grid, kBatchThreads, smem, st, a, tm, tl)
                                 : launch_dependent(batch_kernel<false, 1, P16>, grid, kBatchThreads, smem, st, a, tm, tl);
    if (a.keep == 64) return cos ? launch_dependent(batch_kernel<true, 2, P16>, grid, kBatchThreads, smem, st, a, tm, tl)
                                 : launch_dependent(batch_kernel<false, 2, P16>, grid, kBatchThreads, smem, st, a, tm, tl);
    return cos ? launch_dependent(batch_kernel<true, 4, P16>, grid, kBatchThreads, smem, st, a, tm, tl)
               : launch_dependent(batch_kernel<false, 4, P16>, grid, kBatchThreads, smem, st, a, tm, tl);
}

cudaError_t launch_batch(const BatchArgs &a, cudaStream_t st) {
    CUtensorMap tm, tl;
    cudaError_t e = make_tmap(a, a.codes, &tm);
    if (e != cudaSuccess) return e;
    if ((e = make_tmap(a, a.codes_lo ? a.codes_lo : a.codes, &tl)) != cudaSuccess) return e;
    if (a.stages < 2 || a.stages > (uint32_t)kBatchMaxStages || (a.keep != 32 && a.keep != 64 && a.keep != 128)) return cudaErrorInvalidValue;
    if (a.slice < 2 || a.slice > a.C || (a.slice & 1u)) return cudaErrorInvalidValue;
    const size_t smem = batch_smem_bytes(a.slice, a.stages, a.keep);
    const dim3 grid(a.ngroups * a.nranges);
    return a.codes_lo ? launch_variant<true>(a, tm, tl, grid, smem, st) : launch_variant<false>(a, tm, tl, grid, smem, st);
}

} // namespace szg
