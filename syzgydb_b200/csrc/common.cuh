// common.cuh -- layout constants and device helpers shared by every kernel of the
// SyzgyDB search hot path (sm_100a only).
//
// HBM layout of the mirror ("column-blocked"):
//   rows are grouped in blocks of 32 (one warp lane per row); a row is cut in 16-byte
//   chunks; block b stores chunk c of its 32 rows contiguously:
//       codes[((b * C + c) * 32 + lane)]   (uint4 units, C = chunks per row)
//   so a warp-wide LDG.128 of "chunk c of my row" reads one contiguous 512-byte span,
//   and 8 rows x 16 B is exactly one UMMA K-major core matrix (used by the batched
//   tensor-core path).  Quantized rows (4/8/16-bit) use exactly this; float rows group
//   their chunks by 8 (chunk_index_grouped below).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace szg {

// Launch of the kernel that ends a search pass (finalize) as a programmatic dependent of the scan before it in the stream: its
// CTAs may become resident as soon as every CTA of the scan has executed griddepcontrol.launch_dependents (the scans do so at
// their start) and an SM has room, and they block in griddepcontrol.wait -- the first thing the kernel does -- until the scan
// has completed and its writes are visible.  What it buys: the launch latency of the dependent disappears behind the scan,
// and, with several calls in flight, this call's last kernel is already queued on the SMs when the scan's CTAs retire
// instead of lining up behind the next call's full-device scan.  SZG_PDL=0 turns the attribute off (plain stream order).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    static const bool pdl = !(getenv("SZG_PDL") && atoi(getenv("SZG_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// device side of the above
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int kRowsPerBlock = 32;
constexpr int kChunkBytes = 16;
// The fixed-point query coefficient W_i = round(q_i * 2^F) is split in ND signed base-128 digits
// (|W| < 2^(7 ND)).  ND = 2 is the fast path (8 IDP.4A per 16 codes instead of 12: the integer-dot
// pipe and board power, not HBM, limited the 3-digit kernel); ND = 3 is the precise path a query
// is re-run with when the 2-digit error bound cannot certify its result.
constexpr int kMaxDigits = 3;

enum QuantType : int { Q4 = 0, Q8 = 1, Q16 = 2, F32 = 3, F64 = 4 };
enum Metric : int { EUCLID = 0, COSINE = 1 };

__host__ __device__ inline int quant_bits(int qt) { return qt == Q4 ? 4 : qt == Q8 ? 8 : qt == Q16 ? 16 : qt == F32 ? 32 : 64; }
// elements per 16-byte chunk
__host__ __device__ inline int elems_per_chunk(int qt) { return 128 / quant_bits(qt); }
// bytes of prepared-query payload per chunk (digits or converted query values)
__host__ __device__ inline int pq_bytes_per_chunk(int qt, int nd) {
    return qt == Q4 ? 2 * nd * 16 : qt == Q8 ? nd * 16 : qt == Q16 ? nd * 8 : 16;
}

// Header of one prepared query (device memory, followed by the per-chunk payload).
// Quantized rows: x.q = num * c_dot with num = 2*I + numc, I = sum_i code_i * W_i (exact integer).
//   cosine key  = -(num * c_key) * (1/||x||)            (c_key = c_dot / ||q||)
//   euclid key  = ||x||^2 + qn2 - 2 * num * c_dot       (squared distance)
// |key - true key| <= e_abs + e_rel * |key| is a rigorous bound (fixed-point rounding of the query,
// fp32 rounding of aux values and of the key); finalize uses it to certify the candidate set.
struct __align__(16) PQHeader {
    double numc;       // -M * sum(W) (4/8-bit), +sum(W) (16-bit, codes stored centred)
    double c_dot;      // 1 / (M * 2^F); float rows: 1
    double c_key;      // cosine: c_dot / ||q|| (0 for a zero query); euclid: unused
    double qn2;        // ||q||^2
    double e_abs;      // surrogate error bound, absolute part
    double e_rel;      // surrogate error bound, relative part
    double radius_key; // radius mode: surrogate threshold (key <= radius_key is a candidate)
    double radius;     // radius mode: exact threshold
    int F;
    int zero_query;    // ||q|| == 0
    int nd;
    int pad;
};
static_assert(sizeof(PQHeader) % 16 == 0, "payload must stay 16-byte aligned");

// ---- loads -------------------------------------------------------------------------
// streaming 128-bit load: read-only path, do not allocate in L1 (each byte is used once)
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// coherent load that bypasses L1 (data written by other CTAs of the same launch)
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p) {
    unsigned long long r;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// ---- integer dot products ------------------------------------------------------------
// u8 x s8 -> s32   (SASS IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// s16x2 (a) x s8 bytes {0,1} / {2,3} of b   (SASS IDP.2A.LO/HI.S16.S8)
__device__ __forceinline__ int dp2a_lo_ss(int a, int b, int c) {
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(int a, int b, int c) {
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- selection keys ------------------------------------------------------------------
// Monotone map float -> uint32 (smaller float => smaller uint); NaN => worst.
__device__ __forceinline__ uint32_t ordered_key(float f) {
    uint32_t u = __float_as_uint(f);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
    return (f != f) ? 0xFFFFFFFEu : u;
}
__device__ __forceinline__ float key_to_float(uint32_t u) {
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}
constexpr unsigned long long kNoKey = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ unsigned long long make_key64(float key, uint32_t slot) {
    return ((unsigned long long)ordered_key(key) << 32) | slot;
}

// ---- lexicographic order of decimal ids (sort.Strings, spanfile.go:540-560) -----------
__host__ __device__ inline int dec_digits(unsigned long long v) {
    int n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}
// a * 10^k <= b without overflow surprises (k <= 19)
__host__ __device__ inline bool scaled_le(unsigned long long a, int k, unsigned long long b) {
    unsigned long long p = 1;
    for (int i = 0; i < k; ++i) p *= 10;
    unsigned long long hi;
#ifdef __CUDA_ARCH__
    hi = __umul64hi(a, p);
#else
    hi = (unsigned long long)(((unsigned __int128)a * p) >> 64);
#endif
    if (hi) return false; // a * 10^k >= 2^64 > b
    return a * p <= b;
}
__host__ __device__ inline bool lex_less_u64(unsigned long long a, unsigned long long b) {
    if (a == b) return false;
    const int da = dec_digits(a), db = dec_digits(b);
    if (da == db) return a < b;
    // pad the shorter one with zeros; a proper prefix sorts first
    if (da < db) return scaled_le(a, db - da, b);
    return !scaled_le(b, da - db, a);
}

// ---- synthetic data generator (same as oracle/syzgy_oracle.c orc_rand_u64) ------------
__host__ __device__ inline unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline unsigned long long rand_u64(unsigned long long seed, unsigned long long ctr) {
    return mix64(seed + 0x9E3779B97F4A7C15ull * (ctr + 1));
}

// ---- storage transform of one 16-byte chunk --------------------------------------------
// on-disk bytes (stream 1: 16/32/64-bit big-endian) -> HBM representation:
//   Q4/Q8: verbatim; Q16: little-endian int16 of (u - 32768); F32/F64: little-endian IEEE.
__host__ __device__ inline void chunk_to_device(int qt, unsigned char *b /*16 bytes, in place*/) {
    if (qt == Q16) {
        for (int e = 0; e < 8; ++e) {
            unsigned u = ((unsigned)b[2 * e] << 8) | b[2 * e + 1];
            u ^= 0x8000u;
            b[2 * e] = (unsigned char)(u & 0xFF);
            b[2 * e + 1] = (unsigned char)(u >> 8);
        }
    } else if (qt == F32) {
        for (int e = 0; e < 4; ++e) {
            unsigned char t0 = b[4 * e], t1 = b[4 * e + 1];
            b[4 * e] = b[4 * e + 3]; b[4 * e + 1] = b[4 * e + 2];
            b[4 * e + 2] = t1; b[4 * e + 3] = t0;
        }
    } else if (qt == F64) {
        for (int e = 0; e < 2; ++e)
            for (int i = 0; i < 4; ++i) {
                unsigned char t = b[8 * e + i];
                b[8 * e + i] = b[8 * e + 7 - i];
                b[8 * e + 7 - i] = t;
            }
    }
}
// the transform is an involution for every type (swap / swap+xor), so it also maps back
__host__ __device__ inline void chunk_to_disk(int qt, unsigned char *b) {
    if (qt == Q16) {
        for (int e = 0; e < 8; ++e) {
            unsigned u = ((unsigned)b[2 * e + 1] << 8) | b[2 * e];
            u ^= 0x8000u;
            b[2 * e] = (unsigned char)(u >> 8);
            b[2 * e + 1] = (unsigned char)(u & 0xFF);
        }
    } else {
        chunk_to_device(qt, b);
    }
}

__device__ __forceinline__ size_t chunk_index(uint32_t slot, uint32_t C, uint32_t c) {
    return ((size_t)(slot >> 5) * C + c) * 32 + (slot & 31);
}
// Float rows (32/64-bit collections) are never an MMA operand, but they are what candidate lists of the LSH index gather
// (BASELINE configs[2]: 3072-byte fp64 rows), and in the layout above a row's 16-byte chunks lie 512 bytes apart: every
// chunk of a gathered row costs a 32-byte sector of its own (measured 7.2x the row bytes from DRAM).  So float rows keep
// their chunks in GROUPS of 8 -- 128 contiguous bytes per row, i.e. whole lines for a gather:
//     codes[(((b * C/8 + c/8) * 32 + lane) * 8 + ((c % 8) ^ (lane % 8)))]          (C is rounded up to a multiple of 8)
// A block is still one contiguous span of C * 512 bytes and a tile of 8 n chunks one of 4096 n bytes, so the streaming scan
// moves it with the same bulk copies; the XOR swizzle spreads the 8 lanes of a quarter warp over all 32 banks when they
// read "chunk j of my row" from the staged tile with LDS.128.
constexpr int kGroupChunks = 8;
__host__ __device__ inline bool grouped_layout(int qt) { return qt >= F32; }
__device__ __forceinline__ size_t chunk_index_grouped(uint32_t slot, uint32_t C, uint32_t c) {
    const uint32_t lane = slot & 31;
    return ((((size_t)(slot >> 5) * (C >> 3) + (c >> 3)) * 32 + lane) << 3) + ((c & 7u) ^ (lane & 7u));
}
template <int QT>
__device__ __forceinline__ size_t chunk_at(uint32_t slot, uint32_t C, uint32_t c) {
    return QT >= F32 ? chunk_index_grouped(slot, C, c) : chunk_index(slot, C, c);
}
__device__ __forceinline__ size_t chunk_at_rt(int qt, uint32_t slot, uint32_t C, uint32_t c) {
    return qt >= F32 ? chunk_index_grouped(slot, C, c) : chunk_index(slot, C, c);
}

} // namespace szg
