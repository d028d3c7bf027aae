// tmem_probe.cu -- microbenchmarks behind the design of the batched epilogue (batch_q8.cu):
//   1. which (TMEM lane, column) each register of tcgen05.ld.16x256b holds (found empirically: the PTX
//      manual is not available offline);
//   2. cycles per tcgen05.mma kind::i8 with A in TMEM (TS) or shared memory (SS) for N = 32/128/256, alone
//      and while four other warps stream accumulators out of TMEM with tcgen05.ld (do they share bandwidth?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            s32(bar)),
        "r"(parity)
        : "memory");
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
}

// ---------------------------------------------------------------- 1. layout of 16x256b
__global__ void __launch_bounds__(128, 1) layout_kernel(uint32_t *out /*[2 bases][32 threads][8 regs]*/) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    {   // value = (TMEM lane << 8) | column, written with the plain 32x32b shape (lane = thread, register = column)
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = ((uint32_t)tid << 8) | (uint32_t)i;
        for (int c8 = 0; c8 < 4; ++c8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                             tm + ((uint32_t)(warp * 32) << 16) + c8 * 8),
                         "r"(v[c8 * 8 + 0]), "r"(v[c8 * 8 + 1]), "r"(v[c8 * 8 + 2]), "r"(v[c8 * 8 + 3]), "r"(v[c8 * 8 + 4]),
                         "r"(v[c8 * 8 + 5]), "r"(v[c8 * 8 + 6]), "r"(v[c8 * 8 + 7])
                         : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) { // warp 1 may touch lanes 32..63
        for (int base = 0; base < 2; ++base) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(tm + ((uint32_t)(32 + base * 16) << 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; ++i) out[(base * 32 + lane) * 8 + i] = r[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

// ---------------------------------------------------------------- 2. MMA pacing vs concurrent tcgen05.ld
// warp 0 lane 0 issues `reps` x (K / 32) MMAs; warps 1..4 (if nld > 0) each run nld tcgen05.ld.32x32b.x32.
__global__ void __launch_bounds__(160, 1) pace_kernel(int ts, int N, int K, int reps, int nld, int ld16, long long *cycles, int altd = 0) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = K / 16;
    unsigned char *sA = smem;                     // [c][16 groups][128 B]
    unsigned char *sB = smem + (size_t)C * 2048;  // [c][N/8 groups][128 B]
    for (int i = tid; i < (C * 2048 + C * N * 16) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(i, 1, 2, 3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t tA = tm + 256; // A rows in columns [256, 256 + K / 4) when ts
    const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = clock64();
    if (tid == 0 && reps > 0) {
        for (int rep = 0; rep < reps; ++rep)
            for (int ks = 0; ks < C / 2; ++ks) {
                const uint64_t db = make_desc(s32(sB) + ks * 2 * (N * 16), N * 16, 128);
                const uint32_t acc = ks > 0;
                if (ts) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tm + (altd ? (ks & 1) * 128 : 0)),
                        "r"(tA + ks * 8), "l"(db), "r"(idesc), "r"((uint32_t)(altd ? ks > 1 : acc)), "r"(0u)
                        : "memory");
                } else {
                    const uint64_t da = make_desc(s32(sA) + ks * 2 * 2048, 2048, 128);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tm),
                        "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
                        : "memory");
                }
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        cycles[0] = clock64() - t0;
    }
    if (warp >= 1 && nld > 0) {
        uint32_t sink = 0;
        const uint32_t lq = (uint32_t)(warp & 3);
        for (int i = 0; i < nld; ++i) {
            if (ld16) {
                // 16x256b.x16: 16 lanes x 128 columns, 64 registers per thread
                uint32_t r[32];
                for (int half = 0; half < 2; ++half) {
                    for (int part = 0; part < 2; ++part) {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
                            "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                              "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                              "=r"(r[30]), "=r"(r[31])
                            : "r"(tm + ((lq * 32u + half * 16u) << 16) + (uint32_t)((i & 1) * 128 + part * 64)));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        for (int j = 0; j < 32; ++j) sink += r[j];
                    }
                }
            } else {
                uint32_t r[32];
                tmem_ld32(tm + ((lq * 32u) << 16) + (uint32_t)((i & 7) * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 32; ++j) sink += r[j];
            }
        }
        if (lane == 0) cycles[warp] = clock64() - t0;
        if (sink == 0x12345678u) cycles[8] = sink;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
    uint32_t *dout;
    cudaMalloc(&dout, 2 * 32 * 8 * 4);
    layout_kernel<<<1, 128>>>(dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("layout kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<uint32_t> o(2 * 32 * 8);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    for (int base = 0; base < 2; ++base) {
        printf("16x256b.x2 at lane base %d: thread -> (lane,col) per register\n", 32 + base * 16);
        for (int t = 0; t < 32; ++t) {
            printf("  t%02d:", t);
            for (int i = 0; i < 8; ++i) printf(" (%u,%u)", o[(base * 32 + t) * 8 + i] >> 8, o[(base * 32 + t) * 8 + i] & 255);
            printf("\n");
        }
    }
    long long *dcyc;
    cudaMalloc(&dcyc, 16 * 8);
    const int K = 256, reps = 12; // 96 MMAs
    cudaFuncSetAttribute(pace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int ts = 0; ts < 2; ++ts)
        for (int N : {32, 64, 128, 256})
            for (int mode = 0; mode < 3; ++mode) { // 0: MMA only, 1: + 32x32b loads, 2: + 16x256b loads
                const size_t smem = (size_t)(K / 16) * (2048 + N * 16);
                // loads sized to last about as long as the MMAs at 64 B/cycle: 4 warps x 4 KB per iteration
                const int nld = mode == 0 ? 0 : (mode == 1 ? 96 * 4 : 96);
                cudaMemset(dcyc, 0, 16 * 8);
                pace_kernel<<<1, 160, smem>>>(ts, N, K, reps, nld, mode == 2, dcyc);
                e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("pace kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
                long long c[8];
                cudaMemcpy(c, dcyc, sizeof c, cudaMemcpyDeviceToHost);
                printf("%s N=%3d %-14s: %6.1f cycles/MMA", ts ? "TS" : "SS", N, mode == 0 ? "mma only" : mode == 1 ? "+ld 32x32b" : "+ld 16x256b",
                       (double)c[0] / (reps * (K / 32)));
                if (nld) {
                    const double bytes = mode == 1 ? 4096.0 * nld : 4.0 * 4096.0 * nld;
                    printf("   loader warps: %lld %lld %lld %lld cycles, %.1f B/cycle/warp", c[1], c[2], c[3], c[4], bytes / (double)c[1]);
                }
                printf("\n");
            }
    // long runs (fixed costs amortised), one accumulator vs two alternating accumulators
    for (int ts = 0; ts < 2; ++ts)
        for (int N : {32, 64, 128, 256})
            for (int altd = 0; altd < 2; ++altd) {
                if (altd && (N > 128 || !ts)) continue;
                const size_t smem = (size_t)(K / 16) * (2048 + N * 16);
                cudaMemset(dcyc, 0, 16 * 8);
                const int lreps = 256;
                pace_kernel<<<1, 160, smem>>>(ts, N, K, lreps, 0, 0, dcyc, altd);
                e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("pace kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
                long long c[8];
                cudaMemcpy(c, dcyc, sizeof c, cudaMemcpyDeviceToHost);
                printf("LONG %s N=%3d %s: %6.1f cycles/MMA (%d MMAs)\n", ts ? "TS" : "SS", N, altd ? "2 accumulators" : "1 accumulator ",
                       (double)c[0] / (lreps * (K / 32)), lreps * (K / 32));
            }
    // loads alone
    for (int mode = 1; mode < 3; ++mode) {
        const int nld = mode == 1 ? 96 * 4 : 96;
        cudaMemset(dcyc, 0, 16 * 8);
        pace_kernel<<<1, 160, (size_t)(K / 16) * (2048 + 128 * 16)>>>(1, 128, K, 0, nld, mode == 2, dcyc);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pace kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
        long long c[8];
        cudaMemcpy(c, dcyc, sizeof c, cudaMemcpyDeviceToHost);
        const double bytes = mode == 1 ? 4096.0 * nld : 4.0 * 4096.0 * nld;
        printf("loads only %-12s: %lld %lld %lld %lld cycles, %.1f B/cycle/warp\n", mode == 1 ? "32x32b" : "16x256b", c[1], c[2], c[3], c[4],
               bytes / (double)c[1]);
    }
    return 0;
}
