// umma_test.cu -- validates the tcgen05 building blocks used by the batched path on one CTA:
// D[128 x 32] (s32, TMEM) = A[128 x K] (s8, K-major, no swizzle) * B[32 x K]^T (u8, K-major, no swizzle).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_test tools/umma_test.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B (128 B contiguous); SBO = bytes between 8-row groups,
// LBO = bytes between the two 16-byte K halves of one MMA (K = 32 bytes for 8-bit operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46; // descriptor version (Blackwell)
    return d;               // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128, 1) umma_kernel(const int8_t *A, const uint8_t *B, int32_t *D, int K) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = K / 16; // 16-byte chunks along K
    unsigned char *sA = smem;                      // [c][16 groups][8 rows][16 B]
    unsigned char *sB = smem + (size_t)C * 2048;   // [c][4 groups][8 rows][16 B]  (= the HBM block layout)
    for (int i = tid; i < 128 * C; i += 128) {
        int m = i / C, c = i % C;
        *reinterpret_cast<uint4 *>(sA + ((size_t)c * 16 + m / 8) * 128 + (m % 8) * 16) =
            *reinterpret_cast<const uint4 *>(A + (size_t)m * K + c * 16);
    }
    for (int i = tid; i < 32 * C; i += 128) {
        int n = i / C, c = i % C;
        *reinterpret_cast<uint4 *>(sB + ((size_t)c * 4 + n / 8) * 128 + (n % 8) * 16) =
            *reinterpret_cast<const uint4 *>(B + (size_t)n * K + c * 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy smem writes -> visible to the tensor core
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    // instruction descriptor: c_format S32 (2) @4, a_format S8 (1) @7, b_format U8 (0) @10, K-major both, N>>3 @17, M>>4 @24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        for (int ks = 0; ks < C / 2; ++ks) {
            uint64_t da = make_desc(s32(sA) + ks * 2 * 2048, 2048, 128);
            uint64_t db = make_desc(s32(sB) + ks * 2 * 512, 512, 128);
            uint32_t acc = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tm),
                "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            s32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
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
    for (int n = 0; n < 32; ++n) D[(size_t)(warp * 32 + lane) * 32 + n] = (int32_t)r[n];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tm));
}

// TS mode: A operand in TMEM.  Each of the 128 threads owns TMEM lane = A row m and stores its row's K bytes as
// K/4 consecutive 32-bit columns; an MMA (K = 32 bytes) consumes 8 columns.
__global__ void __launch_bounds__(128, 1) umma_ts_kernel(const int8_t *A, const uint8_t *B, int32_t *D, int K, int reps,
                                                         long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = K / 16;
    unsigned char *sB = smem; // [c][4 groups][8 rows][16 B]
    for (int i = tid; i < 32 * C; i += 128) {
        int n = i / C, c = i % C;
        *reinterpret_cast<uint4 *>(sB + ((size_t)c * 4 + n / 8) * 128 + (n % 8) * 16) =
            *reinterpret_cast<const uint4 *>(B + (size_t)n * K + c * 16);
    }
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
    const uint32_t tA = tm + 256; // columns [256, 256 + K/4)
    {   // row m = tid; columns in groups of 8
        const uint32_t *row = reinterpret_cast<const uint32_t *>(A + (size_t)tid * K);
        for (int c8 = 0; c8 < K / 32; ++c8) {
            uint32_t v[8];
            for (int i = 0; i < 8; ++i) v[i] = row[c8 * 8 + i];
            const uint32_t taddr = tA + ((uint32_t)(warp * 32) << 16) + c8 * 8;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                         "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
            for (int ks = 0; ks < C / 2; ++ks) {
                uint64_t db = make_desc(s32(sB) + ks * 2 * 512, 512, 128);
                uint32_t acc = ks > 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tm + (rep & 3) * 32),
                    "r"(tA + ks * 8), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
                    : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            s32(&bar))
        : "memory");
    if (tid == 0) *cycles = clock64() - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + ((reps - 1) & 3) * 32;
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
    for (int n = 0; n < 32; ++n) D[(size_t)(warp * 32 + lane) * 32 + n] = (int32_t)r[n];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
    const int K = 768;
    std::vector<int8_t> A(128 * K);
    std::vector<uint8_t> B(32 * K);
    srand(1);
    for (auto &x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto &x : B) x = (uint8_t)(rand() % 256);
    int8_t *dA; uint8_t *dB; int32_t *dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, 128 * 32 * 4);
    size_t smem = (size_t)(K / 16) * (2048 + 512);
    cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_kernel<<<1, 128, smem>>>(dA, dB, dD, K);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel status: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<int32_t> D(128 * 32);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            int32_t ref = 0;
            for (int k = 0; k < K; ++k) ref += (int32_t)A[m * K + k] * (int32_t)B[n * K + k];
            if (ref != D[m * 32 + n] && bad++ < 8) printf("mismatch m=%d n=%d got %d want %d\n", m, n, D[m * 32 + n], ref);
        }
    printf("SS %s: %d mismatches of %d\n", bad ? "FAIL" : "PASS", bad, 128 * 32);
    // ---- TS mode (A in TMEM) + timing of the MMA loop
    long long *dcyc; cudaMalloc(&dcyc, 8);
    cudaFuncSetAttribute(umma_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((K / 16) * 512));
    int bad2 = 0;
    for (int reps : {1, 64}) {
        cudaMemset(dD, 0xFF, 128 * 32 * 4);
        umma_ts_kernel<<<1, 128, (K / 16) * 512>>>(dA, dB, dD, K, reps, dcyc);
        e = cudaDeviceSynchronize();
        printf("TS kernel status (reps %d): %s\n", reps, cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        long long cyc = 0; cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int b2 = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                int32_t ref = 0;
                for (int k = 0; k < K; ++k) ref += (int32_t)A[m * K + k] * (int32_t)B[n * K + k];
                if (ref != D[m * 32 + n] && b2++ < 4) printf("TS mismatch m=%d n=%d got %d want %d\n", m, n, D[m * 32 + n], ref);
            }
        printf("TS reps=%d: %s (%d mismatches), %lld cycles for %d MMAs = %.1f cycles/MMA\n", reps, b2 ? "FAIL" : "PASS", b2, cyc,
               reps * (K / 32), (double)cyc / (reps * (K / 32)));
        bad2 += b2;
    }
    return (bad || bad2) ? 1 : 0;
}
