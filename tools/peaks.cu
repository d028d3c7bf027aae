// peaks.cu -- measured ceilings of the two arithmetic pipes the search kernels are judged against on this GPU (diagnostic,
// not part of the library; bench.py reads the committed result profiles/r02_peaks.json):
//   int8 tensor pipe : every SM issues back-to-back tcgen05.mma.kind::i8 (M = 128, N = 128 or 256, K = 32) from resident
//                      shared-memory operands, no loads in the loop  ->  dense int8 TOP/s (2 * M * N * K per MMA)
//   IDP.4A pipe      : every SM runs independent dp4a.u32.s32 chains ->  integer-dot lane-ops/s (what scan_small's
//                      multi-query launches are limited by)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}

template <int N>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int reps, int batches, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem[]; // A: 2 K-halves x 16 groups x 128 B; B: 2 x (N / 8) x 128 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (4096 + 2 * (N / 8) * 128) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01020304u * (i + 1);
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
    const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = make_desc(s32(smem), 2048, 128);
    const uint64_t db = make_desc(s32(smem + 4096), (N / 8) * 128, 128);
    uint32_t phase = 0;
    for (int b = 0; b < batches; ++b) {
        if (tid == 0) {
            for (int r = 0; r < reps; ++r) {
                // two accumulator buffers in turn (N = 256: columns 0 / 256), like a double-buffered epilogue would leave them
                const uint32_t d = tm + (uint32_t)((r & 1) * (N == 256 ? 256 : 128));
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%4, %4, %4, %4}, p;\n\t}\n" ::"r"(d),
                    "l"(da), "l"(db), "r"(idesc), "r"(0u)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        }
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                s32(&bar)),
            "r"(phase)
            : "memory");
        phase ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tm + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v == 0x7FFFFFF1u) *sink = v;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

__global__ void __launch_bounds__(1024, 1) idp_peak_kernel(int iters, unsigned seed, unsigned *sink) {
    unsigned a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 7u;
    int acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a + j), "r"(b));
    }
    int t = 0;
    for (int j = 0; j < 8; ++j) t += acc[j];
    if (t == 0x12345) *sink = (unsigned)t;
}

// dependent fp64 add chain of one warp: cycles per DADD = the latency that bounds the exact (reference-order) distance of
// the survivors of a search: d sequential adds per running sum
__global__ void dadd_latency_kernel(int n, double x, double *out, long long *cycles) {
    double acc = x;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = __dadd_rn(acc, x);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { *out = acc; *cycles = t1 - t0; }
}

template <typename F>
static float time_ms(F launch, int rounds) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < rounds; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned *sink;
    cudaMalloc(&sink, 4);
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    {
        const int reps = 256, batches = 64;
        cudaFuncSetAttribute(mma_peak_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        cudaFuncSetAttribute(mma_peak_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        const float m128 = time_ms([&] { mma_peak_kernel<128><<<sms, 128, 16384>>>(reps, batches, sink); }, 5);
        const float m256 = time_ms([&] { mma_peak_kernel<256><<<sms, 128, 16384>>>(reps, batches, sink); }, 5);
        cudaError_t e = cudaDeviceSynchronize();
        const double mm = (double)sms * reps * batches;
        printf(", \"mma_status\": \"%s\", \"int8_tops_m128_n128\": %.1f, \"int8_tops_m128_n256\": %.1f", cudaGetErrorString(e),
               mm * 2.0 * 128 * 128 * 32 / (m128 * 1e-3) / 1e12, mm * 2.0 * 128 * 256 * 32 / (m256 * 1e-3) / 1e12);
        // a long run (~1 s) of the faster shape: what the pipe sustains under the power cap
        const int long_batches = (int)(1000.0f / m256 * batches);
        const float ml = time_ms([&] { mma_peak_kernel<256><<<sms, 128, 16384>>>(reps, long_batches, sink); }, 1);
        printf(", \"int8_tops_m128_n256_sustained_1s\": %.1f", (double)sms * reps * long_batches * 2.0 * 128 * 256 * 32 / (ml * 1e-3) / 1e12);
    }
    {
        const int iters = 20000;
        const float ms = time_ms([&] { idp_peak_kernel<<<sms * 2, 1024>>>(iters, 1u, sink); }, 5);
        cudaError_t e = cudaDeviceSynchronize();
        // lane-ops: one dp4a instruction of one thread = 1 lane-op (4 multiply-adds)
        const double lane_ops = (double)sms * 2 * 1024 * iters * 8;
        printf(", \"idp_status\": \"%s\", \"idp4a_lane_gops\": %.1f, \"idp4a_lane_ops_per_clk_per_sm\": %.1f", cudaGetErrorString(e),
               lane_ops / (ms * 1e-3) / 1e9, lane_ops / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    }
    {
        double *dout;
        long long *dc, hc = 0;
        cudaMalloc(&dout, 8);
        cudaMalloc(&dc, 8);
        dadd_latency_kernel<<<1, 32>>>(4096, 1e-9, dout, dc);
        dadd_latency_kernel<<<1, 32>>>(4096, 1e-9, dout, dc);
        cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
        printf(", \"dadd_dependent_latency_cycles\": %.2f", (double)hc / (4096.0 * 16));
    }
    printf("}\n");
    return 0;
}
