// readbw.cu -- read-only HBM bandwidth ceilings on this GPU (diagnostic, not part of the library):
//   (a) plain LDG.128 grid-stride sum, (b) per-warp cp.async.bulk ring with a trivial consumer.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/readbw tools/readbw.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void ldg_sum(const uint4 *p, size_t n, unsigned *out) {
    unsigned acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n; i += stride) acc += __ldg(p + i).x;
    if (acc == 0x12345678u) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int LIGHT>
__global__ void __launch_bounds__(512, 1) bulk_ring(const unsigned char *base, size_t ntiles, uint32_t tile_bytes, uint32_t S,
                                                   unsigned *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[16][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) {
        for (uint32_t s = 0; s < S; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[warp][s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned char *ring = smem + (size_t)warp * S * tile_bytes;
    const size_t stride = (size_t)gridDim.x * nw;
    size_t it = (size_t)blockIdx.x * nw + warp;
    auto issue = [&](uint32_t s) {
        if (it < ntiles && lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[warp][s])), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             s32(ring + (size_t)s * tile_bytes)),
                         "l"(base + it * tile_bytes), "r"(tile_bytes), "r"(s32(&bar[warp][s]))
                         : "memory");
        }
        it += stride;
    };
    for (uint32_t s = 0; s < S; ++s) issue(s);
    unsigned acc = 0, phases = 0, cs = 0;
    for (size_t ct = (size_t)blockIdx.x * nw + warp; ct < ntiles; ct += stride) {
        uint32_t par = (phases >> cs) & 1u;
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                         s32(&bar[warp][cs])),
                     "r"(par)
                     : "memory");
        phases ^= 1u << cs;
        const uint4 *sd = reinterpret_cast<const uint4 *>(ring + (size_t)cs * tile_bytes) + lane;
        if (LIGHT) acc += sd[0].x;
        else
            for (uint32_t c = 0; c < tile_bytes / 512; ++c) { uint4 v = sd[c * 32]; acc += v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
        issue(cs);
        cs = cs + 1 == S ? 0 : cs + 1;
    }
    if (acc == 0x12345678u) *out = acc;
}

int main(int argc, char **argv) {
    size_t gb = argc > 1 ? atol(argv[1]) : 8;
    size_t bytes = gb << 30;
    unsigned char *d;
    unsigned *out;
    cudaMalloc(&d, bytes);
    cudaMalloc(&out, 4);
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float ms;
    if (argc > 2) { // sustained mode: argv[2] = iterations; prints the bandwidth of every 8th launch
        int iters = atoi(argv[2]);
        cudaFuncSetAttribute(bulk_ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        for (int mode = 0; mode < 2; ++mode) {
            printf("sustained %s:", mode ? "bulk ring 16w S3 4096B" : "ldg128 4xSM");
            for (int it = 0; it < iters; ++it) {
                cudaEventRecord(e0);
                if (mode == 0) ldg_sum<<<sms * 4, 256>>>((const uint4 *)d, bytes / 16, out);
                else bulk_ring<0><<<sms, 512, 16 * 3 * 4096>>>(d, bytes / 4096, 4096, 3, out);
                cudaEventRecord(e1);
                if (it % 8 == 7) {
                    cudaEventSynchronize(e1);
                    cudaEventElapsedTime(&ms, e0, e1);
                    printf(" %.0f", bytes / ms / 1e6);
                }
            }
            cudaDeviceSynchronize();
            printf("\n");
        }
        return 0;
    }
    for (int blocks_per_sm : {2, 4, 8}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            ldg_sum<<<sms * blocks_per_sm, 256>>>((const uint4 *)d, bytes / 16, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("ldg128 grid=%dxSM x256: %.3f ms  %.1f GB/s\n", blocks_per_sm, ms, bytes / ms / 1e6);
    }
    cudaFuncSetAttribute(bulk_ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(bulk_ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    for (int warps : {4, 8, 16})
        for (int S : {2, 3, 4, 8})
            for (int tile : {2048, 4096, 8192, 16384}) {
                size_t smem = (size_t)warps * S * tile;
                if (smem > 224 * 1024) continue;
                for (int light = 0; light < 2; ++light) {
                    for (int rep = 0; rep < 3; ++rep) {
                        cudaEventRecord(e0);
                        if (light) bulk_ring<1><<<sms, warps * 32, smem>>>(d, bytes / tile, tile, S, out);
                        else bulk_ring<0><<<sms, warps * 32, smem>>>(d, bytes / tile, tile, S, out);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        cudaEventElapsedTime(&ms, e0, e1);
                    }
                    printf("bulk warps=%d S=%d tile=%d inflight=%zuKB %s: %.3f ms  %.1f GB/s\n", warps, S, tile, smem / 1024,
                           light ? "touch" : "read ", ms, bytes / ms / 1e6);
                }
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
