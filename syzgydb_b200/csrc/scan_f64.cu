// scan_f64.cu -- instantiates the scan kernel (scan_impl.cuh) for F64 records.
#include "kernels.h"

namespace szg {

cudaError_t launch_scan_f64(int mode, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    return launch_scan_t<F64>(mode, grid, smem, st, a);
}

cudaError_t scan_attr_f64(size_t max_smem) {
    cudaError_t e;
#define SZG_ATTR(M)                                                                                            \
    e = cudaFuncSetAttribute(scan_kernel<F64, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);  \
    if (e != cudaSuccess) return e;
    SZG_ATTR(0) SZG_ATTR(1) SZG_ATTR(2) SZG_ATTR(3) SZG_ATTR(MODE_RADIUS)
#undef SZG_ATTR
    return cudaSuccess;
}

cudaError_t scan_occ_f64(int mode, size_t smem, int *bps) {
    switch (mode) {
    case 0: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, scan_kernel<F64, 0>, kScanThreads, smem);
    case 1: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, scan_kernel<F64, 1>, kScanThreads, smem);
    case 2: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, scan_kernel<F64, 2>, kScanThreads, smem);
    case 3: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, scan_kernel<F64, 3>, kScanThreads, smem);
    default: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, scan_kernel<F64, MODE_RADIUS>, kScanThreads, smem);
    }
}

} // namespace szg
