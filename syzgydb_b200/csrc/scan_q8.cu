// scan_q8.cu -- instantiates the scan kernel (scan_impl.cuh) for Q8 records.
#include "kernels.h"
#include "scan_small.cuh"

namespace szg {

cudaError_t launch_scan_q8(int mode, int nd, int grid, int threads, size_t smem, cudaStream_t st, const ScanArgs &a) {
    return launch_scan_t<Q8>(mode, nd, grid, threads, smem, st, a);
}

cudaError_t launch_finalize_q8(int mode, uint32_t nq, cudaStream_t st, const FinalizeArgs &a) {
    return launch_finalize_t<Q8>(mode, nq, st, a);
}

cudaError_t scan_attr_q8(size_t max_smem) { return scan_attr_t<Q8>(max_smem); }

cudaError_t launch_scan_small_q8(int nd, uint32_t C, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    return launch_scan_small_t<Q8>(nd, C, grid, smem, st, a);
}

} // namespace szg
