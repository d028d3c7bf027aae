// scan_q4.cu -- instantiates the scan kernel (scan_impl.cuh) for Q4 records.
#include "kernels.h"
#include "scan_small.cuh"

namespace szg {

cudaError_t launch_scan_q4(int mode, int nd, int grid, int threads, size_t smem, cudaStream_t st, const ScanArgs &a) {
    return launch_scan_t<Q4>(mode, nd, grid, threads, smem, st, a);
}

cudaError_t launch_finalize_q4(int mode, uint32_t nq, cudaStream_t st, const FinalizeArgs &a) {
    return launch_finalize_t<Q4>(mode, nq, st, a);
}

cudaError_t scan_attr_q4(size_t max_smem) { return scan_attr_t<Q4>(max_smem); }

cudaError_t launch_scan_small_q4(int nd, uint32_t C, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    return launch_scan_small_t<Q4>(nd, C, grid, smem, st, a);
}

} // namespace szg
