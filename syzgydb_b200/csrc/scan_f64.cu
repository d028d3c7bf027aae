// scan_f64.cu -- instantiates the scan kernel (scan_impl.cuh) for F64 records.
#include "kernels.h"

namespace szg {

cudaError_t launch_scan_f64(int mode, int nd, int grid, int threads, size_t smem, cudaStream_t st, const ScanArgs &a) {
    return launch_scan_t<F64>(mode, nd, grid, threads, smem, st, a);
}

cudaError_t launch_finalize_f64(int mode, uint32_t nq, cudaStream_t st, const FinalizeArgs &a) {
    return launch_finalize_t<F64>(mode, nq, st, a);
}

cudaError_t scan_attr_f64(size_t max_smem) { return scan_attr_t<F64>(max_smem); }

} // namespace szg
