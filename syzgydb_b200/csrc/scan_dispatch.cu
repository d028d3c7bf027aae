// scan_dispatch.cu -- routes a scan launch to the instantiation of its quantization.
#include "kernels.h"

namespace szg {

#define SZG_DECL(l)                                                                              \
    cudaError_t launch_scan_##l(int, int, int, int, size_t, cudaStream_t, const ScanArgs &);     \
    cudaError_t launch_finalize_##l(int, uint32_t, cudaStream_t, const FinalizeArgs &);           \
    cudaError_t scan_attr_##l(size_t);
SZG_DECL(q4) SZG_DECL(q8) SZG_DECL(q16) SZG_DECL(f32) SZG_DECL(f64)
#undef SZG_DECL

cudaError_t launch_scan(int qt, int mode, int nd, int grid, int threads, size_t smem, cudaStream_t st,
                        const ScanArgs &a) {
    switch (qt) {
    case Q4: return launch_scan_q4(mode, nd, grid, threads, smem, st, a);
    case Q8: return launch_scan_q8(mode, nd, grid, threads, smem, st, a);
    case Q16: return launch_scan_q16(mode, nd, grid, threads, smem, st, a);
    case F32: return launch_scan_f32(mode, nd, grid, threads, smem, st, a);
    case F64: return launch_scan_f64(mode, nd, grid, threads, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_scan_small_q4(int, uint32_t, int, size_t, cudaStream_t, const ScanArgs &);
cudaError_t launch_scan_small_q8(int, uint32_t, int, size_t, cudaStream_t, const ScanArgs &);
cudaError_t launch_scan_small_q16(int, uint32_t, int, size_t, cudaStream_t, const ScanArgs &);
cudaError_t launch_scan_small(int qt, int nd, uint32_t C, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    switch (qt) {
    case Q4: return launch_scan_small_q4(nd, C, grid, smem, st, a);
    case Q8: return launch_scan_small_q8(nd, C, grid, smem, st, a);
    case Q16: return launch_scan_small_q16(nd, C, grid, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_finalize(int qt, int mode, uint32_t nq, cudaStream_t st, const FinalizeArgs &a) {
    switch (qt) {
    case Q4: return launch_finalize_q4(mode, nq, st, a);
    case Q8: return launch_finalize_q8(mode, nq, st, a);
    case Q16: return launch_finalize_q16(mode, nq, st, a);
    case F32: return launch_finalize_f32(mode, nq, st, a);
    case F64: return launch_finalize_f64(mode, nq, st, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t scan_configure(int qt, size_t max_smem) {
    switch (qt) {
    case Q4: return scan_attr_q4(max_smem);
    case Q8: return scan_attr_q8(max_smem);
    case Q16: return scan_attr_q16(max_smem);
    case F32: return scan_attr_f32(max_smem);
    case F64: return scan_attr_f64(max_smem);
    }
    return cudaErrorInvalidValue;
}

} // namespace szg
