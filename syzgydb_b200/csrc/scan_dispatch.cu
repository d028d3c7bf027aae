// scan_dispatch.cu -- routes a scan launch to the instantiation of its quantization.
#include "kernels.h"

namespace szg {

#define SZG_DECL(l)                                                                              \
    cudaError_t launch_scan_##l(int, int, size_t, cudaStream_t, const ScanArgs &);               \
    cudaError_t scan_attr_##l(size_t);                                                           \
    cudaError_t scan_occ_##l(int, size_t, int *);
SZG_DECL(q4) SZG_DECL(q8) SZG_DECL(q16) SZG_DECL(f32) SZG_DECL(f64)
#undef SZG_DECL

cudaError_t launch_scan(int qt, int mode, int grid, size_t smem, cudaStream_t st, const ScanArgs &a) {
    switch (qt) {
    case Q4: return launch_scan_q4(mode, grid, smem, st, a);
    case Q8: return launch_scan_q8(mode, grid, smem, st, a);
    case Q16: return launch_scan_q16(mode, grid, smem, st, a);
    case F32: return launch_scan_f32(mode, grid, smem, st, a);
    case F64: return launch_scan_f64(mode, grid, smem, st, a);
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

cudaError_t scan_occupancy(int qt, int mode, size_t smem, int *bps) {
    switch (qt) {
    case Q4: return scan_occ_q4(mode, smem, bps);
    case Q8: return scan_occ_q8(mode, smem, bps);
    case Q16: return scan_occ_q16(mode, smem, bps);
    case F32: return scan_occ_f32(mode, smem, bps);
    case F64: return scan_occ_f64(mode, smem, bps);
    }
    return cudaErrorInvalidValue;
}

} // namespace szg
