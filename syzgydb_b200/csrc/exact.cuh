// exact.cuh -- fp64 distance of one mirrored record to the query, in the reference's
// operation order (sequential over dimensions, separate multiply and add, no FMA):
//   decodeVector + dequantize   collection.go:768-794, quantization.go:25-36
//   euclideanDistance           collection.go:812-819
//   angularDistance             collection.go:821-832
// The __d*_rn intrinsics are never contracted into FMAs by nvcc, which mirrors Go/amd64.
// dequantize's (float64(v)/float64(maxInt))*2-1 is a host-built lookup table (IEEE
// double ops on the host give the same bits as Go), so no division runs per element.
#pragma once
#include "common.cuh"

namespace szg {

struct ExactAcc {
    double dot, m1, m2, sum;
};

template <int METRIC>
__device__ __forceinline__ void exact_step(ExactAcc &s, double qi, double x) {
    if (METRIC == COSINE) {
        s.dot = __dadd_rn(s.dot, __dmul_rn(qi, x));
        s.m1 = __dadd_rn(s.m1, __dmul_rn(qi, qi));
        s.m2 = __dadd_rn(s.m2, __dmul_rn(x, x));
    } else {
        double diff = __dsub_rn(qi, x);
        s.sum = __dadd_rn(s.sum, __dmul_rn(diff, diff));
    }
}

template <int QT, int METRIC>
__device__ double exact_distance_impl(const uint4 *__restrict__ codes, uint32_t C, uint32_t dims,
                                      const double *__restrict__ lut, const double *__restrict__ q,
                                      uint32_t slot) {
    const uint4 *p = codes + chunk_index(slot, C, 0);
    ExactAcc s = {0.0, 0.0, 0.0, 0.0};
    uint32_t i = 0;
    for (uint32_t c = 0; c < C && i < dims; ++c) {
        uint4 v = p[(size_t)c * 32];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (QT == Q4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t byte = (w[k] >> (8 * b)) & 0xFF;
                    if (i < dims) exact_step<METRIC>(s, q[i], lut[byte >> 4]); // even index: high nibble
                    ++i;
                    if (i < dims) exact_step<METRIC>(s, q[i], lut[byte & 0x0F]);
                    ++i;
                }
        } else if (QT == Q8) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (i < dims) exact_step<METRIC>(s, q[i], lut[(w[k] >> (8 * b)) & 0xFF]);
                    ++i;
                }
        } else if (QT == Q16) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t u = ((w[k] >> (16 * h)) & 0xFFFF) ^ 0x8000u; // stored centered
                    if (i < dims) exact_step<METRIC>(s, q[i], lut[u]);
                    ++i;
                }
        } else if (QT == F32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i < dims) exact_step<METRIC>(s, q[i], (double)__uint_as_float(w[k])); // widened, quantization.go:27-28
                ++i;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (i < dims) exact_step<METRIC>(s, q[i], __hiloint2double((int)w[2 * k + 1], (int)w[2 * k]));
                ++i;
            }
        }
    }
    if (METRIC == COSINE) {
        if (s.m1 == 0.0 || s.m2 == 0.0) return 1.0; // collection.go:828-830
        double r = __ddiv_rn(s.dot, __dmul_rn(__dsqrt_rn(s.m1), __dsqrt_rn(s.m2)));
        return __ddiv_rn(acos(r), 3.141592653589793); // acos(r > 1) = NaN, as Go math.Acos
    }
    return __dsqrt_rn(s.sum);
}

template <int QT>
__device__ __forceinline__ double exact_distance(const uint4 *codes, uint32_t C, uint32_t dims, int metric,
                                                 const double *lut, const double *q, uint32_t slot) {
    return metric == COSINE ? exact_distance_impl<QT, COSINE>(codes, C, dims, lut, q, slot)
                            : exact_distance_impl<QT, EUCLID>(codes, C, dims, lut, q, slot);
}

// surrogate key -> distance in the reference's unit (SZG_F_NO_FP64_VERIFY)
__device__ __forceinline__ double key_to_distance(int metric, float key) {
    if (metric == COSINE) {
        double c = -(double)key;
        c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
        return acos(c) / 3.141592653589793;
    }
    return sqrt(key > 0.f ? (double)key : 0.0);
}

} // namespace szg
