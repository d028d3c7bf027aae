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

// Go's math.Acos (go 1.21 standard library, src/math/asin.go + atan.go: the Cephes algorithm, pure Go on amd64), in the
// same operation order with un-fused IEEE operations, so that the angular distance is the reference's to the last bit
// -- including its loss of relative accuracy for nearly parallel vectors (Acos = Pi/2 - Asin cancels; Asin takes
// Sqrt(1 - x*x) of a rounded product).  oracle/syzgy_oracle.c carries the same restatement and tests/test_oracle.py
// pins it against libm.
__device__ __forceinline__ double go_xatan(double x) {
    const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01, P2 = -7.500855792314704667340e+01,
                 P3 = -1.228866684490136173410e+02, P4 = -6.485021904942025371773e+01;
    const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02, Q2 = +4.328810604912902668951e+02,
                 Q3 = +4.853903996359136964868e+02, Q4 = +1.945506571482613964425e+02;
    double z = __dmul_rn(x, x);
    double p = __dadd_rn(__dmul_rn(P0, z), P1);
    p = __dadd_rn(__dmul_rn(p, z), P2);
    p = __dadd_rn(__dmul_rn(p, z), P3);
    p = __dadd_rn(__dmul_rn(p, z), P4);
    double q = __dadd_rn(z, Q0);
    q = __dadd_rn(__dmul_rn(q, z), Q1);
    q = __dadd_rn(__dmul_rn(q, z), Q2);
    q = __dadd_rn(__dmul_rn(q, z), Q3);
    q = __dadd_rn(__dmul_rn(q, z), Q4);
    z = __ddiv_rn(__dmul_rn(z, p), q);
    return __dadd_rn(__dmul_rn(x, z), x);
}
__device__ __forceinline__ double go_satan(double x) {
    const double Morebits = 6.123233995736765886130e-17, Tan3pio8 = 2.41421356237309504880;
    const double Pi = 3.141592653589793;
    if (x <= 0.66) return go_xatan(x);
    if (x > Tan3pio8) return __dadd_rn(__dsub_rn(Pi / 2, go_xatan(__ddiv_rn(1.0, x))), Morebits);
    return __dadd_rn(__dadd_rn(Pi / 4, go_xatan(__ddiv_rn(__dsub_rn(x, 1.0), __dadd_rn(x, 1.0)))), 0.5 * Morebits);
}
__device__ __forceinline__ double go_acos(double x) {
    const double Pi = 3.141592653589793;
    double as;
    if (x == 0.0) as = x;
    else {
        const bool sign = x < 0.0;
        const double ax = sign ? -x : x;
        if (!(ax <= 1.0)) return __longlong_as_double(0x7ff8000000000000ll); // |x| > 1 or NaN -> NaN (collection.go:831)
        double t = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(ax, ax)));
        t = ax > 0.7 ? __dsub_rn(Pi / 2, go_satan(__ddiv_rn(t, ax))) : go_satan(__ddiv_rn(ax, t));
        as = sign ? -t : t;
    }
    return __dsub_rn(Pi / 2, as);
}

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
    ExactAcc s = {0.0, 0.0, 0.0, 0.0};
    uint32_t i = 0;
    for (uint32_t c = 0; c < C && i < dims; ++c) {
        uint4 v = codes[chunk_at<QT>(slot, C, c)];
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
        return __ddiv_rn(go_acos(r), 3.141592653589793); // math.Acos(r > 1) = NaN
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
