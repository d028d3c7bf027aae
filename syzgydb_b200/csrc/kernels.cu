// kernels.cu -- query preparation, mirror upload / layout transform, synthetic fill,
// fp64 re-score (K3) and the cross-rank top-k merge (K5).
#include "kernels.h"

namespace szg {

// ======================================================================= query prep
// One CTA per query.  Turns the float64 query (never quantized, collection.go:596) into
// the payload the scan kernel streams against:
//   4/8/16-bit rows: W_i = round(q_i 2^F), |W_i| < 2^(7 nd), as nd signed base-128 digits laid
//   out per 16-byte chunk of codes (both metrics use the same payload: euclid is evaluated as
//   ||x||^2 - 2 x.q + ||q||^2 with ||x||^2 precomputed per row);
//   32/64-bit rows: the query converted to fp32 / kept fp64, padded per chunk.
// The header carries the scale factors and the rigorous surrogate error bound.
__global__ void __launch_bounds__(256) prep_kernel(const PrepArgs a) {
    grid_launch_dependents(); // batch_kernel (a programmatic dependent) sets itself up and prefetches rows while this runs
    __shared__ double s_red[2][8];
    __shared__ double s_out[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *q = a.queries + (size_t)blockIdx.x * a.dims;
    unsigned char *out = a.pq + (size_t)blockIdx.x * a.pq_stride;
    PQHeader *h = reinterpret_cast<PQHeader *>(out);
    unsigned char *payload = out + sizeof(PQHeader);
    const bool quantized = a.qt <= Q16;
    const int nd = a.nd;
    // payload layout: normally the collection's own; the batched 16-bit path wants the digits laid out like an
    // 8-bit row (16 dimensions per chunk) because its operand is the byte-planar copy of the codes
    const int lqt = a.planar16 ? (int)Q8 : a.qt;

    const uint32_t n16 = (a.C * (uint32_t)pq_bytes_per_chunk(lqt, nd) + 15) / 16;
    for (uint32_t i = tid; i < n16; i += 256) reinterpret_cast<uint4 *>(payload)[i] = make_uint4(0, 0, 0, 0);

    // pass 1: max |q|, sum q^2
    double mq = 0.0, sq = 0.0;
    for (uint32_t i = tid; i < a.dims; i += 256) {
        double qi = q[i];
        double aq = fabs(qi);
        if (aq == aq && aq > mq) mq = aq;
        sq += qi * qi;
    }
    for (int o = 16; o; o >>= 1) {
        mq = fmax(mq, __shfl_xor_sync(0xffffffffu, mq, o));
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) { s_red[0][warp] = mq; s_red[1][warp] = sq; }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0, a1 = 0;
        for (int w = 0; w < 8; ++w) { a0 = fmax(a0, s_red[0][w]); a1 += s_red[1][w]; }
        s_out[0] = a0; s_out[1] = a1;
    }
    __syncthreads();
    mq = s_out[0]; sq = s_out[1];
    __syncthreads();

    int F = 0;
    if (quantized && mq > 0.0 && isfinite(mq)) {
        F = (7 * nd - 1) - ilogb(mq);
        F = F > 900 ? 900 : (F < -900 ? -900 : F);
    }

    // pass 2: payload + sum of W
    double sW = 0.0;
    for (uint32_t i = tid; i < a.dims; i += 256) {
        double qi = q[i];
        if (quantized) {
            long long W = (qi == qi && isfinite(qi)) ? llrint(scalbn(qi, F)) : 0;
            const long long lim = 1ll << (7 * nd);
            W = W >= lim ? lim - 1 : (W < -lim ? -lim : W);
            sW += (double)W;
            signed char dg[kMaxDigits];
            long long t = W;
            for (int j = nd - 1; j >= 1; --j) { dg[j] = (signed char)(t & 127); t >>= 7; }
            dg[0] = (signed char)t; // most significant, signed
            if (lqt == Q8) {
                uint32_t c = i >> 4, b = i & 15;
                for (int j = 0; j < nd; ++j) payload[((size_t)c * nd + j) * 16 + b] = (unsigned char)dg[j];
            } else if (lqt == Q4) {
                uint32_t byte = i >> 1, c = byte >> 4, b = byte & 15, arr = i & 1; // even dim = high nibble
                for (int j = 0; j < nd; ++j) payload[(((size_t)c * 2 + arr) * nd + j) * 16 + b] = (unsigned char)dg[j];
            } else {
                uint32_t c = i >> 3, e = i & 7;
                for (int j = 0; j < nd; ++j) payload[((size_t)c * nd + j) * 8 + e] = (unsigned char)dg[j];
            }
        } else if (a.qt == F32) {
            reinterpret_cast<float *>(payload)[i] = (float)qi;
        } else {
            reinterpret_cast<double *>(payload)[i] = qi;
        }
    }
    for (int o = 16; o; o >>= 1) sW += __shfl_xor_sync(0xffffffffu, sW, o);
    if (lane == 0) s_red[0][warp] = sW;
    __syncthreads();
    if (tid == 0) {
        sW = 0;
        for (int w = 0; w < 8; ++w) sW += s_red[0][w];
        const double M = (double)a.maxint, d = (double)a.dims;
        const double qn = sqrt(sq);
        const double eps = 5.9604644775390625e-08; // 2^-24
        PQHeader hh;
        hh.F = F; hh.zero_query = (sq == 0.0); hh.nd = nd; hh.pad = 0;
        hh.qn2 = sq;
        hh.numc = (a.qt == Q16 ? 1.0 : -M) * sW;
        hh.c_dot = quantized ? scalbn(1.0, -F) / M : 1.0;
        hh.c_key = (sq == 0.0) ? 0.0 : hh.c_dot / qn;
        // ---- rigorous bound on |surrogate key - true key|
        // fixed-point query: |q_i - W_i 2^-F| <= 2^-(F+1)  =>  |x.q error| <= ||x|| sqrt(d) 2^-(F+1), ||x|| <= sqrt(d)
        const double dq = quantized ? scalbn(1.0, -(F + 1)) : 0.0;
        if (a.metric == COSINE) {
            // key = -cos: fixed-point part / (||x|| ||q||) + fp32 roundings (two products, 1/||x|| itself)
            hh.e_rel = 0.0;
            hh.e_abs = (sq == 0.0 ? 0.0 : sqrt(d) * dq / qn) + 12.0 * eps; // incl. the all-fp32 tail of the batched epilogue
            if (a.qt == F32) hh.e_abs += (d + 8.0) * 2.0 * eps; // fp32 accumulation + fp32 copy of the query
        } else if (quantized) {
            // key = ||x||^2 (fp32 aux) + ||q||^2 - 2 x.q, rounded to fp32
            hh.e_abs = 2.0 * d * dq + (d + 8.0 * (d + sq)) * eps + 1e-30; // incl. the fp32 tail of the batched epilogue
            hh.e_rel = 2.0 * eps;
        } else if (a.qt == F32) {
            // fp32 sum of fp32 (q_i - x_i)^2 with an fp32 copy of the query
            hh.e_rel = (d + 8.0) * 4.0 * eps;
            hh.e_abs = 8.0 * eps * (sq + 1.0);
        } else {
            hh.e_rel = 4.0 * eps; // fp64 sum, fp32 key
            hh.e_abs = 1e-13 * (sq + 1.0);
        }
        hh.radius = a.radius;
        hh.radius_key = 0.0;
        if (a.radius_mode) {
            // candidates: every row whose surrogate could belong to a true distance <= radius
            const double r = a.radius;
            double t = a.metric == COSINE ? (r >= 1.0 ? 1.0 : -cospi(r)) : r * r;
            t = t + hh.e_abs + hh.e_rel * fabs(t) * 1.000001 + 1e-7 * fabs(t);
            hh.radius_key = (a.metric == COSINE && r >= 1.0) ? 2.0 : t;
        }
        *h = hh;
    }
}

cudaError_t launch_prep(uint32_t nq, cudaStream_t st, const PrepArgs &a) {
    prep_kernel<<<nq, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// =================================================================== upload / layout
__device__ __forceinline__ void store_chunk(int qt, uint4 *codes, uint32_t slot, uint32_t C, uint32_t c,
                                            const unsigned char *b) {
    uint4 v;
    v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
    v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
    v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
    v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
    codes[chunk_at_rt(qt, slot, C, c)] = v;
}
__device__ __forceinline__ void load_chunk(int qt, const uint4 *codes, uint32_t slot, uint32_t C, uint32_t c,
                                           unsigned char *b) {
    uint4 v = codes[chunk_at_rt(qt, slot, C, c)];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 4; ++i) b[4 * k + i] = (unsigned char)(w[k] >> (8 * i));
}

// thread per (row, chunk): stream-1 bytes -> column-blocked HBM representation
__global__ void scatter_kernel(const RowsArgs a, const unsigned char *__restrict__ staged,
                               const uint32_t *__restrict__ slots, const unsigned long long *__restrict__ ids_in,
                               uint32_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t row = (uint32_t)(t / a.C), c = (uint32_t)(t % a.C);
    if (row >= n) return;
    const uint32_t slot = slots[row];
    if (slot == 0xFFFFFFFFu) return; // superseded duplicate of the same batch
    unsigned char b[16];
    const unsigned char *src = staged + (size_t)row * a.rowbytes + (size_t)c * 16;
    const uint32_t remain = c * 16 < a.rowbytes ? a.rowbytes - c * 16 : 0u; // float rows: C is padded to a multiple of 8 chunks
    for (int i = 0; i < 16; ++i) b[i] = (uint32_t)i < remain ? src[i] : 0;
    chunk_to_device(a.qt, b);
    store_chunk(a.qt, a.codes, slot, a.C, c, b);
    if (c == 0) {
        a.ids[slot] = ids_in[row];
        atomicOr(a.live + (slot >> 5), 1u << (slot & 31));
    }
}

cudaError_t launch_scatter(const RowsArgs &a, const unsigned char *staged, const uint32_t *slots,
                           const unsigned long long *ids_in, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const size_t total = (size_t)n * a.C;
    scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, staged, slots, ids_in, n);
    return cudaGetLastError();
}

// thread per (row, chunk): synthetic rows, generator = oracle/syzgy_oracle.c orc_synth_rows
__global__ void synth_kernel(const RowsArgs a, unsigned long long seed, unsigned long long row0, uint32_t slot0,
                             uint32_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = (uint32_t)(t / a.C), c = (uint32_t)(t % a.C);
    if (r >= n) return;
    const unsigned long long row = row0 + r;
    const uint32_t slot = slot0 + r;
    unsigned char b[16];
    if (a.qt <= Q16) {
        const unsigned long long wpr = (a.rowbytes + 7) / 8;
        for (int half = 0; half < 2; ++half) {
            const unsigned long long j = (unsigned long long)c * 2 + half;
            unsigned long long w = (j < wpr) ? rand_u64(seed, row * wpr + j) : 0ull;
            for (int i = 0; i < 8; ++i) {
                const uint32_t off = c * 16 + half * 8 + i;
                b[half * 8 + i] = off < a.rowbytes ? (unsigned char)(w >> (8 * i)) : 0;
            }
        }
        if (a.qt == Q4 && (a.dims & 1)) {
            const uint32_t last = a.rowbytes - 1;
            if (last / 16 == c) b[last % 16] &= 0xF0;
        }
    } else {
        const int eb = a.qt == F32 ? 4 : 8, epc = 16 / eb;
        for (int e = 0; e < epc; ++e) {
            const uint32_t dim = c * epc + e;
            unsigned long long bits = 0;
            if (dim < a.dims) {
                double v = (double)(rand_u64(seed, row * a.dims + dim) >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
                bits = a.qt == F32 ? (unsigned long long)__float_as_uint((float)v)
                                   : (unsigned long long)__double_as_longlong(v);
            }
            for (int i = 0; i < eb; ++i) b[e * eb + i] = (unsigned char)(bits >> (8 * (eb - 1 - i))); // big-endian
        }
    }
    chunk_to_device(a.qt, b);
    store_chunk(a.qt, a.codes, slot, a.C, c, b);
    if (c == 0) {
        a.ids[slot] = row;
        atomicOr(a.live + (slot >> 5), 1u << (slot & 31));
    }
}

cudaError_t launch_synth(const RowsArgs &a, unsigned long long seed, unsigned long long row0, uint32_t slot0,
                         uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const size_t total = (size_t)n * a.C;
    synth_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, seed, row0, slot0, n);
    return cudaGetLastError();
}

// thread per row: per-row auxiliary values the surrogate needs (not streamed as payload):
//   float2 { 1/||x|| (0 for a zero-norm row -> distance 1.0, collection.go:828-830), ||x||^2 }
// from exact integer code sums for quantized rows, fp64 sums for float rows.
__global__ void aux_kernel(const RowsArgs a, const uint32_t *__restrict__ slots, uint32_t slot0, uint32_t n) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t slot = slots ? slots[r] : slot0 + r;
    if (slot == 0xFFFFFFFFu) return;
    unsigned long long S1 = 0, S2 = 0; // quantized (unsigned code sums; Q16: S1 signed below)
    long long S1s = 0;
    double fs = 0.0;
    uint32_t dim = 0;
    for (uint32_t c = 0; c < a.C && dim < a.dims; ++c) {
        uint4 v = a.codes[chunk_at_rt(a.qt, slot, a.C, c)];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        for (int k = 0; k < 4; ++k) {
            if (a.qt == Q4) {
                for (int i = 0; i < 4; ++i) {
                    uint32_t byte = (w[k] >> (8 * i)) & 0xFF;
                    uint32_t hi = byte >> 4, lo = byte & 15;
                    if (dim < a.dims) { S1 += hi; S2 += hi * hi; }
                    ++dim;
                    if (dim < a.dims) { S1 += lo; S2 += lo * lo; }
                    ++dim;
                }
            } else if (a.qt == Q8) {
                for (int i = 0; i < 4; ++i) {
                    uint32_t u = (w[k] >> (8 * i)) & 0xFF;
                    if (dim < a.dims) { S1 += u; S2 += u * u; }
                    ++dim;
                }
            } else if (a.qt == Q16) {
                for (int i = 0; i < 2; ++i) {
                    int s = (int)(short)((w[k] >> (16 * i)) & 0xFFFF);
                    if (dim < a.dims) { S1s += s; S2 += (unsigned long long)((long long)s * s); }
                    ++dim;
                }
            } else if (a.qt == F32) {
                double x = (double)__uint_as_float(w[k]);
                if (dim < a.dims) fs += x * x;
                ++dim;
            } else if ((k & 1) == 0) {
                double x = __hiloint2double((int)w[k + 1], (int)w[k]);
                if (dim < a.dims) fs += x * x;
                ++dim;
            }
        }
    }
    double nx2;
    const double M = (double)a.maxint, d = (double)a.dims;
    if (a.qt == Q16) nx2 = (4.0 * (double)S2 + 4.0 * (double)S1s + d) / (M * M);                // x = (2s+1)/M
    else if (a.qt <= Q8) nx2 = (4.0 * (double)S2 - 4.0 * M * (double)S1 + M * M * d) / (M * M); // x = (2u-M)/M
    else nx2 = fs;
    float2 o;
    o.x = (nx2 > 0.0 && isfinite(nx2)) ? (float)(1.0 / sqrt(nx2)) : 0.f;
    o.y = (float)nx2;
    reinterpret_cast<float2 *>(a.aux)[slot] = o;
}

cudaError_t launch_aux(const RowsArgs &a, const uint32_t *slots, uint32_t slot0, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    aux_kernel<<<(n + 127) / 128, 128, 0, st>>>(a, slots, slot0, n);
    return cudaGetLastError();
}

__global__ void kill_kernel(uint32_t *live, const uint32_t *__restrict__ slots, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAnd(live + (slots[i] >> 5), ~(1u << (slots[i] & 31)));
}
cudaError_t launch_kill(uint32_t *live, const uint32_t *slots, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    kill_kernel<<<(n + 255) / 256, 256, 0, st>>>(live, slots, n);
    return cudaGetLastError();
}

__global__ void mask_set_kernel(uint32_t *mask, const uint32_t *__restrict__ slots,
                                const unsigned char *__restrict__ pass, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && pass[i] && slots[i] != 0xFFFFFFFFu) atomicOr(mask + (slots[i] >> 5), 1u << (slots[i] & 31));
}
cudaError_t launch_mask_set(uint32_t *mask, const uint32_t *slots, const unsigned char *pass, uint32_t n,
                            cudaStream_t st) {
    if (!n) return cudaSuccess;
    mask_set_kernel<<<(n + 255) / 256, 256, 0, st>>>(mask, slots, pass, n);
    return cudaGetLastError();
}

__global__ void fetch_kernel(const RowsArgs a, const uint32_t *__restrict__ slots, uint32_t n,
                             unsigned char *__restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t row = (uint32_t)(t / a.C), c = (uint32_t)(t % a.C);
    if (row >= n) return;
    unsigned char b[16];
    load_chunk(a.qt, a.codes, slots[row], a.C, c, b);
    chunk_to_disk(a.qt, b);
    const uint32_t remain = c * 16 < a.rowbytes ? a.rowbytes - c * 16 : 0u;
    unsigned char *dst = out + (size_t)row * a.rowbytes + (size_t)c * 16;
    for (int i = 0; i < 16; ++i)
        if ((uint32_t)i < remain) dst[i] = b[i];
}
cudaError_t launch_fetch(const RowsArgs &a, const uint32_t *slots, uint32_t n, unsigned char *out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const size_t total = (size_t)n * a.C;
    fetch_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, slots, n, out);
    return cudaGetLastError();
}

// ============================================================ ingest: encodeDocument on the device
// quantization.go:5-23: clamp to [-1, 1], (value + 1) / 2 * maxInt in that operation order, math.Round (half away from
// zero = round()), conversion to an unsigned integer (NaN converts to 0 here; Go/amd64 gives 1 << 63, whose low
// 4/8/16 bits -- all the reference keeps -- are 0 as well).
__device__ __forceinline__ uint32_t quantize_code(double v, double maxint) {
    if (v < -1.0) v = -1.0;
    else if (v > 1.0) v = 1.0;
    const double q = __dmul_rn(__ddiv_rn(__dadd_rn(v, 1.0), 2.0), maxint);
    return (uint32_t)(unsigned long long)round(q);
}

// thread per (document, 16-byte chunk of its stream-1 row): n vectors of float64 -> the bytes encodeDocument
// (collection.go:713-744) produces: 4-bit pairs (even element in the high nibble), bytes, big-endian 16/32/64-bit words
__global__ void encode_kernel(const RowsArgs a, const double *__restrict__ vec, unsigned char *__restrict__ staged, uint32_t n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t row = (uint32_t)(t / a.C), c = (uint32_t)(t % a.C);
    if (row >= n) return;
    const double *v = vec + (size_t)row * a.dims;
    const double mx = (double)a.maxint;
    __align__(16) unsigned char b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = 0;
    if (a.qt == Q4) {
        for (int i = 0; i < 16; ++i) {
            const uint32_t e = (c * 16 + i) * 2;
            uint32_t hi = 0, lo = 0;
            if (e < a.dims) hi = quantize_code(v[e], mx);
            if (e + 1 < a.dims) lo = quantize_code(v[e + 1], mx);
            b[i] = (unsigned char)((hi << 4) | (lo & 0x0Fu));
        }
    } else if (a.qt == Q8) {
        for (int i = 0; i < 16; ++i) {
            const uint32_t e = c * 16 + i;
            if (e < a.dims) b[i] = (unsigned char)quantize_code(v[e], mx);
        }
    } else if (a.qt == Q16) {
        for (int i = 0; i < 8; ++i) {
            const uint32_t e = c * 8 + i;
            if (e < a.dims) {
                const uint32_t q = quantize_code(v[e], mx);
                b[2 * i] = (unsigned char)(q >> 8);
                b[2 * i + 1] = (unsigned char)q;
            }
        }
    } else if (a.qt == F32) {
        for (int i = 0; i < 4; ++i) {
            const uint32_t e = c * 4 + i;
            if (e < a.dims) {
                const uint32_t q = __float_as_uint(__double2float_rn(v[e])); // math.Float32bits(float32(value))
                for (int k = 0; k < 4; ++k) b[4 * i + k] = (unsigned char)(q >> (24 - 8 * k));
            }
        }
    } else {
        for (int i = 0; i < 2; ++i) {
            const uint32_t e = c * 2 + i;
            if (e < a.dims) {
                const unsigned long long q = (unsigned long long)__double_as_longlong(v[e]);
                for (int k = 0; k < 8; ++k) b[8 * i + k] = (unsigned char)(q >> (56 - 8 * k));
            }
        }
    }
    unsigned char *dst = staged + (size_t)row * a.rowbytes + (size_t)c * 16;
    const uint32_t remain = c * 16 < a.rowbytes ? a.rowbytes - c * 16 : 0u;
    if (remain >= 16 && (((size_t)row * a.rowbytes) & 15u) == 0) {
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(b);
    } else {
        for (uint32_t i = 0; i < 16 && i < remain; ++i) dst[i] = b[i];
    }
}

cudaError_t launch_encode(const RowsArgs &a, const double *vec, unsigned char *staged, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const size_t total = (size_t)n * a.C;
    encode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, vec, staged, n);
    return cudaGetLastError();
}

// ===================================================== K5: merge of row-sharded top-k lists
// One CTA per query: the G*k gathered (distance, id) pairs are ranked by
// (distance, lexicographic decimal id) and the first k written out.
__global__ void __launch_bounds__(256) merge_kernel(const MergeArgs a) {
    extern __shared__ __align__(16) unsigned char sm[];
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    const uint32_t cap = a.G * a.k;
    double *s_d = reinterpret_cast<double *>(sm);
    unsigned long long *s_i = reinterpret_cast<unsigned long long *>(s_d + cap);
    __shared__ uint32_t s_total;
    if (tid == 0) {
        s_total = 0;
        if (a.wait_cnt) {
            // sharded search inside one process (sharded.cu): the shards' finalize kernels store their lists straight into this
            // device's gather buffer over NVLink and then bump wait_cnt[q] (system-scope release).  Wait for all of them
            // (acquire); bounded, so that a failed launch on another device surfaces as an error instead of a hang.
            const volatile uint32_t *cnt = a.wait_cnt + q;
            unsigned long long t0 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            uint32_t seen;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                if (seen >= a.wait_target) break;
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > a.wait_timeout_ns) { if (a.err) atomicOr(a.err, 1u); break; }
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
    for (uint32_t e = tid; e < cap; e += 256) {
        const uint32_t g = e / a.k, j = e % a.k;
        const unsigned long long *ids_g;
        const double *dist_g;
        const uint32_t *n_g;
        if (a.rank_stride) {
            ids_g = reinterpret_cast<const unsigned long long *>(reinterpret_cast<const char *>(a.g_ids) + g * a.rank_stride);
            dist_g = reinterpret_cast<const double *>(reinterpret_cast<const char *>(a.g_dist) + g * a.rank_stride);
            n_g = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(a.g_n) + g * a.rank_stride);
        } else {
            ids_g = a.g_ids + (size_t)g * a.nq * a.k;
            dist_g = a.g_dist + (size_t)g * a.nq * a.k;
            n_g = a.g_n + (size_t)g * a.nq;
        }
        const uint32_t n = __ldcg(n_g + q); // L2: the lists may have been written by peer devices during this launch
        const size_t src = (size_t)q * a.k + j;
        const bool ok = j < n;
        s_d[e] = ok ? __ldcg(dist_g + src) : __longlong_as_double(0x7ff8000000000000ll);
        s_i[e] = ok ? __ldcg(ids_g + src) : 0ull;
        if (ok) atomicAdd(&s_total, 1u);
    }
    __syncthreads();
    for (uint32_t e = tid; e < cap; e += 256) {
        const double d = s_d[e];
        if (!(d == d)) continue;
        const unsigned long long id = s_i[e];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < cap; ++j) {
            const double dj = s_d[j];
            if (j != e && dj == dj && (dj < d || (dj == d && lex_less_u64(s_i[j], id)))) ++rank;
        }
        if (rank < a.k) {
            a.out_ids[(size_t)q * a.k + rank] = id;
            a.out_dist[(size_t)q * a.k + rank] = d;
        }
    }
    if (tid == 0) {
        a.out_n[q] = s_total < a.k ? s_total : a.k;
        if (a.out_flags) {
            uint32_t fl = 0;
            if (a.g_flags)
                for (uint32_t g = 0; g < a.G; ++g) {
                    const uint32_t *f_g = a.rank_stride
                        ? reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(a.g_flags) + g * a.rank_stride)
                        : a.g_flags + (size_t)g * a.nq;
                    fl |= __ldcg(f_g + q);
                }
            a.out_flags[q] = fl;
        }
    }
}

// an empty shard has nothing to finalize: it still reports in (sharded search)
__global__ void bump_kernel(uint32_t *cnt, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        __threadfence_system();
        atomicAdd_system(cnt + i, 1u);
    }
}
cudaError_t launch_bump(uint32_t *cnt, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    bump_kernel<<<(n + 127) / 128, 128, 0, st>>>(cnt, n);
    return cudaGetLastError();
}

cudaError_t launch_merge(const MergeArgs &a, cudaStream_t st) {
    if (!a.nq) return cudaSuccess;
    const size_t smem = (size_t)a.G * a.k * 16;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    merge_kernel<<<a.nq, 256, smem, st>>>(a);
    return cudaGetLastError();
}

// ======================================================================= byte-planar copy of 16-bit codes
// The batched tensor-core path contracts bytes (tcgen05 kind::i8).  A 16-bit collection gets a secondary copy in
// which block b, chunk c (16 dimensions) holds the HIGH bytes of the uncentred codes u = c + 32768 of its 32 rows
// in one array and the LOW bytes in another, both in the column-blocked layout of an 8-bit collection, so that
// I = sum u_i W_i = 256 * sum hi_i W_i + sum lo_i W_i is two 8-bit contractions with the same query operand.
__global__ void __launch_bounds__(256) planar16_kernel(const uint4 *__restrict__ codes, uint32_t C16, uint4 *__restrict__ hi,
                                                       uint4 *__restrict__ lo, uint32_t C8, uint32_t nblk) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)nblk * C8 * 32;
    if (t >= total) return;
    const uint32_t lane = (uint32_t)(t & 31), c8 = (uint32_t)((t >> 5) % C8), b = (uint32_t)((t >> 5) / C8);
    unsigned char h8[16], l8[16];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t c16 = 2 * c8 + half;
        uint4 v = make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u); // centred zero-padding -> u = 0
        if (c16 < C16) v = codes[((size_t)b * C16 + c16) * 32 + lane];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t s16 = (w[e >> 1] >> (16 * (e & 1))) & 0xFFFFu; // little-endian int16 of (u - 32768)
            const uint32_t u = s16 ^ 0x8000u;
            h8[half * 8 + e] = (unsigned char)(u >> 8);
            l8[half * 8 + e] = (unsigned char)(u & 0xFF);
        }
        if (c16 >= C16) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { h8[half * 8 + e] = 0; l8[half * 8 + e] = 0; }
        }
    }
    uint4 H, L;
    memcpy(&H, h8, 16);
    memcpy(&L, l8, 16);
    hi[((size_t)b * C8 + c8) * 32 + lane] = H;
    lo[((size_t)b * C8 + c8) * 32 + lane] = L;
}

// 4-bit collections: a secondary copy with one byte per code (0..15), in the column-blocked layout of an 8-bit
// collection, is the operand of the batched path (tcgen05 has no 4-bit integer kind).  Chunk c8 (16 dimensions) of a
// row is the first or second half of its 4-bit chunk c8 / 2; even dimensions sit in the high nibble (collection.go:774-779).
__global__ void __launch_bounds__(256) expand4_kernel(const uint4 *__restrict__ codes, uint32_t C4, uint4 *__restrict__ out, uint32_t C8,
                                                      uint32_t nblk) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)nblk * C8 * 32;
    if (t >= total) return;
    const uint32_t lane = (uint32_t)(t & 31), c8 = (uint32_t)((t >> 5) % C8), b = (uint32_t)((t >> 5) / C8);
    const uint32_t c4 = c8 >> 1, half = c8 & 1u;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c4 < C4) v = codes[((size_t)b * C4 + c4) * 32 + lane];
    const uint32_t w2[2] = {half ? v.z : v.x, half ? v.w : v.y}; // 8 bytes = 16 codes
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t two = (w2[k >> 1] >> (16 * (k & 1))) & 0xFFFFu; // 2 bytes = 4 codes: dims 4k .. 4k+3
        const uint32_t b0 = two & 0xFF, b1 = two >> 8;
        o[k] = (b0 >> 4) | ((b0 & 0xF) << 8) | ((b1 >> 4) << 16) | ((b1 & 0xF) << 24);
    }
    out[((size_t)b * C8 + c8) * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
}

cudaError_t launch_expand4(const uint4 *codes, uint32_t C4, uint4 *out, uint32_t C8, uint32_t nblk, cudaStream_t st) {
    const size_t total = (size_t)nblk * C8 * 32;
    if (!total) return cudaSuccess;
    expand4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codes, C4, out, C8, nblk);
    return cudaGetLastError();
}

cudaError_t launch_planar16(const uint4 *codes, uint32_t C16, uint4 *hi, uint4 *lo, uint32_t C8, uint32_t nblk, cudaStream_t st) {
    const size_t total = (size_t)nblk * C8 * 32;
    if (!total) return cudaSuccess;
    planar16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codes, C16, hi, lo, C8, nblk);
    return cudaGetLastError();
}

} // namespace szg
