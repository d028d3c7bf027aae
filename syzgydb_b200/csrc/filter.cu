// filter.cu -- metadata filters evaluated on the device.
//
// The reference applies a FilterFn per visited document (collection.go:592-594); filters built by BuildFilter
// (collection.go:204-218) json.Unmarshal the metadata of every document for every query
// (query/compiler.go:477-497) and walk a closure tree (CompileExpression, compiler.go:15-165).  Here the scalar
// top-level metadata fields live in a columnar mirror next to the vectors (one kind byte + one 8-byte payload per
// row and column; strings are dictionary codes), the shim lowers the filter's syntax tree to a postfix program,
// and one kernel evaluates that program for every row into the bitmask the scan kernels already take.
//
// Semantics restated from query/compiler.go (evaluateOperation 167-264, compareValues 266-326, getField 428-444):
//  * every operand is evaluated before its operator, and an error anywhere makes CreateFilterFunction return an
//    error, which BuildFilter turns into "false" (collection.go:211-215): an error flag poisons the row;
//  * ==, != are reflect.DeepEqual over {nil, bool, float64, string}: different kinds are unequal, no error;
//  * <, <=, >, >= dispatch on the LEFT operand: float64 needs a float64 on the right, string a string (bytewise
//    order), anything else is an error;
//  * AND needs two bools; OR needs a bool on the left and, only when that is false, a bool on the right; NOT a bool;
//  * IN / NOT_IN are DeepEqual against each element, never an error;
//  * CONTAINS / STARTS_WITH / ENDS_WITH / MATCHES need strings: the right operand is a literal, so the predicate
//    is a table over the string dictionary built on the host; a non-string left operand is an error;
//  * a missing key reads as nil (getField returns v[key]); a non-object document makes every field access an error;
//  * the result must be a bool ("query result is not a boolean" otherwise).
#include "kernels.h"

namespace szg {

namespace {

struct Val {
    uint32_t kind;
    unsigned long long v;
};

__device__ __forceinline__ bool deep_equal(const Val &l, const Val &r) {
    if (l.kind != r.kind) return false;
    switch (l.kind) {
    case MV_NULL: return true;
    case MV_BOOL: return (l.v != 0) == (r.v != 0);
    case MV_NUMBER: return __longlong_as_double((long long)l.v) == __longlong_as_double((long long)r.v);
    case MV_STRING: return l.v == r.v;
    }
    return false; // arrays / objects never equal a scalar literal
}

} // namespace

__global__ void __launch_bounds__(256) filter_kernel(const FilterArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    if (slot < a.nslots) {
        const uint32_t doc = a.doc_kind[slot];
        bool err = doc == DOC_INVALID; // json.Unmarshal failed: the filter function returns an error
        Val st[kFilterMaxStack];
        int sp = 0;
        for (uint32_t pc = 0; pc < a.nops && !err; ++pc) {
            const FilterOp op = a.prog[pc];
            switch (op.op) {
            case FOP_COL: {
                const uint32_t k = a.col_kind[op.arg][slot];
                if (doc != DOC_OBJECT || k == MV_ERROR) { err = true; break; }
                st[sp].kind = k == MV_MISSING ? (uint32_t)MV_NULL : k;
                st[sp].v = a.col_val[op.arg][slot];
                ++sp;
                break;
            }
            case FOP_NUM: st[sp].kind = MV_NUMBER; st[sp].v = op.bits; ++sp; break;
            case FOP_STR: st[sp].kind = MV_STRING; st[sp].v = op.bits; ++sp; break;
            case FOP_BOOL: st[sp].kind = MV_BOOL; st[sp].v = op.bits; ++sp; break;
            case FOP_NULL: st[sp].kind = MV_NULL; st[sp].v = 0; ++sp; break;
            case FOP_EQ:
            case FOP_NE: {
                const bool eq = deep_equal(st[sp - 2], st[sp - 1]);
                sp -= 1;
                st[sp - 1].kind = MV_BOOL;
                st[sp - 1].v = (op.op == FOP_EQ) ? eq : !eq;
                break;
            }
            case FOP_LT:
            case FOP_LE:
            case FOP_GT:
            case FOP_GE: {
                const Val l = st[sp - 2], r = st[sp - 1];
                sp -= 1;
                int c = 0; // -1, 0, 1; 2 = unordered (a NaN operand: every comparison false)
                if (l.kind == MV_NUMBER) {
                    if (r.kind != MV_NUMBER) { err = true; break; }
                    const double x = __longlong_as_double((long long)l.v), y = __longlong_as_double((long long)r.v);
                    c = x < y ? -1 : (x > y ? 1 : (x == y ? 0 : 2));
                } else if (l.kind == MV_STRING) {
                    if (r.kind != MV_STRING) { err = true; break; }
                    const uint32_t x = l.v < a.nrank ? a.rank[l.v] : 0u, y = r.v < a.nrank ? a.rank[r.v] : 0u;
                    c = x < y ? -1 : (x > y ? 1 : 0);
                } else { err = true; break; }
                bool res;
                if (op.op == FOP_LT) res = c == -1;
                else if (op.op == FOP_LE) res = c == -1 || c == 0;
                else if (op.op == FOP_GT) res = c == 1;
                else res = c == 1 || c == 0;
                st[sp - 1].kind = MV_BOOL;
                st[sp - 1].v = res;
                break;
            }
            case FOP_AND: {
                const Val l = st[sp - 2], r = st[sp - 1];
                sp -= 1;
                if (l.kind != MV_BOOL || r.kind != MV_BOOL) { err = true; break; }
                st[sp - 1].v = (l.v != 0) && (r.v != 0);
                break;
            }
            case FOP_OR: {
                const Val l = st[sp - 2], r = st[sp - 1];
                sp -= 1;
                if (l.kind != MV_BOOL) { err = true; break; }
                if (l.v != 0) { st[sp - 1].v = 1; break; } // the right operand's type is not looked at
                if (r.kind != MV_BOOL) { err = true; break; }
                st[sp - 1].kind = MV_BOOL;
                st[sp - 1].v = r.v != 0;
                break;
            }
            case FOP_NOT:
                if (st[sp - 1].kind != MV_BOOL) { err = true; break; }
                st[sp - 1].v = st[sp - 1].v == 0;
                break;
            case FOP_IN:
            case FOP_NOT_IN: {
                const int n = (int)op.arg;
                const Val l = st[sp - 1 - n];
                bool found = false;
                for (int i = 0; i < n; ++i) found = found || deep_equal(l, st[sp - n + i]);
                sp -= n;
                st[sp - 1].kind = MV_BOOL;
                st[sp - 1].v = (op.op == FOP_IN) ? found : !found;
                break;
            }
            case FOP_STR_TABLE: {
                const Val l = st[sp - 1];
                if (l.kind != MV_STRING) { err = true; break; }
                st[sp - 1].kind = MV_BOOL;
                st[sp - 1].v = (l.v < op.arg) ? op.table[l.v] != 0 : 0;
                break;
            }
            case FOP_EXISTS: // EXISTS(x): "evaluating x gave no error"
                st[sp].kind = MV_BOOL;
                st[sp].v = doc == DOC_OBJECT && a.col_kind[op.arg][slot] != MV_ERROR;
                ++sp;
                break;
            case FOP_NOT_EXISTS: // DOES_NOT_EXIST(x): false for a non-object document, no error
                st[sp].kind = MV_BOOL;
                st[sp].v = doc == DOC_OBJECT && a.col_kind[op.arg][slot] == MV_MISSING;
                ++sp;
                break;
            default: err = true;
            }
        }
        pass = !err && sp == 1 && st[0].kind == MV_BOOL && st[0].v != 0;
    }
    const unsigned w = __ballot_sync(0xffffffffu, pass);
    if ((threadIdx.x & 31) == 0 && (slot >> 5) < a.nwords) a.mask[slot >> 5] = w;
}

cudaError_t launch_filter(const FilterArgs &a, cudaStream_t st) {
    if (!a.nslots) return cudaSuccess;
    const uint32_t n = (a.nslots + 31) / 32 * 32;
    filter_kernel<<<(n + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// one column of a metadata batch into the mirror: kind byte + payload per slot (vals == NULL: kinds only)
__global__ void meta_scatter_kernel(const uint32_t *__restrict__ slots, const unsigned char *__restrict__ kinds,
                                    const unsigned long long *__restrict__ vals, unsigned char *col_kind,
                                    unsigned long long *col_val, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    if (s == 0xFFFFFFFFu) return;
    col_kind[s] = kinds[i];
    if (vals) col_val[s] = vals[i];
}

cudaError_t launch_meta_scatter(const uint32_t *slots, const unsigned char *kinds, const unsigned long long *vals,
                                unsigned char *col_kind, unsigned long long *col_val, uint32_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    meta_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(slots, kinds, vals, col_kind, col_val, n);
    return cudaGetLastError();
}

// removed documents: their slots read as "no metadata" until a new document takes them
__global__ void meta_clear_kernel(const uint32_t *__restrict__ slots, uint32_t n, MetaPtrs p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slots[i];
    if (s == 0xFFFFFFFFu) return;
    if (p.doc_kind) p.doc_kind[s] = DOC_INVALID;
    for (uint32_t c = 0; c < kFilterMaxCols; ++c)
        if (p.col_kind[c]) p.col_kind[c][s] = MV_MISSING;
}

cudaError_t launch_meta_clear(const uint32_t *slots, uint32_t n, const MetaPtrs &p, cudaStream_t st) {
    if (!n) return cudaSuccess;
    meta_clear_kernel<<<(n + 255) / 256, 256, 0, st>>>(slots, n, p);
    return cudaGetLastError();
}

} // namespace szg
