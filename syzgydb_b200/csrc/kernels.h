// kernels.h -- host-callable launchers of every kernel in the library.
#pragma once
#include "scan_impl.cuh"

namespace szg {

struct PrepArgs {
    const double *queries; // nq * dims
    unsigned char *pq;     // nq * pq_stride
    size_t pq_stride;
    uint32_t dims, C, metric, maxint;
    int qt;
    int nd;          // digits (2 or 3) for quantized rows
    int radius_mode;
    double radius;
    int planar16 = 0; // payload laid out like an 8-bit row (batched 16-bit path); C is then the planar chunk count
};
cudaError_t launch_prep(uint32_t nq, cudaStream_t st, const PrepArgs &a);

// scan: one instantiation file per quantization
// nd: digits of the fixed-point query (2 fast / 3 precise; ignored for float rows)
cudaError_t launch_scan(int qt, int mode, int nd, int grid, int threads, size_t smem, cudaStream_t st,
                        const ScanArgs &a);
// short quantized rows (C = 2, 4, 8), k <= 24: scan_small.cuh; a.parts = work items per query, lists per query = parts * 16
cudaError_t launch_scan_small(int qt, int nd, uint32_t C, int grid, size_t smem, cudaStream_t st, const ScanArgs &a);
// one launch for all nq queries of a call: merges each query's per-CTA lists, fp64 re-score, ordered output
cudaError_t launch_finalize(int qt, int mode, uint32_t nq, cudaStream_t st, const FinalizeArgs &a);
cudaError_t scan_configure(int qt, size_t max_smem);

struct RowsArgs {
    uint4 *codes;
    void *aux;
    uint32_t *live;
    unsigned long long *ids;
    uint32_t C, dims, metric, maxint, rowbytes;
    int qt;
};
// staged: n rows of rowbytes bytes (stream-1 format); slots[n]; ids_in[n]
cudaError_t launch_scatter(const RowsArgs &a, const unsigned char *staged, const uint32_t *slots,
                           const unsigned long long *ids_in, uint32_t n, cudaStream_t st);
// rows [slot0, slot0+n) <- synthetic rows [row0, row0+n) of `seed`; ids = row index
cudaError_t launch_synth(const RowsArgs &a, unsigned long long seed, unsigned long long row0, uint32_t slot0,
                         uint32_t n, cudaStream_t st);
// aux of the given slots (slots == NULL: the range [slot0, slot0+n))
cudaError_t launch_aux(const RowsArgs &a, const uint32_t *slots, uint32_t slot0, uint32_t n, cudaStream_t st);
// clears live bits of slots
cudaError_t launch_kill(uint32_t *live, const uint32_t *slots, uint32_t n, cudaStream_t st);
// builds a filter bitmask: bit set for slots[i] with pass[i] != 0 (mask pre-zeroed)
cudaError_t launch_mask_set(uint32_t *mask, const uint32_t *slots, const unsigned char *pass, uint32_t n,
                            cudaStream_t st);
// records back to stream-1 bytes
cudaError_t launch_fetch(const RowsArgs &a, const uint32_t *slots, uint32_t n, unsigned char *out, cudaStream_t st);

// ingest: n float64 vectors -> stream-1 bytes (encodeDocument, collection.go:713-744 + quantize, quantization.go:5-23)
cudaError_t launch_encode(const RowsArgs &a, const double *vec, unsigned char *staged, uint32_t n, cudaStream_t st);

// K3: fp64 distances of gathered rows in the reference's operation order (gather.cu)
struct RescoreArgs {
    const uint4 *codes;
    const unsigned long long *ids;
    const double *lut;
    const double *q;            // nlists x dims
    const double *m1;           // nlists: sum of q_i^2 in dimension order (the reference's m1, collection.go:825), computed by the host
    const uint32_t *list_off;   // nlists + 1 offsets into slots: list l is scored against query l (NULL: one list, one query)
    uint32_t nlists;
    const uint32_t *slots;      // 0xFFFFFFFF = missing
    double *out_dist;
    unsigned long long *out_ids; // optional
    uint32_t C, dims, metric, m;
    int qt;
};
cudaError_t launch_rescore(const RescoreArgs &a, cudaStream_t st);

// K2, second half: exact distances of the rows the radius scan compacted, inclusive test (collection.go:598), ascending
// (distance, lexicographic id) order -- all on the device
constexpr uint32_t kRadiusSortSmall = 2048; // results up to this size are sorted by the launch that scores them
struct RadiusFinishArgs {
    const uint4 *codes;
    const unsigned long long *ids;
    const double *lut;
    const double *q;
    const uint32_t *slots;       // compacted by the scan
    const uint32_t *count_ptr;   // how many (device); at most cap are present
    uint32_t cap;
    double radius;
    double m1;                   // sum of q_i^2 in dimension order (cosine)
    unsigned long long *keys;    // scratch: pairs (distance bits, id); capacity = the power of two >= cap
    uint32_t *out_count;         // exact hits
    double *out_dist;            // [<= cap] ascending
    unsigned long long *out_ids;
    uint32_t C, dims, metric;
    int qt;
};
// scores + filters every compacted row and, when at most kRadiusSortSmall pass, orders them (out_dist / out_ids final)
cudaError_t launch_radius_finish(const RadiusFinishArgs &a, cudaStream_t st);
// larger results: global bitonic sort of the m hits left in a.keys by launch_radius_finish, then out_dist / out_ids
cudaError_t launch_radius_sort_large(const RadiusFinishArgs &a, uint32_t m, cudaStream_t st);

// byte-planar copy of a 16-bit collection's codes (operand of the batched path)
// one-byte-per-code copy of a 4-bit collection's codes (operand of the batched path)
cudaError_t launch_expand4(const uint4 *codes, uint32_t C4, uint4 *out, uint32_t C8, uint32_t nblk, cudaStream_t st);
cudaError_t launch_planar16(const uint4 *codes, uint32_t C16, uint4 *hi, uint4 *lo, uint32_t C8, uint32_t nblk, cudaStream_t st);

// K4: batched-query tensor-core contraction (batch_q8.cu): 8-bit rows, or the byte planes of 16-bit rows
struct BatchArgs {
    const uint4 *codes;        // 8-bit rows; 16-bit: the plane of HIGH bytes
    const uint4 *codes_lo;     // 16-bit: the plane of LOW bytes (NULL for 8-bit)
    const void *aux;
    const uint32_t *live;
    const uint32_t *mask;
    const unsigned char *pq;   // prepared queries (2 digits), pq_stride apart
    size_t pq_stride;
    unsigned long long *cand;  // [nq][nranges][keep] candidate keys for finalize_kernel
    uint32_t C, nblk, metric, nq, dims;
    uint32_t group0, ngroups;  // query groups [group0, group0 + ngroups) run in this launch
    uint32_t nranges, nlists;  // row ranges (CTAs per group); nlists = what finalize sees per query
    uint32_t stages;           // shared-memory ring stages (one K slice of a super tile each)
    uint32_t slice;            // chunks per K slice (even, <= C)
    uint32_t *tile_ctr;        // [groups of the call] next-tile counters (pre-set to 0xFFFFFFFF), or nullptr: fixed row ranges
    uint32_t *gmth;            // [nq][need groups][gm_sp] ordered key of each range's mth-best row so far (range r -> group r % need,
                               // member r / need; 0xFFFFFFFF = none yet / padding), need = ceil(keep / mth)
    uint32_t gm_stride, gm_sp; // words per query (need * gm_sp) and per group (4 or 8)
    uint32_t mth;              // ceil(keep / nranges)
    uint32_t keep;             // candidates per (query, row range) list handed to finalize (32, 64, 128)
    uint32_t debug;            // profiling aid: bit0 skip the epilogue math, bit1 skip aux loads, bit2 skip the MMAs
    uint32_t poll_ns;          // pause of the bound poller warp between rounds once every query has a bound (0 = 2000 ns; SZG_BATCH_POLL_NS)
    long long *trace;          // profiling aid (SZG_OPT_TRACE_BUFFER): clock64 stamps / counters of CTA 0's first epilogue warp
};
uint32_t batch_slice_chunks(uint32_t C, uint32_t want, size_t smem_limit); // chunks per K slice (ring stage); want = 0: automatic
size_t batch_list_bytes(uint32_t keep);                   // shared memory of the 64 candidate lists
size_t batch_smem_bytes(uint32_t slice, uint32_t stages, uint32_t keep);
uint32_t batch_stages(uint32_t slice, size_t smem_limit); // ring stages that fit
uint32_t batch_max_chunks();
cudaError_t batch_configure(size_t max_smem);
size_t batch_dynamic_limit();
cudaError_t launch_batch(const BatchArgs &a, cudaStream_t st);

// ---- metadata filters on the device (filter.cu)
constexpr uint32_t kFilterMaxCols = 32;   // metadata columns per collection
constexpr int kFilterMaxStack = 16;       // value stack of the filter program
enum MetaKind : uint32_t { MV_MISSING = 0, MV_NULL = 1, MV_BOOL = 2, MV_NUMBER = 3, MV_STRING = 4, MV_OTHER = 5, MV_ERROR = 6 };
enum DocKind : uint32_t { DOC_INVALID = 0, DOC_OBJECT = 1, DOC_OTHER = 2 };
enum FilterOpcode : uint32_t {
    FOP_COL = 1, FOP_NUM, FOP_STR, FOP_BOOL, FOP_NULL, FOP_EQ, FOP_NE, FOP_LT, FOP_LE, FOP_GT, FOP_GE, FOP_AND, FOP_OR,
    FOP_NOT, FOP_IN, FOP_NOT_IN, FOP_STR_TABLE, FOP_EXISTS, FOP_NOT_EXISTS
};
struct FilterOp {
    uint32_t op, arg;              // arg: column (COL, EXISTS, NOT_EXISTS), list length (IN, NOT_IN), table length (STR_TABLE)
    unsigned long long bits;       // literal: float64 bits, string code, bool
    const unsigned char *table;    // STR_TABLE: one byte per string code
};
struct FilterArgs {
    const FilterOp *prog;
    uint32_t nops;
    const unsigned char *doc_kind;
    const unsigned char *col_kind[kFilterMaxCols];
    const unsigned long long *col_val[kFilterMaxCols];
    const uint32_t *rank;          // bytewise order of every string code (dictionary + this program's literals)
    uint32_t nrank;
    uint32_t *mask;                // out: bit per slot
    uint32_t nwords, nslots;
};
cudaError_t launch_filter(const FilterArgs &a, cudaStream_t st);
cudaError_t launch_meta_scatter(const uint32_t *slots, const unsigned char *kinds, const unsigned long long *vals,
                                unsigned char *col_kind, unsigned long long *col_val, uint32_t n, cudaStream_t st);
struct MetaPtrs {
    unsigned char *doc_kind;
    unsigned char *col_kind[kFilterMaxCols];
};
cudaError_t launch_meta_clear(const uint32_t *slots, uint32_t n, const MetaPtrs &p, cudaStream_t st);

struct MergeArgs {
    const unsigned long long *g_ids; // [G][nq][k]
    const double *g_dist;
    const uint32_t *g_n; // [G][nq]
    const uint32_t *g_flags; // optional [G][nq]: bit0 = a rank could not certify its local list
    size_t rank_stride;  // bytes between ranks for all three arrays; 0 = each array tightly packed [G][...]
    uint32_t G, nq, k;
    unsigned long long *out_ids; // [nq][k]
    double *out_dist;
    uint32_t *out_n;
    uint32_t *out_flags; // optional [nq]: OR of the ranks' flags
    // sharded search in one process: wait until wait_cnt[q] >= wait_target (peer shards report in), at most wait_timeout_ns
    const uint32_t *wait_cnt;
    uint32_t wait_target;
    unsigned long long wait_timeout_ns;
    uint32_t *err;       // bit0 set when the wait timed out
};
cudaError_t launch_merge(const MergeArgs &a, cudaStream_t st);
cudaError_t launch_bump(uint32_t *cnt, uint32_t n, cudaStream_t st);

} // namespace szg
