/*
 * syzgy_b200.h -- C ABI of the B200-native search hot path of SyzgyDB.
 *
 * The reference (github.com/smhanov/syzgydb, pure Go) has no FFI/plugin interface for
 * this path: the scan sits behind the Go method (*Collection).Search
 * (collection.go:569-711).  This header is the boundary a cgo shim inside package
 * syzgydb binds (INTEGRATION.md shows that shim).  Each entry point cites the
 * reference code it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, a negative SZG_E* code on failure, and never
 *    throws / aborts across the ABI; szg_last_error() gives the message (thread-local).
 *  - plain pointers and sizes only.  Host pointers unless the name ends in _dev.
 *  - cgo rule: the library never retains a caller pointer after returning; inputs are
 *    copied, outputs are caller-allocated (only radius results are library-owned).
 *  - vector bytes are EXACTLY stream 1 of a span (encodeDocument, collection.go:713-744):
 *    4-bit high-nibble-first, 16/32/64-bit big-endian.  rowbytes = getVectorSize
 *    (collection.go:796-811).
 *  - threading: concurrent searches on one handle are safe (Search holds only the RLock,
 *    collection.go:570), and so is building or destroying a filter mask next to them
 *    (szg_mask_create / szg_filter_mask / szg_mask_destroy: Search applies its filter under
 *    the RLock, collection.go:592-594) -- a mask must only outlive the searches that name it.
 *    Mutations (upsert/encode/remove/meta_upsert/reserve/fill/set_option/destroy) must not
 *    overlap anything else on the same handle (they run under the write lock,
 *    collection.go:428, 491, 512).  Different handles are independent.
 *  - there is no CPU fallback: without a CUDA device every call fails with SZG_ECUDA.
 */
#ifndef SYZGY_B200_H
#define SYZGY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SZG_OK 0
#define SZG_EINVAL (-1)   /* bad argument (dimension mismatch, unsupported quantization, k too large) */
#define SZG_ECUDA (-2)    /* CUDA runtime failure or no device */
#define SZG_ENOMEM (-3)   /* host or device allocation failed */
#define SZG_ENOTFOUND (-4)
#define SZG_EINTERNAL (-5)

/* DistanceMethod, collection.go:186-189 */
#define SZG_EUCLIDEAN 0
#define SZG_COSINE 1 /* angularDistance: acos(cos)/pi, collection.go:821-832 */

/* search flags */
#define SZG_F_DEFAULT 0u
/* return the fp32/integer surrogate distance converted to the reference's unit instead of
 * the fp64 re-score of the candidate set (faster by a few microseconds; order may then
 * differ from the reference among near-ties).  Default is fp64 verify ON. */
#define SZG_F_NO_FP64_VERIFY 1u

/* sentinel distance szg_rescore() writes for an id that is not in the mirror, so that the
 * shim can return StopSearch like getDocument's error path (collection.go:584-587). */
#define SZG_MISSING_DISTANCE (-1.0)

#define SZG_MAX_K 224u  /* largest K of one top-k call (candidate set is K rounded up + slack) */
#define SZG_MAX_DIM 16384

typedef struct szg_index szg_index;   /* GPU mirror of one collection (or one row shard of it) */
typedef struct szg_result szg_result; /* library-owned radius result */

/* message of the last failure on this thread ("" if none) */
const char *szg_last_error(void);

/*
 * Creates an empty GPU mirror on CUDA device `device`.
 * Replaces: the per-collection state NewCollection sets up for searching
 * (collection.go:224-314: DimensionCount, Quantization (0 => 64, 254-256), DistanceMethod).
 */
int szg_create(int dim, int quantization, int metric, int device, szg_index **out);

/*
 * One handle over ndev CUDA devices of this box (the reference serves a collection from ONE process,
 * rest.go:20-23, so multi-GPU has to sit behind the same handle the shim already holds).  The rows are dealt to the devices
 * (record id % ndev; a synthetic range is cut in ndev contiguous pieces), every entry point of this header works on the
 * handle as on a single-device one, and results do not depend on ndev: a search runs on every device at once, the devices
 * store their k x 16-byte lists straight into devices[0]'s memory over NVLink (peer stores, then a system-scope release of
 * a per-query arrival counter) and the merge kernel on devices[0] waits on those counters -- no collective library, no
 * host hop.  devices[0] needs peer access from the others (any NVLink / NVSwitch box); a device may be listed more than
 * once (several shards on one GPU: used by the tests on single-GPU machines).  Device-resident calls (*_dev) take their
 * pointers on devices[0].
 */
int szg_create_sharded(int dim, int quantization, int metric, const int *devices, int ndev, szg_index **out);

/* Replaces: Collection.Close for the mirror (collection.go:408-421). */
int szg_destroy(szg_index *h);

/* pre-sizes device storage for nrows records (optional) */
int szg_reserve(szg_index *h, uint64_t nrows);

/*
 * Inserts or replaces n records.  codes = n * rowbytes bytes, record i = stream 1 of the span
 * of ids[i].  Replaces: the vector half of AddDocument (collection.go:427-457, after
 * WriteRecord 446-453) and of the reload loop in NewCollection (298-311).  An existing id is
 * overwritten in place (spanfile.go:459-472 replaces the span).
 */
int szg_upsert(szg_index *h, const uint64_t *ids, const uint8_t *codes, uint64_t n);

/*
 * Ingest-side quantization: encodes n float64 vectors (n * dim values, host memory) on
 * the device exactly as encodeDocument does (collection.go:713-744) with quantize (quantization.go:5-23: clamp to
 * [-1, 1], (v + 1) / 2 * (2^b - 1) in that operation order, math.Round = half away from zero; 32-bit =
 * Float32bits(float32(v)), 64-bit = Float64bits(v); 4-bit pairs with the even element in the high nibble,
 * 16/32/64-bit big-endian).  out_codes (optional, n * rowbytes bytes, caller-allocated) receives the stream-1
 * bytes the caller hands to WriteRecord (collection.go:446-453); upsert != 0 also puts the rows into the mirror
 * under ids[i], like szg_upsert.  Replaces the per-element Go loop of encodeDocument in AddDocument (427-457) for
 * bulk ingest.  A NaN element of a 4/8/16-bit collection encodes as code 0 (what Go/amd64's uint64(NaN) leaves in
 * the low bits); NaN payloads of 32-bit collections are not preserved.
 */
int szg_encode(szg_index *h, const uint64_t *ids, const double *vectors, uint64_t n, uint8_t *out_codes, int upsert);

/* Replaces: removeDocument (collection.go:511-521).  *n_removed (optional) = ids that existed. */
int szg_remove(szg_index *h, const uint64_t *ids, uint64_t n, uint64_t *n_removed);

/* number of live records (SpanFile.GetStats numRecords, spanfile.go:562-566) */
int szg_count(szg_index *h, uint64_t *n);

/*
 * Registers the outcome of a FilterFn (collection.go:184, applied at 592-594) as a bitmask:
 * pass[i] != 0 <=> Filter(ids[i], metadata) returned true.  Ids not listed do not pass.
 * The shim evaluates the Go predicate once per record and caches the mask per filter
 * string; it must be rebuilt after UpdateDocument/AddDocument/removeDocument.
 */
int szg_mask_create(szg_index *h, const uint64_t *ids, const uint8_t *pass, uint64_t n, int *mask_id);
int szg_mask_destroy(szg_index *h, int mask_id);

/*
 * Metadata filters evaluated on the device.  Replaces, for filters built by BuildFilter
 * (collection.go:204-218), the per-document json.Unmarshal + closure walk of query/compiler.go:15-165, 477-497 that
 * Search applies at collection.go:592-594: the shim mirrors the scalar metadata fields it wants to filter on into
 * columns (once per AddDocument/UpdateDocument, where it holds the JSON anyway), lowers the filter's syntax tree
 * (query/parser.go node types) to a postfix program, and szg_filter_mask evaluates it for every row into a mask
 * usable as mask_id of every search call.  Filters the lowering does not cover (ANY/ALL, LENGTH, array indexing,
 * comparisons between two fields) keep using szg_mask_create with the Go predicate.
 *
 * Value kinds mirror what encoding/json puts into an interface{}: a missing key reads as nil (getField,
 * compiler.go:428-444); SZG_MV_OTHER = array or object; SZG_MV_ERROR = evaluating the (nested) field raises an error
 * (compiler.go:222-233: key not found below the top level).
 */
#define SZG_MV_MISSING 0u
#define SZG_MV_NULL 1u
#define SZG_MV_BOOL 2u   /* num = 0 / 1 */
#define SZG_MV_NUMBER 3u /* num */
#define SZG_MV_STRING 4u /* str, str_len */
#define SZG_MV_OTHER 5u
#define SZG_MV_ERROR 6u
/* per document: metadata that json.Unmarshal rejects never passes a filter (compiler.go:480-483) */
#define SZG_DOC_INVALID 0u
#define SZG_DOC_OBJECT 1u
#define SZG_DOC_OTHER 2u /* valid JSON that is not an object: every field access is an error */
#define SZG_MAX_META_COLUMNS 32u

typedef struct szg_meta_value {
    uint32_t kind;
    uint32_t str_len;
    double num;
    const char *str;
} szg_meta_value;

/* Sets, for n documents already in the mirror, the document kind and the values of ncols columns
 * (values[i * ncols + j] = column cols[j] of ids[i]).  Columns not listed keep their values; removing a document
 * clears all of them.  Unknown ids are an error. */
int szg_meta_upsert(szg_index *h, const uint64_t *ids, uint64_t n, const uint8_t *doc_kind, const uint32_t *cols,
                    uint32_t ncols, const szg_meta_value *values);

/* postfix program; operands are pushed left to right (IN: the tested value first, then the list's elements) */
#define SZG_FOP_COL 1u        /* arg = column: IdentifierNode of a mirrored field */
#define SZG_FOP_NUM 2u        /* ValueNode float64 (parser.go:472-479) */
#define SZG_FOP_STR 3u        /* ValueNode string */
#define SZG_FOP_BOOL 4u       /* ValueNode bool: num != 0 */
#define SZG_FOP_NULL 5u       /* ValueNode nil */
#define SZG_FOP_EQ 6u         /* compiler.go:172-175 reflect.DeepEqual */
#define SZG_FOP_NE 7u
#define SZG_FOP_LT 8u         /* compiler.go:266-326 compareValues */
#define SZG_FOP_LE 9u
#define SZG_FOP_GT 10u
#define SZG_FOP_GE 11u
#define SZG_FOP_AND 12u       /* compiler.go:178-184 */
#define SZG_FOP_OR 13u        /* compiler.go:185-197 */
#define SZG_FOP_NOT 14u       /* compiler.go:198-203 */
#define SZG_FOP_IN 15u        /* arg = number of list elements; compiler.go:379-393 */
#define SZG_FOP_NOT_IN 16u
#define SZG_FOP_CONTAINS 17u    /* str = the literal right operand; compiler.go:395-420 */
#define SZG_FOP_STARTS_WITH 18u
#define SZG_FOP_ENDS_WITH 19u
#define SZG_FOP_STR_TABLE 20u /* table[code] for every dictionary code (e.g. MATCHES, evaluated by the shim's regexp
                                 over szg_meta_dictionary_get); a non-string operand is an error like 416-419 */
#define SZG_FOP_EXISTS 21u      /* arg = column; compiler.go:340-345 */
#define SZG_FOP_NOT_EXISTS 22u  /* arg = column; compiler.go:60-75 */

typedef struct szg_filter_op {
    uint32_t op;
    uint32_t arg;
    double num;
    const char *str;
    uint32_t str_len;
    uint32_t table_len;
    const uint8_t *table;
} szg_filter_op;

/* Evaluates the program for every row on the device; *mask_id is then valid for the search calls until
 * szg_mask_destroy.  Like any mask it reflects the metadata at the time of the call. */
int szg_filter_mask(szg_index *h, const szg_filter_op *ops, uint32_t nops, int *mask_id);

/* the string dictionary of the metadata columns (codes 0 .. size-1), for predicates the shim tabulates itself */
int szg_meta_dictionary_size(szg_index *h, uint32_t *size);
int szg_meta_dictionary_get(szg_index *h, uint32_t code, const char **str, uint32_t *len);


/*
 * Exact top-k for nq queries.  A single query is a memory-bound scan (it streams the whole
 * mirror: single-query GEMV semantics); from SZG_OPT_BATCH_MIN_QUERIES queries on (2 to 12 by default, see there) the call is a
 * batch and takes the tensor-core contraction of szg_search_batch when the geometry fits -- same
 * results, bit for bit.  Replaces: Search with Precision=="exact",
 * Radius==0, K>0 (collection.go:672-684 driving consider 583-629 and the drain 693-697).
 *   queries  nq*dim float64, never quantized (collection.go:596)
 *   mask_id  -1 = no filter
 *   out_ids / out_dist   nq*k, ascending distance per query; out_n[q] = results of query q
 *   scanned  (optional) records considered per query = live count (filtered rows count,
 *            collection.go:589 precedes 592) -> PercentSearched
 * Results: ids and order equal the reference scan in lexicographic-decimal-id order
 * (spanfile.go:540-560) except among distances closer than 1e-5 relative; NaN distances
 * (cosine ratio rounding above 1) are never returned.
 */
int szg_search_topk(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id,
                    uint32_t flags, uint64_t *out_ids, double *out_dist, uint32_t *out_n,
                    uint64_t *scanned);

/*
 * Batched exact top-k: the results of nq independent szg_search_topk calls, computed as one dense
 * contraction on the tensor cores (tcgen05.mma kind::i8 over the digit planes of the fixed-point
 * queries, fused threshold top-k epilogue) when the collection is 4-, 8- or 16-bit (4-bit through a
 * one-byte-per-code copy, 16-bit through a byte-planar copy kept next to the mirror), k <= 112 and the
 * row has an even number of 16-dimension chunks (at most 64); otherwise the call is served by the
 * streaming scan.  Replaces: B concurrent Search calls under the RLock (collection.go:569-570).
 * Same outputs, same certification / escalation.
 */
int szg_search_batch(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id,
                     uint32_t flags, uint64_t *out_ids, double *out_dist, uint32_t *out_n,
                     uint64_t *scanned);

/*
 * Radius search: every record with distance <= radius (inclusive, K ignored), ascending.
 * Replaces: Search with Radius>0, Precision=="exact" (collection.go:598-605).
 */
int szg_search_radius(szg_index *h, const double *query, double radius, int mask_id, uint32_t flags,
                      szg_result **out, uint64_t *scanned);
/* nq radius searches in one call (many Search requests arriving together, rest.go:371-487): out[q] receives the result of
 * query q with radius radii[q]; they share the launches and the copies.  Every result is freed with szg_result_free. */
int szg_search_radius_batch(szg_index *h, const double *queries, uint32_t nq, const double *radii, int mask_id,
                            uint32_t flags, szg_result **out, uint64_t *scanned);
int szg_result_count(const szg_result *r, uint64_t *n);
int szg_result_fetch(const szg_result *r, uint64_t offset, uint64_t n, uint64_t *out_ids,
                     double *out_dist);
void szg_result_free(szg_result *r);

/*
 * fp64 distances, in the reference's operation order, of the records ids[0..m) to the
 * query (visit order preserved).  Replaces: decodeVector + c.distance inside consider when
 * it is driven by lshTree.search (lshtree.go:316-335 -> collection.go:584-596): the host
 * traversal emits candidate ids, this call scores them, the host replays the
 * accept/radius/k_counter logic over the returned array.  Missing id =>
 * SZG_MISSING_DISTANCE.  Filter and tombstones are NOT applied here (the replay does it).
 */
int szg_rescore(szg_index *h, const double *query, const uint64_t *ids, uint64_t m, double *out_dist);
/* The candidate lists of nlists queries in one call (one copy each way, one launch): list l = ids[list_offsets[l] ..
 * list_offsets[l + 1]) is scored against queries[l * dim ..]; out_dist is indexed like ids.  What a shim serving several
 * medium-precision searches at once (or one search's speculative batches of several trees) calls. */
int szg_rescore_batch(szg_index *h, const double *queries, uint32_t nlists, const uint64_t *ids,
                      const uint64_t *list_offsets, double *out_dist);

/* ---- span file -> mirror -------------------------------------------------------
 * Direct reader of a collection's .dat file (grammar spanfile.go:1-22).  Replaces, for loading, OpenFile's
 * scanFile (spanfile.go:282-357) and NewCollection's per-record reload loop (collection.go:298-311): the file is
 * mapped read-only, every span is CRC-checked and parsed by all host threads, the highest sequence number per
 * record id wins (first seen wins a tie), corrupt spans are skipped like the reference does, and stream 1 of every
 * live document is uploaded to the mirror in bulk.  The file is never written; pointers returned by
 * szg_spanfile_record point into the mapping and stay valid until szg_spanfile_close. */
typedef struct szg_spanfile szg_spanfile;
typedef struct szg_spanfile_info {
    uint64_t file_bytes;
    uint64_t records;          /* live documents: ids strconv.ParseUint accepts, canonical spelling (GetStats numRecords) */
    uint64_t spans_active;     /* valid 'SPAN' spans, superseded versions included */
    uint64_t spans_free;       /* 'FREE' spans */
    uint64_t spans_corrupt;    /* 'SPAN' spans that failed the checksum or the parse: skipped (spanfile.go:313-329) */
    uint64_t free_bytes;       /* free spans + the zero tail of the file */
    uint64_t foreign_records;  /* record ids that are not documents (ignored by collection.go:299-302) */
    uint32_t next_sequence;    /* highest sequence number + 1 (spanfile.go:355) */
    int32_t has_header;        /* record "" with the CollectionOptions JSON was found (collection.go:241-252) */
    int32_t distance_method;   /* 0 euclidean, 1 cosine; -1 without header */
    int32_t dimension_count;
    int32_t quantization;      /* 0 in the file means 64 (collection.go:254-256) */
    char name[256];
} szg_spanfile_info;
int szg_spanfile_open(const char *path, szg_spanfile **out);
int szg_spanfile_close(szg_spanfile *sf);
int szg_spanfile_get_info(const szg_spanfile *sf, szg_spanfile_info *out);
/* live document ids in IterateSortedRecords order (sort.Strings of the decimal ids, spanfile.go:540-560);
 * *n = total, at most cap are written */
int szg_spanfile_ids(const szg_spanfile *sf, uint64_t *out_ids, uint64_t cap, uint64_t *n);
/* stream 1 (vector bytes) and stream 0 (metadata) of a document, SpanReader.getStream (spanfile.go:67-118);
 * SZG_ENOTFOUND like ReadRecord (spanfile.go:513-519) */
int szg_spanfile_record(const szg_spanfile *sf, uint64_t id, const uint8_t **vector, uint64_t *vector_len,
                        const uint8_t **metadata, uint64_t *metadata_len);
/* bulk szg_upsert of every live document into h (whose dimension / quantization must match the header) */
int szg_spanfile_load(const szg_spanfile *sf, szg_index *h, uint64_t *loaded);

/* ---- device-resident variants (inputs/outputs already in HBM on the handle's device) ----
 * Used by the multi-GPU host and by the benchmark's resident-input leg.  All work is
 * enqueued on `stream` (a cudaStream_t, NULL = default stream) and its helper streams are
 * joined back into it before returning; nothing is synchronised with the host.
 * d_out_ids/d_out_dist are nq*k, d_out_n nq (uint32).  The device variant cannot re-run a query
 * (no host synchronisation): queries whose candidate set could not be certified against the
 * surrogate's error bound are reported in d_out_flags and the caller re-runs them through
 * szg_search_topk (which escalates on its own). */
int szg_search_topk_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id,
                        uint32_t flags, uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                        uint32_t *d_out_flags /* optional, nq: bit0 = result not certified */, void *stream);

/* Device-resident form of szg_search_batch (tensor-core contraction when the geometry fits, the streaming scan
 * otherwise); same contract as szg_search_topk_dev. */
int szg_search_batch_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id,
                         uint32_t flags, uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                         uint32_t *d_out_flags /* optional */, void *stream);

/*
 * Final merge of row-sharded top-k lists: rank g's {ids nq*k, dist nq*k, n nq}
 * as gathered by ncclAllGather; writes the global top-k per query (ascending distance, ties
 * by lexicographic decimal id).  rank_stride_bytes == 0: the three arrays are separate and
 * tight ([G][nq][k], [G][nq][k], [G][nq]); otherwise the pointers name rank 0's arrays
 * inside one packed per-rank record and every rank's record starts rank_stride_bytes after
 * the previous one (one allgather of one buffer).  No reference counterpart: the reference
 * is single-process.
 */
int szg_merge_topk_dev(szg_index *h, const uint64_t *d_gathered_ids, const double *d_gathered_dist,
                       const uint32_t *d_gathered_n, const uint32_t *d_gathered_flags /* optional */,
                       uint64_t rank_stride_bytes, uint32_t nranks, uint32_t nq, uint32_t k,
                       uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n,
                       uint32_t *d_out_flags /* optional: OR of the ranks' flags */, void *stream);

/* ---- synthetic data + introspection (bench/test helpers, no reference counterpart) ---- */

/* Appends rows [row0, row0+nrows) of the synthetic collection `seed` (generator documented
 * in oracle/syzgy_oracle.c orc_synth_rows; ids = row index) directly in HBM. */
int szg_fill_synthetic(szg_index *h, uint64_t seed, uint64_t row0, uint64_t nrows);

/* copies records back in stream-1 byte format (round-trip check of the mirror layout) */
int szg_fetch_codes(szg_index *h, const uint64_t *ids, uint64_t n, uint8_t *out_codes);

typedef struct szg_stats {
    uint64_t kernel_launches;   /* kernels this handle launched since creation */
    uint64_t escalations;       /* top-k calls re-run with a larger candidate set */
    uint64_t uncertain_results; /* top-k queries whose candidate margin stayed below tolerance */
    uint64_t batch_queries;     /* queries served by the tensor-core batched path */
    uint64_t device_bytes;      /* HBM held by the mirror */
    uint64_t live_rows;
    uint64_t slots;             /* rows of HBM layout in use (live + tombstones) */
    uint32_t rowbytes;          /* getVectorSize(quant, dim) */
    uint32_t pitch;             /* bytes per row in the column-blocked layout */
    uint32_t sm_count;
    uint32_t scan_grid;         /* CTAs of one scan launch */
    uint32_t scan_block;        /* threads per CTA */
    uint32_t scan_stages;       /* shared-memory ring stages per warp */
    uint32_t scan_tile_bytes;   /* bytes of one bulk-copied tile */
    uint32_t scan_smem_bytes;   /* dynamic shared memory per CTA */
    uint32_t shards;            /* devices of a sharded handle (1 otherwise) */
    uint64_t combined_queries;  /* queries of concurrent szg_search_topk calls that were answered by a shared launch */
    uint64_t graph_launches;    /* host-buffer top-k calls replayed as one captured launch sequence (CUDA graph) */
} szg_stats;
int szg_get_stats(szg_index *h, szg_stats *out);

/* tuning / test knobs */
#define SZG_OPT_STREAMS 1            /* streams a multi-query call is spread over (1..4, default 2) */
#define SZG_OPT_TIMING 2             /* CUDA events around every scan launch: 0 off (default), 1 keep the last call's,
                                        2 accumulate over calls until szg_last_scan_times_ms reads them.  While it is on,
                                        calls are not replayed as captured launch sequences */
#define SZG_OPT_MIN_CANDIDATE_MODE 3 /* force candidate set >= 32<<v (v in 0..3; -1 = automatic) */
#define SZG_OPT_SCAN_WARPS 4         /* warps per scan CTA: 8 or 16 (default 16) */
#define SZG_OPT_SCAN_STAGES 5        /* ring stages per warp, 2..8 (default 2) */
#define SZG_OPT_SCAN_TILE_CHUNKS 6   /* upper bound of 16-byte chunks per tile, 1..32 (default 8) */
#define SZG_OPT_DIGITS 7             /* fixed-point digits of the query: 0 automatic (2, re-run with 3 when the
                                        result cannot be certified), 2 or 3 forced */
#define SZG_OPT_BATCH_TENSOR 8       /* 0: szg_search_batch uses the streaming scan; 1 (default): tensor cores */
#define SZG_OPT_COMBINE 9            /* 1 (default): concurrent szg_search_topk calls on one handle (Search holds only the
                                        RLock, collection.go:570) with the same k / mask / flags share one launch: whoever
                                        arrives while a launch is running is answered together by the next one; 0: off */
#define SZG_OPT_BATCH_MIN_QUERIES 10 /* szg_search_topk(_dev) calls with at least this many queries are batches: they take
                                        the tensor-core contraction when its geometry fits.  0 (default): chosen by the size
                                        of the mirror from the measured crossovers -- 2 from 2 GB on, 3 from 512 MB on (the
                                        rows stream from HBM once per call instead of once per query), else 12 (cache-resident
                                        scans cost 3-13 us per extra query, the contraction ~100 us per call) */
#define SZG_OPT_GRAPHS 11            /* 1 (default): repeated host-buffer top-k call shapes are replayed as one captured
                                        launch sequence (CUDA graph) instead of launch by launch; 0: off */
#define SZG_OPT_TRACE_BUFFER 12      /* profiling: a device pointer to 8 int64 words that finalize_kernel's first CTA fills with
                                        clock64() at its phase boundaries (0 = off) */
int szg_set_option(szg_index *h, int option, int64_t value);

/* Time of the most recent scan launches on this handle, measured with CUDA events on the
 * launching stream: fills up to cap entries (ms), returns how many via *n. */
int szg_last_scan_times_ms(szg_index *h, float *out_ms, uint32_t cap, uint32_t *n);

#ifdef __cplusplus
}
#endif
#endif /* SYZGY_B200_H */
