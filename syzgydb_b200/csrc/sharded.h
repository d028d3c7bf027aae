// sharded.h -- one handle over several devices (szg_create_sharded): the router behind the C ABI.  index.cu / search.cu send
// every entry point here when h->sh is set; the per-device mirrors are ordinary szg_index objects.
#pragma once
#include "index_internal.h"

namespace szg {

int sharded_create(int dim, int quantization, int metric, const int *devices, int ndev, szg_index **out);
int sharded_destroy(szg_index *h);
szg_index *sharded_root(szg_index *h);         // the shard on devices[0]: gather buffers and merges live there
szg_index *sharded_timing_shard(szg_index *h); // whose scan-launch times szg_last_scan_times_ms reports

// mutations (exclusive, like on a single device)
int sharded_reserve(szg_index *h, uint64_t nrows);
int sharded_upsert(szg_index *h, const uint64_t *ids, const uint8_t *codes, const double *vectors, uint8_t *out_codes, uint64_t n,
                   bool into_mirror);
int sharded_remove(szg_index *h, const uint64_t *ids, uint64_t n, uint64_t *n_removed);
int sharded_fill_synthetic(szg_index *h, uint64_t seed, uint64_t row0, uint64_t nrows);
int sharded_fetch_codes(szg_index *h, const uint64_t *ids, uint64_t n, uint8_t *out_codes);
int sharded_mask_create(szg_index *h, const uint64_t *ids, const uint8_t *pass, uint64_t n, int *mask_id);
int sharded_filter_mask(szg_index *h, const szg_filter_op *ops, uint32_t nops, int *mask_id);
int sharded_mask_destroy(szg_index *h, int mask_id);
int sharded_meta_upsert(szg_index *h, const uint64_t *ids, uint64_t n, const uint8_t *doc_kind, const uint32_t *cols, uint32_t ncols,
                        const szg_meta_value *values);
int sharded_set_option(szg_index *h, int option, int64_t value);
int sharded_get_stats(szg_index *h, szg_stats *out);
uint64_t sharded_count(szg_index *h);

// searches
int sharded_search_topk(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags, uint64_t *out_ids,
                        double *out_dist, uint32_t *out_n, uint64_t *scanned, bool prefer_batch);
int sharded_search_topk_dev(szg_index *h, const double *d_queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags,
                            uint64_t *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, void *stream,
                            bool prefer_batch);
int sharded_search_radius(szg_index *h, const double *queries, uint32_t nq, const double *radii, int mask_id, szg_result **out,
                          uint64_t *scanned);
int sharded_rescore(szg_index *h, const double *queries, uint32_t nlists, const uint64_t *ids, const uint64_t *list_offsets,
                    double *out_dist);

// single-device pieces (index.cu / search.cu) the router calls per shard
int create_single(int dim, int quantization, int metric, int device, std::shared_ptr<MetaDict> dict, szg_index **out);
int destroy_single(szg_index *h);
int grow(szg_index *h, uint64_t want_slots);
int upsert_rows(szg_index *h, const uint64_t *ids, const uint8_t *codes, const double *vectors, uint8_t *out_codes, uint64_t n,
                bool into_mirror);
int radius_device(szg_index *h, const double *queries, uint32_t nq, const double *radii, int mask_id, szg_result **out);
int rescore_device(szg_index *h, const double *queries, uint32_t nl, const uint64_t *ids, const uint64_t *off, double *out_dist);
int ws_for_stream(szg_index *h, void *stream, Workspace **out);
int search_host(szg_index *h, const double *queries, uint32_t nq, uint32_t k, int mask_id, uint32_t flags, uint64_t *out_ids,
                double *out_dist, uint32_t *out_n, uint64_t *scanned, bool prefer_batch);
int run_topk(szg_index *h, Workspace *ws, const double *d_q, uint32_t nq, uint32_t k, const uint32_t *mask, uint32_t flags, int mode,
             int nd, unsigned long long *d_out_ids, double *d_out_dist, uint32_t *d_out_n, uint32_t *d_out_flags, const PeerSink *sink);

} // namespace szg
