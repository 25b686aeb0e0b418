// Host-callable launchers of the device code (internal; the public boundary is include/smafa_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smafa_b200.h"

namespace smafa {

struct ScanParams;

// pack.cu
void launch_pack_planes(const uint64_t *ref, uint32_t n, uint32_t W, uint32_t L, uint32_t row_words, int alphabet,
                        uint32_t *planes, int *invalid, cudaStream_t s);
void launch_init_bound(int *bound, uint32_t Q, int v, cudaStream_t s);
// sets *invalid when a nucleotide word is not made of valid one-hot codes (what pack_planes checks, without the planes)
void launch_check_codes(const uint64_t *ref, uint32_t n, uint32_t W, uint32_t L, int *invalid, cudaStream_t s);

// scan_popc.cu -- return the number of kernels launched
int launch_scan_popc(const ScanParams &p, bool early, uint32_t chunk, cudaStream_t s);
int launch_scan_generic(const ScanParams &p, uint32_t chunk, cudaStream_t s);
// Bound estimator on every tile_stride-th 256-window tile (no emission); see scan_popc.cu
int launch_bound_prepass(const ScanParams &p, uint32_t tile_stride, cudaStream_t s);
void launch_distances(const uint64_t *q_ref, uint32_t Q, const uint64_t *d_ref, uint32_t D, uint32_t W, int alphabet,
                      uint16_t *out, cudaStream_t s);
int popc_tile_rows();

// guess.cu -- optimistic first pass under a guessed bound (see the file header)
int launch_sample_hist(const uint64_t *q_ref, uint32_t Q, uint32_t q_stride, const uint64_t *d_ref, uint32_t d_stride,
                       uint32_t n_d, uint32_t W, int alphabet, unsigned long long *ghist, cudaStream_t s);
int guess_bins();
// Selectivity of the union-row filter of degree 1..3 at need = L - bound on a strided sample: counts[0..2] = passing
// (query, row) pairs, counts[3] = samples (guess.cu)
int launch_union_sample(const uint64_t *q_ref, uint32_t Q, uint32_t q_stride, const uint64_t *d_ref, uint32_t D, uint32_t d_stride,
                        uint32_t n_d, uint32_t W, int need, unsigned long long *counts, cudaStream_t s);
void launch_count_per_query(const uint64_t *cand, uint64_t n, uint32_t Q, uint32_t *per_query, cudaStream_t s);
void launch_list_unfinished(const uint32_t *per_query, uint32_t Q, uint32_t need, uint32_t *list, uint32_t *n_list, cudaStream_t s);
void launch_gather_queries(const uint64_t *q_ref, const uint32_t *list, uint32_t n, uint32_t W, uint64_t *out, cudaStream_t s);
void launch_keep_finished(const uint64_t *cand, uint64_t n, const uint32_t *per_query, uint32_t need, uint64_t *out,
                          unsigned long long *n_out, cudaStream_t s);
void launch_remap_queries(uint64_t *cand, const unsigned long long *begin, const unsigned long long *end, uint64_t cap,
                          const uint32_t *list, cudaStream_t s);

// finalize.cu
struct FinalizeWorkspace {
  uint64_t *keys_sorted = nullptr;  // [cap]
  uint64_t *keys_sel = nullptr;     // [cap]
  uint32_t *seg_start = nullptr;    // [MAX_BATCH_QUERIES]
  uint32_t *seg_end = nullptr;      // [MAX_BATCH_QUERIES]
  unsigned long long *n_selected = nullptr;  // device scalar
  void *cub_temp = nullptr;
  size_t cub_temp_bytes = 0;
  uint64_t cap = 0;
};
size_t finalize_temp_bytes(uint64_t cap);
// Sorts n candidate keys by (query, distance, subject), keeps per query everything <= the k-th
// smallest distance (k = UINT32_MAX keeps all) and writes smafa_hit rows.  Returns kernels launched;
// *n_out_host is valid after the stream has been synchronised by the caller (pinned host scalar).
int launch_finalize(FinalizeWorkspace &ws, const uint64_t *keys, uint64_t n, uint32_t n_queries, uint32_t k,
                    uint32_t q_base, uint64_t subject_offset, smafa_hit *hits_out, uint64_t hits_cap,
                    unsigned long long *n_out_pinned, cudaStream_t s);
// The two halves of launch_finalize: sort + cutoff -> ws.keys_sel / *ws.n_selected (device); keys -> smafa_hit rows.
int launch_finalize_select(FinalizeWorkspace &ws, const uint64_t *keys, uint64_t n, uint32_t n_queries, uint32_t k, cudaStream_t s);
int launch_keys_to_hits(FinalizeWorkspace &ws, uint64_t n_max, uint32_t q_base, uint64_t subject_offset, smafa_hit *hits_out,
                        uint64_t hits_cap, unsigned long long *n_out_pinned, cudaStream_t s);

// Sort-free selection for scans that counted their candidates per query (finalize.cu "buckets"): same result as
// launch_finalize_select, no host-known row count; *fast_ok = 0 (and *ws.n_selected = 0) when the speculation fails.
size_t bucket_temp_bytes(uint32_t Q);
int launch_finalize_buckets(FinalizeWorkspace &ws, const uint64_t *cand, const unsigned long long *cand_count, uint64_t cap,
                            const uint32_t *max_seg, const int *q_invalid, unsigned long long *fast_ok, uint32_t Q, uint32_t k,
                            uint32_t *counters, uint32_t *starts, uint32_t *info, void *temp, size_t temp_bytes,
                            const uint32_t *perm, uint64_t n_hint, cudaStream_t s);

// merge.cu -- multi-GPU merge of per-shard blocks (see the file header)
struct MergeWorkspace {
  uint32_t *seg = nullptr;       // [n_ranks][Q + 1]
  uint64_t *merged = nullptr;    // [n_ranks * cap]
  uint8_t *flags = nullptr;      // [n_ranks * cap]
  uint64_t *selected = nullptr;  // [n_ranks * cap]
  unsigned long long *n_selected = nullptr;  // device scalar
  void *cub_temp = nullptr;
  size_t cub_temp_bytes = 0;
  uint64_t rows = 0;             // n_ranks * cap the buffers are sized for
  uint64_t seg_len = 0;          // n_ranks * (Q + 1) seg is sized for
};
size_t merge_temp_bytes(uint64_t rows);
// block = {rows, status, keys...}: rows := 0, status as given
void launch_block_reset(uint64_t *block, uint64_t status, cudaStream_t s);
// appends the *n_ptr (<= n_max) selected keys of a finished batch to a send block; returns kernels launched
int launch_block_append(const uint64_t *keys, const unsigned long long *n_ptr, uint64_t n_max, uint32_t q_off,
                        uint64_t subject_offset, uint64_t *block, uint64_t cap, cudaStream_t s);
// gathered = n_ranks blocks of 2 + cap u64; info_dev[0..3] = {largest announced row count, first status, its rank, rows kept}
int launch_merge_blocks(MergeWorkspace &ws, const uint64_t *gathered, uint32_t n_ranks, uint64_t cap, uint32_t Q, uint32_t k,
                        uint32_t q_base, smafa_hit *hits_out, uint64_t hits_cap, unsigned long long *info_dev, cudaStream_t s);

// the same for degrees 1, 2, 3, 4, 8, 16 (grouped dbs): counts[0..5] passing pairs, counts[6] samples
int launch_union_sample_wide(const uint64_t *q_ref, uint32_t Q, uint32_t q_stride, const uint64_t *d_ref, uint32_t D,
                             uint32_t d_stride, uint32_t n_d, uint32_t W, int need, unsigned long long *counts, cudaStream_t s);
// candidate keys of a grouped db: row number -> subject number (perm on the device)
void launch_remap_subjects(uint64_t *keys, uint64_t n, const uint32_t *perm, cudaStream_t s);
// smafa_hit rows -> candidate keys (for the multi-GPU merge); q must be < 2^20, d < 2^12
void launch_hits_to_keys(const smafa_hit *hits, uint64_t n, uint64_t *keys, int *bad, cudaStream_t s);

}  // namespace smafa
