// Exact selection over the candidate superset a scan emitted -- the second half of the fused
// epilogue.  Reproduces the reference's print order and cutoffs:
//   Mode B (src/lib.rs:243-265): sort (distance, subject); cutoff = k-th smallest distance, or the
//          largest when fewer than k rows exist; keep every row <= cutoff (ties included).
//   Mode A (src/lib.rs:298-312): all rows at the minimum distance, ascending subject == Mode B, k=1.
// One radix sort on the packed 64-bit key (query | distance | subject), segment boundaries by a
// neighbour compare, a stable stream compaction, then conversion to smafa_hit rows.  The same
// routine is the multi-GPU merge: the union of per-shard supersets is a superset.
// CUB provides the sort and the compaction (library code off the hot path; the scan is ours).
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

namespace smafa {

__global__ void segment_bounds_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ seg_start,
                                      uint32_t *__restrict__ seg_end) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t q = key_q(keys[i]);
  if (i == 0 || key_q(keys[i - 1]) != q) seg_start[q] = (uint32_t)i;
  if (i == n - 1 || key_q(keys[i + 1]) != q) seg_end[q] = (uint32_t)(i + 1);
}

struct KeepWithinKth {
  const uint64_t *keys;
  const uint32_t *seg_start, *seg_end;
  uint32_t k;
  __device__ bool operator()(const uint64_t &key) const {
    uint32_t q = key_q(key);
    uint32_t s = seg_start[q], e = seg_end[q];
    uint64_t idx = (uint64_t)s + k - 1;
    if (idx > e - 1) idx = e - 1;
    return key_d(key) <= key_d(keys[idx]);
  }
};

__global__ void keys_to_hits_kernel(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ n_ptr,
                                    uint64_t hits_cap, uint32_t q_base, uint64_t subject_offset,
                                    smafa_hit *__restrict__ hits, unsigned long long *__restrict__ n_out_pinned) {
  uint64_t n = *n_ptr;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_out_pinned = n;
  if (n > hits_cap) return;  // caller reports SMAFA_E_OOM
  for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t k = keys[i];
    smafa_hit h;
    h.query = key_q(k) + q_base;
    h.subject = (uint32_t)(key_j(k) + subject_offset);
    h.distance = key_d(k);
    hits[i] = h;
  }
}

__global__ void hits_to_keys_kernel(const smafa_hit *__restrict__ hits, uint64_t n, uint64_t *__restrict__ keys,
                                    int *__restrict__ bad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  smafa_hit h = hits[i];
  if (h.query >= MAX_BATCH_QUERIES || h.distance > KEY_D_MASK) atomicOr(bad, 1);
  keys[i] = make_key(h.query, h.distance, h.subject);
}

// Grouped dbs (api.cu group_order): the scan ran on the db in grouped order; perm[row] is the subject number the
// reference knows the window by.  Applied to the candidate keys before the sort, so ties come out in subject order.
__global__ void remap_subjects_kernel(uint64_t *__restrict__ keys, uint64_t n, const uint32_t *__restrict__ perm) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    keys[i] = make_key(key_q(k), key_d(k), perm[key_j(k)]);
  }
}

void launch_remap_subjects(uint64_t *keys, uint64_t n, const uint32_t *perm, cudaStream_t s) {
  if (n == 0) return;
  remap_subjects_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(keys, n, perm);
}

void launch_hits_to_keys(const smafa_hit *hits, uint64_t n, uint64_t *keys, int *bad, cudaStream_t s) {
  if (n == 0) return;
  hits_to_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hits, n, keys, bad);
}

size_t finalize_temp_bytes(uint64_t cap) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, a, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)cap, 0, 64);
  KeepWithinKth pred{nullptr, nullptr, nullptr, 1};
  cub::DeviceSelect::If(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, (unsigned long long *)nullptr,
                        (int64_t)cap, pred);
  return (a > b ? a : b) + 256;
}

int launch_finalize_select(FinalizeWorkspace &ws, const uint64_t *keys, uint64_t n, uint32_t n_queries, uint32_t k, cudaStream_t s) {
  if (n == 0) {
    cudaMemsetAsync(ws.n_selected, 0, sizeof(unsigned long long), s);
    return 0;
  }
  int q_bits = 1;
  while ((1ull << q_bits) < n_queries && q_bits < 20) ++q_bits;
  size_t tb = ws.cub_temp_bytes;
  cub::DeviceRadixSort::SortKeys(ws.cub_temp, tb, keys, ws.keys_sorted, (int64_t)n, 0, KEY_Q_SHIFT + q_bits, s);
  // CUB's onesweep sort: one histogram kernel, one exclusive-sum kernel, one pass per 8-bit digit
  int launches = 2 + (KEY_Q_SHIFT + q_bits + 7) / 8;
  segment_bounds_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.keys_sorted, n, ws.seg_start, ws.seg_end);
  launches += 1;
  KeepWithinKth pred{ws.keys_sorted, ws.seg_start, ws.seg_end, k};
  tb = ws.cub_temp_bytes;
  cub::DeviceSelect::If(ws.cub_temp, tb, ws.keys_sorted, ws.keys_sel, ws.n_selected, (int64_t)n, pred, s);
  launches += 2;  // CUB select: init + sweep
  return launches;
}

int launch_keys_to_hits(FinalizeWorkspace &ws, uint64_t n_max, uint32_t q_base, uint64_t subject_offset, smafa_hit *hits_out,
                        uint64_t hits_cap, unsigned long long *n_out_pinned, cudaStream_t s) {
  unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_max + 255) / 256, 148 * 8));
  keys_to_hits_kernel<<<blocks, 256, 0, s>>>(ws.keys_sel, ws.n_selected, hits_cap, q_base, subject_offset, hits_out,
                                             n_out_pinned);
  return 1;
}

int launch_finalize(FinalizeWorkspace &ws, const uint64_t *keys, uint64_t n, uint32_t n_queries, uint32_t k,
                    uint32_t q_base, uint64_t subject_offset, smafa_hit *hits_out, uint64_t hits_cap,
                    unsigned long long *n_out_pinned, cudaStream_t s) {
  int launches = launch_finalize_select(ws, keys, n, n_queries, k, s);
  return launches + launch_keys_to_hits(ws, n, q_base, subject_offset, hits_out, hits_cap, n_out_pinned, s);
}

// ---- sort-free selection ("buckets") ------------------------------------------------------------------------------
// The same selection as launch_finalize_select for the common shape of a scan's output: a handful of candidates per
// query.  The scan counted them per query (ScanParams::per_query / max_seg), so a prefix sum gives every query a
// bucket, one pass drops the keys into their buckets, and inside a bucket of m keys a key's place in the (distance,
// subject) order and the number of strictly closer keys are two counts over the m keys -- m*m work, trivial for
// m <= BUCKET_MAX -- which is all the reference's rule needs: keep a key iff fewer than k keys of its query are strictly
// closer (src/lib.rs:253-265; k = 1 is Mode A).  Kept keys are a prefix of their bucket's order, so a second prefix
// sum places them.  No sort, and no row count on the host: every kernel takes the candidate count from device memory
// and does nothing unless fast_ok says the speculation holds (no candidate overflow, no bucket above BUCKET_MAX, valid
// query codes) -- the caller reads fast_ok back with the result and re-runs the batch through the sort when it is 0.
constexpr uint32_t BUCKET_MAX = 256;

// max_seg = largest per-query candidate count (the scan only counts: its emission path must not wait on an atomic)
__global__ void bucket_max_kernel(const uint32_t *__restrict__ per_query, uint32_t Q, uint32_t *__restrict__ max_seg) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += gridDim.x * blockDim.x) m = max(m, per_query[i]);
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(max_seg, m);
}

__global__ void fast_ok_kernel(const unsigned long long *__restrict__ cand_count, uint64_t cap, const uint32_t *__restrict__ max_seg,
                               const int *__restrict__ q_invalid, unsigned long long *__restrict__ fast_ok) {
  *fast_ok = (*cand_count <= cap && *max_seg <= BUCKET_MAX && *q_invalid == 0) ? 1ull : 0ull;
}

__global__ void bucket_scatter_kernel(const uint64_t *__restrict__ cand, const unsigned long long *__restrict__ n_ptr,
                                      const unsigned long long *__restrict__ fast_ok, const uint32_t *__restrict__ seg_start,
                                      uint32_t *__restrict__ fill, const uint32_t *__restrict__ perm, uint64_t *__restrict__ bucket) {
  if (!*fast_ok) return;
  const uint64_t n = *n_ptr;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t k = cand[i];
    if (perm != nullptr) k = make_key(key_q(k), key_d(k), perm[key_j(k)]);  // grouped / mapped db: row -> subject number
    const uint32_t q = key_q(k);
    bucket[seg_start[q] + atomicAdd(fill + q, 1u)] = k;
  }
}

__global__ void bucket_select_kernel(const uint64_t *__restrict__ bucket, const unsigned long long *__restrict__ n_ptr,
                                     const unsigned long long *__restrict__ fast_ok, const uint32_t *__restrict__ seg_start,
                                     const uint32_t *__restrict__ per_query, uint32_t k, uint32_t *__restrict__ info,
                                     uint32_t *__restrict__ kept) {
  if (!*fast_ok) return;
  const uint64_t n = *n_ptr;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = bucket[i];
    const uint32_t q = key_q(key), d = key_d(key);
    const uint32_t s = seg_start[q], m = per_query[q];
    uint32_t before = 0, closer = 0;
    for (uint32_t j = s; j < s + m; ++j) {
      const uint64_t o = bucket[j];
      before += o < key;
      closer += key_d(o) < d;
    }
    const bool keep = closer < k;
    info[i] = before | (keep ? 0x80000000u : 0u);
    if (keep) atomicAdd(kept + q, 1u);
  }
}

__global__ void bucket_write_kernel(const uint64_t *__restrict__ bucket, const unsigned long long *__restrict__ n_ptr,
                                    const unsigned long long *__restrict__ fast_ok, const uint32_t *__restrict__ info,
                                    const uint32_t *__restrict__ kept_start, uint32_t Q, uint64_t *__restrict__ keys_sel,
                                    unsigned long long *__restrict__ n_selected) {
  const bool ok = *fast_ok != 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_selected = ok ? kept_start[Q] : 0ull;
  if (!ok) return;
  const uint64_t n = *n_ptr;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t f = info[i];
    if (f & 0x80000000u) keys_sel[kept_start[key_q(bucket[i])] + (f & 0x7fffffffu)] = bucket[i];
  }
}

size_t bucket_temp_bytes(uint32_t Q) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)(Q + 1));
  return b + 256;
}

// counters = [per_query | fill | kept], each Q + 1 u32 (zeroed by the caller before the scan; per_query filled by it);
// starts = [seg_start | kept_start], each Q + 1.  Result: ws.keys_sel / *ws.n_selected like launch_finalize_select.
int launch_finalize_buckets(FinalizeWorkspace &ws, const uint64_t *cand, const unsigned long long *cand_count, uint64_t cap,
                            const uint32_t *max_seg, const int *q_invalid, unsigned long long *fast_ok, uint32_t Q, uint32_t k,
                            uint32_t *counters, uint32_t *starts, uint32_t *info, void *temp, size_t temp_bytes,
                            const uint32_t *perm, uint64_t n_hint, cudaStream_t s) {
  uint32_t *per_query = counters, *fill = counters + (Q + 1), *kept = counters + 2 * (size_t)(Q + 1);
  uint32_t *seg_start = starts, *kept_start = starts + (Q + 1);
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_hint + 255) / 256, 148 * 8));
  bucket_max_kernel<<<(Q + 1023) / 1024 > 296 ? 296 : (Q + 1023) / 1024, 1024, 0, s>>>(per_query, Q, const_cast<uint32_t *>(max_seg));
  fast_ok_kernel<<<1, 1, 0, s>>>(cand_count, cap, max_seg, q_invalid, fast_ok);
  size_t tb = temp_bytes;
  cub::DeviceScan::ExclusiveSum(temp, tb, per_query, seg_start, (int)(Q + 1), s);
  bucket_scatter_kernel<<<grid, 256, 0, s>>>(cand, cand_count, fast_ok, seg_start, fill, perm, ws.keys_sorted);
  bucket_select_kernel<<<grid, 256, 0, s>>>(ws.keys_sorted, cand_count, fast_ok, seg_start, per_query, k, info, kept);
  tb = temp_bytes;
  cub::DeviceScan::ExclusiveSum(temp, tb, kept, kept_start, (int)(Q + 1), s);
  bucket_write_kernel<<<grid, 256, 0, s>>>(ws.keys_sorted, cand_count, fast_ok, info, kept_start, Q, ws.keys_sel, ws.n_selected);
  return 9;  // ours: 5, CUB exclusive sum: 2 x (init + scan)
}

}  // namespace smafa
