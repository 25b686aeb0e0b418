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

}  // namespace smafa
