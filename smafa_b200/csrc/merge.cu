// Multi-GPU merge: the reference's selection (src/lib.rs:243-265 Mode B, 298-312 Mode A) applied to the
// union of per-shard answers, without a sort and without a host round trip.
//
// Every shard contributes one BLOCK: [0] = rows it found, [1] = its status (0 = fine), [2 ..] = its rows as candidate
// keys query | distance | GLOBAL subject (common.cuh), already in the reference's print order -- they are the output of
// the shard's own finalize.  A shard's local cutoff is never below the global one, so the union of the blocks is a
// superset of the answer and holds every row at or below the global cutoff.  Blocks arrive from ncclAllGather (one
// process per GPU) or from peer copies (one process driving several GPUs); `gathered` = n_ranks blocks, `stride` u64 apart.
//
// Because every block is sorted, a row's place in the merged order is a sum of binary searches:
//   merge_segments_kernel   seg[r][q] = first row of query q in block r                        (one search per (q, r))
//   merge_rank_kernel       per row (q, d, s) of block r:
//                             less_d = sum_r' #rows of q in r' with distance < d                -> keep  <=>  less_d < k
//                             pos    = sum_r' #rows of r' that sort before (q, d, s)            -> merged[pos] = key, flag[pos] = keep
//                           (Mode B keeps everything <= the k-th smallest distance, ties included: a row is kept iff fewer
//                           than k rows of its query are strictly closer; Mode A is k = 1; k = UINT32_MAX keeps all.)
//   cub::DeviceSelect::Flagged + keys_to_hits                                                  -> smafa_hit rows, count
// All row counts are read on the device (the block headers), so the host only learns the result: rows kept, the largest
// block's need (an overflowed block is re-sent by the caller with a larger capacity) and the first failing status.
#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

namespace smafa {

__device__ __forceinline__ uint32_t lower_bound_key(const uint64_t *__restrict__ keys, uint32_t lo, uint32_t hi, uint64_t key) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(keys + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint32_t block_rows(const uint64_t *__restrict__ block, uint64_t cap) {
  const uint64_t n = block[0];
  return (uint32_t)(n < cap ? n : cap);
}

// seg[r * (Q + 1) + q], q in [0, Q]: first row of query q in block r (seg[r][Q] = rows of block r).
// info[0] = largest row count any block announces (may exceed cap), info[1] = first non-zero status, info[2] = its rank.
__global__ void merge_segments_kernel(const uint64_t *__restrict__ gathered, uint32_t n_ranks, uint64_t stride, uint64_t cap,
                                      uint32_t Q, uint32_t *__restrict__ seg, unsigned long long *__restrict__ info) {
  const uint64_t total = (uint64_t)n_ranks * (Q + 1);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(i / (Q + 1)), q = (uint32_t)(i % (Q + 1));
    const uint64_t *block = gathered + (uint64_t)r * stride;
    const uint32_t n = block_rows(block, cap);
    seg[i] = q == Q ? n : lower_bound_key(block + 2, 0, n, (uint64_t)q << KEY_Q_SHIFT);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long need = 0, status = 0, who = 0;
    for (uint32_t r = 0; r < n_ranks; ++r) {
      const uint64_t *block = gathered + (uint64_t)r * stride;
      need = max(need, (unsigned long long)block[0]);
      if (status == 0 && block[1] != 0) { status = block[1]; who = r; }
    }
    info[0] = need;
    info[1] = status;
    info[2] = who;
  }
}

__global__ void merge_rank_kernel(const uint64_t *__restrict__ gathered, uint32_t n_ranks, uint64_t stride, uint64_t cap, uint32_t Q,
                                  const uint32_t *__restrict__ seg, uint32_t k, uint64_t *__restrict__ merged,
                                  uint8_t *__restrict__ flags) {
  const uint64_t total = (uint64_t)n_ranks * cap;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(t / cap), i = (uint32_t)(t % cap);
    const uint64_t *mine = gathered + (uint64_t)r * stride;
    if (i >= block_rows(mine, cap)) continue;
    const uint64_t key = mine[2 + i];
    const uint32_t q = key_q(key);
    if (q >= Q) continue;  // cannot happen for blocks built by this library; never write out of bounds on foreign input
    const uint64_t key_d0 = key & ~(uint64_t)0xFFFFFFFFu;  // (q, d, subject 0): first row at this distance
    uint64_t pos = 0;
    uint32_t less_d = 0;
    for (uint32_t o = 0; o < n_ranks; ++o) {
      const uint64_t *keys = gathered + (uint64_t)o * stride + 2;
      const uint32_t s = seg[(uint64_t)o * (Q + 1) + q], e = seg[(uint64_t)o * (Q + 1) + q + 1];
      const uint32_t at_d = lower_bound_key(keys, s, e, key_d0);
      less_d += at_d - s;
      pos += o == r ? i : lower_bound_key(keys, at_d, e, key);
    }
    merged[pos] = key;
    flags[pos] = less_d < k;
  }
}

// Appends the n rows of a finished batch (the output of a shard's finalize: sorted, batch-local query numbers, local
// subjects) to the shard's send block, as keys with slab-relative query numbers and global subjects.  rows beyond the
// block's capacity are counted but not stored (the caller re-sends with a larger block).
__global__ void block_append_kernel(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ n_ptr, uint32_t q_off,
                                    uint64_t subject_offset, uint64_t *__restrict__ block, uint64_t cap) {
  const uint64_t n = *n_ptr, base = block[0];
  const uint64_t add = ((uint64_t)q_off << KEY_Q_SHIFT) + subject_offset;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (base + i < cap) block[2 + base + i] = keys[i] + add;
}
__global__ void block_count_kernel(const unsigned long long *__restrict__ n_ptr, uint64_t *__restrict__ block) { block[0] += *n_ptr; }

__global__ void merged_to_hits_kernel(const uint64_t *__restrict__ keys, const unsigned long long *__restrict__ n_ptr, uint64_t hits_cap,
                                      uint32_t q_base, smafa_hit *__restrict__ hits, unsigned long long *__restrict__ info) {
  const uint64_t n = *n_ptr;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) info[3] = n;
  if (n > hits_cap) return;  // caller reports SMAFA_E_OOM
  for (; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    smafa_hit h;
    h.query = key_q(k) + q_base;
    h.subject = key_j(k);
    h.distance = key_d(k);
    hits[i] = h;
  }
}

static unsigned grid_for(uint64_t n) { return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, 148 * 8)); }

void launch_block_reset(uint64_t *block, uint64_t status, cudaStream_t s) {
  // header = {rows, status}; a status is a small positive number (-smafa_status), so its low byte is all of it
  cudaMemsetAsync(block, 0, 2 * sizeof(uint64_t), s);
  if (status) cudaMemsetAsync(block + 1, (int)(status & 0xff), 1, s);
}

int launch_block_append(const uint64_t *keys, const unsigned long long *n_ptr, uint64_t n_max, uint32_t q_off,
                        uint64_t subject_offset, uint64_t *block, uint64_t cap, cudaStream_t s) {
  block_append_kernel<<<grid_for(n_max), 256, 0, s>>>(keys, n_ptr, q_off, subject_offset, block, cap);
  block_count_kernel<<<1, 1, 0, s>>>(n_ptr, block);
  return 2;
}

size_t merge_temp_bytes(uint64_t rows) {
  size_t b = 0;
  cub::DeviceSelect::Flagged(nullptr, b, (const uint64_t *)nullptr, (const uint8_t *)nullptr, (uint64_t *)nullptr,
                             (unsigned long long *)nullptr, (int64_t)rows);
  return b + 256;
}

int launch_merge_blocks(MergeWorkspace &ws, const uint64_t *gathered, uint32_t n_ranks, uint64_t cap, uint32_t Q, uint32_t k,
                        uint32_t q_base, smafa_hit *hits_out, uint64_t hits_cap, unsigned long long *info_dev, cudaStream_t s) {
  const uint64_t stride = cap + 2, rows = (uint64_t)n_ranks * cap;
  merge_segments_kernel<<<grid_for((uint64_t)n_ranks * (Q + 1)), 256, 0, s>>>(gathered, n_ranks, stride, cap, Q, ws.seg, info_dev);
  cudaMemsetAsync(ws.flags, 0, rows, s);
  merge_rank_kernel<<<grid_for(rows), 256, 0, s>>>(gathered, n_ranks, stride, cap, Q, ws.seg, k, ws.merged, ws.flags);
  size_t tb = ws.cub_temp_bytes;
  cub::DeviceSelect::Flagged(ws.cub_temp, tb, ws.merged, ws.flags, ws.selected, ws.n_selected, (int64_t)rows, s);
  merged_to_hits_kernel<<<grid_for(std::min<uint64_t>(rows, hits_cap)), 256, 0, s>>>(ws.selected, ws.n_selected, hits_cap, q_base, hits_out,
                                                                                  info_dev);
  return 5;  // segments, ranks, select (2), hits
}

}  // namespace smafa
