// Optimistic first pass for queries that start without a useful bound (no or a loose --max-divergence).
//
// The selection cutoff of such a query (its minimum, src/lib.rs:298, or its k-th smallest distance,
// src/lib.rs:253-256) is only known once enough of the db has been seen; until then a scan has to emit -- and
// verify, and histogram -- every window below a loose running bound (ncu, unbounded top-10 on 100 k x 1 M: 6.6 M
// candidates for 1.1 M rows, tensor pipe 48 % active).  Instead the batch is first scanned under a GUESSED bound g:
// the largest distance at which the whole batch is still expected to emit no more than a budget of candidates,
// read off the distance distribution of a strided sample of (query, window) pairs.  A query that finds its k
// windows within g is finished exactly as if --max-divergence g had been given (its cutoff is <= g).  The others
// -- none at all when every query has relatives in the db -- are gathered into a compact batch and scanned again
// under the caller's bound; their first-pass candidates are dropped and the second pass' candidates are mapped
// back to the original query numbers.  Results are identical either way (tests compare against the oracle with
// the pass forced on and off); only the number of candidates changes.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace smafa {

constexpr int GUESS_THREADS = 128;
constexpr int GUESS_BINS = 66;  // distances 0..65 (the guessed pass is used for L <= 64)

// ghist[d] += number of sampled pairs at distance d.  One thread = one sampled query (every q_stride-th), the
// sampled windows (every d_stride-th, n_d of them) pass through shared memory 128 at a time; per-thread
// histogram columns in shared memory, so no atomics and no bank conflicts until the final reduction.
template <int W>
__global__ void __launch_bounds__(GUESS_THREADS) sample_hist_kernel(const uint64_t *__restrict__ q_ref, uint32_t Q,
                                                                    uint32_t q_stride, const uint64_t *__restrict__ d_ref,
                                                                    uint32_t d_stride, uint32_t n_d, int alphabet,
                                                                    unsigned long long *__restrict__ ghist) {
  __shared__ uint64_t tile[GUESS_THREADS * W];
  __shared__ uint16_t hist[GUESS_BINS * GUESS_THREADS];
  const uint32_t tid = threadIdx.x;
  const uint64_t q = (uint64_t)(blockIdx.x * GUESS_THREADS + tid) * q_stride;
  uint64_t qw[W];
#pragma unroll
  for (int x = 0; x < W; ++x) qw[x] = q < Q ? q_ref[q * W + x] : 0;
  for (int b = 0; b < GUESS_BINS; ++b) hist[b * GUESS_THREADS + tid] = 0;
  for (uint32_t j0 = 0; j0 < n_d; j0 += GUESS_THREADS) {
    const uint32_t n = min((uint32_t)GUESS_THREADS, n_d - j0);
    __syncthreads();
    if (tid < n) {
#pragma unroll
      for (int x = 0; x < W; ++x) tile[tid * W + x] = d_ref[(uint64_t)(j0 + tid) * d_stride * W + x];
    }
    __syncthreads();
    if (q < Q) {
      for (uint32_t w = 0; w < n; ++w) {
        const int d = ref_distance(qw, tile + w * W, W, alphabet);
        hist[min(d, GUESS_BINS - 1) * GUESS_THREADS + tid]++;  // n_d <= 65535: no overflow
      }
    }
  }
  __syncthreads();
  if (tid < GUESS_BINS) {
    unsigned long long s = 0;
    // thread t sums bin t over the 128 columns (rotated start: conflict-free)
    for (int c = 0; c < GUESS_THREADS; ++c) s += hist[tid * GUESS_THREADS + ((c + tid) & (GUESS_THREADS - 1))];
    if (s) atomicAdd(ghist + tid, s);
  }
}

int launch_sample_hist(const uint64_t *q_ref, uint32_t Q, uint32_t q_stride, const uint64_t *d_ref, uint32_t d_stride,
                       uint32_t n_d, uint32_t W, int alphabet, unsigned long long *ghist, cudaStream_t s) {
  const uint32_t n_q = (Q + q_stride - 1) / q_stride;
  const uint32_t blocks = (n_q + GUESS_THREADS - 1) / GUESS_THREADS;
  cudaMemsetAsync(ghist, 0, GUESS_BINS * sizeof(unsigned long long), s);
  switch (W) {
    case 1: sample_hist_kernel<1><<<blocks, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, d_stride, n_d, alphabet, ghist); break;
    case 2: sample_hist_kernel<2><<<blocks, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, d_stride, n_d, alphabet, ghist); break;
    case 3: sample_hist_kernel<3><<<blocks, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, d_stride, n_d, alphabet, ghist); break;
    case 4: sample_hist_kernel<4><<<blocks, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, d_stride, n_d, alphabet, ghist); break;
    case 5: sample_hist_kernel<5><<<blocks, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, d_stride, n_d, alphabet, ghist); break;
    default: sample_hist_kernel<6><<<blocks, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, d_stride, n_d, alphabet, ghist); break;
  }
  return 1;
}
int guess_bins() { return GUESS_BINS; }

// ---- selectivity of the union-row filter (scan_mma.cu, UPR) on this batch -------------------------------------------
// counts[u-1] += number of sampled (query, db operand row of u windows) pairs that pass the filter at need = L - bound,
// u = 1, 2, 3: base matches with ANY window of the row >= need - nN_q (nN_q = the query's N/gap positions), which is
// exactly the sign test of the one-hot tcgen05 operands (the +-1 feature operands of u = 1 are at least as tight).
// On the reference's 5-bit one-hot codes (src/lib.rs:167-184) the base matches of q and w are popcount(q & w) over
// the four base bits of every group, and a union row is the OR of its windows.  The sampled row of degree u is the one
// that holds window j (every d_stride-th window), so the sample sees the db's real neighbour structure.
// counts[3] += samples.  Thread = sampled query; blockIdx.y = slice of the sampled windows (uniform loads).
template <int W>
__global__ void __launch_bounds__(GUESS_THREADS) union_sample_kernel(const uint64_t *__restrict__ q_ref, uint32_t Q, uint32_t q_stride,
                                                                     const uint64_t *__restrict__ d_ref, uint32_t D,
                                                                     uint32_t d_stride, uint32_t n_d, uint32_t per_block, int need,
                                                                     unsigned long long *__restrict__ counts) {
  constexpr uint64_t NBITS = 0x0084210842108421ull;  // bit 0 of every 5-bit group: code 1 = N / gap / IUPAC
  const uint64_t q = (uint64_t)(blockIdx.x * GUESS_THREADS + threadIdx.x) * q_stride;
  uint64_t qw[W];
  int nN = 0;
#pragma unroll
  for (int x = 0; x < W; ++x) {
    const uint64_t v = q < Q ? q_ref[q * W + x] : 0;
    nN += __popcll(v & NBITS);
    qw[x] = v & ~NBITS;
  }
  const int thr = q < Q ? need - nN : 1 << 20;
  uint32_t n1 = 0, n2 = 0, n3 = 0, ns = 0;
  const uint32_t i_end = min(n_d, (blockIdx.y + 1) * per_block);
  for (uint32_t i = blockIdx.y * per_block; i < i_end; ++i) {
    const uint32_t j = i * d_stride, p0 = j & ~1u, t0 = j / 3 * 3;
    int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int x = 0; x < W; ++x) {
      const uint64_t a = d_ref[(uint64_t)p0 * W + x], b = p0 + 1 < D ? d_ref[(uint64_t)(p0 + 1) * W + x] : 0;
      uint64_t t = d_ref[(uint64_t)t0 * W + x];
      if (t0 + 1 < D) t |= d_ref[(uint64_t)(t0 + 1) * W + x];
      if (t0 + 2 < D) t |= d_ref[(uint64_t)(t0 + 2) * W + x];
      c1 += __popcll(qw[x] & ((j & 1u) ? b : a));
      c2 += __popcll(qw[x] & (a | b));
      c3 += __popcll(qw[x] & t);
    }
    n1 += c1 >= thr;
    n2 += c2 >= thr;
    n3 += c3 >= thr;
    ns += q < Q;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, d);
    n2 += __shfl_xor_sync(0xffffffffu, n2, d);
    n3 += __shfl_xor_sync(0xffffffffu, n3, d);
    ns += __shfl_xor_sync(0xffffffffu, ns, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n1) atomicAdd(counts + 0, (unsigned long long)n1);
    if (n2) atomicAdd(counts + 1, (unsigned long long)n2);
    if (n3) atomicAdd(counts + 2, (unsigned long long)n3);
    if (ns) atomicAdd(counts + 3, (unsigned long long)ns);
  }
}

int launch_union_sample(const uint64_t *q_ref, uint32_t Q, uint32_t q_stride, const uint64_t *d_ref, uint32_t D, uint32_t d_stride,
                        uint32_t n_d, uint32_t W, int need, unsigned long long *counts, cudaStream_t s) {
  const uint32_t n_q = (Q + q_stride - 1) / q_stride;
  const uint32_t per_block = 32;
  const dim3 grid((n_q + GUESS_THREADS - 1) / GUESS_THREADS, (n_d + per_block - 1) / per_block);
  cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), s);
  switch (W) {
    case 1: union_sample_kernel<1><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 2: union_sample_kernel<2><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 3: union_sample_kernel<3><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 4: union_sample_kernel<4><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 5: union_sample_kernel<5><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    default: union_sample_kernel<6><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
  }
  return 1;
}

// The same for a grouped db (api.cu group_order), whose rows can be 1, 2, 3, 4, 8 or 16 windows wide: counts[0..5] =
// passing (query, row) pairs of those degrees, counts[6] = samples.  The rows of degree 2, 4, 8, 16 that hold window j
// nest inside the aligned block of 16 windows around j; the row of degree 3 is read separately.
template <int W>
__global__ void __launch_bounds__(GUESS_THREADS) union_sample_wide_kernel(const uint64_t *__restrict__ q_ref, uint32_t Q,
                                                                          uint32_t q_stride, const uint64_t *__restrict__ d_ref,
                                                                          uint32_t D, uint32_t d_stride, uint32_t n_d,
                                                                          uint32_t per_block, int need,
                                                                          unsigned long long *__restrict__ counts) {
  constexpr uint64_t NBITS = 0x0084210842108421ull;
  const uint64_t q = (uint64_t)(blockIdx.x * GUESS_THREADS + threadIdx.x) * q_stride;
  uint64_t qw[W];
  int nN = 0;
#pragma unroll
  for (int x = 0; x < W; ++x) {
    const uint64_t v = q < Q ? q_ref[q * W + x] : 0;
    nN += __popcll(v & NBITS);
    qw[x] = v & ~NBITS;
  }
  const int thr = q < Q ? need - nN : 1 << 20;
  uint32_t n[7] = {0, 0, 0, 0, 0, 0, 0};
  const uint32_t i_end = min(n_d, (blockIdx.y + 1) * per_block);
  for (uint32_t i = blockIdx.y * per_block; i < i_end; ++i) {
    const uint32_t j = i * d_stride, b0 = j & ~15u, t0 = j / 3 * 3;
    int c[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int x = 0; x < W; ++x) {
      uint64_t o1 = 0, o2 = 0, o4 = 0, o8 = 0, o16 = 0;
#pragma unroll 1
      for (uint32_t w = 0; w < 16; ++w) {
        const uint32_t jj = b0 + w;
        const uint64_t v = jj < D ? d_ref[(uint64_t)jj * W + x] : 0;
        o16 |= v;
        if ((jj ^ j) < 8) o8 |= v;
        if ((jj ^ j) < 4) o4 |= v;
        if ((jj ^ j) < 2) o2 |= v;
        if (jj == j) o1 = v;
      }
      uint64_t o3 = d_ref[(uint64_t)t0 * W + x];
      if (t0 + 1 < D) o3 |= d_ref[(uint64_t)(t0 + 1) * W + x];
      if (t0 + 2 < D) o3 |= d_ref[(uint64_t)(t0 + 2) * W + x];
      c[0] += __popcll(qw[x] & o1);
      c[1] += __popcll(qw[x] & o2);
      c[2] += __popcll(qw[x] & o3);
      c[3] += __popcll(qw[x] & o4);
      c[4] += __popcll(qw[x] & o8);
      c[5] += __popcll(qw[x] & o16);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) n[k] += c[k] >= thr;
    n[6] += q < Q;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n[k] += __shfl_xor_sync(0xffffffffu, n[k], d);
    if ((threadIdx.x & 31) == 0 && n[k]) atomicAdd(counts + k, (unsigned long long)n[k]);
  }
}

int launch_union_sample_wide(const uint64_t *q_ref, uint32_t Q, uint32_t q_stride, const uint64_t *d_ref, uint32_t D,
                             uint32_t d_stride, uint32_t n_d, uint32_t W, int need, unsigned long long *counts, cudaStream_t s) {
  const uint32_t n_q = (Q + q_stride - 1) / q_stride;
  const uint32_t per_block = 16;
  const dim3 grid((n_q + GUESS_THREADS - 1) / GUESS_THREADS, (n_d + per_block - 1) / per_block);
  cudaMemsetAsync(counts, 0, 7 * sizeof(unsigned long long), s);
  switch (W) {
    case 1: union_sample_wide_kernel<1><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 2: union_sample_wide_kernel<2><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 3: union_sample_wide_kernel<3><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 4: union_sample_wide_kernel<4><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    case 5: union_sample_wide_kernel<5><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
    default: union_sample_wide_kernel<6><<<grid, GUESS_THREADS, 0, s>>>(q_ref, Q, q_stride, d_ref, D, d_stride, n_d, per_block, need, counts); break;
  }
  return 1;
}

// per_query[q] = number of candidates of query q
__global__ void count_per_query_kernel(const uint64_t *__restrict__ cand, uint64_t n, uint32_t *__restrict__ per_query) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(per_query + key_q(cand[i]), 1u);
}

// Queries with fewer than `need` candidates under the guessed bound are unfinished: their numbers go to
// `list` (ascending: one thread block walks the batch in order, ballot + prefix per warp).
__global__ void __launch_bounds__(1024) list_unfinished_kernel(const uint32_t *__restrict__ per_query, uint32_t Q, uint32_t need,
                                                               uint32_t *__restrict__ list, uint32_t *__restrict__ n_list) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t base;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (uint32_t q0 = 0; q0 < Q; q0 += 1024) {
    const uint32_t q = q0 + tid;
    const bool open = q < Q && per_query[q] < need;
    const uint32_t m = __ballot_sync(0xffffffffu, open);
    if (lane == 0) warp_sum[warp] = __popc(m);
    __syncthreads();
    uint32_t before = base;
    for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
    if (open) list[before + __popc(m & ((1u << lane) - 1))] = q;
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0;
      for (int w = 0; w < 32; ++w) t += warp_sum[w];
      base += t;
    }
    __syncthreads();
  }
  if (tid == 0) *n_list = base;
}

// out[i] = q_ref[list[i]]
__global__ void gather_queries_kernel(const uint64_t *__restrict__ q_ref, const uint32_t *__restrict__ list, uint32_t n,
                                      uint32_t W, uint64_t *__restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n * W) return;
  out[i] = q_ref[(uint64_t)list[i / W] * W + i % W];
}

// Keeps the candidates of finished queries (order is irrelevant: finalize sorts).
__global__ void keep_finished_kernel(const uint64_t *__restrict__ cand, uint64_t n, const uint32_t *__restrict__ per_query,
                                     uint32_t need, uint64_t *__restrict__ out, unsigned long long *__restrict__ n_out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = cand[i];
    if (per_query[key_q(k)] >= need) out[atomicAdd(n_out, 1ull)] = k;
  }
}

// Second-pass candidates [*begin, min(*end, cap)) carry compact query numbers: map them back.
__global__ void remap_queries_kernel(uint64_t *__restrict__ cand, const unsigned long long *__restrict__ begin,
                                     const unsigned long long *__restrict__ end, uint64_t cap,
                                     const uint32_t *__restrict__ list) {
  const uint64_t b = *begin, e = min((uint64_t)*end, cap);
  for (uint64_t i = b + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = cand[i];
    cand[i] = make_key(list[key_q(k)], key_d(k), key_j(k));
  }
}

static unsigned grid_for(uint64_t n) { return (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8); }

void launch_count_per_query(const uint64_t *cand, uint64_t n, uint32_t Q, uint32_t *per_query, cudaStream_t s) {
  cudaMemsetAsync(per_query, 0, (size_t)Q * sizeof(uint32_t), s);
  if (n) count_per_query_kernel<<<grid_for(n), 256, 0, s>>>(cand, n, per_query);
}
void launch_list_unfinished(const uint32_t *per_query, uint32_t Q, uint32_t need, uint32_t *list, uint32_t *n_list, cudaStream_t s) {
  list_unfinished_kernel<<<1, 1024, 0, s>>>(per_query, Q, need, list, n_list);
}
void launch_gather_queries(const uint64_t *q_ref, const uint32_t *list, uint32_t n, uint32_t W, uint64_t *out, cudaStream_t s) {
  const uint64_t t = (uint64_t)n * W;
  if (t) gather_queries_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(q_ref, list, n, W, out);
}
void launch_keep_finished(const uint64_t *cand, uint64_t n, const uint32_t *per_query, uint32_t need, uint64_t *out,
                          unsigned long long *n_out, cudaStream_t s) {
  cudaMemsetAsync(n_out, 0, sizeof(unsigned long long), s);
  if (n) keep_finished_kernel<<<grid_for(n), 256, 0, s>>>(cand, n, per_query, need, out, n_out);
}
void launch_remap_queries(uint64_t *cand, const unsigned long long *begin, const unsigned long long *end, uint64_t cap,
                          const uint32_t *list, cudaStream_t s) {
  remap_queries_kernel<<<148 * 4, 256, 0, s>>>(cand, begin, end, cap, list);
}

}  // namespace smafa
