// Formulation (a): CUDA-core scan.  Replaces WindowSet::get_distances (reference
// src/lib.rs:71-89) fused with the first stage of the selection (src/lib.rs:243-265,298-312).
//
// Layout/mapping (B200: 148 SMs, 4 SMSPs each, ALU pipe 16 lanes/clk/SMSP, POPC 4 lanes/clk/SMSP):
//  * one thread owns R queries (bit planes in registers, running bound in a register: the
//    per-query selection state is thread-private, no shared-memory atomics);
//  * the db is streamed through shared memory in 256-window tiles (one aligned 16/32-byte row per
//    thread per tile, register-prefetched one tile ahead) and read back as warp-wide broadcasts;
//  * fast path per pair: 2 LOP3 + 1 POPC per 32 positions on the H/Lo planes only (a lower bound of
//    the distance; the N plane joins in the exact re-check of the rare survivors);
//  * EARLY: only the first 32 positions are examined by the fast path (mismatches only add up, so
//    this is again a lower bound) -- one POPC per pair when --max-divergence is small;
//  * grid = query tiles x db chunks, chunk-major, so co-resident blocks stream the same db chunk
//    (L2/L1 hits) and later chunks start from bounds tightened by earlier ones.
#define SMAFA_KTH_SCAN_ATTR __noinline__
#define SMAFA_KTH_SCAN_SMALL
#include "common.cuh"
#include "kernels.h"

namespace smafa {

static constexpr int POPC_THREADS = 256;
static constexpr int POPC_TILE = 256;  // windows per shared-memory tile

// Row layout (pack.cu): PW=2 -> [H0 Lo0 H1 Lo1 | N0 N1 0 0], PW=1 -> [H0 Lo0 N0 0].
template <int PW>
struct Planes {
  uint32_t h[PW], l[PW], n[PW];
};

template <int PW>
__device__ __forceinline__ Planes<PW> load_row(const uint32_t *__restrict__ base, size_t row) {
  Planes<PW> r;
  if constexpr (PW == 2) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base) + row * 2;
    uint4 a = __ldg(p), b = __ldg(p + 1);
    r.h[0] = a.x; r.l[0] = a.y; r.h[1] = a.z; r.l[1] = a.w; r.n[0] = b.x; r.n[1] = b.y;
  } else {
    uint4 a = __ldg(reinterpret_cast<const uint4 *>(base) + row);
    r.h[0] = a.x; r.l[0] = a.y; r.n[0] = a.z;
  }
  return r;
}

// Out-of-line slow path; returns the tightened bound (by value, so the bounds stay in registers).
// For protein windows `d` is the distance of the 4-class filter image (a lower bound): the exact distance is
// taken from the symbol words here.
// The alphabet is a template parameter of the kernel and of this function: with a run-time branch the 64-bit word
// loop of the protein case raised the registers the nucleotide hot loop has to save around the call (64 -> 160
// bytes of spills per thread, 4.3e12 -> 2.2e12 comparisons/s at --max-divergence 5).
template <bool AA>
__device__ __noinline__ int popc_hit(const ScanParams *p, uint32_t q, uint32_t j, int d, int bound) {
  if constexpr (AA) {
    d = ref_distance(p->q_ref + (size_t)q * p->W, p->d_ref + (size_t)j * p->W, p->W, ALPHA_AA);
    if (d > bound) return bound;
  }
  emit_candidate(*p, q, j, d, bound);
  return bound;
}

// Fast path = a LOWER bound of the distance: the N plane is left out (N is stored as (0,0) in the
// H/Lo planes), so  popc((H^H')|(Lo^Lo')) <= distance  with equality unless exactly one side has
// an N at a position where the other has A.  A pair whose lower bound already exceeds the query's
// bound is rejected with 2 LOP3 + 1 POPC per 32 positions; survivors get the exact 3-plane distance.
template <int PW, int R, bool EARLY, bool AA>
__global__ void __launch_bounds__(POPC_THREADS, 3) scan_popc_kernel(const __grid_constant__ ScanParams p, uint32_t n_qtiles, uint32_t chunk) {
  constexpr int ROW4 = PW;                      // uint4 per row
  constexpr int FW = (PW == 2 && !EARLY) ? 2 : 1;  // plane words examined by the fast path
  constexpr int WU = 4;                         // windows per inner iteration
  __shared__ uint4 tile[POPC_TILE * ROW4];

  const uint32_t qt = blockIdx.x % n_qtiles, ck = blockIdx.x / n_qtiles;
  const uint32_t tid = threadIdx.x;

  Planes<PW> q[R];
  int bound[R];
  uint32_t qi[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    qi[r] = (qt * R + r) * POPC_THREADS + tid;
    if (qi[r] < p.Q) {
      q[r] = load_row<PW>(p.q_planes, qi[r]);
      bound[r] = __ldcg(p.bound + qi[r]);
    } else {
      q[r] = Planes<PW>{};
      bound[r] = -1;  // never a candidate
    }
  }

  const uint32_t w_begin = p.d_begin + ck * chunk;
  const uint32_t w_end = min(w_begin + chunk, p.d_end);
  const uint4 *drows = reinterpret_cast<const uint4 *>(p.d_planes);

  // register prefetch of this thread's row of the first tile (the plane matrix is padded by two
  // tiles of rows, so the load is always in bounds)
  uint4 pre[ROW4];
#pragma unroll
  for (int v = 0; v < ROW4; ++v) pre[v] = __ldg(drows + (size_t)(w_begin + tid) * ROW4 + v);

  for (uint32_t t0 = w_begin; t0 < w_end; t0 += POPC_TILE) {
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int v = 0; v < ROW4; ++v) tile[tid * ROW4 + v] = pre[v];
    __syncthreads();
    if (t0 + POPC_TILE < w_end) {
#pragma unroll
      for (int v = 0; v < ROW4; ++v) pre[v] = __ldg(drows + (size_t)(t0 + POPC_TILE + tid) * ROW4 + v);
    }
    const int nw = (int)min((uint32_t)POPC_TILE, w_end - t0);
    // WU windows x R queries per iteration: all POPCs are issued back to back (ILP); per query one
    // min over the WU lower bounds and one compare; one branch per WU*R pairs.
    for (int w = 0; w < nw; w += WU) {
      uint32_t dh[WU][FW], dl[WU][FW];
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        if constexpr (FW == 2) {
          uint4 a = tile[(w + u) * 2];
          dh[u][0] = a.x; dl[u][0] = a.y; dh[u][1] = a.z; dl[u][1] = a.w;
        } else {
          uint2 a = *reinterpret_cast<const uint2 *>(&tile[(w + u) * ROW4]);
          dh[u][0] = a.x; dl[u][0] = a.y;
        }
      }
      int c[WU][R];
      bool any = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int mn = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          int v = __popc((q[r].h[0] ^ dh[u][0]) | (q[r].l[0] ^ dl[u][0]));
          if constexpr (FW == 2) v += __popc((q[r].h[1] ^ dh[u][1]) | (q[r].l[1] ^ dl[u][1]));
          c[u][r] = v;
          mn = min(mn, v);
        }
        any |= (mn <= bound[r]);
      }
      if (any) {
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          if (w + u < nw) {  // rows past the chunk end are other windows (or padding): skip
            Planes<PW> d;
            if constexpr (PW == 2) {
              uint4 a = tile[(w + u) * 2], b = tile[(w + u) * 2 + 1];
              d.h[0] = a.x; d.l[0] = a.y; d.h[1] = a.z; d.l[1] = a.w; d.n[0] = b.x; d.n[1] = b.y;
            } else {
              uint4 a = tile[w + u];
              d.h[0] = a.x; d.l[0] = a.y; d.n[0] = a.z;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (c[u][r] <= bound[r]) {
                int dist = 0;
#pragma unroll
                for (int x = 0; x < PW; ++x)
                  dist += __popc((q[r].h[x] ^ d.h[x]) | (q[r].l[x] ^ d.l[x]) | (q[r].n[x] ^ d.n[x]));
                if (dist <= bound[r]) bound[r] = popc_hit<AA>(&p, qi[r], t0 + w + u, dist, bound[r]);
              }
            }
          }
        }
      }
    }
  }
}

// Bound estimator ("pre-pass"): when a query starts without a useful bound (no --max-divergence),
// the first tiles of a scan would emit almost every pair.  This kernel scans a strided sample of the
// db WITHOUT emitting anything and only tightens bound[q]: the minimum distance over the sample
// (MODE_MIN) or the k-th smallest (MODE_KTH, per-query histogram in shared memory, column-per-thread so
// no atomics and no bank conflicts).  Any distance seen on a subset of the db is a valid upper bound
// of the final cutoff, so the main scan (which re-visits the sample) stays exact.
// One block = 256 queries x the whole sample.
template <int PW>
__global__ void __launch_bounds__(POPC_THREADS) bound_prepass_kernel(const __grid_constant__ ScanParams p, uint32_t tile_stride) {
  constexpr int ROW4 = PW;
  constexpr int HB = 66;  // bins 0..65 (L <= 64)
  __shared__ uint4 tile[POPC_TILE * ROW4];
  extern __shared__ uint16_t hist[];  // [HB][256], MODE_KTH only
  const uint32_t tid = threadIdx.x, q = blockIdx.x * POPC_THREADS + tid;
  const bool kth = p.mode == MODE_KTH;
  Planes<PW> qp{};
  int best = -1;
  if (q < p.Q) { qp = load_row<PW>(p.q_planes, q); best = __ldcg(p.bound + q); }
  if (kth)
    for (int b = 0; b < HB; ++b) hist[b * POPC_THREADS + tid] = 0;
  const uint4 *drows = reinterpret_cast<const uint4 *>(p.d_planes);
  const uint32_t n_tiles = (p.d_end - p.d_begin + POPC_TILE - 1) / POPC_TILE;
  for (uint32_t t = 0; t < n_tiles; t += tile_stride) {
    const uint32_t t0 = p.d_begin + t * POPC_TILE;
    __syncthreads();
#pragma unroll
    for (int v = 0; v < ROW4; ++v) tile[tid * ROW4 + v] = __ldg(drows + (size_t)(t0 + tid) * ROW4 + v);
    __syncthreads();
    const int nw = (int)min((uint32_t)POPC_TILE, p.d_end - t0);
    for (int w = 0; w < nw; ++w) {
      Planes<PW> d;
      if constexpr (PW == 2) {
        uint4 a = tile[w * 2], b = tile[w * 2 + 1];
        d.h[0] = a.x; d.l[0] = a.y; d.h[1] = a.z; d.l[1] = a.w; d.n[0] = b.x; d.n[1] = b.y;
      } else {
        uint4 a = tile[w];
        d.h[0] = a.x; d.l[0] = a.y; d.n[0] = a.z;
      }
      int dist = 0;
#pragma unroll
      for (int x = 0; x < PW; ++x) dist += __popc((qp.h[x] ^ d.h[x]) | (qp.l[x] ^ d.l[x]) | (qp.n[x] ^ d.n[x]));
      if (kth) {
        if (dist <= best) hist[dist * POPC_THREADS + tid]++;  // a sample is < 65536 windows: no overflow
      } else {
        best = min(best, dist);
      }
    }
  }
  if (q >= p.Q) return;
  if (kth) {
    uint32_t cum = 0;
    for (int t = 0; t <= best; ++t) {
      cum += hist[t * POPC_THREADS + tid];
      if (cum >= p.k) { best = t; break; }
    }
  }
  atomicMin(p.bound + q, best);
}

// The same estimator on the reference words (exact for every alphabet): used for protein windows, whose
// planes hold the 4-class filter image -- plane distances there are lower bounds and would give bounds that
// are too tight.  One block = 256 queries x the whole sample; tile = 256 windows x W words.
template <int W>
__global__ void __launch_bounds__(POPC_THREADS) bound_prepass_ref_kernel(const __grid_constant__ ScanParams p, uint32_t tile_stride) {
  constexpr int HB = 66;
  __shared__ uint64_t tile[POPC_TILE * W];
  extern __shared__ uint16_t hist[];  // [HB][256], MODE_KTH only
  const uint32_t tid = threadIdx.x, q = blockIdx.x * POPC_THREADS + tid;
  const bool kth = p.mode == MODE_KTH;
  uint64_t qw[W];
  int best = -1;
#pragma unroll
  for (int x = 0; x < W; ++x) qw[x] = 0;
  if (q < p.Q) {
#pragma unroll
    for (int x = 0; x < W; ++x) qw[x] = p.q_ref[(size_t)q * W + x];
    best = __ldcg(p.bound + q);
  }
  if (kth)
    for (int b = 0; b < HB; ++b) hist[b * POPC_THREADS + tid] = 0;
  const uint32_t n_tiles = (p.d_end - p.d_begin + POPC_TILE - 1) / POPC_TILE;
  for (uint32_t t = 0; t < n_tiles; t += tile_stride) {
    const uint32_t t0 = p.d_begin + t * POPC_TILE;
    const int nw = (int)min((uint32_t)POPC_TILE, p.d_end - t0);
    __syncthreads();
    for (int i = tid; i < nw * W; i += POPC_THREADS) tile[i] = p.d_ref[(size_t)t0 * W + i];
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
      const int dist = ref_distance(qw, tile + w * W, W, p.alphabet);
      if (kth) {
        if (dist <= best) hist[dist * POPC_THREADS + tid]++;  // a sample is < 65536 windows: no overflow
      } else {
        best = min(best, dist);
      }
    }
  }
  if (q >= p.Q) return;
  if (kth) {
    uint32_t cum = 0;
    for (int t = 0; t <= best; ++t) {
      cum += hist[t * POPC_THREADS + tid];
      if (cum >= p.k) { best = t; break; }
    }
  }
  atomicMin(p.bound + q, best);
}

template <int W>
static void launch_prepass_ref(const ScanParams &p, uint32_t tile_stride, uint32_t blocks, size_t smem, cudaStream_t s) {
  if (smem) cudaFuncSetAttribute(bound_prepass_ref_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bound_prepass_ref_kernel<W><<<blocks, POPC_THREADS, smem, s>>>(p, tile_stride);
}

int launch_bound_prepass(const ScanParams &p, uint32_t tile_stride, cudaStream_t s) {
  if (p.Q == 0 || p.d_end <= p.d_begin || p.L > 64) return 0;
  const uint32_t blocks = (p.Q + POPC_THREADS - 1) / POPC_THREADS;
  const size_t smem = p.mode == MODE_KTH ? (size_t)66 * POPC_THREADS * sizeof(uint16_t) : 0;
  if (p.alphabet != ALPHA_NUC) {
    switch (p.W) {
      case 1: launch_prepass_ref<1>(p, tile_stride, blocks, smem, s); break;
      case 2: launch_prepass_ref<2>(p, tile_stride, blocks, smem, s); break;
      case 3: launch_prepass_ref<3>(p, tile_stride, blocks, smem, s); break;
      case 4: launch_prepass_ref<4>(p, tile_stride, blocks, smem, s); break;
      case 5: launch_prepass_ref<5>(p, tile_stride, blocks, smem, s); break;
      default: launch_prepass_ref<6>(p, tile_stride, blocks, smem, s); break;
    }
    return 1;
  }
  if (p.L <= 32) {
    if (smem) cudaFuncSetAttribute(bound_prepass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bound_prepass_kernel<1><<<blocks, POPC_THREADS, smem, s>>>(p, tile_stride);
  } else {
    if (smem) cudaFuncSetAttribute(bound_prepass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bound_prepass_kernel<2><<<blocks, POPC_THREADS, smem, s>>>(p, tile_stride);
  }
  return 1;
}

// Generic fallback on the reference word layout: any L <= 4095 and arbitrary (even invalid) words,
// computing popcount(a^b)/2 exactly like src/lib.rs:80-88.  One thread per query.
__global__ void __launch_bounds__(128) scan_generic_kernel(const ScanParams p, uint32_t n_qtiles, uint32_t chunk) {
  const uint32_t qt = blockIdx.x % n_qtiles, ck = blockIdx.x / n_qtiles;
  const uint32_t q = qt * 128 + threadIdx.x;
  if (q >= p.Q) return;
  int bound = __ldcg(p.bound + q);
  const uint64_t *qw = p.q_ref + (size_t)q * p.W;
  const uint32_t w_begin = p.d_begin + ck * chunk;
  const uint32_t w_end = min(w_begin + chunk, p.d_end);
  for (uint32_t j = w_begin; j < w_end; ++j) {
    int d = ref_distance(qw, p.d_ref + (size_t)j * p.W, p.W, p.alphabet);
    if (d <= bound) emit_candidate(p, q, j, d, bound);
  }
}

// get_distances for parity/debug: out[q*D + j]
__global__ void distances_kernel(const uint64_t *__restrict__ q_ref, uint32_t Q, const uint64_t *__restrict__ d_ref,
                                 uint32_t D, uint32_t W, int alphabet, uint16_t *__restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t q = blockIdx.y;
  if (j >= D || q >= Q) return;
  out[(size_t)q * D + j] = (uint16_t)ref_distance(q_ref + (size_t)q * W, d_ref + (size_t)j * W, W, alphabet);
}

void launch_distances(const uint64_t *q_ref, uint32_t Q, const uint64_t *d_ref, uint32_t D, uint32_t W, int alphabet,
                      uint16_t *out, cudaStream_t s) {
  if (Q == 0 || D == 0) return;
  for (uint32_t q0 = 0; q0 < Q; q0 += 32768) {
    uint32_t nq = min(Q - q0, 32768u);
    dim3 grid((D + 255) / 256, nq);
    distances_kernel<<<grid, 256, 0, s>>>(q_ref + (size_t)q0 * W, nq, d_ref, D, W, alphabet, out + (size_t)q0 * D);
  }
}

template <int PW, int R, bool EARLY>
static void launch_one(const ScanParams &p, uint32_t chunk, cudaStream_t s) {
  uint32_t n_qtiles = (p.Q + POPC_THREADS * R - 1) / (POPC_THREADS * R);
  uint32_t n_chunks = (p.d_end - p.d_begin + chunk - 1) / chunk;
  if (p.alphabet == ALPHA_NUC) scan_popc_kernel<PW, R, EARLY, false><<<n_qtiles * n_chunks, POPC_THREADS, 0, s>>>(p, n_qtiles, chunk);
  else scan_popc_kernel<PW, R, EARLY, true><<<n_qtiles * n_chunks, POPC_THREADS, 0, s>>>(p, n_qtiles, chunk);
}

// chunk: windows per block, a multiple of POPC_TILE.
int launch_scan_popc(const ScanParams &p, bool early, uint32_t chunk, cudaStream_t s) {
  if (p.Q == 0 || p.d_end <= p.d_begin) return 0;
  chunk = (chunk + POPC_TILE - 1) / POPC_TILE * POPC_TILE;
  const bool wide = p.Q >= 148u * 256u * 2u;  // enough queries to give every thread R=4
  if (p.L <= 32) {
    if (wide) launch_one<1, 4, false>(p, chunk, s); else launch_one<1, 1, false>(p, chunk, s);
  } else if (early) {
    if (wide) launch_one<2, 4, true>(p, chunk, s); else launch_one<2, 1, true>(p, chunk, s);
  } else {
    if (wide) launch_one<2, 4, false>(p, chunk, s); else launch_one<2, 1, false>(p, chunk, s);
  }
  return 1;
}

int launch_scan_generic(const ScanParams &p, uint32_t chunk, cudaStream_t s) {
  if (p.Q == 0 || p.d_end <= p.d_begin) return 0;
  uint32_t n_qtiles = (p.Q + 127) / 128;
  uint32_t n_chunks = (p.d_end - p.d_begin + chunk - 1) / chunk;
  scan_generic_kernel<<<n_qtiles * n_chunks, 128, 0, s>>>(p, n_qtiles, chunk);
  return 1;
}

int popc_tile_rows() { return POPC_TILE; }

}  // namespace smafa
