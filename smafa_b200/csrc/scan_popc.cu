// Formulation (a): CUDA-core scan.  Replaces WindowSet::get_distances (reference
// src/lib.rs:71-89) fused with the first stage of the selection (src/lib.rs:243-265,298-312).
//
// Layout/mapping (B200: 148 SMs, 4 SMSPs each, ALU pipe 16 lanes/clk/SMSP, POPC 4 lanes/clk/SMSP):
//  * one thread owns R queries (bit planes in registers, running bound in a register: the
//    per-query selection state is thread-private, no shared-memory atomics);
//  * the db is streamed through shared memory in 256-window tiles (one aligned 16/32-byte row per
//    thread per tile, register-prefetched one tile ahead) and read back as warp-wide broadcasts;
//  * per pair: 3 LOP3 + 1 POPC per 32 positions, then one compare against the bound;
//  * EARLY: the second 32 positions are only evaluated when the first 32 already fit the bound
//    (exact: mismatches only add up) -- halves the POPC count when --max-divergence is small;
//  * grid = query tiles x db chunks, chunk-major, so co-resident blocks stream the same db chunk
//    (L2/L1 hits) and later chunks start from bounds tightened by earlier ones.
#include "common.cuh"
#include "kernels.h"

namespace smafa {

static constexpr int POPC_THREADS = 256;
static constexpr int POPC_TILE = 256;  // windows per shared-memory tile

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
    r.h[0] = a.x; r.l[0] = a.y; r.n[0] = a.z; r.h[1] = a.w; r.l[1] = b.x; r.n[1] = b.y;
  } else {
    uint4 a = __ldg(reinterpret_cast<const uint4 *>(base) + row);
    r.h[0] = a.x; r.l[0] = a.y; r.n[0] = a.z;
  }
  return r;
}

// Out-of-line slow path; returns the tightened bound (by value, so the bounds stay in registers).
__device__ __noinline__ int popc_hit(const ScanParams *p, uint32_t q, uint32_t j, int d, int bound) {
  emit_candidate(*p, q, j, d, bound);
  return bound;
}

template <int PW, int R, bool EARLY>
__global__ void __launch_bounds__(POPC_THREADS) scan_popc_kernel(const __grid_constant__ ScanParams p, uint32_t n_qtiles, uint32_t chunk) {
  constexpr int ROW4 = PW;  // uint4 per row
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

  // register prefetch of this thread's row of the first tile (the plane matrix is padded to a
  // multiple of POPC_TILE rows, so the load is always in bounds)
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
    // WU windows x R queries per iteration: all POPCs are issued back to back (ILP), one combined
    // "anything within its bound?" branch per WU*R pairs, and the rare slow path re-checks.
    constexpr int WU = 2;
    for (int w = 0; w < nw; w += WU) {
      uint32_t dh0[WU], dl0[WU], dn0[WU], dh1[WU], dl1[WU], dn1[WU];
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        if constexpr (PW == 2) {
          uint4 a = tile[(w + u) * 2];
          uint2 b = *reinterpret_cast<const uint2 *>(&tile[(w + u) * 2 + 1]);
          dh0[u] = a.x; dl0[u] = a.y; dn0[u] = a.z; dh1[u] = a.w; dl1[u] = b.x; dn1[u] = b.y;
        } else {
          uint4 a = tile[w + u];
          dh0[u] = a.x; dl0[u] = a.y; dn0[u] = a.z; dh1[u] = dl1[u] = dn1[u] = 0;
        }
      }
      int c[WU][R];
      bool any = false;
#pragma unroll
      for (int u = 0; u < WU; ++u) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          uint32_t x0 = (q[r].h[0] ^ dh0[u]) | (q[r].l[0] ^ dl0[u]) | (q[r].n[0] ^ dn0[u]);
          c[u][r] = __popc(x0);
          if constexpr (PW == 2 && !EARLY) {
            uint32_t x1 = (q[r].h[1] ^ dh1[u]) | (q[r].l[1] ^ dl1[u]) | (q[r].n[1] ^ dn1[u]);
            c[u][r] += __popc(x1);
          }
          any |= (c[u][r] <= bound[r]);
        }
      }
      if (any) {
#pragma unroll
        for (int u = 0; u < WU; ++u) {
          if (w + u < nw) {  // the tile is padded in shared memory, not in the db
#pragma unroll
            for (int r = 0; r < R; ++r) {
              int d = c[u][r];
              if (d <= bound[r]) {
                if constexpr (PW == 2 && EARLY) {
                  uint32_t x1 = (q[r].h[1] ^ dh1[u]) | (q[r].l[1] ^ dl1[u]) | (q[r].n[1] ^ dn1[u]);
                  d += __popc(x1);
                }
                if (d <= bound[r]) bound[r] = popc_hit(&p, qi[r], t0 + w + u, d, bound[r]);
              }
            }
          }
        }
      }
    }
  }
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
    int d = ref_distance(qw, p.d_ref + (size_t)j * p.W, p.W);
    if (d <= bound) emit_candidate(p, q, j, d, bound);
  }
}

// get_distances for parity/debug: out[q*D + j]
__global__ void distances_kernel(const uint64_t *__restrict__ q_ref, uint32_t Q, const uint64_t *__restrict__ d_ref,
                                 uint32_t D, uint32_t W, uint16_t *__restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t q = blockIdx.y;
  if (j >= D || q >= Q) return;
  out[(size_t)q * D + j] = (uint16_t)ref_distance(q_ref + (size_t)q * W, d_ref + (size_t)j * W, W);
}

void launch_distances(const uint64_t *q_ref, uint32_t Q, const uint64_t *d_ref, uint32_t D, uint32_t W,
                      uint16_t *out, cudaStream_t s) {
  if (Q == 0 || D == 0) return;
  for (uint32_t q0 = 0; q0 < Q; q0 += 32768) {
    uint32_t nq = min(Q - q0, 32768u);
    dim3 grid((D + 255) / 256, nq);
    distances_kernel<<<grid, 256, 0, s>>>(q_ref + (size_t)q0 * W, nq, d_ref, D, W, out + (size_t)q0 * D);
  }
}

template <int PW, int R, bool EARLY>
static void launch_one(const ScanParams &p, uint32_t chunk, cudaStream_t s) {
  uint32_t n_qtiles = (p.Q + POPC_THREADS * R - 1) / (POPC_THREADS * R);
  uint32_t n_chunks = (p.d_end - p.d_begin + chunk - 1) / chunk;
  scan_popc_kernel<PW, R, EARLY><<<n_qtiles * n_chunks, POPC_THREADS, 0, s>>>(p, n_qtiles, chunk);
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
