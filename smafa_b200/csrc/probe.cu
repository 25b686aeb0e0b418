// Measurement probes for the tcgen05 formulation (not on the product path; reached only through the
// smafa_debug_* entry points).  They answer two questions the scan kernel's roofline depends on:
//   * issue-rate probes: what the tensor pipe sustains for the instruction shapes a scan could use -- dense
//     kind::i8 M128xN256xK32 (the shape scan_mma_kernel issues; `smafa_debug_mma_peak` is the roofline denominator
//     bench.py reports), the same k-step as two N = 128 instructions, and the 2:4-sparse kind::i8 M128xN256xK64
//     (`tcgen05.mma.sp`): a one-hot window operand has one non-zero per 4 K slots, so it is a legal sparse A operand
//     and 4 sparse instructions would cover the K = 256 one-hot contraction that takes 8 dense ones;
//   * sparse_decode_kernel: how the hardware reads the sparsity metadata (TMEM placement, nibble order, the
//     shared-memory image tcgen05.cp expects).  B is the identity, so D[m][k] = the logical A row the tensor core
//     reconstructed from (compressed values, metadata): the layout is read off the result instead of guessed.
#include <string>

#include "internal.h"
#include "tcgen05.cuh"

namespace smafa {

__device__ __forceinline__ void tc_mma_i8_sp(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t tmem_e,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.sp.cta_group::1.kind::i8 [%0], %1, %2, [%6], %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(tmem_e)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_dyn(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_st2(uint32_t taddr, uint32_t v0, uint32_t v1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(v0), "r"(v1) : "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_cp_128x128b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10);  // D = S32, A = B = signed 8 bit, K-major
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t M, uint32_t N, bool sparse) {
  return IDESC_I8 | (sparse ? (1u << 2) : 0u) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- issue-rate probe ------------------------------------------------------------------------------------------
// SHAPE 0: dense M128 x N256 x K32, one accumulator, 4 k-step operand offsets (an ENC 2-like tile)
// SHAPE 1: dense, every k-step as two M128 x N128 x K32 instructions into the two halves of the accumulator
// SHAPE 2: sparse M128 x N256 x K64 (A compressed to 32 bytes per row and k-step, metadata in TMEM columns 256..263)
// n_steps k-steps per CTA; grid = all SMs, 128 threads.
template <int SHAPE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(uint32_t n_steps) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr uint32_t KA = 128;                       // A bytes per row over 4 k-steps (dense K = 128; sparse: compressed K = 256)
  constexpr uint32_t KBB = SHAPE == 2 ? 256 : 128;   // B bytes per row over 4 k-steps
  constexpr uint32_t A_BYTES = 128 * KA, B_BYTES = 256 * KBB;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + A_BYTES + B_BYTES);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t warp = threadIdx.x >> 5;
  if (SHAPE == 2) {  // metadata: every group keeps elements 0 and 1 (nibble 0b0100), 64 bits per row and k-step
    const uint32_t taddr = tmem + ((warp * 32u) << 16) + 256;
    tc_st4(taddr, 0x44444444u, 0x44444444u, 0x44444444u, 0x44444444u);
    tc_st4(taddr + 4, 0x44444444u, 0x44444444u, 0x44444444u, 0x44444444u);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + A_BYTES;
    const uint32_t sbo_a = (8 * KA) >> 4, sbo_b = (8 * KBB) >> 4;
    for (uint32_t i = 0; i < n_steps; ++i) {
      const uint32_t ks = i & 3;
      if (SHAPE == 0) {
        tc_mma_i8_dyn(tmem, smem_desc(a0 + ks * 256, 8, sbo_a), smem_desc(b0 + ks * 256, 8, sbo_b), idesc_i8(128, 256, false), i != 0);
      } else if (SHAPE == 1) {
        const uint64_t ad = smem_desc(a0 + ks * 256, 8, sbo_a);
        tc_mma_i8_dyn(tmem, ad, smem_desc(b0 + ks * 256, 8, sbo_b), idesc_i8(128, 128, false), i != 0);
        tc_mma_i8_dyn(tmem + 128, ad, smem_desc(b0 + 128 * KBB + ks * 256, 8, sbo_b), idesc_i8(128, 128, false), i != 0);
      } else {
        tc_mma_i8_sp(tmem, smem_desc(a0 + ks * 256, 8, sbo_a), smem_desc(b0 + ks * 512, 8, sbo_b), idesc_i8(128, 256, true),
                     tmem + 256 + 2 * ks, i != 0);
      }
    }
    tc_commit(smem_u32(bar));
    mbar_wait(smem_u32(bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- sparse metadata decode ------------------------------------------------------------------------------------
// One CTA.  For step s (one M128 x N64 x K64 sparse instruction, not accumulated: each step has its own 64 accumulator columns)
//   A compressed row m, byte c  = a_comp[m][32 s + c]          (caller: distinct non-zero values)
//   metadata row m              = meta[m][2 s .. 2 s + 1]      (caller: random valid nibbles)
//   B = identity over the step's 64 logical K slots
// so out[s][m][k] is the value the hardware placed at logical slot k of row m.
// meta_path 0: metadata written to TMEM with tcgen05.st (lane = row, 2 columns per step);
// meta_path 1: metadata staged in shared memory as 16-byte rows (two steps per row; 8 rows = one 128-byte core
//              matrix, 8-row groups 128 bytes apart) and moved with one tcgen05.cp.128x128b per two steps.
__global__ void __launch_bounds__(128, 1) sparse_decode_kernel(const uint8_t *a_comp, const uint32_t *meta, uint32_t n_steps,
                                                               int meta_path, int32_t *out) {
  __shared__ __align__(128) uint8_t sA[128 * 64];   // compressed A, canonical K-major image, KB = 64 (2 steps)
  __shared__ __align__(128) uint8_t sB[64 * 64];    // identity, canonical K-major image, KB = 64
  __shared__ __align__(128) uint8_t sE[128 * 16];   // metadata rows for tcgen05.cp
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  for (uint32_t i = threadIdx.x; i < 128 * 64; i += 128) {
    const uint32_t r = i / 64, kb = i % 64;
    sA[tile_offset(r, kb, 64)] = kb < 32 * n_steps ? a_comp[r * 64 + kb] : 0;
  }
  for (uint32_t i = threadIdx.x; i < 64 * 64; i += 128) {
    const uint32_t r = i / 64, kb = i % 64;
    sB[tile_offset(r, kb, 64)] = r == kb ? 1 : 0;
  }
  for (uint32_t i = threadIdx.x; i < 128 * 4; i += 128) reinterpret_cast<uint32_t *>(sE)[i] = meta[i];  // row-major 16-byte rows
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t e_col = 128;  // accumulators: columns [64 s, 64 s + 64); metadata from column 128
  if (meta_path == 0) {
    tc_st4(tmem + ((warp * 32u) << 16) + e_col, meta[row * 4 + 0], meta[row * 4 + 1], meta[row * 4 + 2], meta[row * 4 + 3]);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    if (meta_path == 1) tc_cp_128x128b(tmem + e_col, smem_desc(smem_u32(sE), 8, 8));
    for (uint32_t s = 0; s < n_steps; ++s) {
      // A: 32 compressed bytes of this step = two 16-byte k-chunks (128 bytes apart), 8-row groups 8 * 64 bytes apart
      // B: 64 bytes = four k-chunks; the same image serves both steps (identity over the step's own 64 slots)
      tc_mma_i8_sp(tmem + 64 * s, smem_desc(smem_u32(sA) + s * 256, 8, (8 * 64) >> 4), smem_desc(smem_u32(sB), 8, (8 * 64) >> 4),
                   idesc_i8(128, 64, true), tmem + e_col + 2 * s, 0);
    }
    tc_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (uint32_t s = 0; s < n_steps; ++s) {
    uint32_t v[32];
    for (uint32_t c = 0; c < 2; ++c) {
      tc_ld32(tmem + ((warp * 32u) << 16) + 64 * s + 32 * c, v);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) out[(s * 128 + row) * 64 + 32 * c + i] = (int32_t)v[i];
    }
  }
  (void)lane;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

}  // namespace smafa

using namespace smafa;

template <int SHAPE>
static cudaError_t run_rate(smafa_ctx *ctx, uint32_t n_steps, float *ms) {
  const size_t smem = 128 * 128 + 256 * (SHAPE == 2 ? 256 : 128) + 64;
  cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaStream_t s = ctx->stream;
  mma_rate_kernel<SHAPE><<<ctx->num_sms, 128, smem, s>>>(n_steps / 10 + 8);  // warm-up
  cudaEventRecord(ctx->ev[0], s);
  mma_rate_kernel<SHAPE><<<ctx->num_sms, 128, smem, s>>>(n_steps);
  cudaEventRecord(ctx->ev[1], s);
  e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]);
  return e;
}

// shape: 0 dense N256, 1 dense 2 x N128, 2 sparse N256 K64.  *ns_per_step = time per k-step and SM.
int mma_rate_probe(smafa_ctx *ctx, int shape, uint32_t n_steps, double *ns_per_step) {
  float ms = 0;
  cudaError_t e = shape == 0 ? run_rate<0>(ctx, n_steps, &ms) : shape == 1 ? run_rate<1>(ctx, n_steps, &ms) : run_rate<2>(ctx, n_steps, &ms);
  if (e != cudaSuccess) {
    ctx->err = std::string("mma_rate_kernel: ") + cudaGetErrorString(e);
    return SMAFA_E_CUDA;
  }
  *ns_per_step = (double)ms * 1e6 / n_steps;
  return SMAFA_OK;
}

// a_comp [128][64] bytes (32 per step), meta [128][4] words (2 per step), out [n_steps][128][64] int32; host pointers.
int sparse_decode_probe(smafa_ctx *ctx, const uint8_t *a_comp, const uint32_t *meta, uint32_t n_steps, int meta_path, int32_t *out) {
  uint8_t *d_a = nullptr;
  uint32_t *d_m = nullptr;
  int32_t *d_o = nullptr;
  const size_t out_bytes = (size_t)n_steps * 128 * 64 * sizeof(int32_t);
  cudaError_t e = cudaMalloc((void **)&d_a, 128 * 64);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_m, 128 * 16);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_o, out_bytes);
  cudaStream_t s = ctx->stream;
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_a, a_comp, 128 * 64, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_m, meta, 128 * 16, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_o, 0x7f, out_bytes, s);
  if (e == cudaSuccess) {
    sparse_decode_kernel<<<1, 128, 0, s>>>(d_a, d_m, n_steps, meta_path, d_o);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, out_bytes, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_a);
  cudaFree(d_m);
  cudaFree(d_o);
  if (e != cudaSuccess) {
    ctx->err = std::string("sparse_decode_kernel: ") + cudaGetErrorString(e);
    return SMAFA_E_CUDA;
  }
  return SMAFA_OK;
}
