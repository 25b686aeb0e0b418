// Re-packing of reference-layout windows into the GPU-resident matrices (north_star: "makedb
// encoding is re-packed ... into a GPU-resident, coalesced, 16-byte-aligned db matrix").
//
// Input  : reference words, ceil(L/12) u64 per window, 5-bit one-hot code of symbol p at bit
//          5*(p%12) of word p/12 (reference src/lib.rs:29-52; codes src/lib.rs:167-184).
// Output : (1) bit planes for the POPC kernel.  Three planes per window -- H, Lo, N -- with one
//              bit per position:   A=(0,0,0) C=(0,1,0) G=(1,0,0) T=(1,1,0) N/gap=(0,0,1).
//              Two windows differ at position p  <=>  (H^H')|(Lo^Lo')|(N^N') has bit p set, which
//              is exactly the reference's "codes differ" rule (N==N is a match, N vs base is a
//              mismatch).  Row = [H0 Lo0 H1 Lo1 | N0 N1 0 0] (32 B, L<=64) or
//              [H0 Lo0 N0 0] (16 B, L<=32): one or two aligned 16-byte loads per window.
//          (2) one-hot int8 rows for the tcgen05 kernel (see scan_mma.cu for the tile layout).
// Protein windows (ALPHA_AA, common.cuh) are packed through their 4-class filter image.
// Windows holding anything but the five valid codes (possible only in a hand-made db file) set
// the `invalid` flag; the caller then uses the generic reference-layout kernel, which computes
// popcount(a^b)/2 like the reference for arbitrary words.
#include "common.cuh"
#include "kernels.h"

namespace smafa {

__global__ void pack_planes_kernel(const uint64_t *__restrict__ ref, uint32_t n, uint32_t W, uint32_t L,
                                   uint32_t row_words, int alphabet, uint32_t *__restrict__ planes,
                                   int *__restrict__ invalid) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t *w = ref + (size_t)i * W;
  uint32_t h[2] = {0, 0}, lo[2] = {0, 0}, nn[2] = {0, 0};
  bool bad = false;
  for (uint32_t p = 0; p < L; ++p) {
    uint32_t code = filter_code((uint32_t)(w[p / 12] >> (5 * (p % 12))) & 31u, alphabet);
    uint32_t bit = 1u << (p & 31), s = p >> 5;
    switch (code) {
      case 16: break;
      case 8: lo[s] |= bit; break;
      case 4: h[s] |= bit; break;
      case 2: h[s] |= bit; lo[s] |= bit; break;
      case 1: nn[s] |= bit; break;
      default: bad = true;
    }
  }
  // stray bits above the window (would count in the reference's popcount) also disqualify
  for (uint32_t x = 0; x < W; ++x) {
    uint32_t used = (L >= 12 * (x + 1)) ? 12 : (L > 12 * x ? L - 12 * x : 0);
    uint64_t mask = used == 12 ? ((1ull << 60) - 1) : ((1ull << (5 * used)) - 1);
    if (w[x] & ~mask) bad = true;
  }
  if (bad) atomicOr(invalid, 1);
  uint32_t *row = planes + (size_t)i * row_words;
  if (row_words == 8) {
    reinterpret_cast<uint4 *>(row)[0] = make_uint4(h[0], lo[0], h[1], lo[1]);
    reinterpret_cast<uint4 *>(row)[1] = make_uint4(nn[0], nn[1], 0u, 0u);
  } else {
    reinterpret_cast<uint4 *>(row)[0] = make_uint4(h[0], lo[0], nn[0], 0u);
  }
}

void launch_pack_planes(const uint64_t *ref, uint32_t n, uint32_t W, uint32_t L, uint32_t row_words, int alphabet,
                        uint32_t *planes, int *invalid, cudaStream_t s) {
  if (n == 0) return;
  pack_planes_kernel<<<(n + 255) / 256, 256, 0, s>>>(ref, n, W, L, row_words, alphabet, planes, invalid);
}

// The validity half of pack_planes_kernel alone (nucleotide words): every used 5-bit group holds exactly one bit and
// nothing sits above the window.  For scans that need no bit planes (tcgen05 kernel without a bound pre-pass).
__global__ void check_codes_kernel(const uint64_t *__restrict__ ref, uint32_t n, uint32_t W, uint32_t L, int *__restrict__ invalid) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)n * W) return;
  const uint32_t x = (uint32_t)(idx % W);
  const uint64_t w = ref[idx];
  const uint32_t used = (L >= 12 * (x + 1)) ? 12 : (L > 12 * x ? L - 12 * x : 0);
  const uint64_t mask = used == 12 ? ((1ull << 60) - 1) : ((1ull << (5 * used)) - 1);
  constexpr uint64_t LOW = 0x0084210842108421ull;  // bit 0 of each group
  // per-group bit count, one 5-bit field per group (at most 5: no carry into the neighbour)
  const uint64_t cnt = (w & LOW) + ((w >> 1) & LOW) + ((w >> 2) & LOW) + ((w >> 3) & LOW) + ((w >> 4) & LOW);
  if ((w & ~mask) != 0 || cnt != (LOW & mask)) atomicOr(invalid, 1);
}

void launch_check_codes(const uint64_t *ref, uint32_t n, uint32_t W, uint32_t L, int *invalid, cudaStream_t s) {
  const uint64_t t = (uint64_t)n * W;
  if (t == 0) return;
  check_codes_kernel<<<(unsigned)((t + 255) / 256), 256, 0, s>>>(ref, n, W, L, invalid);
}

__global__ void init_bound_kernel(int *bound, uint32_t Q, int v) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Q) bound[i] = v;
}

void launch_init_bound(int *bound, uint32_t Q, int v, cudaStream_t s) {
  if (Q == 0) return;
  init_bound_kernel<<<(Q + 255) / 256, 256, 0, s>>>(bound, Q, v);
}

}  // namespace smafa
