// Shared device-side definitions of the smafa B200 hot path.
//
// Protocol shared by both scan formulations (POPC and tcgen05 MMA):
//   every query q of the current batch has a running `bound[q]` = the largest distance that can
//   still be part of the answer.  A scan kernel must emit EVERY db window whose distance is
//   <= the final bound (it may emit more: the bound only ever decreases, emission tests against
//   a possibly stale, i.e. larger, value).  finalize.cu then applies the reference's exact
//   cutoff (src/lib.rs:253-265,298-312) to the emitted superset.
//
//   MODE_FIXED : bound never moves (cluster's in-batch pass: everything within t)
//   MODE_MIN   : bound = min(bound, d)                      -> "Mode A", src/lib.rs:296-314
//   MODE_KTH   : bound = smallest t with #(d <= t) >= k     -> "Mode B", src/lib.rs:242-265
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smafa {

enum ScanMode : int { MODE_FIXED = 0, MODE_MIN = 1, MODE_KTH = 2 };

// Alphabets.  NUC is the reference's encoding (5-bit ONE-HOT codes, src/lib.rs:167-184).  AA is this
// build's protein extension (the reference panics on amino-acid bytes, src/lib.rs:35-42):
// same word geometry -- 12 five-bit groups per u64 -- but each group holds a symbol NUMBER 1..23
// (20 amino acids, X, '-', '*'), and the distance is the number of positions whose groups differ.
enum Alphabet : int { ALPHA_NUC = 0, ALPHA_AA = 1 };

// Filter class of a protein symbol, expressed as a nucleotide one-hot code (16, 8, 4, 2; 1 = the N-like
// class of X, '-' and '*'; 0 = not a symbol).  Both scan kernels filter protein windows on this 4-class
// image with their nucleotide machinery -- equal symbols are in equal classes, so class matches >= true
// matches and the filter stays conservative -- and re-evaluate survivors exactly on the symbol words.
// The classes split the usual substitution pairs (I/L/V, D/E, K/R, S/T, N/Q, F/Y) and are roughly balanced
// by background frequency.
__host__ __device__ __forceinline__ uint32_t aa_class_code(uint32_t sym) {
  //                         -   A  C  D  E  F  G  H  I  K   L   M  N   P  Q  R  S   T  V  W  Y  X  -  *
  constexpr uint8_t T[32] = {0, 8, 16, 2, 8, 8, 4, 8, 4, 16, 16, 4, 16, 2, 4, 4, 16, 8, 2, 4, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0};
  return T[sym & 31u];
}
__host__ __device__ __forceinline__ uint32_t filter_code(uint32_t code, int alphabet) {
  return alphabet == ALPHA_NUC ? code : aa_class_code(code);
}

// Candidate rows are stored as one sortable 64-bit key:
//   bits 63..44 query (batch-local, < 2^20) | bits 43..32 distance (< 2^12) | bits 31..0 subject
constexpr int KEY_Q_SHIFT = 44;
constexpr int KEY_D_SHIFT = 32;
constexpr uint32_t KEY_D_MASK = 0xFFFu;
constexpr uint32_t MAX_BATCH_QUERIES = 1u << 20;
constexpr uint32_t MAX_WINDOW_LEN = 4095;

__host__ __device__ inline uint64_t make_key(uint32_t q, uint32_t d, uint32_t j) {
  return ((uint64_t)q << KEY_Q_SHIFT) | ((uint64_t)d << KEY_D_SHIFT) | (uint64_t)j;
}
__host__ __device__ inline uint32_t key_q(uint64_t k) { return (uint32_t)(k >> KEY_Q_SHIFT); }
__host__ __device__ inline uint32_t key_d(uint64_t k) { return (uint32_t)(k >> KEY_D_SHIFT) & KEY_D_MASK; }
__host__ __device__ inline uint32_t key_j(uint64_t k) { return (uint32_t)k; }

struct ScanParams {
  // queries of this batch
  const uint32_t *q_planes;  // [Qpad][row_words] bit planes (see pack.cu)
  const uint64_t *q_ref;     // [Q][W] reference-layout words
  uint32_t Q;
  // db (or db shard)
  const uint32_t *d_planes;  // [Dpad][row_words]
  const uint64_t *d_ref;     // [D][W]
  uint32_t D;
  uint32_t d_begin, d_end;   // window range covered by this launch
  uint32_t W, L;
  int alphabet;              // Alphabet of q_ref / d_ref (planes and MMA operands hold the filter classes)
  // running state
  int mode;
  uint32_t k;                // MODE_KTH
  int *bound;                // [Q]; signed so that "no candidates" can be expressed as -1
  uint32_t *hist;            // [Q][hist_stride] (MODE_KTH) else nullptr
  uint32_t hist_stride;
  uint64_t *cand;            // candidate keys
  unsigned long long *cand_count;
  uint64_t cand_cap;
  // optional (tcgen05 scan only, else nullptr): candidates emitted per query, for the sort-free selection of
  // finalize.cu (launch_finalize_buckets)
  uint32_t *per_query;
};

// Tightens bound[q] after a candidate at distance d (<= bound) has been recorded.  `bound` is the
// caller's private copy; the global copy is updated with atomicMin so that blocks working on other
// db ranges of the same query start from the tightened value.
__device__ __forceinline__ void tighten_bound(const ScanParams &p, uint32_t q, int d, int &bound);

// Slow path shared by all scan kernels: record a candidate and tighten the query's bound.
__device__ __forceinline__ void emit_candidate(const ScanParams &p, uint32_t q, uint32_t j, int d, int &bound) {
  unsigned long long slot = atomicAdd(p.cand_count, 1ull);
  if (slot < p.cand_cap) p.cand[slot] = make_key(q, (uint32_t)d, j);
  tighten_bound(p, q, d, bound);
}

// Smallest t < bound with #(d <= t) >= k in the histogram row h (bound itself when there is none).
// SMAFA_KTH_SCAN_ATTR: the POPC kernels keep this out of line (scan_popc.cu) -- inlined into their slow path it
// doubled the registers the hot loop has to save around the call.
#ifndef SMAFA_KTH_SCAN_ATTR
#define SMAFA_KTH_SCAN_ATTR __forceinline__
#endif
__device__ SMAFA_KTH_SCAN_ATTR int kth_scan(const uint32_t *h, int bound, uint32_t k) {
  uint32_t cum = 0;
  int nb = bound;
  const int nvec = (bound + 3) >> 2;  // covers bins [0, bound)
  const uint4 *hv = reinterpret_cast<const uint4 *>(h);
#ifdef SMAFA_KTH_SCAN_SMALL
  // Register-lean form for kernels whose hot loop is register bound (scan_popc.cu): one 16-byte load at a time.
  if (nvec <= 16) {
    for (int v = 0; v < nvec; ++v) {
      const uint4 c = __ldcg(hv + v);
      const uint32_t s0 = cum + c.x, s1 = s0 + c.y, s2 = s1 + c.z, s3 = s2 + c.w;
      if (s3 >= k) {
        const int t = v * 4 + (s0 >= k ? 0 : (s1 >= k ? 1 : (s2 >= k ? 2 : 3)));
        if (t < bound) nb = t;
        break;
      }
      cum = s3;
    }
  } else
#endif
  if (nvec <= 16) {
    // 16 bins (four independent 16-byte loads) per round trip, stopping at the crossing bin: tight bounds --
    // the common case -- need one trip, and only 16 registers are live
    bool open = true;  // the crossing bin has not been reached yet
#pragma unroll 1
    for (int v0 = 0; v0 < nvec && open; v0 += 4) {
      uint4 c[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) c[v] = v0 + v < nvec ? __ldcg(hv + v0 + v) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (open && v0 + v < nvec) {
          const uint32_t s0 = cum + c[v].x, s1 = s0 + c[v].y, s2 = s1 + c[v].z, s3 = s2 + c[v].w;
          if (s3 >= k) {
            // bins >= bound of the last vector may hold stale emissions: a crossing there tightens nothing
            const int t = (v0 + v) * 4 + (s0 >= k ? 0 : (s1 >= k ? 1 : (s2 >= k ? 2 : 3)));
            if (t < bound) nb = t;
            open = false;
          }
          cum = s3;
        }
      }
    }
  } else {  // long windows (generic kernel only)
    for (int t = 0; t < bound; ++t) {
      cum += __ldcg(h + t);
      if (cum >= k) { nb = t; break; }
    }
  }
  return nb;
}

__device__ __forceinline__ void tighten_bound(const ScanParams &p, uint32_t q, int d, int &bound) {
  if (p.mode == MODE_MIN) {
    if (d < bound) {
      bound = d;
      atomicMin(p.bound + q, d);
    }
  } else if (p.mode == MODE_KTH) {
    // hist rows are 16-byte aligned and hist_stride is a multiple of 4 (run_batch)
    uint32_t *h = p.hist + (size_t)q * p.hist_stride;
    atomicAdd(h + d, 1u);
    // Only a candidate strictly below the bound can lower it (the bound drops to t < bound once
    // #(d <= t) >= k); ties AT the bound -- the common case -- skip the scan.
    if (d < bound) {
      const int nb = kth_scan(h, bound, p.k);
      if (nb < bound) {
        bound = nb;
        atomicMin(p.bound + q, nb);
      }
    }
  }
}

// Exact distance on the reference word layout.  NUC: popcount(a^b)/2 (src/lib.rs:80-88).  AA: number of
// 5-bit groups in which the two words differ.
__device__ __forceinline__ int ref_distance(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint32_t W,
                                            int alphabet) {
  int s = 0;
  if (alphabet == ALPHA_NUC) {
    for (uint32_t w = 0; w < W; ++w) s += __popcll(a[w] ^ b[w]);
    return s >> 1;
  }
  for (uint32_t w = 0; w < W; ++w) {
    const uint64_t x = a[w] ^ b[w];
    s += __popcll((x | (x >> 1) | (x >> 2) | (x >> 3) | (x >> 4)) & 0x0084210842108421ull);  // bit 0 of each group
  }
  return s;
}

}  // namespace smafa
