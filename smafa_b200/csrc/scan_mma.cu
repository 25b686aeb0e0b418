// Formulation (b): tcgen05 int8 MMA scan (north_star: "matches = one-hot(query) . one-hot(db)^T over
// 5 symbols x L positions, with the TMEM accumulator drained into the threshold/top-k epilogue").
// Replaces WindowSet::get_distances + the first selection stage (reference src/lib.rs:71-89,
// 243-265, 298-312).
//
// The literal formulation (ENC = 5), which the other operand encodings below refine:
//   Operands (int8, K-major, UMMA canonical no-swizzle layout = 8-row x 16-byte core matrices):
//   K index = symbol * PB + position, symbol in {A,C,G,T,N}, PB = 64 (L <= 63) or 32 (L <= 31)
//   => K = 320 (10 MMA k-steps of 32) or 160 (5 k-steps).
//   A = db tile, M = 128 operand rows (TMEM lanes), streamed: one contiguous image per tile in global memory,
//       fetched with a single cp.async.bulk (TMA bulk copy, no tensor map needed because pack_operand_kernel
//       already writes the shared-memory image).
//   B = query tile, N = 256 queries (TMEM columns), resident in shared memory for a whole work item.
//   D = A.B^T in TMEM, int32, two 256-column buffers (MMA of tile t+1 overlaps the drain of tile t).
// Threshold folded into the MMA: the spare K slot (symbol A, position PB-1) holds 1 in every db row
// and  -need_q  in query q, need_q = L - bound_q, so  D = matches - need_q  and
//     D >= 0  <=>  distance <= bound_q.
// The epilogue therefore only ANDs sign bits (one LOP3 per two accumulators); a surviving pair is
// re-evaluated exactly on the reference words (popcount(a^b)/2) before it is emitted, so the MMA
// is a conservative filter and bit-exactness never depends on it.  The bias bytes are refreshed from
// the global bounds every tile by the otherwise idle producer warp (stale = looser = still a
// superset).
// What runs by default: +-1 character-feature operands (ENC = 3, K = 192, one window per
// row) when the bound is loose; one-hot union rows (ENC = 4, K = 256, UPR = 2 or 3 windows per row, i.e. per
// accumulator) when it is tight -- the caller picks per scan (api.cu pick_union_degree).
//
// Warp roles (448 threads, 1 CTA/SM, persistent over work items = query tile x db chunk):
//   warp 0 : bulk-copy producer (B once per item, A tiles through a 2..4-stage mbarrier ring) + bias refresh
//   warp 1 : TMEM allocation, single-thread tcgen05.mma issue, tcgen05.commit -> mbarriers
//   warps 2-9 : epilogue, two warps per TMEM lane quarter (128 columns each, tcgen05.ld 32x32b.x32.pack::16b); survivors
//               of the sign filter go to a per-warp shared-memory ring as (query, window) pairs
//   warps 10-13 : verifiers, drain the rings (exact distance, candidate emission, bound tightening)
#include <algorithm>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "common.cuh"
#include "internal.h"
#include "tcgen05.cuh"

namespace smafa {

constexpr int MMA_M = 128;      // db windows per tile
constexpr int MMA_N = 256;      // queries per tile
// Epilogue warps: EPI_WARPS/4 per TMEM lane quarter (a warp may only read the 32 lanes of quarter warp%4),
// each draining MMA_N / (EPI_WARPS/4) accumulator columns of every tile.
// Verifier warps: consume the survivor rings the epilogue warps fill (exact re-check + emission), so that the
// latency of verification never sits between two accumulator drains.
constexpr int MMA_VER_WARPS = 4;
constexpr int mma_threads(int epi_warps) { return 64 + 32 * epi_warps + 32 * MMA_VER_WARPS; }

struct MmaParams {
  ScanParams sp;
  const uint8_t *a_tiles;  // [n_db_tiles][128 * KB]
  const uint8_t *b_tiles;  // [n_qtiles][256 * KB]
  uint32_t n_qtiles, n_chunks, tiles_per_chunk, n_db_tiles;
  uint32_t qt_major;            // work-item order, see scan_mma_kernel ("work items")
  uint32_t desc_lbo, desc_sbo;  // smem descriptor strides in 16-byte units
  int need0;                    // initial L - bound (the bias stored in b_tiles)
  int32_t *dump;                // debug: raw accumulators of work item 0, tile 0 ([128][256]) or nullptr
  const int16_t *q_meta;        // per query: N/gap count (ENC 4) or bias base (ENC 2/3, see HadCoef), else nullptr
};

// ---- operand encodings ------------------------------------------------------------------------
// ENC = 5 : one-hot over A,C,G,T,N           K = 5*PB   D = matches - need                    (exact threshold)
// ENC = 4 : one-hot over A,C,G,T             K = 4*PB   D = base matches - (need - nN_q)      (conservative)
// ENC = 3 : +-1 character features (h, l, h*l) of the 2-bit base code, N/gap = (0,0,0)
//           K = 3*PB.  Per position with two bases  h*h' + l*l' + hl*hl' = 4*[match] - 1,  so with
//           S = feature dot product, n_bb = #positions where both are bases = L - nN_q - nN_d + nNN:
//               4*matches = S + (L - nN_q - nN_d) + 5*nNN            (nNN = positions where both are N)
// ENC = 2 : features (h, l) only, K = 2*PB.  h*h' + l*l' is 2 / 0 / 0 / -2, so [match] <= (u + 2)/4 and
//               4*matches <= S + 2*(L - nN_q - nN_d) + 6*nNN         (lossy but conservative)
// For ENC 2/3 the spare K slots (positions >= L of each feature block) carry the rest of the inequality
//     D = S + c_q - alpha*nN_d + (alpha+4)*min'(nN_q, nN_d) >= 0,   c_q = alpha*(L - nN_q) - 4*need_q
//   spare 0,1 : db 1,            query c_q split into two int8 (refreshed as bound_q tightens)
//   spare 2   : db -alpha*nN_d,  query 1
//   spare 3+t : thermometer code of the N counts, t < T: db [nN_d >= t+1], query (alpha+4)*[nN_q >= t+1];
//               the last level carries (alpha+4)*max(0, nN_q-(T-1)) so the sum is >= min(nN_q, nN_d) >= nNN.
// ENC = 20: protein windows of L <= 20 (ALPHA_AA): one-hot over the 20 amino acids, K index = symbol*20 + position
//           (400) + 16 spare slots = 416 (13 k-steps).  X / '-' / '*' rows are all-zero like N in ENC 4:
//           D = amino-acid matches - (need - nX_q).  Protein windows longer than 20 (and the POPC kernel) filter
//           on the 4-class image of the symbols instead (common.cuh aa_class_code) with ENC 2..5.
// Every variant is a conservative filter; survivors are re-evaluated exactly before they are emitted.
constexpr uint32_t MMA_ENC_AA = 20, MMA_AA_POS = 20, MMA_AA_KB = 416;
__host__ __device__ __forceinline__ uint32_t mma_pb(uint32_t enc, uint32_t L) {
  return enc == MMA_ENC_AA ? MMA_AA_POS : (L <= (enc <= 3 ? 30u : 31u) ? 32u : 64u);
}
__host__ __device__ __forceinline__ bool mma_enc_ok(uint32_t enc, uint32_t L) {
  return L >= 1 && L <= (enc == MMA_ENC_AA ? MMA_AA_POS : (enc <= 3 ? 62u : 63u));
}
// k index of spare slot si (feature-block major)
__host__ __device__ __forceinline__ uint32_t had_spare_k(uint32_t si, uint32_t PB, uint32_t L) {
  const uint32_t gap = PB - L;
  return (si / gap) * PB + L + si % gap;
}
__host__ __device__ __forceinline__ int clamp8(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }

// Survivors of the sign filter are not verified by the warp that finds them: verification is a chain of
// dependent L2 round trips (reference words, bound, candidate slot, histogram), and with only two accumulator
// buffers one epilogue warp stuck in it stalls the MMA pipeline of the whole CTA within a tile (ncu on the
// unbounded top-k scan: epilogue warps 45 % of their samples on TFULL, tensor pipe 48 % active).  Each epilogue
// warp instead appends (query, window) pairs to a private shared-memory ring; MMA_VER_WARPS verifier warps
// drain the rings -- one survivor per lane, all exact distances in flight together, one global atomicAdd per
// warp-full of accepted candidates.  Everything here is inlined: a call inside the epilogue loop makes ptxas
// keep loop state in local memory (seen in the v5 SASS).
constexpr int MMA_LIST_CAP = 256;  // ring entries per epilogue warp (power of two)

__device__ __forceinline__ uint32_t ld_shared_volatile(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_shared_volatile(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// Records the candidates of a warp (lane-private (q, j, d); `ok` = this lane holds one that passed the exact check):
// one global atomicAdd per call for all of them, then the per-query bound tightening.  Warp-converged.
__device__ __forceinline__ void mma_emit_warp(const ScanParams &sp, bool ok, uint32_t q, uint32_t j, int d, int &bnd, uint32_t lane) {
  const uint32_t mask = __ballot_sync(0xffffffffu, ok);
  if (mask) {
    const int leader = __ffs(mask) - 1;
    unsigned long long slot0 = 0;
    if ((int)lane == leader) slot0 = atomicAdd(sp.cand_count, (unsigned long long)__popc(mask));
    slot0 = __shfl_sync(0xffffffffu, slot0, leader);
    if (ok) {
      const unsigned long long slot = slot0 + __popc(mask & ((1u << lane) - 1));
      if (slot < sp.cand_cap) sp.cand[slot] = make_key(q, (uint32_t)d, j);
      if (sp.per_query != nullptr) atomicAdd(sp.per_query + q, 1u);  // result unused: a fire-and-forget RED, no round trip
      tighten_bound(sp, q, d, bnd);
    }
  }
}

// Exact re-check of one operand ROW per lane (`have` = the lane holds a survivor (q, row)): the row's UPR windows
// row * UPR .. row * UPR + UPR - 1 are consecutive in the reference words, the query words are loaded once, and the
// windows are evaluated four at a time with all their loads in flight together -- verification is a chain of L2
// round trips whose latency, not whose work, is the cost (a wide row of a grouped db sends 16 windows here at once).
// Warp-converged: every lane runs the same number of steps.
template <int UPR>
__device__ __forceinline__ void mma_verify_rows_warp(const ScanParams &sp, bool have, uint32_t q, uint32_t row, uint32_t lane) {
  constexpr int WMAX = 6;  // mma_supported: L <= 63
  have = have && q < sp.Q;
  uint64_t qw[WMAX];
#pragma unroll
  for (int w = 0; w < WMAX; ++w) qw[w] = (have && (uint32_t)w < sp.W) ? __ldg(sp.q_ref + (size_t)q * sp.W + w) : 0;
  int bnd = have ? __ldcg(sp.bound + q) : -1;
  constexpr int STEP = UPR < 4 ? UPR : 4;
#pragma unroll 1
  for (int c = 0; c < UPR; c += STEP) {
    int d[STEP];
    bool valid[STEP];
#pragma unroll
    for (int k = 0; k < STEP; ++k) {
      const uint32_t j = row * UPR + c + k;
      valid[k] = have && j < sp.d_end;
      const uint64_t *dw = sp.d_ref + (size_t)(valid[k] ? j : 0) * sp.W;
      int acc = 0;
      if (sp.alphabet == ALPHA_NUC) {
#pragma unroll
        for (int w = 0; w < WMAX; ++w)
          if ((uint32_t)w < sp.W) acc += __popcll(qw[w] ^ __ldg(dw + w));
        acc >>= 1;
      } else {
#pragma unroll
        for (int w = 0; w < WMAX; ++w)
          if ((uint32_t)w < sp.W) {
            const uint64_t x = qw[w] ^ __ldg(dw + w);
            acc += __popcll((x | (x >> 1) | (x >> 2) | (x >> 3) | (x >> 4)) & 0x0084210842108421ull);
          }
      }
      d[k] = acc;
    }
    if (c > 0 && have) bnd = min(bnd, __ldcg(sp.bound + q));  // other warps may have tightened it meanwhile
#pragma unroll
    for (int k = 0; k < STEP; ++k) mma_emit_warp(sp, valid[k] && d[k] <= bnd, q, row * UPR + c + k, d[k], bnd, lane);
  }
}

__device__ __forceinline__ uint32_t and3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;  // opaque to the optimiser: keeps the reduction a tree instead of one dependent LOP3 chain
  asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// AND of 32 registers, depth 4.  The intermediate levels are kept: the slow path descends through them to the
// registers that hold a survivor instead of testing all 64 sign bits.
struct AndTree {
  uint32_t t[11];  // t[g] = AND of registers 3g..3g+2 (t[10]: registers 30, 31)
  uint32_t u[4];   // u[h] = AND of t[3h..3h+2] (u[3]: t[9], t[10])
  uint32_t all;
};
__device__ __forceinline__ AndTree and_tree32(const uint32_t (&v)[32]) {
  AndTree a;
#pragma unroll
  for (int i = 0; i < 10; ++i) a.t[i] = and3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  a.t[10] = v[30] & v[31];
#pragma unroll
  for (int h = 0; h < 3; ++h) a.u[h] = and3(a.t[3 * h], a.t[3 * h + 1], a.t[3 * h + 2]);
  a.u[3] = a.t[9] & a.t[10];
  a.all = and3(a.u[0], a.u[1], a.u[2]) & a.u[3];
  return a;
}

// NSYM = 5: operands hold all five symbols, D = matches - need exactly.
// NSYM = 4: only A,C,G,T are encoded (K = 256 instead of 320: 20 % less tensor work and operand traffic).
//           N/gap rows are all-zero, so D counts base-base matches only; because N-N matches are at most
//           the query's N count nN_q, the bias uses need_q - nN_q and the filter stays conservative:
//           matches >= need  =>  base matches >= need - nN_q  =>  D >= 0.  Survivors are verified exactly.
// B_BUFS = 2: the query operand of the next work item is fetched while the current item computes.
// SPLIT_N: every k-step is issued as two M128xN128 instructions, query half h into accumulator columns
//          [128h, 128h+128) of the buffer, and each half has its own full/empty barrier pair: the epilogue warps of
//          half 0 drain while the MMAs of half 1 run, and an accumulator half is refilled as soon as ITS four warps
//          have read it (four half-buffers in flight instead of two whole ones).  Same tensor cycles per tile
//          (128*N/256 per instruction), finer hand-over.
// UPR (windows per db operand row) = 2 or 3: "union rows".  Row r of the db operand is the OR of the one-hot images
//          of windows UPR*r .. UPR*r + UPR-1, so D counts the positions where the query base equals ANY of their
//          bases -- an upper bound of every one of their match counts.  One accumulator then filters UPR comparisons:
//          1/UPR of the tensor work and of the accumulators to drain per comparison.  A survivor row sends all of its
//          windows to the exact re-check.  Between unrelated windows a position passes the union test with probability
//          1 - (3/4)^UPR (7/16, 37/64) instead of 1/4, so the filter is selective only while need = L - bound is
//          large; the caller picks UPR per scan from a sample of the batch (api.cu pick_union_degree).  One-hot
//          operands only (NSYM = 4).
template <int KSTEPS, int NSYM, int STAGES, int EPI_WARPS, bool PACK16, int B_BUFS, bool SPLIT_N, int UPR>
__global__ void __launch_bounds__(mma_threads(EPI_WARPS), 1) scan_mma_kernel(const __grid_constant__ MmaParams P) {
  constexpr int MMA_EPI_WARPS = EPI_WARPS;
  static_assert(UPR == 1 || ((UPR == 2 || UPR == 3 || UPR == 4 || UPR == 8 || UPR == 16) && NSYM == 4), "union rows need one-hot operands");
  constexpr uint32_t KB = KSTEPS * 32;       // operand bytes per row
  constexpr uint32_t PB = KB / NSYM;         // positions per symbol / feature block
  constexpr uint32_t BIAS_K = NSYM == (int)MMA_ENC_AA ? MMA_ENC_AA * MMA_AA_POS : PB - 1;  // one-hot: symbol A, position PB-1
  constexpr bool HAD = NSYM <= 3;            // +-1 feature encodings: bias in spare slots 0 and 1
  constexpr uint32_t A_BYTES = MMA_M * KB, B_BYTES = MMA_N * KB;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sB0 = smem;
  uint8_t *sA = smem + B_BUFS * B_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + B_BUFS * B_BYTES + STAGES * A_BYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 24);
  uint32_t *ring_ctl = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(bars) + 256);  // tail[16] head[16] done[16]
  uint2 *lists = reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(bars) + 512);  // [MMA_EPI_WARPS][MMA_LIST_CAP]
  uint32_t *ring_tail = ring_ctl, *ring_head = ring_ctl + 16, *ring_done = ring_ctl + 32;
  static_assert(EPI_WARPS <= 16 && EPI_WARPS % MMA_VER_WARPS == 0, "ring control block");

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  static_assert(2 * STAGES + 14 <= 24, "barrier block holds 24 mbarriers");
  static_assert(!SPLIT_N || EPI_WARPS == 8, "SPLIT_N: one epilogue warp per (lane quarter, accumulator half)");
  constexpr uint32_t ACC_UNITS = SPLIT_N ? 4 : 2;  // accumulator hand-over units: (buffer, half) or whole buffers
  auto FULL = [&](uint32_t s) { return bar0 + 8 * s; };
  auto EMPTY = [&](uint32_t s) { return bar0 + 8 * (STAGES + s); };
  auto TFULL = [&](uint32_t b) { return bar0 + 8 * (2 * STAGES + b); };        // b = buffer, or 2*buffer + half
  auto TEMPTY = [&](uint32_t b) { return bar0 + 8 * (2 * STAGES + 4 + b); };
  auto B_FULL = [&](uint32_t b) { return bar0 + 8 * (2 * STAGES + 8 + b); };
  auto B_EMPTY = [&](uint32_t b) { return bar0 + 8 * (2 * STAGES + 10 + b); };
  auto B_READY = [&](uint32_t b) { return bar0 + 8 * (2 * STAGES + 12 + b); };

  if (threadIdx.x < 48) ring_ctl[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < (uint32_t)STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (uint32_t b = 0; b < ACC_UNITS; ++b) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), MMA_EPI_WARPS * 2 / ACC_UNITS); }
    for (uint32_t b = 0; b < 2; ++b) { mbar_init(B_FULL(b), 1); mbar_init(B_EMPTY(b), 1); mbar_init(B_READY(b), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const ScanParams &sp = P.sp;
  const uint32_t n_items = P.n_qtiles * P.n_chunks;
  // Work items = (query tile, db chunk).  Chunk-major (qt_major = 0): CTA c takes items c, c + grid, ... of the order
  // "all query tiles of chunk 0, then chunk 1, ..." -- the CTAs stream the same db chunk together, so a db image larger
  // than L2 is read from DRAM once.  Query-tile-major (qt_major = 1, db images that fit L2: wide union rows): CTA c takes
  // a contiguous range of the order "all chunks of query tile 0, then tile 1, ..." -- consecutive items share their query
  // operand, which is then fetched once per tile instead of once per item, and with small chunks the CTAs finish within
  // one chunk of each other.  The three roles walk the same item sequence.
  const uint32_t item_begin = P.qt_major ? (uint32_t)((uint64_t)n_items * blockIdx.x / gridDim.x) : blockIdx.x;
  const uint32_t my_items = P.qt_major ? (uint32_t)((uint64_t)n_items * (blockIdx.x + 1) / gridDim.x) - item_begin
                                       : (n_items > blockIdx.x ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u);
  auto item_of = [&](uint32_t it) { return P.qt_major ? item_begin + it : blockIdx.x + it * gridDim.x; };
  auto qt_of = [&](uint32_t item) { return P.qt_major ? item / P.n_chunks : item % P.n_qtiles; };
  auto chunk_of = [&](uint32_t item) { return P.qt_major ? item % P.n_chunks : item / P.n_qtiles; };
  // does item `it` of this CTA need a query operand of its own (always, unless it continues the previous item's tile)
  auto new_b_at = [&](uint32_t it) { return !P.qt_major || it == 0 || qt_of(item_of(it)) != qt_of(item_of(it - 1)); };

  if (warp == 0) {
    // ===== producer: bulk copies + bias refresh =====
    uint32_t stage = 0, phase = 0, b_ord = 0xffffffffu;  // b_ord = ordinal of the query operand in use (this CTA's n-th)
    int cur[8];
    int meta[8];  // per-query constant of this lane's queries (see MmaParams::q_meta)
    // Fetches the query operand of tile qt as this CTA's n-th into buffer n % B_BUFS once the MMAs that used the
    // buffer before are done with it.
    auto fetch_B = [&](uint32_t qt, uint32_t n) {
      if (elect_one()) {
        const uint32_t b = n % B_BUFS, use = n / B_BUFS;
        if (use > 0) mbar_wait(B_EMPTY(b), (use - 1) & 1);
        mbar_expect_tx(B_FULL(b), B_BYTES);
        bulk_g2s(smem_u32(sB0 + b * B_BYTES), P.b_tiles + (size_t)qt * B_BYTES, B_BYTES, B_FULL(b));
      }
      __syncwarp();
    };
    if (my_items) fetch_B(qt_of(item_of(0)), 0);
    for (uint32_t it = 0; it < my_items; ++it) {
      const uint32_t item = item_of(it), chunk = chunk_of(item), qt = qt_of(item);
      const uint32_t t_begin = chunk * P.tiles_per_chunk, t_end = min(t_begin + P.tiles_per_chunk, P.n_db_tiles);
      const bool new_b = new_b_at(it);
      if (new_b) ++b_ord;
      const uint32_t bb_ = b_ord % B_BUFS;
      uint8_t *sB = sB0 + bb_ * B_BYTES;
      const bool has_next = it + 1 < my_items && new_b_at(it + 1);  // the next item needs an operand fetch
      const uint32_t next_qt = it + 1 < my_items ? qt_of(item_of(it + 1)) : 0u;
      // with two buffers the next operand is requested once the ring has turned over (the item that
      // used that buffer is then certainly finished, so the wait inside fetch_B does not stall A loads)
      const uint32_t t_fetch = min(t_begin + (uint32_t)STAGES, t_end - 1);
      // bias value of a query at bound b: one-hot = need (stored negated), +-1 features = c_q
      auto bias_of = [&](int m, int b) -> int {
        const int need = max(0, min((int)sp.L - b, (int)sp.L));
        if constexpr (HAD) return min(254, m - 4 * need);
        else return max(0, min(need - m, 127));
      };
      const uint32_t k0 = HAD ? had_spare_k(0, PB, sp.L) : 0, k1 = HAD ? had_spare_k(1, PB, sp.L) : 0;
      if (new_b) {  // a continued tile keeps the bias bytes (and cur[]) it has reached
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t q = qt * MMA_N + lane * 8 + i;
          meta[i] = (NSYM != 5 && q < sp.Q) ? (int)P.q_meta[q] : 0;
          cur[i] = bias_of(meta[i], (int)sp.L - P.need0);  // what pack_operand_kernel stored
        }
      }
      // Bias refresh: each lane owns 8 consecutive queries of the tile.  The two 16-byte bound
      // loads are issued BEFORE the barrier wait so their L2 latency hides behind it (the bound
      // array is padded to a multiple of the tile width, see run_batch).
      const int4 *bptr = reinterpret_cast<const int4 *>(sp.bound + (size_t)qt * MMA_N + lane * 8);
      const bool dyn = sp.mode != MODE_FIXED;
      auto apply = [&](const int4 &b0, const int4 &b1) {
        const int bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        bool wrote = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t col = lane * 8 + i;
          const int v = bias_of(meta[i], bb[i]);
          if (qt * MMA_N + col < sp.Q && v != cur[i]) {
            cur[i] = v;
            if constexpr (HAD) {
              // both halves are monotone in c, so a reader that sees one old and one new byte still
              // sees a value between the old and the new c: looser, never tighter
              const int c0 = clamp8(v);
              sB[tile_offset(col, k0, KB)] = (uint8_t)(int8_t)c0;
              sB[tile_offset(col, k1, KB)] = (uint8_t)(int8_t)(v - c0);
            } else {
              sB[tile_offset(col, BIAS_K, KB)] = (uint8_t)(int8_t)(-v);
            }
            wrote = true;
          }
        }
        if (wrote) fence_proxy_async();  // generic-proxy writes -> visible to the UMMA (async proxy) reads
        __syncwarp();
      };
      int4 b0 = make_int4(0, 0, 0, 0), b1 = b0;
      if (dyn) { b0 = __ldcg(bptr); b1 = __ldcg(bptr + 1); }
      if (new_b) mbar_wait(B_FULL(bb_), (b_ord / B_BUFS) & 1);
      if (dyn) apply(b0, b1);
      if (new_b && lane == 0) mbar_arrive(B_READY(bb_));
      for (uint32_t t = t_begin; t < t_end; ++t) {
        if (dyn) { b0 = __ldcg(bptr); b1 = __ldcg(bptr + 1); }
        if (elect_one()) {
          mbar_wait(EMPTY(stage), phase ^ 1);
          mbar_expect_tx(FULL(stage), A_BYTES);
          bulk_g2s(smem_u32(sA + stage * A_BYTES), P.a_tiles + (size_t)t * A_BYTES, A_BYTES, FULL(stage));
        }
        __syncwarp();
        if (B_BUFS == 2 && has_next && t == t_fetch) fetch_B(next_qt, b_ord + 1);
        if (dyn) apply(b0, b1);
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
      }
      if (B_BUFS == 1 && has_next) fetch_B(next_qt, b_ord + 1);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=S32 [4,6)=2, A=S8 [7,10)=1, B=S8 [10,13)=1,
    // K-major A/B (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t N_INST = SPLIT_N ? MMA_N / 2 : MMA_N;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_INST >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
    uint32_t stage = 0, phase = 0, tcount = 0, b_ord = 0xffffffffu;
    // descriptors: only the 14-bit start-address field changes (stage and k-step offsets, in 16-byte units)
    const uint64_t desc_hi = ((uint64_t)(P.desc_lbo & 0x3FFFu) << 16) | ((uint64_t)(P.desc_sbo & 0x3FFFu) << 32) | (1ull << 46);
    const uint64_t adesc_base = desc_hi | (uint64_t)((smem_u32(sA) >> 4) & 0x3FFFu);
    const uint64_t bdesc_base0 = desc_hi | (uint64_t)((smem_u32(sB0) >> 4) & 0x3FFFu);
    for (uint32_t it = 0; it < my_items; ++it) {
      const uint32_t chunk = chunk_of(item_of(it));
      const uint32_t t_begin = chunk * P.tiles_per_chunk, t_end = min(t_begin + P.tiles_per_chunk, P.n_db_tiles);
      const bool new_b = new_b_at(it);
      if (new_b) ++b_ord;
      const bool last_of_b = it + 1 >= my_items || new_b_at(it + 1);  // the operand is free after this item
      const uint32_t bb_ = b_ord % B_BUFS;
      const uint64_t bdesc_base = bdesc_base0 + (uint64_t)(bb_ * (B_BYTES >> 4));
      if (new_b) mbar_wait(B_READY(bb_), (b_ord / B_BUFS) & 1);
      for (uint32_t t = t_begin; t < t_end; ++t, ++tcount) {
        const uint32_t buf = tcount & 1, use = tcount >> 1;
        if constexpr (SPLIT_N) {
          mbar_wait(FULL(stage), phase);  // A tile has landed
          const uint64_t ad = adesc_base + (uint64_t)(stage * (A_BYTES >> 4));
#pragma unroll
          for (uint32_t h = 0; h < 2; ++h) {
            mbar_wait(TEMPTY(2 * buf + h), (use & 1) ^ 1);  // the four warps of this half have read it
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + buf * MMA_N + h * N_INST;
              const uint64_t bd = bdesc_base + (uint64_t)(h * (N_INST * KB >> 4));  // query rows 128h.. of the tile image
              tc_mma_i8<false>(d_tmem, ad, bd, idesc);
#pragma unroll
              for (uint32_t ks = 1; ks < (uint32_t)KSTEPS; ++ks) tc_mma_i8<true>(d_tmem, ad + ks * 16, bd + ks * 16, idesc);
              if (h == 1) tc_commit(EMPTY(stage));
              tc_commit(TFULL(2 * buf + h));
              if (h == 1 && t + 1 == t_end && last_of_b) tc_commit(B_EMPTY(bb_));
            }
            __syncwarp();
          }
        } else {
          mbar_wait(TEMPTY(buf), (use & 1) ^ 1);  // epilogue has drained this accumulator buffer
          mbar_wait(FULL(stage), phase);          // A tile has landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + buf * MMA_N;
            const uint64_t ad = adesc_base + (uint64_t)(stage * (A_BYTES >> 4));
            tc_mma_i8<false>(d_tmem, ad, bdesc_base, idesc);
#pragma unroll
            for (uint32_t ks = 1; ks < (uint32_t)KSTEPS; ++ks)
              tc_mma_i8<true>(d_tmem, ad + ks * 16, bdesc_base + ks * 16, idesc);  // +256 bytes per k-step
            tc_commit(EMPTY(stage));  // smem stage reusable once these MMAs have read it
            tc_commit(TFULL(buf));    // accumulator ready for the epilogue
            if (t + 1 == t_end && last_of_b) tc_commit(B_EMPTY(bb_));
          }
          __syncwarp();
        }
        if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 2 + MMA_EPI_WARPS) {
    // ===== epilogue: TMEM -> registers, sign-AND filter, survivors -> ring =====
    const uint32_t quarter = warp & 3;          // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const uint32_t part = (warp - 2) >> 2;      // which column range of the tile this warp drains
    constexpr uint32_t COLS_PER_WARP = MMA_N / (MMA_EPI_WARPS / 4);
    constexpr uint32_t COLS_PER_LD = PACK16 ? 64 : 32;
    constexpr uint32_t CHUNKS = COLS_PER_WARP / COLS_PER_LD;
    uint2 *my_list = lists + (warp - 2) * MMA_LIST_CAP;
    uint32_t tail = 0, head_seen = 0;  // ring positions (free-running, warp-uniform); head_seen = last head read
    uint32_t tcount = 0;
    // Slow path, entered by the whole warp when any lane saw a non-negative accumulator in chunk c:
    // per-lane survivor bit masks, a warp scan for the ring offsets, then a short per-lane store loop.
    auto slow = [&](const uint32_t (&v)[32], const AndTree &tr, uint32_t hitmask, uint32_t qc, uint32_t row) {
      constexpr uint32_t SIGNS = PACK16 ? 0x80008000u : 0x80000000u;
      uint32_t m0 = 0, m1 = 0;  // bit i: accumulator i (pack16: low / high half of register i) is >= 0
      // Survivors are sparse whenever this path matters, so the 64 sign bits are not all tested: only the lanes
      // that hold one descend the AND tree (9-register blocks, then 3-register groups) to the registers involved.
      if ((tr.all & SIGNS) != SIGNS) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if ((tr.u[h] & SIGNS) != SIGNS) {
#pragma unroll
            for (int g = 3 * h; g < 3 * h + 3 && g < 11; ++g) {
              if ((tr.t[g] & SIGNS) != SIGNS) {
#pragma unroll
                for (int i = 3 * g; i < 3 * g + 3 && i < 32; ++i) {
                  if constexpr (PACK16) {
                    m0 |= ((~v[i] >> 15) & 1u) << i;
                    m1 |= (~v[i] >> 31) << i;
                  } else {
                    m0 |= (~v[i] >> 31) << i;
                  }
                }
              }
            }
          }
        }
      }
      __syncwarp();
      const uint32_t n = __popc(m0) + __popc(m1);  // ring entries: one per (query, operand row); the verifier expands the row
      uint32_t incl, total;
      if ((hitmask & (hitmask - 1)) == 0) {  // one lane holds all the survivors (the common case): no scan
        total = __shfl_sync(0xffffffffu, n, __ffs(hitmask) - 1);
        incl = n;  // the holder's exclusive offset is incl - n = 0; the other lanes store nothing
      } else {
        incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
          if ((int)lane >= d) incl += y;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
      }
      // Survivors normally go to the ring.  They are verified right here instead -- straight from the masks, warp-wide
      // per bit -- in a flood (the bound admits > 1/8 of the chunk: more survivors than the ring holds) and when the
      // ring stays full for ~0.1 s (a verifier warp that is not being scheduled: time-slicing, a profiler replay).
      // Either way the answer is the same; nothing here can hang or poison the context.
      bool inline_verify = total > (uint32_t)MMA_LIST_CAP;
      if (!inline_verify && tail + total - head_seen > (uint32_t)MMA_LIST_CAP) {  // ring full: wait for the verifier warp
        const long long t0 = clock64();
        do {
          head_seen = ld_shared_volatile(ring_head + (warp - 2));
          if (clock64() - t0 > 200000000ll) { inline_verify = true; break; }
        } while (tail + total - head_seen > (uint32_t)MMA_LIST_CAP);
        __threadfence_block();  // the slots were read before the head moved
      }
      if (inline_verify) {
#pragma unroll 1
        for (int i = 0; i < 32; ++i) {
          mma_verify_rows_warp<UPR>(sp, (m0 >> i) & 1u, qc + (PACK16 ? 2 * i : i), row, lane);
          if constexpr (PACK16) mma_verify_rows_warp<UPR>(sp, (m1 >> i) & 1u, qc + 2 * i + 1, row, lane);
        }
      } else {
        uint32_t off = tail + incl - n;
        while (m0) {
          const int i = __ffs(m0) - 1;
          m0 &= m0 - 1;
          my_list[off++ & (MMA_LIST_CAP - 1)] = make_uint2(qc + (PACK16 ? 2 * i : i), row);
        }
        if constexpr (PACK16) {
          while (m1) {
            const int i = __ffs(m1) - 1;
            m1 &= m1 - 1;
            my_list[off++ & (MMA_LIST_CAP - 1)] = make_uint2(qc + 2 * i + 1, row);
          }
        }
        tail += total;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();  // entries before the new tail
          st_shared_volatile(ring_tail + (warp - 2), tail);
        }
      }
      __syncwarp();
    };
    for (uint32_t it = 0; it < my_items; ++it) {
      const uint32_t item = item_of(it), chunk = chunk_of(item), qt = qt_of(item);
      const uint32_t t_begin = chunk * P.tiles_per_chunk, t_end = min(t_begin + P.tiles_per_chunk, P.n_db_tiles);
      const uint32_t qbase = qt * MMA_N + part * COLS_PER_WARP;
      for (uint32_t t = t_begin; t < t_end; ++t, ++tcount) {
        const uint32_t buf = tcount & 1, use = tcount >> 1;
        const uint32_t unit = SPLIT_N ? 2 * buf + part : buf;  // the barrier pair this warp hands over on
        mbar_wait(TFULL(unit), use & 1);
        tc_fence_after();
        const uint32_t row = t * MMA_M + quarter * 32 + lane;  // db operand row = windows [row * UPR, row * UPR + UPR)
        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * MMA_N + part * COLS_PER_WARP;
        auto load = [&](uint32_t c, uint32_t (&v)[32]) {
          if constexpr (PACK16) tc_ld32_pack16(taddr + c * 64, v);
          else tc_ld32(taddr + c * 32, v);
        };
        auto process = [&](const uint32_t (&v)[32], uint32_t c) {
          const AndTree tr = and_tree32(v);
          const bool hit = PACK16 ? (tr.all & 0x80008000u) != 0x80008000u : (int)tr.all >= 0;
          const uint32_t hitmask = __ballot_sync(0xffffffffu, hit);
          if (hitmask) slow(v, tr, hitmask, qbase + c * COLS_PER_LD, row);
        };
        uint32_t va[32];
        if (P.dump != nullptr && item == 0 && t == t_begin) {  // debug hook, off the hot path
          for (uint32_t c = 0; c < CHUNKS; ++c) {
            load(c, va);
            tc_wait_ld();
            int32_t *drow = P.dump + (quarter * 32 + lane) * MMA_N + part * COLS_PER_WARP + c * COLS_PER_LD;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if constexpr (PACK16) {
                drow[2 * i] = (int32_t)(int16_t)(va[i] & 0xffffu);
                drow[2 * i + 1] = (int32_t)(int16_t)(va[i] >> 16);
              } else {
                drow[i] = (int32_t)va[i];
              }
            }
          }
        }
        if constexpr (CHUNKS == 1) {
          // the whole column range fits one load: the accumulator buffer is released as soon as the
          // registers hold it, before any filtering
          load(0, va);
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(TEMPTY(unit));
          process(va, 0);
        } else {
          // two loads in flight; the buffer is released after the last wait
          uint32_t vb[32];
          load(0, va);
#pragma unroll
          for (uint32_t c = 0; c < CHUNKS; c += 2) {
            load(c + 1, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c + 2 == CHUNKS) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(TEMPTY(unit));
            }
            process(va, c);
            if (c + 2 < CHUNKS) load(c + 2, va);
            process(vb, c + 1);
            if (c + 2 < CHUNKS) tc_wait_ld();
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      st_shared_volatile(ring_done + (warp - 2), 1u);
    }
  } else {
    // ===== verifier: exact re-check + emission of the survivors the epilogue warps queued =====
    constexpr uint32_t RPV = MMA_EPI_WARPS / MMA_VER_WARPS;  // rings per verifier warp
    const uint32_t first = (warp - 2 - MMA_EPI_WARPS) * RPV;
    uint32_t head[RPV], tl[RPV];
#pragma unroll
    for (uint32_t r = 0; r < RPV; ++r) head[r] = 0;
    uint32_t waited = 0, nap = 128;
    for (;;) {
      // Verification is a chain of dependent L2 round trips whose cost does not depend on how many lanes take
      // part, so a batch is gathered across this warp's rings and started only when it is full -- or when the
      // survivors have waited for a while, or their producers are done.
      bool all_done = true;
      uint32_t total = 0;
#pragma unroll
      for (uint32_t r = 0; r < RPV; ++r) {
        // `done` is read before the tail: a ring seen done has its final tail published
        const uint32_t done = ld_shared_volatile(ring_done + first + r);
        __threadfence_block();
        tl[r] = ld_shared_volatile(ring_tail + first + r);
        total += tl[r] - head[r];
        if (!done) all_done = false;
      }
      if (total == 0) {
        if (all_done) break;
        // Idle is the normal state (tight bounds: a few survivors per thousand tiles), so the poll backs off to
        // ~4 us: frequent polling cost the tile loop 3 % (the verifiers share issue slots with the epilogue warps).
        // A hang is impossible here: the epilogue warps' own waits are bounded.
        waited = 0;
        nap = min(nap * 2u, 4096u);
        __nanosleep(nap);
        continue;
      }
      // give a partial batch ~10 us to fill up (wide rows: a single entry is already UPR windows of work)
      if (total < (UPR >= 4 ? 8u : 32u) && !all_done && waited < (UPR >= 4 ? 4u : 16u)) {
        ++waited;
        __nanosleep(512);
        continue;
      }
      waited = 0;
      nap = 128;
      __threadfence_block();  // entries are read after the tails
      uint2 e = make_uint2(0, 0);
      uint32_t taken = 0;
#pragma unroll
      for (uint32_t r = 0; r < RPV; ++r) {
        const uint32_t n = min(tl[r] - head[r], 32u - taken);
        if (lane >= taken && lane < taken + n) e = lists[(first + r) * MMA_LIST_CAP + ((head[r] + lane - taken) & (MMA_LIST_CAP - 1))];
        head[r] += n;
        taken += n;
      }
      __syncwarp();
      if (lane < RPV) {
        __threadfence_block();  // the slots are in registers: the producers may reuse them
#pragma unroll
        for (uint32_t r = 0; r < RPV; ++r)
          if (lane == r) st_shared_volatile(ring_head + first + r, head[r]);
      }
      mma_verify_rows_warp<UPR>(sp, lane < taken, e.x, e.y, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// One thread per (row, 16-byte k-chunk): writes the int8 operand image of rows [row_begin,row_end) in
// encoding `enc` (see "operand encodings" above).  Rows >= n_valid are padding and can never pass the
// sign filter.  Query rows (is_query) get their bias for the batch's initial need0 = L - bound0 and
// their per-query constant in meta[].
// ENC_C / IS_QUERY_C / ALPHA_C >= 0 fix the encoding, the side and the alphabet at compile time (the per-step query
// operand and the nucleotide db images: the branches below fold away, ~3x faster than the generic instantiation).
template <int ENC_C, int IS_QUERY_C, int ALPHA_C>
__global__ void pack_operand_kernel(const uint64_t *__restrict__ ref, uint32_t n_valid, uint32_t row_begin, uint32_t row_end,
                                    uint32_t W, uint32_t L, uint32_t rows_per_tile, uint32_t KB, uint32_t enc_rt, int is_query_rt,
                                    int need0, int alphabet_rt, int16_t *__restrict__ meta, uint8_t *__restrict__ out,
                                    uint32_t upr) {
  const uint32_t enc = ENC_C >= 0 ? (uint32_t)ENC_C : enc_rt;
  const int is_query = IS_QUERY_C >= 0 ? IS_QUERY_C : is_query_rt;
  const int alphabet = ALPHA_C >= 0 ? ALPHA_C : alphabet_rt;
  const bool aa_exact = enc == MMA_ENC_AA;
  const uint32_t chunks = KB / 16, PB = aa_exact ? MMA_AA_POS : KB / enc, gap = PB - L;
  // thread -> (row, k-chunk): eight consecutive threads take the eight rows of one core-matrix column, whose 16-byte
  // pieces are contiguous in the tile image (128-byte stores instead of 16-byte ones 128 bytes apart)
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t t = (uint32_t)(idx % (8 * chunks));
  const uint32_t row = row_begin + (uint32_t)(idx / (8 * chunks)) * 8 + (t & 7), c = t >> 3;
  if (row >= row_end) return;
  // upr > 1 (db side of the one-hot encodings only): the row is the union of windows upr*row .. upr*row + upr-1;
  // n_valid counts windows
  const bool valid = (uint64_t)row * upr < n_valid, had = enc <= 3;
  const uint64_t *w = ref + (size_t)row * upr * W;
  const uint32_t extra = valid ? (uint32_t)min((uint64_t)upr, (uint64_t)n_valid - (uint64_t)row * upr) - 1 : 0;  // further windows of the row
  int nN = 0;  // N/gap positions: code 1 = bit 0 of a 5-bit group (protein: symbols of the N-like filter class)
  if (valid) {
    if (alphabet == ALPHA_NUC)
      for (uint32_t i = 0; i < W; ++i) nN += __popcll(w[i] & 0x0084210842108421ull);
    else
      for (uint32_t x = 0; x < L; ++x) nN += aa_class_code((uint32_t)(w[x / 12] >> (5 * (x % 12))) & 31u) == 1u;
  }
  if (aa_exact) {  // one-hot over the symbol numbers 1..20; bias in slot 400
    uint32_t oo[4] = {0, 0, 0, 0};
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) {
      const uint32_t k = c * 16 + i, sy = k / MMA_AA_POS, p = k % MMA_AA_POS;
      int v = 0;
      if (k < MMA_ENC_AA * MMA_AA_POS) {
        if (valid && p < L) v = ((uint32_t)(w[p / 12] >> (5 * (p % 12))) & 31u) == sy + 1;
      } else if (k == MMA_ENC_AA * MMA_AA_POS) {
        v = !is_query ? 1 : (valid ? -max(0, min(need0 - nN, 127)) : -128);
      }
      oo[i >> 2] |= ((uint32_t)v & 0xffu) << (8 * (i & 3));
    }
    if (is_query && meta != nullptr && c == 0) meta[row] = (int16_t)(valid ? nN : 0);
    const uint32_t tile = row / rows_per_tile, r = row % rows_per_tile;
    *reinterpret_cast<uint4 *>(out + (size_t)tile * rows_per_tile * KB + tile_offset(r, c * 16, KB)) =
        make_uint4(oo[0], oo[1], oo[2], oo[3]);
    return;
  }
  const uint32_t pb_shift = 31u - (uint32_t)__clz(PB);  // PB = 32 or 64 here (the protein one-hot returned above)
  const int alpha = enc == 2 ? 2 : 1, w5 = alpha + 4, T = (int)(enc * gap) - 3;
  const int over = max(0, nN - (T - 1));
  int qbase = 0, cq = 0;
  if (had) {
    qbase = alpha * ((int)L - nN) + max(0, w5 * over - 127);
    cq = min(254, qbase - 4 * need0);
  }
  uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
  for (uint32_t i = 0; i < 16; ++i) {
    const uint32_t k = c * 16 + i, f = k >> pb_shift, p = k & (PB - 1);
    int v = 0;
    if (p < L) {
      if (valid) {
        const uint32_t code = filter_code((uint32_t)(w[p / 12] >> (5 * (p % 12))) & 31u, alphabet);
        if (!had) {
          v = code == (16u >> f);  // A C G T N
          for (uint32_t u = 1; u <= extra; ++u)
            v |= filter_code((uint32_t)(w[u * W + p / 12] >> (5 * (p % 12))) & 31u, alphabet) == (16u >> f);
        } else if (code >= 2 && (code & (code - 1)) == 0) {
          // A=16 (+,+,+)  C=8 (+,-,-)  G=4 (-,+,-)  T=2 (-,-,+): h, l, h*l
          const uint32_t plus = f == 0 ? (16u | 8u) : (f == 1 ? (16u | 4u) : (16u | 2u));
          v = (code & plus) ? 1 : -1;
        }
      }
    } else if (!had) {
      if (f == 0 && p == PB - 1)  // the bias slot
        v = !is_query ? 1 : (valid ? -max(0, min(need0 - (enc == 4 ? nN : 0), 127)) : -128);
    } else {
      const int si = (int)(f * gap + (p - L));
      if (is_query) {
        if (si == 0) v = valid ? clamp8(cq) : -128;
        else if (si == 1) v = valid ? cq - clamp8(cq) : -128;
        else if (si == 2) v = 1;
        else if (valid) v = (si - 3) < T - 1 ? (nN >= si - 2 ? w5 : 0) : min(127, w5 * over);
      } else {
        if (si <= 1) v = 1;
        else if (si == 2) v = valid ? -alpha * nN : -128;
        else v = (valid && nN >= si - 2) ? 1 : 0;
      }
    }
    o[i >> 2] |= ((uint32_t)v & 0xffu) << (8 * (i & 3));
  }
  if (is_query && meta != nullptr && c == 0) meta[row] = (int16_t)(valid ? (had ? qbase : nN) : 0);
  const uint32_t tile = row / rows_per_tile, r = row % rows_per_tile;
  uint8_t *dst = out + (size_t)tile * rows_per_tile * KB + tile_offset(r, c * 16, KB);
  *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
}

// The per-step case of pack_operand_kernel written out: nucleotide QUERY rows of the 4-symbol one-hot operand (the B
// side of every union-row scan).  A thread writes the 16 bytes of (row, k-chunk c): symbol f = c / (PB/16), positions
// p0 .. p0+15 with p0 = (c % (PB/16)) * 16 -- the symbol's bit of position p is bit 5 (p % 12) + 4 - f of word p / 12, all
// of it known per chunk at compile time (the switch below), so a byte costs a shift and a mask.  Only the threads of the
// chunks that hold the bias byte (symbol A, position PB-1) and the per-query constant count the query's N's.
template <uint32_t PB>
__global__ void pack_query_onehot_kernel(const uint64_t *__restrict__ ref, uint32_t n_valid, uint32_t n_rows, uint32_t W, uint32_t L,
                                         int need0, int16_t *__restrict__ meta, uint8_t *__restrict__ out) {
  constexpr uint32_t KB = 4 * PB, CHUNKS = KB / 16, CPS = PB / 16;  // chunks per symbol
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t t = (uint32_t)(idx % (8 * CHUNKS));
  const uint32_t row = (uint32_t)(idx / (8 * CHUNKS)) * 8 + (t & 7), c = t >> 3;
  if (row >= n_rows) return;
  const bool valid = row < n_valid;
  uint64_t w[6] = {0, 0, 0, 0, 0, 0};
  if (valid) {
#pragma unroll
    for (uint32_t x = 0; x < 6; ++x)
      if (x < W) w[x] = __ldg(ref + (size_t)row * W + x);
  }
  const uint32_t f = c / CPS, part = c % CPS;
  uint32_t o[4] = {0, 0, 0, 0};
  auto fill = [&](auto P0) {
    constexpr uint32_t p0 = decltype(P0)::value;
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) {
      constexpr uint32_t dummy = 0;
      (void)dummy;
      const uint32_t p = p0 + i;
      if (p < 72) {  // six words hold 72 positions; p < L is applied below
        const uint32_t bit = (uint32_t)(w[p / 12] >> (5 * (p % 12) + 4 - f)) & 1u;
        o[i >> 2] |= (p < L ? bit : 0u) << (8 * (i & 3));
      }
    }
  };
  switch (part) {
    case 0: fill(std::integral_constant<uint32_t, 0>()); break;
    case 1: fill(std::integral_constant<uint32_t, 16>()); break;
    case 2: fill(std::integral_constant<uint32_t, 32>()); break;
    default: fill(std::integral_constant<uint32_t, 48>()); break;
  }
  const bool bias_chunk = f == 0 && part == CPS - 1, meta_chunk = c == 0;
  if (bias_chunk || meta_chunk) {
    int nN = 0;
#pragma unroll
    for (uint32_t x = 0; x < 6; ++x) nN += __popcll(w[x] & 0x0084210842108421ull);
    if (bias_chunk) {  // byte 15 of the chunk = position PB - 1 of symbol A: the negated need (see the file header)
      const int v = valid ? -max(0, min(need0 - nN, 127)) : -128;
      o[3] = (o[3] & 0x00ffffffu) | (((uint32_t)v & 0xffu) << 24);
    }
    if (meta_chunk && meta != nullptr) meta[row] = (int16_t)(valid ? nN : 0);
  }
  const uint32_t tile = row / MMA_N, r = row % MMA_N;
  *reinterpret_cast<uint4 *>(out + (size_t)tile * MMA_N * KB + tile_offset(r, c * 16, KB)) = make_uint4(o[0], o[1], o[2], o[3]);
}

static void launch_pack_operand(const uint64_t *ref, uint32_t n_valid, uint32_t row_begin, uint32_t row_end, uint32_t W,
                                uint32_t L, uint32_t rows_per_tile, uint32_t KB, uint32_t enc, int is_query, int need0,
                                int alphabet, int16_t *meta, uint8_t *out, cudaStream_t s, uint32_t upr = 1) {
  if (row_end <= row_begin) return;
  const uint64_t n = (uint64_t)((row_end - row_begin + 7) / 8) * 8 * (KB / 16);
  const unsigned grid = (unsigned)((n + 255) / 256);
  static const bool fast_query_pack = getenv("SMAFA_NO_FAST_PACK") == nullptr;
  if (fast_query_pack && alphabet == ALPHA_NUC && enc == 4 && is_query && upr == 1 && row_begin == 0 && rows_per_tile == MMA_N &&
      W <= 6 && (KB == 256 || KB == 128)) {
    if (KB == 256) pack_query_onehot_kernel<64><<<grid, 256, 0, s>>>(ref, n_valid, row_end, W, L, need0, meta, out);
    else pack_query_onehot_kernel<32><<<grid, 256, 0, s>>>(ref, n_valid, row_end, W, L, need0, meta, out);
    return;
  }
#define SMAFA_PACK_ARGS ref, n_valid, row_begin, row_end, W, L, rows_per_tile, KB, enc, is_query, need0, alphabet, meta, out, upr
  if (alphabet == ALPHA_NUC && enc == 4 && is_query) pack_operand_kernel<4, 1, ALPHA_NUC><<<grid, 256, 0, s>>>(SMAFA_PACK_ARGS);
  else if (alphabet == ALPHA_NUC && enc == 4) pack_operand_kernel<4, 0, ALPHA_NUC><<<grid, 256, 0, s>>>(SMAFA_PACK_ARGS);
  else if (alphabet == ALPHA_NUC && enc == 3 && is_query) pack_operand_kernel<3, 1, ALPHA_NUC><<<grid, 256, 0, s>>>(SMAFA_PACK_ARGS);
  else if (alphabet == ALPHA_NUC && enc == 3) pack_operand_kernel<3, 0, ALPHA_NUC><<<grid, 256, 0, s>>>(SMAFA_PACK_ARGS);
  else pack_operand_kernel<-1, -1, -1><<<grid, 256, 0, s>>>(SMAFA_PACK_ARGS);
#undef SMAFA_PACK_ARGS
}

// Peak probe: one thread per CTA issues back-to-back int8 MMAs (two alternating accumulators, ten
// k-step operand offsets like the real kernel); no loads, no epilogue.
__global__ void __launch_bounds__(64, 1) mma_peak_kernel(uint32_t n_mma) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 128 * 320 + 256 * 320);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  for (uint32_t i = threadIdx.x; i < (128 * 320 + 256 * 320) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
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
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MMA_N >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 320;
    for (uint32_t i = 0; i < n_mma; ++i) {
      const uint32_t ks = i % 10, buf = (i / 10) & 1;
      if (ks == 0) tc_mma_i8<false>(tmem + buf * MMA_N, smem_desc(a0, 8, 160), smem_desc(b0, 8, 160), idesc);
      else tc_mma_i8<true>(tmem + buf * MMA_N, smem_desc(a0 + ks * 256, 8, 160), smem_desc(b0 + ks * 256, 8, 160), idesc);
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

}  // namespace smafa

using namespace smafa;

int mma_peak_probe(smafa_ctx *ctx, uint32_t mmas_per_cta, float *ms) {
  const size_t smem = 128 * 320 + 256 * 320 + 64;
  cudaError_t e = cudaFuncSetAttribute(mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SMAFA_E_CUDA; }
  cudaStream_t s = ctx->stream;
  mma_peak_kernel<<<ctx->num_sms, 64, smem, s>>>(mmas_per_cta / 10 + 10);  // warm-up
  cudaEventRecord(ctx->ev[0], s);
  mma_peak_kernel<<<ctx->num_sms, 64, smem, s>>>(mmas_per_cta);
  cudaEventRecord(ctx->ev[1], s);
  e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("mma_peak_kernel: ") + cudaGetErrorString(e); return SMAFA_E_CUDA; }
  cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]);
  return SMAFA_OK;
}

bool mma_supported(const smafa_db *db);
static uint32_t mma_kb(const smafa_db *db) {
  return db->mma_nsym == MMA_ENC_AA ? MMA_AA_KB : mma_pb(db->mma_nsym, db->L) * db->mma_nsym;
}

// The encoding a db of window length L gets when `want` is requested (+-1 features need two spare
// positions per feature block: L = 63 falls back to the 4-symbol one-hot operands).
uint32_t mma_pick_encoding(uint32_t want, uint32_t L, int alphabet) {
  // protein: the 4-class filter image is too permissive for loose bounds (random 20-aa windows sit at class
  // distance ~15 of 20), so short protein windows get exact one-hot operands unless an ablation asks otherwise
  if (alphabet == ALPHA_AA && mma_enc_ok(MMA_ENC_AA, L) && !getenv("SMAFA_MMA_NSYM")) return MMA_ENC_AA;
  if (want < 2 || want > 5) want = 3;
  return mma_enc_ok(want, L) ? want : 4;
}

extern "C" uint32_t smafa_db_mma_k(const smafa_db *db) { return db && mma_supported(db) ? mma_kb(db) : 0; }

bool mma_supported(const smafa_db *db) { return !db->generic_only && mma_enc_ok(db->mma_nsym, db->L); }

static int mma_fail(smafa_ctx *ctx, int code, const char *what, cudaError_t e) {
  ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
  smafa_set_global_error(ctx->err);
  return code;
}

// Union-row operand images (kernel template parameter UPR = 2, 3): 4-symbol one-hot, UPR windows per row.
static uint32_t union_kb(const smafa_db *db) { return 4 * mma_pb(4, db->L); }
static uint32_t union_max_degree(const smafa_ctx *ctx, const smafa_db *db) {
  if (db->alphabet != ALPHA_NUC || !mma_enc_ok(4, db->L)) return 1;
  if (db->grouped && ctx->mma_union >= 3) return 16;  // grouped db: near-copies share a row, wide rows stay selective
  return std::min<uint32_t>(ctx->mma_union, 3);
}

static int union_reserve(smafa_ctx *ctx, smafa_db *db, uint32_t upr, uint64_t rows) {
  const uint64_t per_tile = (uint64_t)upr * MMA_M;
  const uint64_t tiles = (rows + per_tile - 1) / per_tile;
  uint8_t *&img = db->union_img[union_slot(upr)];
  if (tiles <= db->union_cap[union_slot(upr)]) return SMAFA_OK;
  const size_t tile_bytes = (size_t)MMA_M * union_kb(db);
  uint8_t *n = nullptr;
  cudaError_t e = cudaMalloc((void **)&n, tiles * tile_bytes);
  if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_OOM, "cudaMalloc(union-row db operand)", e);
  if (img && db->D) cudaMemcpyAsync(n, img, ((db->D + per_tile - 1) / per_tile) * tile_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(img);
  img = n;
  db->union_cap[union_slot(upr)] = tiles;
  return SMAFA_OK;
}

int mma_db_reserve(smafa_ctx *ctx, smafa_db *db, uint64_t rows) {
  if (db->L == 0 || db->L > 63) return SMAFA_OK;
  for (uint32_t upr : UNION_DEGREES) {
    if (upr > union_max_degree(ctx, db)) break;
    int rc = union_reserve(ctx, db, upr, rows);
    if (rc) return rc;
  }
  const uint64_t tiles = (rows + MMA_M - 1) / MMA_M;
  if (tiles <= db->onehot_cap) return SMAFA_OK;
  const size_t tile_bytes = (size_t)MMA_M * mma_kb(db);
  uint8_t *n = nullptr;
  cudaError_t e = cudaMalloc((void **)&n, tiles * tile_bytes);
  if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_OOM, "cudaMalloc(one-hot db)", e);
  if (db->onehot && db->D)
    cudaMemcpyAsync(n, db->onehot, ((db->D + MMA_M - 1) / MMA_M) * tile_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(db->onehot);
  db->onehot = n;
  db->onehot_cap = tiles;
  return SMAFA_OK;
}

// Packs rows [first, first+n) and re-pads the tail of the last tile.  db->ref already holds the rows.
int mma_db_pack(smafa_ctx *ctx, smafa_db *db, uint64_t first, uint64_t n) {
  if (db->L == 0 || db->L > 63 || n == 0) return SMAFA_OK;
  const uint32_t end = (uint32_t)(first + n);
  const uint32_t padded = (end + MMA_M - 1) / MMA_M * MMA_M;
  launch_pack_operand(db->ref, end, (uint32_t)first, padded, db->W, db->L, MMA_M, mma_kb(db), db->mma_nsym, 0, 0,
                      db->alphabet, nullptr, db->onehot, ctx->stream);
  for (uint32_t upr : UNION_DEGREES) {  // rows that hold a window of [first, end), then the padding of the last tile
    if (db->union_img[union_slot(upr)] == nullptr) continue;
    const uint32_t r_begin = (uint32_t)(first / upr), r_end = (end + upr - 1) / upr;
    launch_pack_operand(db->ref, end, r_begin, (r_end + MMA_M - 1) / MMA_M * MMA_M, db->W, db->L, MMA_M, union_kb(db), 4, 0, 0,
                        db->alphabet, nullptr, db->union_img[union_slot(upr)], ctx->stream, upr);
  }
  return SMAFA_OK;
}

void mma_db_free(smafa_db *db) {
  cudaFree(db->onehot);
  for (int i = 0; i < 5; ++i) {
    cudaFree(db->union_img[i]);
    db->union_img[i] = nullptr;
    db->union_cap[i] = 0;
  }
  db->onehot = nullptr;
  db->onehot_cap = 0;
}

template <int KSTEPS, int NSYM, int STAGES, int EPI_WARPS, bool PACK16, int B_BUFS = 2, bool SPLIT_N = false, int UPR = 1>
static cudaError_t launch_mma(const MmaParams &P, uint32_t grid, cudaStream_t s) {
  constexpr size_t smem = (size_t)B_BUFS * MMA_N * KSTEPS * 32 + (size_t)STAGES * MMA_M * KSTEPS * 32 + 512 +
                          (size_t)EPI_WARPS * MMA_LIST_CAP * sizeof(uint2);
  static_assert(smem <= 232448, "more than 227 KB of shared memory");
  auto kern = scan_mma_kernel<KSTEPS, NSYM, STAGES, EPI_WARPS, PACK16, B_BUFS, SPLIT_N, UPR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, mma_threads(EPI_WARPS), smem, s>>>(P);
  return cudaGetLastError();
}

int mma_scan(smafa_ctx *ctx, const smafa_db *db, ScanParams &p, cudaStream_t s, int32_t *dump) {
  // Operand choice for this scan: ctx->mma_union_pick windows per db operand row (api.cu pick_union_degree), if this
  // db holds that image; the db's own encoding (+-1 features by default) with one window per row otherwise.
  uint32_t upr = ctx->mma_union_pick;
  if (union_slot(upr) < 0 || db->union_img[union_slot(upr)] == nullptr) upr = 1;
  const bool use_union = upr > 1;
  const uint32_t enc = use_union ? 4u : db->mma_nsym;
  const uint32_t KB = use_union ? union_kb(db) : mma_kb(db);
  ctx->last_mma_k = KB / upr;
  ctx->mma_union_used = upr;
  const uint32_t n_qtiles = (p.Q + MMA_N - 1) / MMA_N;
  const size_t b_bytes = (size_t)n_qtiles * MMA_N * KB + (size_t)n_qtiles * MMA_N * sizeof(int16_t);  // operand tiles + q_meta
  if (ctx->q_onehot_cap < b_bytes) {
    const size_t want = std::max(b_bytes, ctx->q_onehot_cap * 2);  // geometric: see ensure_buf (api.cu)
    cudaStreamSynchronize(s);
    cudaFree(ctx->q_onehot);
    ctx->q_onehot = nullptr;
    ctx->q_onehot_cap = 0;
    cudaError_t e = cudaMalloc((void **)&ctx->q_onehot, want);
    if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_OOM, "cudaMalloc(one-hot queries)", e);
    ctx->q_onehot_cap = want;
  }
  int16_t *meta = reinterpret_cast<int16_t *>(ctx->q_onehot + (size_t)n_qtiles * MMA_N * KB);
  MmaParams P{};
  P.sp = p;
  P.need0 = std::max(0, (int)p.L - ctx->mma_bound0);  // the initial bound is uniform over the batch
  launch_pack_operand(p.q_ref, p.Q, 0, n_qtiles * MMA_N, p.W, p.L, MMA_N, KB, enc, 1, P.need0, db->alphabet, meta,
                      ctx->q_onehot, s);
  P.dump = dump;
  P.q_meta = meta;
  P.a_tiles = use_union ? db->union_img[union_slot(upr)] : db->onehot;
  P.b_tiles = ctx->q_onehot;
  P.n_qtiles = n_qtiles;
  P.n_db_tiles = (uint32_t)((db->D + MMA_M * upr - 1) / (MMA_M * upr));
  // Work items = query tiles x db chunks, dealt to the CTAs round-robin (chunk-major).  The scan ends with its slowest
  // CTA, so the chunk count is chosen to minimise  rounds x (tiles per chunk + the hand-over between two items), rounds =
  // ceil(items / CTAs): 100 k x 1 M at degree 16 is 391 query tiles x 489 db tiles -- 8 chunks give 3128 items = 22
  // rounds of 62 tiles (1364 + hand-overs), 3 chunks give 1173 items = 8 rounds of 163 (1304): the ncu capture of the
  // 8-chunk schedule had the SMs active 94 % of the time (profiles/r02_ncu_mma_u16_summary.txt).  Chunks stay below
  // 1024 tiles so that a chunk's tiles remain L2-resident while all CTAs stream it.
  uint32_t tiles_per_chunk = P.n_db_tiles;
  {
    const uint64_t ctas = std::min<uint64_t>((uint64_t)ctx->num_sms, (uint64_t)n_qtiles * P.n_db_tiles);
    const uint32_t handover = 6;  // tile times lost between two items of a CTA (query operand swap)
    uint64_t best_cost = UINT64_MAX;
    for (uint32_t n = 1; n <= 96 && n <= P.n_db_tiles; ++n) {
      const uint32_t per = (P.n_db_tiles + n - 1) / n;
      if (per > 1024 && n < 96 && n < P.n_db_tiles) continue;
      const uint32_t n_eff = (P.n_db_tiles + per - 1) / per;  // chunks of `per` tiles that are really needed
      const uint64_t rounds = ((uint64_t)n_qtiles * n_eff + ctas - 1) / ctas;
      const uint64_t cost = rounds * (per + handover);
      if (cost < best_cost) { best_cost = cost; tiles_per_chunk = per; }
    }
  }
  // a db image that fits L2 with room to spare (wide union rows): query-tile-major order, small chunks (see the kernel)
  static const int qt_major_env = getenv("SMAFA_MMA_QT_MAJOR") ? atoi(getenv("SMAFA_MMA_QT_MAJOR")) : -1;
  const size_t image_bytes = (size_t)P.n_db_tiles * MMA_M * KB;
  P.qt_major = qt_major_env >= 0 ? (uint32_t)(qt_major_env != 0 && use_union && image_bytes <= (48u << 20)) : 0u;
  if (P.qt_major) tiles_per_chunk = 16;
  P.tiles_per_chunk = tiles_per_chunk;
  P.n_chunks = (P.n_db_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
  // K-major, no swizzle: LBO = distance between the two 16-byte k-chunks of one k-step (128 B),
  // SBO = distance between 8-row groups (8*KB bytes)
  P.desc_lbo = 128 >> 4;
  P.desc_sbo = (8 * KB) >> 4;
  if (const char *e = getenv("SMAFA_MMA_SWAP_LBO_SBO")) {
    if (e[0] == '1') std::swap(P.desc_lbo, P.desc_sbo);
  }
  const uint32_t n_items = P.n_qtiles * P.n_chunks;
  const uint32_t grid = std::min<uint32_t>((uint32_t)ctx->num_sms, n_items);
  cudaError_t e;
  const bool wide = mma_pb(enc, db->L) == 64;
  if (use_union) {
    // One query operand buffer and four db tile stages (default since round 2: 4.39 -> 4.15 ms per 1e11 comparisons at
    // degree 3, 6.5 -> 6.16 at degree 2, whole parity suite green under it: profiles/r02_union_calib_stages4.log,
    // r02_pytest_gpu_stages4.log); SMAFA_MMA_UNION_STAGES4=0 keeps the two-and-two shape reachable
    static const bool stages4 = getenv("SMAFA_MMA_UNION_STAGES4") ? atoi(getenv("SMAFA_MMA_UNION_STAGES4")) != 0 : true;
    if (stages4 && wide) {
      switch (upr) {
        case 2: e = launch_mma<8, 4, 4, 8, true, 1, false, 2>(P, grid, s); break;
        case 3: e = launch_mma<8, 4, 4, 8, true, 1, false, 3>(P, grid, s); break;
        case 4: e = launch_mma<8, 4, 4, 8, true, 1, false, 4>(P, grid, s); break;
        case 8: e = launch_mma<8, 4, 4, 8, true, 1, false, 8>(P, grid, s); break;
        default: e = launch_mma<8, 4, 4, 8, true, 1, false, 16>(P, grid, s); break;
      }
      if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_CUDA, "scan_mma_kernel launch", e);
      return 2;
    }
    switch (upr) {
      case 2: e = wide ? launch_mma<8, 4, 2, 8, true, 2, false, 2>(P, grid, s) : launch_mma<4, 4, 4, 8, true, 2, false, 2>(P, grid, s); break;
      case 3: e = wide ? launch_mma<8, 4, 2, 8, true, 2, false, 3>(P, grid, s) : launch_mma<4, 4, 4, 8, true, 2, false, 3>(P, grid, s); break;
      case 4: e = wide ? launch_mma<8, 4, 2, 8, true, 2, false, 4>(P, grid, s) : launch_mma<4, 4, 4, 8, true, 2, false, 4>(P, grid, s); break;
      case 8: e = wide ? launch_mma<8, 4, 2, 8, true, 2, false, 8>(P, grid, s) : launch_mma<4, 4, 4, 8, true, 2, false, 8>(P, grid, s); break;
      default: e = wide ? launch_mma<8, 4, 2, 8, true, 2, false, 16>(P, grid, s) : launch_mma<4, 4, 4, 8, true, 2, false, 16>(P, grid, s); break;
    }
    if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_CUDA, "scan_mma_kernel launch", e);
    return 2;
  }
  // Epilogue shape: 8 warps + .pack::16b TMEM loads measured best (profiles/r01_epilogue_variants.txt; the 16-warp
  // shape measured there is gone since the verifier warps took its register budget); SMAFA_MMA_PACK16=0 keeps the
  // unpacked loads of the default encoding reachable.
  static const bool pack16 = getenv("SMAFA_MMA_PACK16") ? atoi(getenv("SMAFA_MMA_PACK16")) != 0 : true;
  // SMAFA_MMA_SPLIT_N=1: two N = 128 instructions per k-step with per-half accumulator hand-over (see the kernel)
  static const bool split_n = getenv("SMAFA_MMA_SPLIT_N") ? atoi(getenv("SMAFA_MMA_SPLIT_N")) != 0 : false;
  if (split_n && pack16 && (db->mma_nsym == 2 || db->mma_nsym == 3 || (db->mma_nsym == 4 && wide))) {
    if (db->mma_nsym == 4) e = launch_mma<8, 4, 2, 8, true, 2, true>(P, grid, s);
    else if (db->mma_nsym == 2) e = wide ? launch_mma<4, 2, 4, 8, true, 2, true>(P, grid, s) : launch_mma<2, 2, 4, 8, true, 2, true>(P, grid, s);
    else e = wide ? launch_mma<6, 3, 4, 8, true, 2, true>(P, grid, s) : launch_mma<3, 3, 4, 8, true, 2, true>(P, grid, s);
    if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_CUDA, "scan_mma_kernel launch", e);
    return 2;
  }
  switch (db->mma_nsym) {
    case MMA_ENC_AA: e = launch_mma<13, (int)MMA_ENC_AA, 2, 8, true, 1>(P, grid, s); break;
    case 5: e = wide ? launch_mma<10, 5, 3, 8, true, 1>(P, grid, s) : launch_mma<5, 5, 4, 8, true>(P, grid, s); break;
    case 4: e = wide ? launch_mma<8, 4, 2, 8, true>(P, grid, s) : launch_mma<4, 4, 4, 8, true>(P, grid, s); break;
    case 2: e = wide ? launch_mma<4, 2, 4, 8, true>(P, grid, s) : launch_mma<2, 2, 4, 8, true>(P, grid, s); break;
    default:
      if (!wide) e = launch_mma<3, 3, 4, 8, true>(P, grid, s);
      else e = pack16 ? launch_mma<6, 3, 4, 8, true>(P, grid, s) : launch_mma<6, 3, 4, 8, false>(P, grid, s);
      break;
  }
  if (e != cudaSuccess) return mma_fail(ctx, SMAFA_E_CUDA, "scan_mma_kernel launch", e);
  return 2;
}
