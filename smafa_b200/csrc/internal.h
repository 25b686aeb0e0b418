// Internal host-side state behind the opaque handles of include/smafa_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kernels.h"

constexpr uint64_t DEFAULT_CAND_CAP = 32ull << 20;     // largest candidate workspace allocated up front (rows, 8 B keys)
// Default of smafa_ctx::mma_union: the largest union degree (windows per accumulator) dbs are packed for.
constexpr uint32_t SMAFA_MMA_UNION_DEFAULT = 3;
constexpr uint64_t MAX_AUTO_CAND_CAP = 256ull << 20;  // largest the workspace grows to on its own before a batch is split

struct smafa_ctx {
  int device = 0;
  int kernel = 0;
  int num_sms = 148;
  // AUTO picks the tcgen05 formulation when the db is eligible (L <= 63, valid codes): measured on
  // B200 at 6.3e12 cmp/s vs 4.3e12 for the POPC kernel with early exit (profiles/r01_*), and its
  // rate does not depend on how tight the bound is.  Tiny query batches stay on the POPC kernel.
  bool auto_prefers_mma = true;
  uint32_t mma_nsym = 3;          // MMA operand encoding (scan_mma.cu): 3 = +-1 features (default); SMAFA_MMA_NSYM=2/4/5: ablations
  uint32_t mma_union = 1;         // largest union degree dbs get operand images for (1..3 windows per accumulator, scan_mma.cu); SMAFA_MMA_UNION
  double union_verify_ns = 1.5;   // cost of one verified window in pick_union_degree's model; SMAFA_UNION_VERIFY_NS (calibration)
  bool db_group = false;          // SMAFA_DB_GROUP=1 (experimental, off): large nucleotide dbs are stored in similarity-grouped order (api.cu group_order)
  int mma_union_force = 0;        // SMAFA_MMA_UNION_FORCE=u: every tcgen05 scan that starts at need >= L/2 uses degree u whatever the sample says (tests)
  uint32_t mma_union_pick = 1;    // degree of the next mma_scan (set per scan by run_batch)
  uint32_t last_mma_k = 0;        // int8 contraction depth per WINDOW of the last tcgen05 scan (K / windows per row)
  int alphabet = 0;               // Alphabet of the dbs uploaded next and of smafa_cluster input (smafa_ctx_set_alphabet)
  bool disable_prepass = false;   // SMAFA_NO_PREPASS=1 (ablation)
  bool disable_guess = false;     // SMAFA_NO_GUESS=1: no optimistic first pass under a guessed bound (guess.cu)
  int force_guess = -1;           // SMAFA_FORCE_GUESS=g: use g as the guessed bound whatever the sample says (tests)
  int32_t *mma_dump = nullptr;    // debug hook (smafa_debug_mma_dump)
  int mma_bound0 = 0;             // initial bound of the batch being scanned (bias of the query operand)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {};
  std::string err;
  // candidate + finalize workspace
  uint64_t cand_cap_request = 0;
  uint64_t cand_needed = 0;      // candidate rows the last batch emitted (may exceed ws_cap: overflow)
  uint64_t ws_cap = 0;
  uint64_t *cand = nullptr;
  smafa::FinalizeWorkspace fw;
  smafa_hit *hits = nullptr;
  // per-batch query state
  int *bound = nullptr;           size_t bound_cap = 0;
  uint32_t *hist = nullptr;       size_t hist_cap = 0;
  uint64_t *q_ref = nullptr;      size_t q_ref_cap = 0;
  uint32_t *q_planes = nullptr;   size_t q_planes_cap = 0;
  uint8_t *q_onehot = nullptr;    size_t q_onehot_cap = 0;
  // optimistic first pass (guess.cu)
  uint32_t *per_query = nullptr;  size_t per_query_cap = 0;   // candidates per query after the first pass
  uint32_t *unfinished = nullptr; size_t unfinished_cap = 0;  // query numbers that need the second pass
  uint64_t *q_ref2 = nullptr;     size_t q_ref2_cap = 0;      // their words, compacted
  // scalars: [0] candidate count, [1] selected count, [2] scratch flag, [3] candidates kept after the first
  // pass, [4] unfinished queries, [8..8+guess_bins) sampled distance histogram
  static constexpr int N_SCALARS = 8 + 72;
  unsigned long long *d_scalars = nullptr;
  unsigned long long *h_scalars = nullptr;  // pinned mirror
  int *d_scratch_flag() { return reinterpret_cast<int *>(d_scalars + 2); }
};

struct smafa_db {
  smafa_ctx *ctx = nullptr;
  uint64_t D = 0, cap = 0;
  uint32_t L = 0, W = 0, row_words = 0;
  uint64_t subject_offset = 0;
  int alphabet = 0;           // smafa::Alphabet of `ref` (and of every query batch run against this db)
  bool generic_only = false;  // invalid codes or L > 64: reference-layout kernel only
  uint64_t *ref = nullptr;    // [cap][W]
  uint32_t *planes = nullptr; // [cap + pad][row_words]
  int *invalid_flag = nullptr;
  // tcgen05 operand (scan_mma.cu)
  uint8_t *onehot = nullptr;
  uint32_t mma_nsym = 3;      // encoding of `onehot` (mma_pick_encoding)
  uint64_t onehot_cap = 0;
  uint32_t *perm = nullptr;      // grouped dbs only: device row -> subject number (nullptr: rows are in subject order)
  std::vector<uint32_t> perm_host;
  // union-row images, one per degree of UNION_DEGREES: [tiles of 128*u windows][128 rows * 4 PB bytes]; degrees above
  // 3 only for grouped dbs (perm != nullptr)
  uint8_t *union_img[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t union_cap[5] = {0, 0, 0, 0, 0};     // in tiles
};

// Degrees (windows per accumulator) a union-row image can have, and the slot of a degree in smafa_db::union_img (-1: none)
constexpr uint32_t UNION_DEGREES[5] = {2, 3, 4, 8, 16};
inline int union_slot(uint32_t u) { return u == 2 ? 0 : u == 3 ? 1 : u == 4 ? 2 : u == 8 ? 3 : u == 16 ? 4 : -1; }

const char *smafa_global_error();
void smafa_set_global_error(const std::string &s);

// scan_mma.cu
bool mma_supported(const smafa_db *db);
uint32_t mma_pick_encoding(uint32_t want, uint32_t L, int alphabet);
int mma_db_reserve(smafa_ctx *ctx, smafa_db *db, uint64_t rows);
int mma_db_pack(smafa_ctx *ctx, smafa_db *db, uint64_t first, uint64_t n);
void mma_db_free(smafa_db *db);
// returns kernels launched (>= 0) or a negative smafa_status
int mma_scan(smafa_ctx *ctx, const smafa_db *db, smafa::ScanParams &p, cudaStream_t s, int32_t *dump = nullptr);
int mma_peak_probe(smafa_ctx *ctx, uint32_t mmas_per_cta, float *ms);
// probe.cu
int mma_rate_probe(smafa_ctx *ctx, int shape, uint32_t n_steps, double *ns_per_step);
int sparse_decode_probe(smafa_ctx *ctx, const uint8_t *a_comp, const uint32_t *meta, uint32_t n_steps, int meta_path, int32_t *out);
