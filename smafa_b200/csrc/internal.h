// Internal host-side state behind the opaque handles of include/smafa_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <string>
#include <vector>

#include "kernels.h"

constexpr uint64_t DEFAULT_CAND_CAP = 32ull << 20;     // largest candidate workspace allocated up front (rows, 8 B keys)
// Default of smafa_ctx::mma_union: the largest union degree (windows per accumulator) dbs are packed for.
constexpr uint32_t SMAFA_MMA_UNION_DEFAULT = 3;
constexpr uint64_t MAX_AUTO_CAND_CAP = 256ull << 20;  // largest the workspace grows to on its own before a batch is split

// The candidate exchange of a sharded query (sharded.cu): this shard's send block, the receive area of all blocks and
// the merge workspace (merge.cu), kept between calls; `cap` rows per block adapts to the workload.
struct SmafaExchange {
  uint64_t *block = nullptr;      // [2 + cap]: {rows, status, keys...}
  uint64_t *gathered = nullptr;   // [n_ranks][2 + cap]
  uint64_t cap = 0;
  uint32_t n_ranks = 0;
  uint64_t cap_hint = 0;          // capacity the next exchange starts with (twice the fullest block's last need)
  smafa::MergeWorkspace mw;
  smafa_hit *hits = nullptr;      // [n_ranks * cap] merged rows when the caller gave no device buffer
  unsigned long long *info_dev = nullptr;  // [0..3] merge results (merge.cu), [4] rows selected
  unsigned long long *info_host = nullptr; // pinned mirror
  cudaEvent_t ev[2] = {nullptr, nullptr};  // local part done / merge done (smafa_stats.exchange_ms)
};

struct SmafaComm;   // NCCL communicator of a one-process-per-GPU run (sharded.cu)
struct SmafaMulti;  // the per-device contexts behind a multi-device context (sharded.cu)

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
  double union_verify_ns = 0.03;  // cost of one verified window in pick_union_degree's model (profiles/r02_union_calib.log); SMAFA_UNION_VERIFY_NS
  bool db_group = true;           // large nucleotide dbs are stored in similarity-grouped order (api.cu group_order); SMAFA_DB_GROUP=0: plain order
  int mma_union_force = 0;        // SMAFA_MMA_UNION_FORCE=u: every tcgen05 scan that starts at need >= L/2 uses degree u whatever the sample says (tests)
  uint32_t mma_union_pick = 1;    // degree of the next mma_scan (set per scan by run_batch)
  uint32_t mma_union_used = 0;    // degree the last mma_scan really ran with (1 when the db lacks the picked image)
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
  // sort-free selection (finalize.cu launch_finalize_buckets)
  uint32_t *fz_counters = nullptr; size_t fz_counters_cap = 0;  // [per_query | fill | kept], Q + 1 each
  uint32_t *fz_starts = nullptr;   size_t fz_starts_cap = 0;    // [seg_start | kept_start], Q + 1 each
  uint32_t *fz_info = nullptr;     size_t fz_info_cap = 0;      // per candidate: place in its bucket's order, keep flag
  uint8_t *fz_temp = nullptr;      size_t fz_temp_cap = 0;      // CUB scan scratch
  bool disable_fast = false;      // SMAFA_NO_FAST_FINALIZE=1: always the sort (ablation, tests)
  uint32_t fast_skip = 0;         // batches that skip the speculative path after it failed (api.cu run_batch_fast)
  // optimistic first pass (guess.cu)
  uint32_t *per_query = nullptr;  size_t per_query_cap = 0;   // candidates per query after the first pass
  uint32_t *unfinished = nullptr; size_t unfinished_cap = 0;  // query numbers that need the second pass
  uint64_t *q_ref2 = nullptr;     size_t q_ref2_cap = 0;      // their words, compacted
  // scalars: [0] candidate count, [1] selected count, [2] scratch flag, [3] candidates kept after the first
  // pass, [4] unfinished queries, [5] largest per-query candidate count, [6] fast_ok (finalize.cu buckets),
  // [8..8+guess_bins) sampled distance histogram
  static constexpr int N_SCALARS = 8 + 72;
  unsigned long long *d_scalars = nullptr;
  unsigned long long *h_scalars = nullptr;  // pinned mirror
  int *d_scratch_flag() { return reinterpret_cast<int *>(d_scalars + 2); }
  // sharded queries (sharded.cu)
  SmafaExchange xchg;
  SmafaComm *comm = nullptr;    // one process per GPU: smafa_ctx_comm_init
  SmafaMulti *multi = nullptr;  // one process, several GPUs: smafa_ctx_create_multi (this context then owns no device state of its own)
};

struct smafa_db {
  smafa_ctx *ctx = nullptr;
  uint64_t D = 0, cap = 0;
  uint32_t L = 0, W = 0, row_words = 0;
  uint64_t subject_offset = 0;
  uint64_t global_rows = 0;   // rows of the whole db this shard belongs to (0: not a shard; smafa_db_upload_shard)
  std::vector<smafa_db *> shards;  // db of a multi-device context: one shard per device, nothing else
  int alphabet = 0;           // smafa::Alphabet of `ref` (and of every query batch run against this db)
  bool generic_only = false;  // invalid codes or L > 64: reference-layout kernel only
  uint64_t *ref = nullptr;    // [cap][W]
  uint32_t *planes = nullptr; // [cap + pad][row_words]
  int *invalid_flag = nullptr;
  // tcgen05 operand (scan_mma.cu)
  uint8_t *onehot = nullptr;
  uint32_t mma_nsym = 3;      // encoding of `onehot` (mma_pick_encoding)
  uint64_t onehot_cap = 0;
  uint32_t *perm = nullptr;      // device row -> subject number (nullptr: rows are in subject order); grouped and mapped dbs
  uint64_t perm_cap = 0;
  std::vector<uint32_t> perm_host;
  bool mapped = false;           // the caller gave the subject numbers (smafa_db_upload_mapped, shards of a multi-device context)
  bool grouped = false;          // rows are in similarity-grouped order (api.cu group_order, or stated by the caller): wide union rows
  // union-row images, one per degree of UNION_DEGREES: [tiles of 128*u windows][128 rows * 4 PB bytes]; degrees above
  // 3 only for grouped dbs (perm != nullptr)
  uint8_t *union_img[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t union_cap[5] = {0, 0, 0, 0, 0};     // in tiles
  // last sampled choice of pick_union_degree (api.cu), reused for the next scans of similar batches at the same need
  mutable int pick_need = -1;
  mutable uint32_t pick_degree = 0, pick_age = 0, pick_nq = 0;
  mutable uint64_t pick_rows = 0;
};

// Degrees (windows per accumulator) a union-row image can have, and the slot of a degree in smafa_db::union_img (-1: none)
constexpr uint32_t UNION_DEGREES[5] = {2, 3, 4, 8, 16};
inline int union_slot(uint32_t u) { return u == 2 ? 0 : u == 3 ? 1 : u == 4 ? 2 : u == 8 ? 3 : u == 16 ? 4 : -1; }

// ---- api.cu internals shared with sharded.cu ----
struct QueryPlan {
  int mode;        // smafa::ScanMode
  uint32_t k_scan; // MODE_KTH tightening parameter
  uint32_t k_fin;  // finalize: keep rows <= k-th smallest distance (UINT32_MAX: keep all)
  int bound0;      // initial bound
};
// Where a batch's answer goes.  Default: smafa_hit rows in ctx->hits, count on the host.  hits != nullptr: rows are
// written there instead (capacity hits_cap).  block != nullptr: no rows at all -- the selected keys are appended to a
// send block (merge.cu) without any host read-back; the row count then stays on the device.
struct BatchOut {
  smafa_hit *hits = nullptr;
  uint64_t hits_cap = 0;
  uint64_t *block = nullptr;
  uint64_t block_cap = 0;
};
// Checks the arguments of a query in the reference's order and fills the plan; D_total = rows of the whole db.
int validate_query_plan(smafa_ctx *ctx, uint64_t D_total, uint32_t db_len, uint64_t Q, uint32_t q_len, int64_t m, int64_t k, QueryPlan *plan);
// Runs queries [0, n) (device words, on db's device) against db, splitting on candidate overflow.  Per finished batch:
// sink(first query of the batch, queries in it, rows) -- rows = UINT64_MAX in block mode (count on the device).
typedef std::function<int(uint64_t, uint64_t, uint64_t)> BatchSink;
int run_query_range(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_dev, uint64_t n, uint64_t q_base, const QueryPlan &plan,
                    cudaStream_t s, smafa_stats *st, const BatchOut &out, const BatchSink &sink);
int smafa_fail(smafa_ctx *ctx, int code, const char *fmt, ...);
// subjects != nullptr: row r is reported as subject subjects[r] (+ subject_offset); grouped: the rows are in similarity-
// grouped order (wide union rows are packed); try_group: let the library find such an order (api.cu group_order)
int db_upload_rows(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint64_t subject_offset, const uint32_t *subjects,
                   bool grouped, bool try_group, smafa_db **out);
int db_append_rows(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n, uint64_t first_subject);
int ensure_query_words(smafa_ctx *ctx, size_t words);  // ctx->q_ref

// ---- sharded.cu: entry points api.cu forwards to for a multi-device context ----
void multi_destroy(smafa_ctx *ctx);
int multi_set(smafa_ctx *ctx, int what, int64_t value);  // 0 kernel, 1 alphabet, 2 candidate capacity
int multi_db_upload(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint64_t subject_offset, smafa_db **out);
int multi_db_append(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n);
void multi_db_free(smafa_db *db);
int multi_distances(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len, uint16_t *out);
int multi_query(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len, int64_t m, int64_t k,
                smafa_hit **hits, uint64_t *n_hits, smafa_stats *stats);
smafa_ctx *multi_first(smafa_ctx *ctx);  // the context of the first device (cluster and the debug hooks run there)
void exchange_free(smafa_ctx *ctx);
void comm_free(smafa_ctx *ctx);

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
