// C ABI of libsmafa_b200.so: context / db management, query batching, overflow retry.
// See include/smafa_b200.h for the contract and the reference code each entry point replaces.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "kernels.h"

using namespace smafa;

static thread_local std::string g_err;  // ctx-less failures

const char *smafa_global_error() { return g_err.c_str(); }
void smafa_set_global_error(const std::string &s) { g_err = s; }

static int fail(smafa_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  g_err = buf;
  return code;
}

int smafa_fail(smafa_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? SMAFA_E_OOM : SMAFA_E_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
  } while (0)

extern "C" int smafa_abi_version(void) { return SMAFA_B200_ABI_VERSION; }

extern "C" const char *smafa_status_name(int s) {
  switch (s) {
    case SMAFA_OK: return "SMAFA_OK";
    case SMAFA_E_LENGTH_MISMATCH: return "SMAFA_E_LENGTH_MISMATCH";
    case SMAFA_E_EMPTY_DB: return "SMAFA_E_EMPTY_DB";
    case SMAFA_E_BAD_K: return "SMAFA_E_BAD_K";
    case SMAFA_E_LIMIT_NEEDS_K: return "SMAFA_E_LIMIT_NEEDS_K";
    case SMAFA_E_CUDA: return "SMAFA_E_CUDA";
    case SMAFA_E_OOM: return "SMAFA_E_OOM";
    case SMAFA_E_INVALID: return "SMAFA_E_INVALID";
    case SMAFA_E_UNSUPPORTED: return "SMAFA_E_UNSUPPORTED";
    case SMAFA_E_NCCL: return "SMAFA_E_NCCL";
    case SMAFA_E_PEER: return "SMAFA_E_PEER";
    case SMAFA_E_IO: return "SMAFA_E_IO";
    case SMAFA_E_PANIC: return "SMAFA_E_PANIC";
    default: return "SMAFA_E_?";
  }
}

extern "C" const char *smafa_last_error(const smafa_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" void smafa_free(void *p) { free(p); }

// ------------------------------------------------------------------------------- context

extern "C" int smafa_ctx_create(smafa_ctx **out, int device, int kernel) {
  smafa_ctx *ctx = nullptr;
  if (!out) return fail(nullptr, SMAFA_E_INVALID, "smafa_ctx_create: null out pointer");
  *out = nullptr;
  // SMAFA_TIMING=1: where context creation spends its 1-3 s (stderr)
  const bool timing = getenv("SMAFA_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[smafa timing] ctx: %-23s %9.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  lap("cudaGetDeviceCount");
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, SMAFA_E_CUDA,
                "no CUDA device available (%s); libsmafa_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, SMAFA_E_INVALID, "device %d out of range (0..%d)", device, n - 1);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, SMAFA_E_CUDA, "device %d is sm_%d%d; libsmafa_b200 holds sm_100a code only", device,
                prop.major, prop.minor);
  lap("cudaGetDeviceProperties");
  CU(cudaSetDevice(device));
  CU(cudaFree(nullptr));  // forces the primary context into existence here
  lap("primary context");
  ctx = new smafa_ctx();
  ctx->device = device;
  ctx->kernel = kernel;
  ctx->num_sms = prop.multiProcessorCount;
  if (const char *e = getenv("SMAFA_NO_PREPASS")) ctx->disable_prepass = e[0] == '1';
  if (const char *e = getenv("SMAFA_NO_GUESS")) ctx->disable_guess = e[0] == '1';
  if (const char *e = getenv("SMAFA_NO_FAST_FINALIZE")) ctx->disable_fast = e[0] == '1';
  if (const char *e = getenv("SMAFA_FORCE_GUESS")) ctx->force_guess = atoi(e);
  if (const char *e = getenv("SMAFA_MMA_NSYM")) ctx->mma_nsym = (e[0] >= '2' && e[0] <= '5') ? (uint32_t)(e[0] - '0') : 3;
  // union rows are on unless an ablation pins the operand encoding (SMAFA_MMA_NSYM) or asks for single rows
  ctx->mma_union = SMAFA_MMA_UNION_DEFAULT;
  if (getenv("SMAFA_MMA_NSYM") != nullptr) ctx->mma_union = 1;
  if (const char *e = getenv("SMAFA_MMA_UNION")) ctx->mma_union = (e[0] >= '1' && e[0] <= '3') ? (uint32_t)(e[0] - '0') : 1u;
  if (const char *e = getenv("SMAFA_MMA_UNION_FORCE")) ctx->mma_union_force = atoi(e);
  if (const char *e = getenv("SMAFA_DB_GROUP")) ctx->db_group = e[0] == '1';
  if (const char *e = getenv("SMAFA_UNION_VERIFY_NS")) { const double v = atof(e); if (v > 0) ctx->union_verify_ns = v; }
  cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  const char *what = "cudaStreamCreate";
  for (auto &ev : ctx->ev)
    if (e2 == cudaSuccess) { e2 = cudaEventCreate(&ev); what = "cudaEventCreate"; }
  if (e2 == cudaSuccess) {
    e2 = cudaHostAlloc((void **)&ctx->h_scalars, smafa_ctx::N_SCALARS * sizeof(unsigned long long), cudaHostAllocDefault);
    what = "cudaHostAlloc(scalars)";
  }
  if (e2 == cudaSuccess) {
    e2 = cudaMalloc((void **)&ctx->d_scalars, smafa_ctx::N_SCALARS * sizeof(unsigned long long));
    what = "cudaMalloc(scalars)";
  }
  if (e2 == cudaSuccess) {
    e2 = cudaMemsetAsync(ctx->d_scalars, 0, smafa_ctx::N_SCALARS * sizeof(unsigned long long), ctx->stream);
    what = "cudaMemset(scalars)";
  }
  if (e2 != cudaSuccess) {
    const int code = e2 == cudaErrorMemoryAllocation ? SMAFA_E_OOM : SMAFA_E_CUDA;
    smafa_ctx_destroy(ctx);  // releases whatever was created (the CUDA destroy/free calls accept null handles)
    return fail(nullptr, code, "smafa_ctx_create: %s: %s", what, cudaGetErrorString(e2));
  }
  lap("stream, events, scalars");
  *out = ctx;
  return SMAFA_OK;
}

static void free_workspace(smafa_ctx *ctx) {
  cudaFree(ctx->cand); ctx->cand = nullptr;
  cudaFree(ctx->fw.keys_sorted); cudaFree(ctx->fw.keys_sel); cudaFree(ctx->fw.cub_temp);
  cudaFree(ctx->fw.seg_start); cudaFree(ctx->fw.seg_end);
  ctx->fw = FinalizeWorkspace();
  cudaFree(ctx->hits); ctx->hits = nullptr;
  ctx->ws_cap = 0;
}

extern "C" void smafa_ctx_destroy(smafa_ctx *ctx) {
  if (!ctx) return;
  if (ctx->multi) { multi_destroy(ctx); delete ctx; return; }
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_free(ctx);
  exchange_free(ctx);
  free_workspace(ctx);
  cudaFree(ctx->bound); cudaFree(ctx->hist); cudaFree(ctx->q_ref); cudaFree(ctx->q_planes);
  cudaFree(ctx->q_onehot);
  cudaFree(ctx->per_query); cudaFree(ctx->unfinished); cudaFree(ctx->q_ref2);
  cudaFree(ctx->fz_counters); cudaFree(ctx->fz_starts); cudaFree(ctx->fz_info); cudaFree(ctx->fz_temp);
  cudaFree(ctx->d_scalars);
  cudaFreeHost(ctx->h_scalars);
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int smafa_ctx_set_kernel(smafa_ctx *ctx, int kernel) {
  if (!ctx || kernel < 0 || kernel > 2) return fail(ctx, SMAFA_E_INVALID, "bad kernel selector %d", kernel);
  if (ctx->multi) return multi_set(ctx, 0, kernel);
  ctx->kernel = kernel;
  return SMAFA_OK;
}

extern "C" int smafa_ctx_set_alphabet(smafa_ctx *ctx, int alphabet) {
  if (!ctx || (alphabet != SMAFA_ALPHABET_NUCLEOTIDE && alphabet != SMAFA_ALPHABET_PROTEIN))
    return fail(ctx, SMAFA_E_INVALID, "bad alphabet selector %d", alphabet);
  ctx->alphabet = alphabet;
  if (ctx->multi) return multi_set(ctx, 1, alphabet);
  return SMAFA_OK;
}

extern "C" int smafa_ctx_set_candidate_capacity(smafa_ctx *ctx, uint64_t rows) {
  if (!ctx) return SMAFA_E_INVALID;
  if (ctx->multi) return multi_set(ctx, 2, (int64_t)rows);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_workspace(ctx);
  ctx->cand_cap_request = rows;
  return SMAFA_OK;
}

// Candidate / finalize workspace for `rows` candidate rows.
static int ensure_workspace(smafa_ctx *ctx, uint64_t rows) {
  if (ctx->ws_cap >= rows) return SMAFA_OK;
  cudaStreamSynchronize(ctx->stream);
  free_workspace(ctx);
  ctx->fw.cub_temp_bytes = finalize_temp_bytes(rows);
  cudaError_t e = cudaMalloc((void **)&ctx->cand, rows * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->fw.keys_sorted, rows * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->fw.keys_sel, rows * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->fw.seg_start, MAX_BATCH_QUERIES * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->fw.seg_end, MAX_BATCH_QUERIES * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->fw.cub_temp, ctx->fw.cub_temp_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->hits, rows * sizeof(smafa_hit));
  if (e != cudaSuccess) {
    free_workspace(ctx);  // nothing half-allocated stays behind
    return fail(ctx, e == cudaErrorMemoryAllocation ? SMAFA_E_OOM : SMAFA_E_CUDA, "candidate workspace for %llu rows: %s",
                (unsigned long long)rows, cudaGetErrorString(e));
  }
  ctx->fw.n_selected = ctx->d_scalars + 1;
  ctx->fw.cap = rows;
  ctx->ws_cap = rows;
  return SMAFA_OK;
}

template <class T>
static int ensure_buf(smafa_ctx *ctx, T *&ptr, size_t &cap, size_t need) {
  if (cap >= need) return SMAFA_OK;
  // geometric growth: cluster's batches grow a little every round, and a cudaFree + cudaMalloc pair per round
  // (milliseconds each) dominated small runs
  need = std::max(need, cap * 2);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  CU(cudaMalloc((void **)&ptr, need * sizeof(T)));
  cap = need;
  return SMAFA_OK;
}

// ------------------------------------------------------------------------------- db

static uint32_t words_for(uint32_t L) { return (L + 11) / 12; }
static uint32_t row_words_for(uint32_t L) { return L <= 32 ? 4 : (L <= 64 ? 8 : 0); }

static int db_reserve(smafa_ctx *ctx, smafa_db *db, uint64_t rows) {
  if (rows <= db->cap) return SMAFA_OK;
  uint64_t ncap = std::max<uint64_t>(rows, db->cap * 2);
  ncap = (ncap + 255) / 256 * 256;
  uint64_t *nref = nullptr;
  uint32_t *npl = nullptr;
  cudaError_t e = cudaMalloc((void **)&nref, std::max<uint64_t>(1, ncap * db->W) * sizeof(uint64_t));
  if (e == cudaSuccess && db->row_words) {
    // one extra tile of rows so the POPC kernel's register prefetch never leaves the allocation
    size_t bytes = (ncap + 512) * db->row_words * sizeof(uint32_t);
    e = cudaMalloc((void **)&npl, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(npl, 0, bytes, ctx->stream);
  }
  if (e == cudaSuccess && db->D) {
    e = cudaMemcpyAsync(nref, db->ref, db->D * db->W * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && db->row_words)
      e = cudaMemcpyAsync(npl, db->planes, db->D * db->row_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {  // the db keeps its old arrays
    cudaFree(nref);
    cudaFree(npl);
    return fail(ctx, e == cudaErrorMemoryAllocation ? SMAFA_E_OOM : SMAFA_E_CUDA, "db storage for %llu rows: %s",
                (unsigned long long)ncap, cudaGetErrorString(e));
  }
  cudaFree(db->ref);
  cudaFree(db->planes);
  db->ref = nref;
  db->planes = npl;
  db->cap = ncap;
  int rc = mma_db_reserve(ctx, db, ncap);
  return rc;
}

static int db_add_rows(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n) {
  if (n == 0) return SMAFA_OK;
  if (db->D + n >= (1ull << 32)) return fail(ctx, SMAFA_E_UNSUPPORTED, "db larger than 2^32-1 windows");
  int rc = db_reserve(ctx, db, db->D + n);
  if (rc) return rc;
  CU(cudaMemcpyAsync(db->ref + db->D * db->W, enc, n * db->W * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  if (db->row_words)
    launch_pack_planes(db->ref + db->D * db->W, (uint32_t)n, db->W, db->L, db->row_words, db->alphabet,
                       db->planes + db->D * db->row_words, db->invalid_flag, ctx->stream);
  rc = mma_db_pack(ctx, db, db->D, n);
  if (rc) return rc;
  int flag = 0;
  CU(cudaMemcpyAsync(&flag, db->invalid_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  if (flag) db->generic_only = true;
  db->D += n;
  return SMAFA_OK;
}

constexpr int RC_TOO_MANY_CENTROIDS = 2;  // internal (cluster_impl)
static int cluster_impl(smafa_ctx *ctx, const uint64_t *enc, uint64_t n, uint32_t L, uint32_t t, uint32_t *centroid_of,
                        uint64_t *n_centroids, uint64_t *n_comparisons, smafa_stats *stats, uint64_t max_centroids);

// The layout half of group_order as a pure host function (no GPU; tests/test_host_abi.py drives it): centroid_of[i] =
// index of the window that founded window i's cluster (itself for a founder; founders precede their members, as in
// smafa_cluster's output), perm_out[row] = window stored at `row`.
extern "C" int smafa_group_layout(const uint32_t *centroid_of, uint64_t D, uint32_t *perm_out) {
  if (D && (!centroid_of || !perm_out)) return SMAFA_E_INVALID;
  const uint32_t *cof = centroid_of;
  for (uint64_t i = 0; i < D; ++i)
    if (cof[i] > i || cof[cof[i]] != cof[i]) return SMAFA_E_INVALID;  // a founder precedes its members and founded itself
  // cluster number (founding order) of every window, cluster sizes
  std::vector<uint32_t> cid(D), size;
  for (uint64_t i = 0; i < D; ++i) {
    if (cof[i] == i) { cid[i] = (uint32_t)size.size(); size.push_back(0); }
    else cid[i] = cid[cof[i]];  // a centroid precedes its members
    size[cid[i]]++;
  }
  // layout: order of the clusters.  A cluster of s windows ends s mod 16 rows past a row boundary when it starts on one,
  // so what is packed are the remainders: clusters whose size is a multiple of 16 first (they keep the alignment), then
  // chains of clusters whose remainders fill a row -- largest remainder first, then the best fit for what is left of
  // the row -- so that a chain ends on a row boundary again.  Clusters that fit nowhere simply continue the chain.
  constexpr uint32_t ROW = 16;
  std::vector<std::vector<uint32_t>> rem(ROW);  // rem[r] = clusters with size % 16 == r, founding order (used from the back)
  std::vector<uint32_t> order;
  order.reserve(size.size());
  for (uint32_t c = (uint32_t)size.size(); c-- > 0;) rem[size[c] % ROW].push_back(c);
  while (!rem[0].empty()) { order.push_back(rem[0].back()); rem[0].pop_back(); }
  uint32_t pos = 0;  // fill of the current row
  for (;;) {
    uint32_t r = ROW - 1;
    if (pos != 0) {  // best fit for the rest of the row, else whatever is largest
      r = ROW - pos;
      while (r > 0 && rem[r].empty()) --r;
      if (r == 0) r = ROW - 1;
    }
    while (r > 0 && rem[r].empty()) --r;
    if (r == 0) break;
    order.push_back(rem[r].back());
    rem[r].pop_back();
    pos = (pos + r) % ROW;
  }
  std::vector<uint64_t> start(size.size() + 1, 0);
  for (uint32_t c : order) start[c] = 0;
  uint64_t at = 0;
  for (uint32_t c : order) { start[c] = at; at += size[c]; }
  for (uint64_t i = 0; i < D; ++i) perm_out[start[cid[i]]++] = (uint32_t)i;
  return SMAFA_OK;
}

// Similarity-grouped db order (DESIGN.md section 3b "Grouped rows").  Union rows (scan_mma.cu) filter several windows
// with one accumulator, and how many they can hold is set by how often the union of a row's windows matches an
// unrelated query -- hardly more often than one window when the windows of a row are near-copies of each other.
// SingleM-style dbs are highly redundant, so the db is stored on the device in an order that puts similar windows
// next to each other: perm[row] = input index of the window stored at `row`.  Everything on the device then works on
// that order; candidates are mapped back to subject numbers before finalize (launch_remap_subjects), so results do
// not change.
//   1. clusters = the library's own greedy clustering (src/cluster.rs semantics, cluster_impl) at 2L/5: far below
//      the distance of unrelated windows (3L/4 +- a few), wide enough to keep a family of near-copies together;
//   2. clusters are laid out so that they start at multiples of 16 rows where possible -- a cluster that straddles two
//      16-wide operand rows makes both of them pass for its queries (see "layout" below).  Members keep their input
//      order inside a cluster.
// Leaves perm empty when the db has too little structure to gain from it (more clusters than half its windows).
static int group_order(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, std::vector<uint32_t> &perm, uint64_t *n_clusters) {
  perm.clear();
  if (n_clusters) *n_clusters = 0;
  std::vector<uint32_t> cof(D);
  uint64_t nc = 0;
  const bool saved = ctx->db_group;
  ctx->db_group = false;  // the greedy's own dbs are plain
  uint32_t t_group = 2 * L / 5;
  if (const char *e = getenv("SMAFA_DB_GROUP_T")) t_group = (uint32_t)atoi(e);
  int rc = cluster_impl(ctx, enc, D, L, t_group, cof.data(), &nc, nullptr, nullptr, D / 2);
  ctx->db_group = saved;
  if (getenv("SMAFA_UNION_DEBUG"))
    fprintf(stderr, "[smafa group] %llu windows, threshold %u: %llu clusters (rc %d)\n", (unsigned long long)D, t_group, (unsigned long long)nc, rc);
  if (rc == RC_TOO_MANY_CENTROIDS) return SMAFA_OK;
  if (rc) return rc;
  if (n_clusters) *n_clusters = nc;
  perm.resize(D);
  return smafa_group_layout(cof.data(), D, perm.data());
}

extern "C" int smafa_group_order(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint32_t *perm_out, uint64_t *n_clusters) {
  if (!ctx || (D && (!enc || !perm_out)) || L == 0) return fail(ctx, SMAFA_E_INVALID, "smafa_group_order: bad argument");
  if (ctx->multi) ctx = multi_first(ctx);
  if (D >= (1ull << 32)) return fail(ctx, SMAFA_E_UNSUPPORTED, "db larger than 2^32-1 windows");
  CU(cudaSetDevice(ctx->device));
  std::vector<uint32_t> perm;
  int rc = SMAFA_OK;
  uint64_t nc = 0;
  if (ctx->db_group && ctx->alphabet == ALPHA_NUC && L <= 63 && D >= 65536) rc = group_order(ctx, enc, D, L, perm, &nc);
  if (rc) return rc;
  if (n_clusters) *n_clusters = perm.empty() ? 0 : nc;
  if (perm.empty())
    for (uint64_t i = 0; i < D; ++i) perm_out[i] = (uint32_t)i;
  else
    memcpy(perm_out, perm.data(), D * sizeof(uint32_t));
  return SMAFA_OK;
}

// subjects != nullptr: row r is reported as subject subjects[r] (+ subject_offset); grouped: the caller states that the
// rows are in similarity-grouped order (wide union rows are packed).  try_group: let the library find such an order.
int db_upload_rows(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint64_t subject_offset, const uint32_t *subjects,
                   bool grouped, bool try_group, smafa_db **out) {
  if (D > 0 && (!enc || L == 0)) return fail(ctx, SMAFA_E_INVALID, "smafa_db_upload: D > 0 needs enc and L > 0");
  if (L > MAX_WINDOW_LEN) return fail(ctx, SMAFA_E_UNSUPPORTED, "window length %u > %u", L, MAX_WINDOW_LEN);
  CU(cudaSetDevice(ctx->device));
  smafa_db *db = new smafa_db();
  db->ctx = ctx;
  db->L = L;
  db->W = words_for(L);
  db->row_words = row_words_for(L);
  db->generic_only = (db->row_words == 0);
  db->subject_offset = subject_offset;
  db->alphabet = ctx->alphabet;
  db->mma_nsym = mma_pick_encoding(ctx->mma_nsym, L, db->alphabet);
  cudaError_t e = cudaMalloc((void **)&db->invalid_flag, sizeof(int));
  if (e != cudaSuccess) { delete db; return fail(ctx, SMAFA_E_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
  cudaMemsetAsync(db->invalid_flag, 0, sizeof(int), ctx->stream);
  std::vector<uint64_t> reordered;
  const bool eligible = db->alphabet == ALPHA_NUC && !db->generic_only && L <= 63;
  if (subjects) {
    db->perm_host.assign(subjects, subjects + D);
    db->mapped = true;
    db->grouped = grouped && eligible;
  } else if (try_group && eligible && D >= 65536) {
    int rc = group_order(ctx, enc, D, L, db->perm_host, nullptr);
    if (rc) { smafa_db_free(db); return rc; }
    if (!db->perm_host.empty()) {
      reordered.resize(D * db->W);
      const uint32_t W = db->W;
      const uint32_t *pm = db->perm_host.data();
      for (uint64_t r = 0; r < D; ++r) memcpy(reordered.data() + r * W, enc + (uint64_t)pm[r] * W, W * sizeof(uint64_t));
      enc = reordered.data();
      db->grouped = true;
    }
  }
  if (!db->perm_host.empty()) {
    db->perm_cap = std::max<uint64_t>(D, 256);
    e = cudaMalloc((void **)&db->perm, db->perm_cap * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(db->perm, db->perm_host.data(), D * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { smafa_db_free(db); return fail(ctx, SMAFA_E_OOM, "db row -> subject table: %s", cudaGetErrorString(e)); }
  }
  int rc = db_add_rows(ctx, db, enc, D);
  if (rc) { smafa_db_free(db); return rc; }
  *out = db;
  return SMAFA_OK;
}

extern "C" int smafa_db_upload(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint64_t subject_offset,
                               smafa_db **out) {
  if (!ctx || !out) return fail(ctx, SMAFA_E_INVALID, "smafa_db_upload: null argument");
  *out = nullptr;
  if (ctx->multi) return multi_db_upload(ctx, enc, D, L, subject_offset, out);
  return db_upload_rows(ctx, enc, D, L, subject_offset, nullptr, false, ctx->db_group, out);
}

// Rows in a caller-chosen order with explicit subject numbers (a shard of a db that was grouped as a whole).
extern "C" int smafa_db_upload_mapped(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, const uint32_t *subjects,
                                      uint64_t D_total, int grouped, smafa_db **out) {
  if (!ctx || !out || (D && !subjects)) return fail(ctx, SMAFA_E_INVALID, "smafa_db_upload_mapped: null argument");
  *out = nullptr;
  if (ctx->multi) return fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_db_upload_mapped: a multi-device context shards its dbs itself");
  if (D > D_total || D_total >= (1ull << 32)) return fail(ctx, SMAFA_E_INVALID, "smafa_db_upload_mapped: %llu rows of a db of %llu",
                                                            (unsigned long long)D, (unsigned long long)D_total);
  for (uint64_t r = 0; r < D; ++r)
    if (subjects[r] >= D_total) return fail(ctx, SMAFA_E_INVALID, "smafa_db_upload_mapped: subject %u of row %llu is outside the db", subjects[r], (unsigned long long)r);
  int rc = db_upload_rows(ctx, enc, D, L, 0, subjects, grouped != 0, false, out);
  if (rc == SMAFA_OK) (*out)->global_rows = D_total;
  return rc;
}

// A shard of a row-sharded db: rows [subject_offset, subject_offset + D) of a db of D_total rows.
extern "C" int smafa_db_upload_shard(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint64_t subject_offset,
                                     uint64_t D_total, smafa_db **out) {
  if (ctx && ctx->multi) return fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_db_upload_shard: a multi-device context shards its dbs itself");
  if (subject_offset + D > D_total) return fail(ctx, SMAFA_E_INVALID, "smafa_db_upload_shard: rows [%llu, %llu) exceed the db's %llu rows",
                                                (unsigned long long)subject_offset, (unsigned long long)(subject_offset + D), (unsigned long long)D_total);
  if (D_total >= (1ull << 32)) return fail(ctx, SMAFA_E_UNSUPPORTED, "db larger than 2^32-1 windows");
  int rc = smafa_db_upload(ctx, enc, D, L, subject_offset, out);
  if (rc == SMAFA_OK) (*out)->global_rows = D_total;
  return rc;
}

extern "C" int smafa_db_append(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n) {
  if (!ctx || !db || (n && !enc)) return fail(ctx, SMAFA_E_INVALID, "smafa_db_append: null argument");
  if (ctx->multi) return multi_db_append(ctx, db, enc, n);
  if (db->global_rows) return fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_db_append: the db is one shard of a larger db");
  return db_append_rows(ctx, db, enc, n, db->D);
}

// first_subject: subject number of the first new window when the db has a row -> subject table (a db stored in grouped
// order: the new windows go behind it under the next subject numbers)
int db_append_rows(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n, uint64_t first_subject) {
  CU(cudaSetDevice(ctx->device));
  if (db->perm != nullptr && n) {
    const uint64_t D = db->D;
    if (D + n > db->perm_cap) {
      const uint64_t ncap = std::max<uint64_t>(D + n, db->perm_cap * 2);
      uint32_t *np = nullptr;
      CU(cudaMalloc((void **)&np, ncap * sizeof(uint32_t)));
      cudaError_t e = cudaMemcpyAsync(np, db->perm, D * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) { cudaFree(np); return fail(ctx, SMAFA_E_CUDA, "smafa_db_append: %s", cudaGetErrorString(e)); }
      cudaFree(db->perm);
      db->perm = np;
      db->perm_cap = ncap;
    }
    db->perm_host.resize(D + n);
    for (uint64_t i = 0; i < n; ++i) db->perm_host[D + i] = (uint32_t)(first_subject + i);
    CU(cudaMemcpyAsync(db->perm + D, db->perm_host.data() + D, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  return db_add_rows(ctx, db, enc, n);
}

extern "C" uint64_t smafa_db_size(const smafa_db *db) { return db ? db->D : 0; }
extern "C" uint32_t smafa_db_window_len(const smafa_db *db) { return db ? db->L : 0; }

extern "C" void smafa_db_free(smafa_db *db) {
  if (!db) return;
  if (db->ctx && db->ctx->multi) { multi_db_free(db); return; }
  if (db->ctx) { cudaSetDevice(db->ctx->device); cudaStreamSynchronize(db->ctx->stream); }
  cudaFree(db->ref);
  cudaFree(db->planes);
  cudaFree(db->invalid_flag);
  cudaFree(db->perm);
  mma_db_free(db);
  delete db;
}

// ------------------------------------------------------------------------------- distances

extern "C" int smafa_distances(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len,
                               uint16_t *out) {
  if (!ctx || !db) return fail(ctx, SMAFA_E_INVALID, "smafa_distances: null argument");
  if (ctx->multi) return multi_distances(ctx, db, q_enc, Q, q_len, out);
  if (Q == 0 || db->D == 0) return SMAFA_OK;
  if (!q_enc || !out) return fail(ctx, SMAFA_E_INVALID, "smafa_distances: null buffer");
  if (q_len != db->L)
    return fail(ctx, SMAFA_E_LENGTH_MISMATCH,
                "Cannot compute distances between seq of length %u and windows of lengths %u", q_len, db->L);
  CU(cudaSetDevice(ctx->device));
  const uint64_t D = db->D;
  uint64_t qb = std::max<uint64_t>(1, std::min<uint64_t>(Q, (1ull << 29) / D));  // <= 1 GiB of u16 per batch
  uint64_t *dq = nullptr;
  uint16_t *dout = nullptr;
  CU(cudaMalloc((void **)&dq, qb * db->W * sizeof(uint64_t)));
  cudaError_t e = cudaMalloc((void **)&dout, qb * D * sizeof(uint16_t));
  if (e != cudaSuccess) { cudaFree(dq); return fail(ctx, SMAFA_E_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
  int rc = SMAFA_OK;
  for (uint64_t q0 = 0; q0 < Q && rc == SMAFA_OK; q0 += qb) {
    uint64_t nq = std::min(qb, Q - q0);
    cudaMemcpyAsync(dq, q_enc + q0 * db->W, nq * db->W * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    launch_distances(dq, (uint32_t)nq, db->ref, (uint32_t)D, db->W, db->alphabet, dout, ctx->stream);
    cudaMemcpyAsync(out + q0 * D, dout, nq * D * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 != cudaSuccess) rc = fail(ctx, SMAFA_E_CUDA, "smafa_distances: %s", cudaGetErrorString(e2));
    if (rc == SMAFA_OK && !db->perm_host.empty() && !db->mapped) {  // grouped db: column r of the device result is subject perm[r]
      std::vector<uint16_t> row(D);
      for (uint64_t q = q0; q < q0 + nq; ++q) {
        memcpy(row.data(), out + q * D, D * sizeof(uint16_t));
        for (uint64_t r = 0; r < D; ++r) out[q * D + db->perm_host[r]] = row[r];
      }
    }
  }
  cudaFree(dq);
  cudaFree(dout);
  return rc;
}

// ------------------------------------------------------------------------------- query

namespace {

constexpr int RC_OVERFLOW = 1;  // internal: candidate buffer too small for this batch

}  // namespace

int validate_query_plan(smafa_ctx *ctx, uint64_t D, uint32_t L, uint64_t Q, uint32_t q_len, int64_t m, int64_t k, QueryPlan *plan) {
  if (Q == 0) return SMAFA_OK;
  // order of the reference's checks for the first record: length (src/lib.rs:72-79), then the
  // selection's unwrap()/underflow panics (src/lib.rs:253-255,298)
  if (D > 0 && q_len != L)
    return fail(ctx, SMAFA_E_LENGTH_MISMATCH,
                "Cannot compute distances between seq of length %u and windows of lengths %u", q_len, L);
  const bool mode_b = (k >= 0 && k != 1);  // src/lib.rs:224
  if (mode_b && k == 0 && D > 0) return fail(ctx, SMAFA_E_BAD_K, "attempt to subtract with overflow");
  if (D == 0) return fail(ctx, SMAFA_E_EMPTY_DB, "called `Option::unwrap()` on a `None` value");
  plan->bound0 = (int)L;
  if (m >= 0 && m < (int64_t)L) plan->bound0 = (int)m;
  if (!mode_b) {
    plan->mode = MODE_MIN;
    plan->k_scan = 1;
    plan->k_fin = 1;
  } else if ((uint64_t)k >= D) {  // cutoff = largest distance (src/lib.rs:254): nothing to tighten
    plan->mode = MODE_FIXED;
    plan->k_scan = 0;
    plan->k_fin = UINT32_MAX;
  } else {
    plan->mode = MODE_KTH;
    plan->k_scan = (uint32_t)k;
    plan->k_fin = (uint32_t)k;
  }
  return SMAFA_OK;
}

static int validate_query(smafa_ctx *ctx, const smafa_db *db, uint64_t Q, uint32_t q_len, int64_t m, int64_t k, QueryPlan *plan) {
  return validate_query_plan(ctx, db->D, db->L, Q, q_len, m, k, plan);
}

static uint32_t pick_chunk(const smafa_ctx *ctx, uint64_t D, uint64_t n_qtiles, uint32_t tile) {
  // aim at >= 8 blocks per SM-slot so the tail wave is short; never below one tile
  uint64_t target_blocks = (uint64_t)ctx->num_sms * 3 * 8;
  uint64_t chunks = std::max<uint64_t>(1, target_blocks / std::max<uint64_t>(1, n_qtiles));
  uint64_t chunk = (D + chunks - 1) / chunks;
  chunk = std::min<uint64_t>(std::max<uint64_t>(chunk, tile), 16384);
  return (uint32_t)((chunk + tile - 1) / tile * tile);
}

// Bound for the optimistic first pass (guess.cu): the largest distance g < bound0 at which the batch as a whole is
// expected to emit no more than a budget of candidates, from the distance distribution of a strided sample of
// (query, window) pairs.  Returns -1 when no such pass is worth running.
static int guess_bound(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_ref_dev, uint32_t Qb, const QueryPlan &plan,
                       uint32_t need, cudaStream_t s, int *launches, int *rc_out) {
  *rc_out = SMAFA_OK;
  if (ctx->force_guess >= 0) return ctx->force_guess < plan.bound0 ? ctx->force_guess : -1;
  const double pairs = (double)Qb * (double)db->D;
  if (pairs < 2e9 || db->D < 65536 || Qb < 1024) return -1;  // small batches: the pre-pass alone is cheaper
  const uint32_t q_stride = (Qb + 8191) / 8192;
  const uint32_t n_d = (uint32_t)std::min<uint64_t>(2048, db->D);
  const uint32_t d_stride = (uint32_t)(db->D / n_d);
  unsigned long long *ghist = ctx->d_scalars + 8;
  const int bins = guess_bins();
  *launches += launch_sample_hist(q_ref_dev, Qb, q_stride, db->ref, d_stride, n_d, db->W, db->alphabet, ghist, s);
  cudaError_t e = cudaMemcpyAsync(ctx->h_scalars + 8, ghist, bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { *rc_out = fail(ctx, SMAFA_E_CUDA, "sample histogram: %s", cudaGetErrorString(e)); return -1; }
  const double n_q = (double)((Qb + q_stride - 1) / q_stride);
  const double scale = pairs / (n_q * (double)n_d);  // real pairs per sampled pair
  // Candidates cost ~1.5 ns of GPU time each against ~0.1 ns per comparison: 1e-5 of the pairs keeps them below
  // ~15 % of the scan; the rows every query is owed anyway (need per query) come on top.
  const double budget = 1e-5 * pairs + 2.0 * (double)need * (double)Qb;
  double cum = 0;
  int g = -1;
  for (int t = 0; t < bins && t < plan.bound0; ++t) {
    cum += (double)ctx->h_scalars[8 + t];
    if (cum * scale > budget) break;
    g = t;
  }
  return g;
}

// How many db windows share one accumulator in the next tcgen05 scan (scan_mma.cu, UPR).  A union row of u windows
// costs 1/u of the tensor work and of the accumulator drain per comparison, but every row that passes its (looser)
// filter sends u windows to the exact re-check.  Per comparison:  cost(u) = t_u + c_v * f_u.
//   t_u = time per comparison of a scan of degree u that verifies nothing: one 128-row tile costs its MMAs (68 ns per
//         k-step per SM at M128 x N256 x K32) or the ~400 ns accumulator drain, whichever is longer, at the ~90 % the
//         kernel sustains, and holds 128 u windows x 256 queries.  L = 60: 9.7e-5, 6.2e-5, 4.2e-5 ns (measured).
//   f_u = fraction of the (query, row) pairs of degree u that pass at need = L - b0 -- measured on a strided sample of
//         THIS batch against THIS db (union_sample_kernel), so related windows, skewed base composition or a loose
//         bound simply show up as a larger f_u and a smaller degree.
//   c_v = 0.03 ns per window sent to the re-check.  Calibrated on forced degrees over six db shapes x several bounds
//         (profiles/r02_union_calib.log, scripts/union_calib.py): the measured cost per verified window is 0.01-0.03 ns
//         while the verifier warps keep up and rises towards 0.1 ns once rows pass by the percent (ring-full waits);
//         with 0.03 the model picks the fastest forced degree in all 17 (db, bound) cases of that log.  (Round 1 used
//         1.5 ns, the cost of an EMITTED candidate -- those are the same at every degree and cancel out.)
// The sample costs a kernel and a host read-back, so its verdict is kept with the db and reused for the next scans at
// the same need with a similar batch size (re-sampled every 32 scans and whenever the db has grown by a quarter).
// Batches too small to sample use the thresholds this model gives for unrelated uniform windows: a position passes a
// union of u windows with probability p_u = 1 - (3/4)^u, a row passes when Binomial(L, p_u) >= need; the degree is
// worth it while f_u stays below (t_{u-1} - t_u) / c_v ~ 1e-3, i.e. need >= L p_u + 3.1 sqrt(L p_u (1 - p_u)).
static double union_scan_ns(const smafa_ctx *ctx, uint32_t L, uint32_t u) {
  // k-steps per tile: union rows are one-hot over 4 bases (K = 256, or 128 for L <= 31); single rows use the db's own
  // encoding -- +-1 features by default (K = 192 / 96)
  const uint32_t ksteps = u > 1 ? (L <= 31 ? 4 : 8) : (L <= 30 ? 3 : 6);
  const double tile_ns = std::max(68.0 * ksteps, 400.0) / 0.9;
  return tile_ns / (128.0 * u * 256.0 * (double)ctx->num_sms);
}

static uint32_t pick_union_degree(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_dev, uint32_t nq, int b0, cudaStream_t s,
                                  int *launches, int *rc_out) {
  *rc_out = SMAFA_OK;
  uint32_t max_u = 1;
  for (uint32_t u : UNION_DEGREES)
    if (db->union_img[union_slot(u)] != nullptr) max_u = u;
  const int need = (int)db->L - b0;
  if (max_u == 1 || 2 * need < (int)db->L) return 1;  // a union of two windows differs from a query in < L/2 positions far too often
  if (ctx->mma_union_force >= 1) return std::min<uint32_t>((uint32_t)ctx->mma_union_force, max_u);
  const double pairs = (double)nq * (double)db->D;
  const bool debug = getenv("SMAFA_UNION_DEBUG") != nullptr;
  if (pairs >= 2e9 && db->D >= 65536 && nq >= 1024 && !debug && db->pick_degree && db->pick_need == need && db->pick_age < 32 &&
      nq >= db->pick_nq / 2 && nq <= db->pick_nq * 2 && db->D >= db->pick_rows && db->D <= db->pick_rows + db->pick_rows / 4) {
    db->pick_age++;
    return std::min(db->pick_degree, max_u);
  }
  if (db->grouped && max_u > 3) {
    // Grouped db (experimental): rows up to 16 windows wide; always sampled (such a db has >= 65536 windows).
    const uint32_t q_stride = (nq + 4095) / 4096;
    const uint32_t n_d = (uint32_t)std::min<uint64_t>(1024, db->D);
    const uint32_t d_stride = (uint32_t)(db->D / n_d);
    unsigned long long *counts = ctx->d_scalars + 8;
    *launches += launch_union_sample_wide(q_dev, nq, q_stride, db->ref, (uint32_t)db->D, d_stride, n_d, db->W, need, counts, s);
    cudaError_t e = cudaMemcpyAsync(ctx->h_scalars + 8, counts, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { *rc_out = fail(ctx, SMAFA_E_CUDA, "union sample: %s", cudaGetErrorString(e)); return 1; }
    const double n_s = std::max<double>(1.0, (double)ctx->h_scalars[14]);
    static const uint32_t deg[6] = {1, 2, 3, 4, 8, 16};
    if (debug)
      fprintf(stderr, "[smafa union] need %d: passing fraction of rows of degree 1/2/3/4/8/16 = %.3e / %.3e / %.3e / %.3e / %.3e / %.3e (%.0f samples)\n",
              need, ctx->h_scalars[8] / n_s, ctx->h_scalars[9] / n_s, ctx->h_scalars[10] / n_s, ctx->h_scalars[11] / n_s,
              ctx->h_scalars[12] / n_s, ctx->h_scalars[13] / n_s, n_s);
    uint32_t best = 1;
    double best_cost = 0;
    for (int i = 0; i < 6 && deg[i] <= max_u; ++i) {
      const double cost = union_scan_ns(ctx, db->L, deg[i]) + ctx->union_verify_ns * (double)ctx->h_scalars[8 + i] / n_s;
      if (i == 0 || cost < best_cost) { best = deg[i]; best_cost = cost; }
    }
    db->pick_need = need; db->pick_degree = best; db->pick_age = 0; db->pick_nq = nq; db->pick_rows = db->D;
    return best;
  }
  if (pairs < 2e9 || db->D < 65536 || nq < 1024) {
    const double Ld = (double)db->L;
    auto worth = [&](double p) { return (double)need >= Ld * p + 3.1 * sqrt(Ld * p * (1.0 - p)); };
    if (max_u >= 3 && worth(37.0 / 64.0)) return 3;
    return worth(7.0 / 16.0) ? 2 : 1;
  }
  // 4096 x 1024 = 4.2 M sampled pairs: one count is 2.4e-7 of the rows, i.e. 7e-9 ns in the cost below (the t_u are
  // 2e-5 apart); ncu: 81 us per launch at 4096 x 2048 (profiles/r01_launches_v9_summary.txt)
  const uint32_t q_stride = (nq + 4095) / 4096;
  const uint32_t n_d = (uint32_t)std::min<uint64_t>(1024, db->D);
  const uint32_t d_stride = (uint32_t)(db->D / n_d);
  unsigned long long *counts = ctx->d_scalars + 8;  // shared with the guess histogram: both are read back before reuse
  *launches += launch_union_sample(q_dev, nq, q_stride, db->ref, (uint32_t)db->D, d_stride, n_d, db->W, need, counts, s);
  cudaError_t e = cudaMemcpyAsync(ctx->h_scalars + 8, counts, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { *rc_out = fail(ctx, SMAFA_E_CUDA, "union sample: %s", cudaGetErrorString(e)); return 1; }
  const double n_s = std::max<double>(1.0, (double)ctx->h_scalars[11]);
  if (debug)
    fprintf(stderr, "[smafa union] need %d: passing fraction of rows of degree 1/2/3 = %.3e / %.3e / %.3e (%.0f samples)\n", need,
            ctx->h_scalars[8] / n_s, ctx->h_scalars[9] / n_s, ctx->h_scalars[10] / n_s, n_s);
  uint32_t best = 1;
  double best_cost = 0;
  for (uint32_t u = 1; u <= max_u; ++u) {
    const double cost = union_scan_ns(ctx, db->L, u) + ctx->union_verify_ns * (double)ctx->h_scalars[8 + u - 1] / n_s;
    if (u == 1 || cost < best_cost) { best = u; best_cost = cost; }
  }
  db->pick_need = need; db->pick_degree = best; db->pick_age = 0; db->pick_nq = nq; db->pick_rows = db->D;
  return best;
}

constexpr int RC_SLOW_PATH = 3;  // internal: the speculative batch did not hold, run it the ordinary way

// The common batch, enqueued in one go: a tcgen05 scan under a useful starting bound, the sort-free selection of
// finalize.cu ("buckets") and the output conversion, every kernel taking the candidate count from device memory --
// then ONE read-back.  Nothing waits on the host between the kernels, so their launch latencies hide behind the scan.
// The speculation -- candidates fit the workspace, no query holds more than BUCKET_MAX of them, the query words are
// valid codes -- is checked on the device (fast_ok); when it fails nothing was written and the caller runs the batch
// again through run_batch's ordinary path (sort-based selection, overflow handling, generic kernel).
static int run_batch_fast(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_ref_dev, uint32_t Qb, uint32_t q_base,
                          const QueryPlan &plan, uint64_t *n_rows, cudaStream_t s, smafa_stats *st, const BatchOut &out) {
  int rc;
  const uint32_t hist_stride = (db->L + 1 + 3) / 4 * 4;
  const size_t q1 = (size_t)Qb + 1;
  if ((rc = ensure_buf(ctx, ctx->fz_counters, ctx->fz_counters_cap, 3 * q1))) return rc;
  if ((rc = ensure_buf(ctx, ctx->fz_starts, ctx->fz_starts_cap, 2 * q1))) return rc;
  if ((rc = ensure_buf(ctx, ctx->fz_info, ctx->fz_info_cap, (size_t)ctx->ws_cap))) return rc;
  if ((rc = ensure_buf(ctx, ctx->fz_temp, ctx->fz_temp_cap, bucket_temp_bytes(Qb)))) return rc;
  unsigned long long *cand_count = ctx->d_scalars + 0, *fast_ok = ctx->d_scalars + 6;
  uint32_t *max_seg = reinterpret_cast<uint32_t *>(ctx->d_scalars + 5);
  int *q_invalid = ctx->d_scratch_flag();
  CU(cudaMemsetAsync(ctx->d_scalars, 0, 8 * sizeof(unsigned long long), s));
  CU(cudaMemsetAsync(ctx->fz_counters, 0, 3 * q1 * sizeof(uint32_t), s));
  int launches = 2;
  cudaEventRecord(ctx->ev[0], s);
  launch_init_bound(ctx->bound, Qb + 512, plan.bound0, s);
  if (plan.mode == MODE_KTH) CU(cudaMemsetAsync(ctx->hist, 0, (size_t)Qb * hist_stride * sizeof(uint32_t), s));
  if (db->alphabet == ALPHA_NUC) {
    launch_check_codes(q_ref_dev, Qb, db->W, db->L, q_invalid, s);
  } else {  // protein: validity comes with the class planes
    if ((rc = ensure_buf(ctx, ctx->q_planes, ctx->q_planes_cap, ((size_t)Qb + 256) * db->row_words))) return rc;
    launch_pack_planes(q_ref_dev, Qb, db->W, db->L, db->row_words, db->alphabet, ctx->q_planes, q_invalid, s);
  }
  launches += 2;
  ScanParams p{};
  p.d_planes = db->planes;
  p.d_ref = db->ref;
  p.D = (uint32_t)db->D;
  p.d_begin = 0;
  p.d_end = (uint32_t)db->D;
  p.W = db->W;
  p.L = db->L;
  p.alphabet = db->alphabet;
  p.mode = plan.mode;
  p.k = plan.k_scan;
  p.bound = ctx->bound;
  p.hist = ctx->hist;
  p.hist_stride = hist_stride;
  p.cand = ctx->cand;
  p.cand_count = cand_count;
  p.cand_cap = ctx->ws_cap;
  p.per_query = ctx->fz_counters;
  p.q_ref = q_ref_dev;
  p.Q = Qb;
  ctx->mma_bound0 = plan.bound0;
  ctx->mma_union_pick = pick_union_degree(ctx, db, q_ref_dev, Qb, plan.bound0, s, &launches, &rc);
  if (rc) return rc;
  int l = mma_scan(ctx, db, p, s, nullptr);
  if (l < 0) return l;
  launches += l;
  cudaEventRecord(ctx->ev[1], s);
  // a rough upper bound of the candidate count sizes the grids (grid-stride kernels: any value is correct)
  const uint64_t n_hint = std::min<uint64_t>(ctx->ws_cap, std::max<uint64_t>(4ull * Qb, ctx->cand_needed));
  launches += launch_finalize_buckets(ctx->fw, ctx->cand, cand_count, ctx->ws_cap, max_seg, q_invalid, fast_ok, Qb, plan.k_fin,
                                      ctx->fz_counters, ctx->fz_starts, ctx->fz_info, ctx->fz_temp, ctx->fz_temp_cap, db->perm, n_hint, s);
  if (out.block)
    launches += launch_block_append(ctx->fw.keys_sel, ctx->fw.n_selected, n_hint, q_base, db->subject_offset, out.block, out.block_cap, s);
  else
    launches += launch_keys_to_hits(ctx->fw, n_hint, q_base, db->subject_offset, out.hits ? out.hits : ctx->hits,
                                    out.hits ? out.hits_cap : ctx->ws_cap, ctx->h_scalars + 1, s);
  CU(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaGetLastError());
  const uint64_t n_cand = ctx->h_scalars[0];
  if (st) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    st->scan_ms += ms;
    st->kernel_launches += launches;
  }
  if (ctx->h_scalars[6] == 0) {  // nothing was written downstream (n_selected = 0): the ordinary path takes over
    if (st) st->retries++;
    ctx->fast_skip = 8;  // and keeps the next batches too: a workload that breaks the speculation once tends to do it again
    return RC_SLOW_PATH;
  }
  ctx->cand_needed = n_cand;
  if (st) {
    st->kernel_used = SMAFA_KERNEL_MMA;
    st->guess_bound = -1;
    st->union_degree = ctx->mma_union_used;
    st->candidates += n_cand;
  }
  *n_rows = out.block ? UINT64_MAX : ctx->h_scalars[1];
  return SMAFA_OK;
}

// One batch (<= 2^20 queries, words already on the device).  Leaves *n_rows rows in ctx->hits.
static int run_batch(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_ref_dev, uint32_t Qb, uint32_t q_base,
                     const QueryPlan &plan, uint64_t *n_rows, cudaStream_t s, smafa_stats *st, const BatchOut &out = BatchOut()) {
  int rc;
  // Candidate rows: a tight bound emits about a row per query, so the workspace starts at 32 rows per query
  // (1 Mi..32 Mi; a 32 Mi-row workspace is 1.4 GB of cudaMalloc, most of a small run's start-up) and run_query_range
  // grows it to what a batch turned out to need.
  uint64_t want_cap = ctx->cand_cap_request;
  if (!want_cap) {
    want_cap = 1ull << 20;
    while (want_cap < 32ull * Qb && want_cap < DEFAULT_CAND_CAP) want_cap <<= 1;
  }
  if ((rc = ensure_workspace(ctx, std::max<uint64_t>(want_cap, ctx->ws_cap)))) return rc;
  if ((rc = ensure_buf(ctx, ctx->bound, ctx->bound_cap, (size_t)Qb + 512))) return rc;  // padded: tile-wide vector loads
  const uint32_t hist_stride = (db->L + 1 + 3) / 4 * 4;  // rows 16-byte aligned for the vector scan in emit_candidate
  if (plan.mode == MODE_KTH && (rc = ensure_buf(ctx, ctx->hist, ctx->hist_cap, (size_t)Qb * hist_stride))) return rc;
  unsigned long long *cand_count = ctx->d_scalars + 0;

  ScanParams p{};
  p.d_planes = db->planes;
  p.d_ref = db->ref;
  p.D = (uint32_t)db->D;
  p.d_begin = 0;
  p.d_end = (uint32_t)db->D;
  p.W = db->W;
  p.L = db->L;
  p.alphabet = db->alphabet;
  p.mode = plan.mode;
  p.k = plan.k_scan;
  p.bound = ctx->bound;
  p.hist = ctx->hist;
  p.hist_stride = hist_stride;
  p.cand = ctx->cand;
  p.cand_count = cand_count;
  p.cand_cap = ctx->ws_cap;

  int kernel = ctx->kernel;
  if (db->generic_only) kernel = -1;
  else if (kernel == SMAFA_KERNEL_AUTO) kernel = mma_supported(db) && ctx->auto_prefers_mma && Qb >= 64 ? SMAFA_KERNEL_MMA : SMAFA_KERNEL_POPC;
  if (kernel == SMAFA_KERNEL_MMA && !mma_supported(db)) kernel = SMAFA_KERNEL_POPC;
  // a fixed bound that admits everything would send every accumulator down the MMA slow path
  if (kernel == SMAFA_KERNEL_MMA && plan.mode == MODE_FIXED && plan.bound0 >= (int)db->L) kernel = SMAFA_KERNEL_POPC;

  int *q_invalid = ctx->d_scratch_flag();
  int launches = 0;
  // One scan of `nq` queries (words at q_dev) under the starting bound b0; candidates are appended at *cand_count.
  auto scan_pass = [&](const uint64_t *q_dev, uint32_t nq, int b0, bool prepass, int kern) -> int {
    int r;
    launch_init_bound(ctx->bound, nq + 512, b0, s);
    if (plan.mode == MODE_KTH) CU(cudaMemsetAsync(ctx->hist, 0, (size_t)nq * hist_stride * sizeof(uint32_t), s));
    launches += 1;
    p.q_ref = q_dev;
    p.Q = nq;
    if (kern == -1) {
      launches += launch_scan_generic(p, pick_chunk(ctx, db->D, (nq + 127) / 128, 256), s);
      return SMAFA_OK;
    }
    if (kern == SMAFA_KERNEL_MMA && nq < 64) kern = SMAFA_KERNEL_POPC;
    const uint32_t n_tiles = (uint32_t)((db->D + 255) / 256);
    const bool want_prepass = prepass && n_tiles >= 32;
    if (kern == SMAFA_KERNEL_MMA && !want_prepass && db->alphabet == ALPHA_NUC) {
      // the tcgen05 scan reads no bit planes: only the validity check of the query words is needed
      launch_check_codes(q_dev, nq, db->W, db->L, q_invalid, s);
      p.q_planes = nullptr;
    } else {
      if ((r = ensure_buf(ctx, ctx->q_planes, ctx->q_planes_cap, ((size_t)nq + 256) * db->row_words))) return r;
      launch_pack_planes(q_dev, nq, db->W, db->L, db->row_words, db->alphabet, ctx->q_planes, q_invalid, s);
      p.q_planes = ctx->q_planes;
    }
    launches += 1;
    // No useful starting bound (no or a loose --max-divergence): estimate one on a strided db
    // sample first, otherwise the first tiles of the scan would emit nearly every pair.
    if (want_prepass) {
      uint32_t sample_tiles = std::min<uint32_t>(std::max<uint32_t>(n_tiles / 32, 16), 224);
      launches += launch_bound_prepass(p, std::max<uint32_t>(1, n_tiles / sample_tiles), s);
    }
    if (kern == SMAFA_KERNEL_MMA) {
      ctx->mma_bound0 = b0;
      ctx->mma_union_pick = pick_union_degree(ctx, db, q_dev, nq, b0, s, &launches, &r);
      if (r) return r;
      int l = mma_scan(ctx, db, p, s, ctx->mma_dump);
      if (l < 0) return l;
      launches += l;
    } else {
      const bool early = b0 * 4 <= (int)db->L;
      uint32_t r4 = nq >= 148u * 256u * 2u ? 4 : 1;
      launches += launch_scan_popc(p, early, pick_chunk(ctx, db->D, (nq + 256 * r4 - 1) / (256 * r4), popc_tile_rows()), s);
    }
    return SMAFA_OK;
  };

  const bool loose = plan.mode != MODE_FIXED && plan.bound0 * 3 > (int)db->L && !ctx->disable_prepass;
  const uint32_t need = plan.mode == MODE_KTH ? plan.k_scan : 1;  // rows that finish a query under a guessed bound
  if (ctx->fast_skip > 0) {
    ctx->fast_skip--;
  } else if (kernel == SMAFA_KERNEL_MMA && Qb >= 64 && !loose && !ctx->disable_fast && ctx->mma_dump == nullptr && db->L <= 63) {
    rc = run_batch_fast(ctx, db, q_ref_dev, Qb, q_base, plan, n_rows, s, st, out);
    if (rc != RC_SLOW_PATH) return rc;
  }
  uint64_t n_cand = 0;
  for (;;) {
    CU(cudaMemsetAsync(q_invalid, 0, sizeof(int), s));
    CU(cudaMemsetAsync(cand_count, 0, sizeof(unsigned long long), s));
    launches = 2;
    cudaEventRecord(ctx->ev[0], s);
    int g = -1;
    if (kernel != -1 && loose && !ctx->disable_guess && db->L <= 64) {
      g = guess_bound(ctx, db, q_ref_dev, Qb, plan, need, s, &launches, &rc);
      if (rc) return rc;
    }
    if (g >= 0) {
      // ---- first pass under the guessed bound; second pass for the queries it did not finish (guess.cu) ----
      if ((rc = scan_pass(q_ref_dev, Qb, g, false, kernel))) return rc;
      CU(cudaMemcpyAsync(ctx->h_scalars + 0, cand_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      const uint64_t n1 = ctx->h_scalars[0];
      if (n1 > ctx->ws_cap) { n_cand = n1; break; }  // overflow: the caller retries with a smaller batch
      if ((rc = ensure_buf(ctx, ctx->per_query, ctx->per_query_cap, (size_t)Qb))) return rc;
      if ((rc = ensure_buf(ctx, ctx->unfinished, ctx->unfinished_cap, (size_t)Qb))) return rc;
      uint32_t *n_list = reinterpret_cast<uint32_t *>(ctx->d_scalars + 4);
      launch_count_per_query(ctx->cand, n1, Qb, ctx->per_query, s);
      launch_list_unfinished(ctx->per_query, Qb, need, ctx->unfinished, n_list, s);
      launches += 2;
      CU(cudaMemcpyAsync(ctx->h_scalars + 4, n_list, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      const uint32_t n_open = *reinterpret_cast<uint32_t *>(ctx->h_scalars + 4);
      if (st) st->rescanned += n_open;
      if (n_open) {
        if ((rc = ensure_buf(ctx, ctx->q_ref2, ctx->q_ref2_cap, (size_t)n_open * db->W))) return rc;
        launch_gather_queries(q_ref_dev, ctx->unfinished, n_open, db->W, ctx->q_ref2, s);
        // the unfinished queries' first-pass candidates would come again in the second pass: drop them
        unsigned long long *n_keep = ctx->d_scalars + 3;
        launch_keep_finished(ctx->cand, n1, ctx->per_query, need, ctx->fw.keys_sorted, n_keep, s);
        CU(cudaMemcpyAsync(ctx->cand, ctx->fw.keys_sorted, n1 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
        CU(cudaMemcpyAsync(cand_count, n_keep, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        launches += 3;
        if ((rc = scan_pass(ctx->q_ref2, n_open, plan.bound0, true, kernel))) return rc;
        launch_remap_queries(ctx->cand, n_keep, cand_count, ctx->ws_cap, ctx->unfinished, s);
        launches += 1;
      }
    } else {
      if ((rc = scan_pass(q_ref_dev, Qb, plan.bound0, loose, kernel))) return rc;
    }
    cudaEventRecord(ctx->ev[1], s);
    CU(cudaMemcpyAsync(ctx->h_scalars + 0, cand_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ctx->h_scalars + 2, ctx->d_scalars + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    CU(cudaGetLastError());
    n_cand = ctx->h_scalars[0];
    if (st) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
      st->scan_ms += ms;
      st->kernel_used = kernel == -1 ? 0 : (uint32_t)kernel;
      st->kernel_launches += launches;
      st->guess_bound = g;
      if (kernel == SMAFA_KERNEL_MMA) st->union_degree = ctx->mma_union_used;
    }
    if (kernel != -1 && (int)(ctx->h_scalars[2] & 0xffffffffu) != 0) {
      // a query holds words that are not valid one-hot codes: redo the batch on the reference
      // word layout, which reproduces popcount(a^b)/2 for arbitrary words
      kernel = -1;
      continue;
    }
    break;
  }
  if (st) st->candidates += n_cand;
  ctx->cand_needed = n_cand;
  if (n_cand > ctx->ws_cap) return RC_OVERFLOW;
  if (db->perm != nullptr) launch_remap_subjects(ctx->cand, n_cand, db->perm, s);  // grouped db: rows -> subject numbers
  int fl = launch_finalize_select(ctx->fw, ctx->cand, n_cand, Qb, plan.k_fin, s);
  if (out.block) {
    // sharded query: the selected keys join this shard's send block; their number stays on the device (merge.cu)
    fl += launch_block_append(ctx->fw.keys_sel, ctx->fw.n_selected, n_cand, q_base, db->subject_offset, out.block, out.block_cap, s);
    CU(cudaGetLastError());
    if (st) st->kernel_launches += fl;
    *n_rows = UINT64_MAX;
    return SMAFA_OK;
  }
  fl += launch_keys_to_hits(ctx->fw, n_cand, q_base, db->subject_offset, out.hits ? out.hits : ctx->hits,
                            out.hits ? out.hits_cap : ctx->ws_cap, ctx->h_scalars + 1, s);
  CU(cudaStreamSynchronize(s));
  CU(cudaGetLastError());
  if (st) st->kernel_launches += fl;
  *n_rows = ctx->h_scalars[1];
  return SMAFA_OK;
}

// Runs queries [0, n) of the device-resident words at q_dev, splitting on candidate overflow.  `sink` gets each
// finished batch (see BatchOut for where its rows are).  Reported query numbers start at q_base.
int run_query_range(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_dev, uint64_t n, uint64_t q_base, const QueryPlan &plan,
                    cudaStream_t s, smafa_stats *st, const BatchOut &out, const BatchSink &sink) {
  uint64_t done = 0;
  while (done < n) {
    uint64_t nb = std::min<uint64_t>(n - done, MAX_BATCH_QUERIES);
    for (;;) {
      uint64_t rows = 0;
      int rc = run_batch(ctx, db, q_dev + done * db->W, (uint32_t)nb, (uint32_t)(q_base + done), plan, &rows, s, st, out);
      if (rc == RC_OVERFLOW) {
        if (st) st->retries++;
        // the counter kept counting past the capacity, so the need is known (a lower bound of it when the batch
        // stopped after its first pass): grow to it while that stays reasonable, else split the batch
        if (!ctx->cand_cap_request && ctx->cand_needed <= MAX_AUTO_CAND_CAP) {
          int r2 = ensure_workspace(ctx, std::max<uint64_t>(ctx->ws_cap * 2, ctx->cand_needed + ctx->cand_needed / 4));
          if (r2) return r2;
          continue;
        }
        if (nb == 1) {  // a single query can emit at most D rows
          int r2 = ensure_workspace(ctx, std::max<uint64_t>(ctx->ws_cap * 2, db->D + 1024));
          if (r2) return r2;
          ctx->cand_cap_request = 0;
        } else {
          nb = (nb + 1) / 2;
        }
        continue;
      }
      if (rc) return rc;
      rc = sink(done, nb, rows);
      if (rc) return rc;
      break;
    }
    done += nb;
  }
  return SMAFA_OK;
}

int ensure_query_words(smafa_ctx *ctx, size_t words) { return ensure_buf(ctx, ctx->q_ref, ctx->q_ref_cap, words); }

extern "C" int smafa_query(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len,
                           int64_t m, int64_t k, smafa_hit **hits, uint64_t *n_hits, smafa_stats *stats) {
  if (!ctx || !db || !hits || !n_hits) return fail(ctx, SMAFA_E_INVALID, "smafa_query: null argument");
  *hits = nullptr;
  *n_hits = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  if (ctx->multi) return multi_query(ctx, db, q_enc, Q, q_len, m, k, hits, n_hits, stats);
  QueryPlan plan{};
  int rc = validate_query(ctx, db, Q, q_len, m, k, &plan);
  if (rc || Q == 0) return rc;
  if (!q_enc) return fail(ctx, SMAFA_E_INVALID, "smafa_query: null query buffer");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  cudaEventRecord(ctx->ev[2], s);
  // the answer grows in the buffer the caller will own: rows go device -> that buffer, nothing in between
  smafa_hit *all = nullptr;
  uint64_t n_all = 0, cap_all = 0;
  const uint64_t slab = MAX_BATCH_QUERIES;
  for (uint64_t q0 = 0; q0 < Q && rc == SMAFA_OK; q0 += slab) {
    uint64_t nq = std::min(slab, Q - q0);
    if ((rc = ensure_buf(ctx, ctx->q_ref, ctx->q_ref_cap, nq * db->W))) break;
    cudaError_t e = cudaMemcpyAsync(ctx->q_ref, q_enc + q0 * db->W, nq * db->W * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { rc = fail(ctx, SMAFA_E_CUDA, "H2D of queries: %s", cudaGetErrorString(e)); break; }
    rc = run_query_range(ctx, db, ctx->q_ref, nq, q0, plan, s, stats, BatchOut(), [&](uint64_t, uint64_t, uint64_t rows) -> int {
      if (n_all + rows > cap_all) {
        cap_all = std::max<uint64_t>({n_all + rows, cap_all * 2, 1024});
        smafa_hit *grown = (smafa_hit *)realloc(all, cap_all * sizeof(smafa_hit));
        if (!grown) return fail(ctx, SMAFA_E_OOM, "realloc of %llu hits failed", (unsigned long long)cap_all);
        all = grown;
      }
      if (rows) {
        cudaError_t e2 = cudaMemcpyAsync(all + n_all, ctx->hits, rows * sizeof(smafa_hit), cudaMemcpyDeviceToHost, s);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(s);
        if (e2 != cudaSuccess) return fail(ctx, SMAFA_E_CUDA, "D2H of hits: %s", cudaGetErrorString(e2));
      }
      n_all += rows;
      return SMAFA_OK;
    });
  }
  if (rc) { free(all); return rc; }
  cudaEventRecord(ctx->ev[3], s);
  cudaEventSynchronize(ctx->ev[3]);
  if (stats) {
    cudaEventElapsedTime(&stats->total_ms, ctx->ev[2], ctx->ev[3]);
    stats->pairs = Q * db->D;
  }
  if (!all && !(all = (smafa_hit *)malloc(sizeof(smafa_hit)))) return fail(ctx, SMAFA_E_OOM, "malloc failed");
  *hits = all;
  *n_hits = n_all;
  return SMAFA_OK;
}

extern "C" int smafa_query_dev(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc_dev, uint64_t Q, uint32_t q_len,
                               int64_t m, int64_t k, smafa_hit *hits_dev, uint64_t hits_capacity, uint64_t *n_hits,
                               void *stream, smafa_stats *stats) {
  if (!ctx || !db || !n_hits) return fail(ctx, SMAFA_E_INVALID, "smafa_query_dev: null argument");
  *n_hits = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  if (ctx->multi) return fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_query_dev: device pointers belong to one device; use smafa_query on a multi-device context");
  QueryPlan plan{};
  int rc = validate_query(ctx, db, Q, q_len, m, k, &plan);
  if (rc || Q == 0) return rc;
  if (!q_enc_dev) return fail(ctx, SMAFA_E_INVALID, "smafa_query_dev: null query buffer");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;  // NULL = the legacy default stream, like any CUDA API
  cudaEventRecord(ctx->ev[2], s);
  uint64_t total = 0;
  // Every batch writes its rows behind the previous one's, straight into the caller's buffer; a batch that does not
  // fit writes nothing (keys_to_hits_kernel) and only reports its row count.
  for (uint64_t done = 0; done < Q && rc == SMAFA_OK;) {
    const uint64_t nq = std::min<uint64_t>(Q - done, MAX_BATCH_QUERIES);
    BatchOut out;
    const uint64_t room = total < hits_capacity && hits_dev ? hits_capacity - total : 0;
    out.hits = room ? hits_dev + total : ctx->hits;
    out.hits_cap = room;  // 0: count only
    rc = run_query_range(ctx, db, q_enc_dev + done * db->W, nq, done, plan, s, stats, out, [&](uint64_t, uint64_t, uint64_t rows) -> int {
      // run_query_range may split a slab into several batches: the next one must land behind this one
      total += rows;
      const uint64_t left = total < hits_capacity && hits_dev ? hits_capacity - total : 0;
      out.hits = left ? hits_dev + total : ctx->hits;
      out.hits_cap = left;
      return SMAFA_OK;
    });
    done += nq;
  }
  if (rc) return rc;
  cudaEventRecord(ctx->ev[3], s);
  cudaEventSynchronize(ctx->ev[3]);
  if (stats) {
    cudaEventElapsedTime(&stats->total_ms, ctx->ev[2], ctx->ev[3]);
    stats->pairs = Q * db->D;
  }
  *n_hits = total;
  if (total > hits_capacity)
    return fail(ctx, SMAFA_E_OOM, "hits buffer too small: %llu rows needed, capacity %llu", (unsigned long long)total,
                (unsigned long long)hits_capacity);
  return SMAFA_OK;
}

extern "C" int smafa_merge_dev(smafa_ctx *ctx, smafa_hit *cands_dev, uint64_t n, int64_t m, int64_t k, uint64_t *n_out,
                               void *stream) {
  if (!ctx || !n_out) return fail(ctx, SMAFA_E_INVALID, "smafa_merge_dev: null argument");
  *n_out = 0;
  if (ctx->multi) return fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_merge_dev: device pointers belong to one device");
  if (n == 0) return SMAFA_OK;
  if (!cands_dev) return fail(ctx, SMAFA_E_INVALID, "smafa_merge_dev: null buffer");
  (void)m;  // every shard already applied --max-divergence
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = (cudaStream_t)stream;  // NULL = the legacy default stream, like any CUDA API
  int rc = ensure_workspace(ctx, std::max<uint64_t>({n, ctx->ws_cap, (uint64_t)1 << 20}));
  if (rc) return rc;
  const bool mode_b = (k >= 0 && k != 1);
  uint32_t k_fin = mode_b ? (uint32_t)std::min<int64_t>(k, UINT32_MAX) : 1;
  if (mode_b && k == 0) return fail(ctx, SMAFA_E_BAD_K, "attempt to subtract with overflow");
  int *bad = ctx->d_scratch_flag();
  CU(cudaMemsetAsync(bad, 0, sizeof(int), s));
  launch_hits_to_keys(cands_dev, n, ctx->cand, bad, s);
  launch_finalize(ctx->fw, ctx->cand, n, MAX_BATCH_QUERIES, k_fin, 0, 0, ctx->hits, ctx->ws_cap, ctx->h_scalars + 1, s);
  int hbad = 0;
  CU(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  CU(cudaGetLastError());
  if (hbad) return fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_merge_dev: query index >= 2^20 or distance >= 4096 in one call");
  uint64_t rows = ctx->h_scalars[1];
  // stream-ordered like any *_dev result: valid for work queued on `stream` after this call
  CU(cudaMemcpyAsync(cands_dev, ctx->hits, rows * sizeof(smafa_hit), cudaMemcpyDeviceToDevice, s));
  *n_out = rows;
  return SMAFA_OK;
}

extern "C" uint64_t smafa_apply_limit_per_sequence(smafa_hit *hits, uint64_t n, const uint64_t *db_enc, uint32_t W,
                                                   uint64_t subject_offset, uint32_t limit) {
  // src/lib.rs:259-260,269-289: the run is keyed on the decoded subject string, i.e. on the
  // encoding; a skipped hit does not reset the run; a new query starts with no run.
  uint64_t o = 0;
  const uint64_t *last = nullptr;
  uint32_t count = 0, last_q = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t *enc = db_enc + ((uint64_t)hits[i].subject - subject_offset) * W;
    if (i == 0 || hits[i].query != last_q) { last = nullptr; count = 0; last_q = hits[i].query; }
    if (last && memcmp(last, enc, W * sizeof(uint64_t)) == 0) {
      if (count >= limit) continue;
      count++;
    } else {
      last = enc;
      count = 1;
    }
    hits[o++] = hits[i];
  }
  return o;
}

// ------------------------------------------------------------------------------- cluster

// src/cluster.rs:45-74 with the distance evaluation batched on the GPU.
//
// The greedy is order dependent: sequence i sees every centroid founded by sequences < i.  For a
// batch [b0,b1) the GPU evaluates (1) batch x existing centroids -- Mode A with --max-divergence t,
// i.e. per sequence the minimum distance and the lowest centroid index at it, only if <= t -- and
// (2) batch x batch with a fixed bound t (every pair within t).  The host then replays the batch
// sequentially: a sequence joins the nearest centroid among the old ones and the in-batch
// sequences that became centroids before it (ties -> lowest centroid index, src/cluster.rs:62-68:
// old centroids always have lower indices than in-batch ones), else founds a new centroid.
extern "C" int smafa_cluster(smafa_ctx *ctx, const uint64_t *enc, uint64_t n, uint32_t L, uint32_t t,
                             uint32_t *centroid_of, uint64_t *n_centroids, uint64_t *n_comparisons, smafa_stats *stats) {
  // the greedy is strictly sequential (src/cluster.rs:45-74): a multi-device context runs it on its first device
  if (ctx && ctx->multi) ctx = multi_first(ctx);
  return cluster_impl(ctx, enc, n, L, t, centroid_of, n_centroids, n_comparisons, stats, UINT64_MAX);
}

// The greedy of src/cluster.rs:45-74.  max_centroids: give up with RC_TOO_MANY_CENTROIDS once more centroids than
// that exist (group_order: a db without near-duplicates is not worth grouping, and its greedy would be quadratic).
static int cluster_impl(smafa_ctx *ctx, const uint64_t *enc, uint64_t n, uint32_t L, uint32_t t, uint32_t *centroid_of,
                        uint64_t *n_centroids, uint64_t *n_comparisons, smafa_stats *stats, uint64_t max_centroids) {
  if (!ctx || (n && (!enc || !centroid_of))) return fail(ctx, SMAFA_E_INVALID, "smafa_cluster: null argument");
  if (stats) memset(stats, 0, sizeof *stats);
  if (n_centroids) *n_centroids = 0;
  if (n_comparisons) *n_comparisons = 0;
  if (n == 0) return SMAFA_OK;
  if (L == 0) return fail(ctx, SMAFA_E_INVALID, "smafa_cluster: L == 0");
  if (n >= (1ull << 32)) return fail(ctx, SMAFA_E_UNSUPPORTED, "more than 2^32-1 sequences");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const uint32_t W = words_for(L);
  smafa_db *cdb = nullptr, *bdb = nullptr;
  const auto t_setup = std::chrono::steady_clock::now();
  int rc = smafa_db_upload(ctx, nullptr, 0, L, 0, &cdb);
  if (!rc) rc = smafa_db_upload(ctx, nullptr, 0, L, 0, &bdb);
  // The centroid set can only grow to n rows (~264 B each with all three images): reserving it up front avoids
  // the realloc + copy + cudaFree cycle of a growing db, which cost 0.78 s of a 1.27 s run at n = 3.9 M.
  if (!rc) rc = db_reserve(ctx, cdb, std::min<uint64_t>(n, 64ull << 20));
  std::vector<uint32_t> cent_input;          // centroid number -> input index
  std::vector<int64_t> cent_of_input_batch;  // in-batch: centroid number founded by batch row, or -1
  std::vector<smafa_hit> old_hits, in_hits;
  std::vector<uint64_t> new_words;
  uint64_t comparisons = 0, pairs = 0;
  cudaEventRecord(ctx->ev[2], s);

  // SMAFA_TIMING=1: where the wall time of the greedy goes (host clock, stderr)
  const bool timing = getenv("SMAFA_TIMING") != nullptr;
  double t_stage[7] = {0, 0, 0, 0, 0, 0, 0};  // batch upload, scan vs centroids, scan in-batch, host replay, centroid append, set-up, tear-down
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto tick = now();
  auto lap = [&](int i) {
    const auto n = now();
    t_stage[i] += std::chrono::duration<double, std::milli>(n - tick).count();
    tick = n;
  };
  uint64_t n_batches = 0;
  t_stage[5] = std::chrono::duration<double, std::milli>(tick - t_setup).count();

  QueryPlan plan_old{MODE_MIN, 1, 1, (int)std::min<uint32_t>(t, L)};
  QueryPlan plan_in{MODE_FIXED, 0, UINT32_MAX, (int)std::min<uint32_t>(t, L)};

  for (uint64_t b0 = 0; b0 < n && !rc;) {
    const uint64_t C = cent_input.size();
    uint64_t B = std::min<uint64_t>(std::max<uint64_t>(4096, C / 2), 65536);
    B = std::min(B, n - b0);
    const uint64_t *benc = enc + b0 * W;
    // batch rows as queries (device copy) and as a db
    if ((rc = ensure_buf(ctx, ctx->q_ref, ctx->q_ref_cap, B * W))) break;
    cudaMemcpyAsync(ctx->q_ref, benc, B * W * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
    bdb->D = 0;
    if ((rc = db_add_rows(ctx, bdb, benc, B))) break;
    ++n_batches;
    lap(0);

    old_hits.clear();
    in_hits.clear();
    auto collect = [&](std::vector<smafa_hit> &dst) {
      return [&dst, ctx, s](uint64_t, uint64_t, uint64_t rows) -> int {
        size_t old = dst.size();
        dst.resize(old + rows);
        if (rows) {
          cudaError_t e = cudaMemcpyAsync(dst.data() + old, ctx->hits, rows * sizeof(smafa_hit), cudaMemcpyDeviceToHost, s);
          if (e == cudaSuccess) e = cudaStreamSynchronize(s);
          if (e != cudaSuccess) return SMAFA_E_CUDA;
        }
        return SMAFA_OK;
      };
    };
    if (C > 0) {
      rc = run_query_range(ctx, cdb, ctx->q_ref, B, 0, plan_old, s, stats, BatchOut(), collect(old_hits));
      if (rc) break;
      pairs += B * C;
    }
    lap(1);
    rc = run_query_range(ctx, bdb, ctx->q_ref, B, 0, plan_in, s, stats, BatchOut(), collect(in_hits));
    if (rc) break;
    pairs += B * B;
    lap(2);

    // sequential replay (hits are sorted by (query, distance, subject))
    cent_of_input_batch.assign(B, -1);
    new_words.clear();
    size_t po = 0, pi = 0;
    for (uint64_t i = 0; i < B; ++i) {
      comparisons += cent_input.size();
      int64_t best_c = -1;
      uint32_t best_d = UINT32_MAX;
      while (po < old_hits.size() && old_hits[po].query < i) ++po;
      if (po < old_hits.size() && old_hits[po].query == i) {  // first row = lowest centroid at the minimum
        best_d = old_hits[po].distance;
        best_c = old_hits[po].subject;
      }
      while (pi < in_hits.size() && in_hits[pi].query < i) ++pi;
      for (size_t h = pi; h < in_hits.size() && in_hits[h].query == i; ++h) {
        const smafa_hit &x = in_hits[h];
        if (x.distance >= best_d) break;  // sorted by distance; old centroids win ties
        if (x.subject < i && cent_of_input_batch[x.subject] >= 0) {
          best_d = x.distance;  // first in (distance, subject) order = lowest in-batch centroid
          best_c = cent_of_input_batch[x.subject];
          break;
        }
      }
      if (best_c >= 0 && best_d <= t) {
        centroid_of[b0 + i] = cent_input[(size_t)best_c];
      } else {
        cent_of_input_batch[i] = (int64_t)cent_input.size();
        cent_input.push_back((uint32_t)(b0 + i));
        centroid_of[b0 + i] = (uint32_t)(b0 + i);
        new_words.insert(new_words.end(), benc + i * W, benc + (i + 1) * W);
      }
    }
    lap(3);
    if (cent_input.size() > max_centroids) { rc = RC_TOO_MANY_CENTROIDS; break; }
    if (!new_words.empty()) rc = db_add_rows(ctx, cdb, new_words.data(), new_words.size() / W);
    lap(4);
    b0 += B;
  }
  cudaEventRecord(ctx->ev[3], s);
  cudaEventSynchronize(ctx->ev[3]);
  if (stats) {
    cudaEventElapsedTime(&stats->total_ms, ctx->ev[2], ctx->ev[3]);
    stats->pairs = pairs;
  }
  tick = now();
  if (cdb) smafa_db_free(cdb);
  if (bdb) smafa_db_free(bdb);
  lap(6);
  if (timing)
    fprintf(stderr, "[smafa timing] cluster: %llu batches; set-up %.1f ms, batch upload %.1f ms, scan vs centroids %.1f ms, "
                    "scan in-batch %.1f ms, host replay %.1f ms, centroid append %.1f ms, tear-down %.1f ms\n",
            (unsigned long long)n_batches, t_stage[5], t_stage[0], t_stage[1], t_stage[2], t_stage[3], t_stage[4], t_stage[6]);
  if (rc == SMAFA_E_CUDA && ctx->err.empty()) fail(ctx, rc, "smafa_cluster: CUDA failure");
  if (rc) return rc;
  if (n_centroids) *n_centroids = cent_input.size();
  if (n_comparisons) *n_comparisons = comparisons;
  return SMAFA_OK;
}

// Debug/parity hook for the tcgen05 formulation: runs one MMA scan of (up to 256) queries against the
// db with a fixed bound and returns the raw int32 accumulators of the first tile
// (out[row * 256 + col] = matches(db row, query col) - (L - bound)).
extern "C" int smafa_debug_mma_dump(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t bound,
                                    int32_t *out) {
  if (!ctx || !db || !q_enc || !out || Q == 0 || Q > 256) return fail(ctx, SMAFA_E_INVALID, "smafa_debug_mma_dump: bad argument");
  if (ctx->multi) return fail(ctx, SMAFA_E_UNSUPPORTED, "debug hooks need a single-device context");
  if (!mma_supported(db) || db->D == 0) return fail(ctx, SMAFA_E_UNSUPPORTED, "db not eligible for the MMA kernel");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  int rc;
  if ((rc = ensure_buf(ctx, ctx->q_ref, ctx->q_ref_cap, Q * db->W))) return rc;
  CU(cudaMemcpyAsync(ctx->q_ref, q_enc, Q * db->W * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  int32_t *dump = nullptr;
  CU(cudaMalloc((void **)&dump, 128 * 256 * sizeof(int32_t)));
  CU(cudaMemsetAsync(dump, 0x7f, 128 * 256 * sizeof(int32_t), s));
  QueryPlan plan{MODE_FIXED, 0, UINT32_MAX, (int)std::min<uint32_t>(bound, db->L)};
  const int saved_kernel = ctx->kernel;
  ctx->kernel = SMAFA_KERNEL_MMA;
  ctx->mma_dump = dump;
  uint64_t rows = 0;
  rc = run_batch(ctx, db, ctx->q_ref, (uint32_t)Q, 0, plan, &rows, s, nullptr);
  ctx->mma_dump = nullptr;
  ctx->kernel = saved_kernel;
  if (rc == SMAFA_OK) {
    cudaError_t e = cudaMemcpy(out, dump, 128 * 256 * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(ctx, SMAFA_E_CUDA, "dump copy: %s", cudaGetErrorString(e));
  } else if (rc == RC_OVERFLOW) {
    rc = fail(ctx, SMAFA_E_OOM, "candidate overflow in debug dump");
  }
  cudaFree(dump);
  return rc;
}

// Measures the dense int8 tcgen05 rate of this GPU (the roofline denominator of the MMA formulation):
// every SM issues `mmas_per_cta` back-to-back M128xN256xK32 kind::i8 MMAs on resident operands.
extern "C" int smafa_debug_mma_peak(smafa_ctx *ctx, uint32_t mmas_per_cta, double *tops) {
  if (!ctx || !tops || mmas_per_cta == 0) return fail(ctx, SMAFA_E_INVALID, "smafa_debug_mma_peak: bad argument");
  if (ctx->multi) ctx = multi_first(ctx);
  CU(cudaSetDevice(ctx->device));
  float ms = 0;
  int rc = mma_peak_probe(ctx, mmas_per_cta, &ms);
  if (rc) return rc;
  *tops = 2.0 * 128 * 256 * 32 * (double)mmas_per_cta * ctx->num_sms / (ms * 1e-3) / 1e12;
  return SMAFA_OK;
}

extern "C" uint32_t smafa_ctx_last_mma_k(const smafa_ctx *ctx) { return ctx ? ctx->last_mma_k : 0; }
