// Row-sharded queries across GPUs, behind the C ABI of include/smafa_b200.h:
//
//   one process per GPU   smafa_ctx_comm_init + smafa_db_upload_shard + smafa_query_sharded[_dev]  (torchrun, MPI-style
//                         launchers; the exchange is ONE ncclAllGather of fixed-capacity blocks on the caller's stream)
//   one process, n GPUs   smafa_ctx_create_multi: smafa_db_upload / smafa_query / smafa_query_file ... work as on one
//                         device; the db is row-sharded over the devices, every device is driven by its own host
//                         thread, blocks reach the first device by peer copies over NVLink (the `smafa` CLI: --devices)
//
// Either way every shard runs the ordinary scan + selection on its rows (api.cu run_query_range) and leaves its answer
// -- a superset of its part of the global answer, already in print order -- as a block of candidate keys; merge.cu
// turns the gathered blocks into the reference's rows (src/lib.rs:243-265, 298-312) with global subject numbers.  The
// host reads four numbers back per exchange: rows kept, the fullest block's need, the first failing status, its rank.
// A block that overflowed is re-sent with a larger capacity after the shards have run again (rare: the capacity
// adapts to twice the last need); a shard that failed still takes part in the exchange with its status in the block
// header, so no rank is left waiting in a collective and all of them report the failure.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy a host process such as PyTorch has already loaded,
// else the system's), so the library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "kernels.h"

using namespace smafa;

// ------------------------------------------------------------------------------- NCCL (run-time binding)

namespace {

struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string why;  // why it is unavailable
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("SMAFA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
      api.why = dlerror();
    }
    if (!api.handle) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
      api.why = "libnccl lacks a required symbol";
      api.handle = nullptr;
    }
  });
  return &api;
}

uint64_t pow2_at_least(uint64_t n) {
  uint64_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

struct SmafaComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

static_assert(SMAFA_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "smafa_comm_unique_id hands out an ncclUniqueId");

extern "C" int smafa_comm_unique_id(uint8_t *id) {
  if (!id) return smafa_fail(nullptr, SMAFA_E_INVALID, "smafa_comm_unique_id: null argument");
  NcclApi *api = nccl_api();
  if (!api->handle) return smafa_fail(nullptr, SMAFA_E_NCCL, "NCCL is not available: %s", api->why.c_str());
  ncclUniqueId uid;
  ncclResult_t r = api->GetUniqueId(&uid);
  if (r != ncclSuccess) return smafa_fail(nullptr, SMAFA_E_NCCL, "ncclGetUniqueId: %s", api->GetErrorString(r));
  memcpy(id, uid.internal, SMAFA_COMM_ID_BYTES);
  return SMAFA_OK;
}

void comm_free(smafa_ctx *ctx) {
  if (!ctx || !ctx->comm) return;
  NcclApi *api = nccl_api();
  if (api->handle && ctx->comm->comm) api->CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
}

extern "C" int smafa_ctx_comm_init(smafa_ctx *ctx, const uint8_t *id, int rank, int world_size) {
  if (!ctx || !id || world_size < 1 || rank < 0 || rank >= world_size)
    return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_ctx_comm_init: bad argument (rank %d of %d)", rank, world_size);
  if (ctx->multi) return smafa_fail(ctx, SMAFA_E_UNSUPPORTED, "smafa_ctx_comm_init: a multi-device context needs no communicator");
  comm_free(ctx);
  NcclApi *api = nccl_api();
  if (!api->handle) return smafa_fail(ctx, SMAFA_E_NCCL, "NCCL is not available: %s", api->why.c_str());
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return smafa_fail(ctx, SMAFA_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  ncclUniqueId uid;
  memcpy(uid.internal, id, SMAFA_COMM_ID_BYTES);
  SmafaComm *c = new SmafaComm();
  c->rank = rank;
  c->world = world_size;
  ncclResult_t r = api->CommInitRank(&c->comm, world_size, uid, rank);
  if (r != ncclSuccess) {
    delete c;
    return smafa_fail(ctx, SMAFA_E_NCCL, "ncclCommInitRank(rank %d of %d): %s", rank, world_size, api->GetErrorString(r));
  }
  ctx->comm = c;
  return SMAFA_OK;
}

extern "C" void smafa_ctx_comm_free(smafa_ctx *ctx) { comm_free(ctx); }

// ------------------------------------------------------------------------------- exchange buffers

void exchange_free(smafa_ctx *ctx) {
  if (!ctx) return;
  SmafaExchange &x = ctx->xchg;
  cudaFree(x.block); cudaFree(x.gathered);
  cudaFree(x.mw.seg); cudaFree(x.mw.merged); cudaFree(x.mw.flags); cudaFree(x.mw.selected); cudaFree(x.mw.cub_temp);
  cudaFree(x.hits);
  cudaFree(x.info_dev);
  if (x.info_host) cudaFreeHost(x.info_host);
  for (auto &ev : x.ev)
    if (ev) cudaEventDestroy(ev);
  x = SmafaExchange();
}

// Send block of `cap` rows; with n_ranks > 0 also the receive area, the merge workspace for Q queries and a device
// buffer for the merged rows (this context merges).
static int exchange_reserve(smafa_ctx *ctx, uint32_t n_ranks, uint64_t cap, uint32_t Q, cudaStream_t s) {
  SmafaExchange &x = ctx->xchg;
  cudaError_t e = cudaSuccess;
  if (!x.info_dev) {
    e = cudaMalloc((void **)&x.info_dev, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&x.info_host, 8 * sizeof(unsigned long long), cudaHostAllocDefault);
    for (auto &ev : x.ev)
      if (e == cudaSuccess) e = cudaEventCreate(&ev);
  }
  const bool regrow = x.cap != cap || x.n_ranks != n_ranks;
  if (e == cudaSuccess && regrow) {
    cudaStreamSynchronize(s);
    cudaFree(x.block); cudaFree(x.gathered); cudaFree(x.mw.merged); cudaFree(x.mw.flags); cudaFree(x.mw.selected);
    cudaFree(x.mw.cub_temp); cudaFree(x.hits);
    x.block = x.gathered = x.mw.merged = x.mw.selected = nullptr;
    x.mw.flags = nullptr; x.mw.cub_temp = nullptr; x.hits = nullptr;
    x.cap = 0; x.n_ranks = 0; x.mw.rows = 0;
    e = cudaMalloc((void **)&x.block, (2 + cap) * sizeof(uint64_t));
    if (e == cudaSuccess && n_ranks) {
      const uint64_t rows = (uint64_t)n_ranks * cap;
      e = cudaMalloc((void **)&x.gathered, (uint64_t)n_ranks * (2 + cap) * sizeof(uint64_t));
      if (e == cudaSuccess) e = cudaMalloc((void **)&x.mw.merged, rows * sizeof(uint64_t));
      if (e == cudaSuccess) e = cudaMalloc((void **)&x.mw.selected, rows * sizeof(uint64_t));
      if (e == cudaSuccess) e = cudaMalloc((void **)&x.mw.flags, rows);
      x.mw.cub_temp_bytes = merge_temp_bytes(rows);
      if (e == cudaSuccess) e = cudaMalloc(&x.mw.cub_temp, x.mw.cub_temp_bytes);
      if (e == cudaSuccess) e = cudaMalloc((void **)&x.hits, rows * sizeof(smafa_hit));
      x.mw.rows = rows;
    }
    if (e == cudaSuccess) { x.cap = cap; x.n_ranks = n_ranks; }
  }
  const uint64_t seg_len = (uint64_t)n_ranks * ((uint64_t)Q + 1);
  if (e == cudaSuccess && n_ranks && x.mw.seg_len < seg_len) {
    cudaStreamSynchronize(s);
    cudaFree(x.mw.seg);
    x.mw.seg = nullptr; x.mw.seg_len = 0;
    e = cudaMalloc((void **)&x.mw.seg, seg_len * sizeof(uint32_t));
    if (e == cudaSuccess) x.mw.seg_len = seg_len;
  }
  if (e != cudaSuccess) {
    exchange_free(ctx);
    return smafa_fail(ctx, e == cudaErrorMemoryAllocation ? SMAFA_E_OOM : SMAFA_E_CUDA, "exchange buffers (%u blocks of %llu rows): %s",
                      n_ranks, (unsigned long long)cap, cudaGetErrorString(e));
  }
  x.mw.n_selected = x.info_dev + 4;
  return SMAFA_OK;
}

static uint64_t initial_block_cap(uint64_t nq) {
  // SMAFA_XCHG_CAP=rows: a small first capacity, so that tests reach the overflow re-send
  if (const char *e = getenv("SMAFA_XCHG_CAP")) {
    const long long v = atoll(e);
    if (v > 0) return (uint64_t)v;
  }
  return pow2_at_least(std::max<uint64_t>(4096, 4 * nq));
}
// capacity for the next exchange: twice the fullest block's last need (the same on every rank: all saw the same headers)
static uint64_t adapted_block_cap(uint64_t cap, uint64_t need) {
  const uint64_t want = pow2_at_least(std::max<uint64_t>(4096, 2 * need));
  return want < cap / 2 || want > cap ? want : cap;
}

// Local part of one slab on one shard: this shard's answer to queries [0, nq) (device words) lands in ctx->xchg.block.
// Failures are recorded in the block header (the shard still takes part in the exchange) and returned.
static int shard_local(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_dev, uint64_t nq, int64_t m, int64_t k, cudaStream_t s,
                       smafa_stats *st) {
  SmafaExchange &x = ctx->xchg;
  launch_block_reset(x.block, 0, s);
  int rc = SMAFA_OK;
  if (db->D > 0) {  // an empty shard of a non-empty db contributes no rows
    QueryPlan plan{};
    rc = validate_query_plan(ctx, db->D, db->L, nq, db->L, m, k, &plan);  // the local plan: k against the shard's own rows
    BatchOut out;
    out.block = x.block;
    out.block_cap = x.cap;
    if (rc == SMAFA_OK)
      rc = run_query_range(ctx, db, q_dev, nq, 0, plan, s, st, out, [](uint64_t, uint64_t, uint64_t) { return SMAFA_OK; });
  }
  if (rc != SMAFA_OK) launch_block_reset(x.block, (uint64_t)(-rc), s);
  return rc;
}

static uint32_t merge_k(int64_t k) {
  const bool mode_b = (k >= 0 && k != 1);  // src/lib.rs:224
  return mode_b ? (uint32_t)std::min<int64_t>(k, UINT32_MAX) : 1u;
}

// ------------------------------------------------------------------------------- one process per GPU (NCCL)

// One slab (<= 2^20 queries) of a sharded query: local scan, all-gather, merge.  Rows go to hits_out (device, capacity
// hits_cap rows) with query numbers starting at q_base; *rows_out = rows of the merged answer.
static int sharded_slab_nccl(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_dev, uint64_t nq, uint64_t q_base, int64_t m, int64_t k,
                             smafa_hit *hits_out, uint64_t hits_cap, uint64_t *rows_out, cudaStream_t s, smafa_stats *st) {
  NcclApi *api = nccl_api();
  SmafaComm *c = ctx->comm;
  SmafaExchange &x = ctx->xchg;
  uint64_t cap = x.n_ranks == (uint32_t)c->world && x.cap_hint ? x.cap_hint : initial_block_cap(nq);
  for (int attempt = 0;; ++attempt) {
    int rc = exchange_reserve(ctx, (uint32_t)c->world, cap, (uint32_t)nq, s);
    if (rc) return rc;  // (a rank that cannot even allocate leaves the others waiting: NCCL's own timeout ends that)
    const int local_rc = shard_local(ctx, db, q_dev, nq, m, k, s, st);
    cudaEventRecord(x.ev[0], s);
    ncclResult_t r = api->AllGather(x.block, x.gathered, 2 + cap, ncclUint64, c->comm, s);
    if (r != ncclSuccess) return smafa_fail(ctx, SMAFA_E_NCCL, "ncclAllGather: %s", api->GetErrorString(r));
    smafa_hit *dst = hits_out ? hits_out : x.hits;
    const uint64_t dst_cap = hits_out ? hits_cap : x.mw.rows;
    const int launches = 1 + launch_merge_blocks(x.mw, x.gathered, (uint32_t)c->world, cap, (uint32_t)nq, merge_k(k), (uint32_t)q_base, dst,
                                                 dst_cap, x.info_dev, s);
    if (st) st->kernel_launches += launches;
    cudaEventRecord(x.ev[1], s);
    cudaError_t e = cudaMemcpyAsync(x.info_host, x.info_dev, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return smafa_fail(ctx, SMAFA_E_CUDA, "sharded query: %s", cudaGetErrorString(e));
    if (st) {
      float ms = 0;
      cudaEventElapsedTime(&ms, x.ev[0], x.ev[1]);
      st->exchange_ms += ms;
    }
    if (local_rc) return local_rc;
    if (x.info_host[1])
      return smafa_fail(ctx, SMAFA_E_PEER, "rank %llu of the sharded query failed with %s", x.info_host[2],
                        smafa_status_name(-(int)x.info_host[1]));
    const uint64_t need = x.info_host[0];
    if (need > cap) {  // some block overflowed (every rank sees that): larger blocks, the shards run again
      if (st) st->retries++;
      if (attempt > 8) return smafa_fail(ctx, SMAFA_E_OOM, "sharded query: block capacity keeps overflowing");
      cap = pow2_at_least(need + need / 4);
      continue;
    }
    *rows_out = x.info_host[3];
    x.cap_hint = adapted_block_cap(cap, need);
    return SMAFA_OK;
  }
}

static int sharded_check(smafa_ctx *ctx, const smafa_db *db, uint64_t Q, uint32_t q_len, int64_t m, int64_t k, const char *who) {
  if (!ctx->comm) return smafa_fail(ctx, SMAFA_E_INVALID, "%s: no communicator (smafa_ctx_comm_init)", who);
  if (!db->global_rows && db->D) return smafa_fail(ctx, SMAFA_E_INVALID, "%s: the db is not a shard (smafa_db_upload_shard)", who);
  // the reference's checks, against the WHOLE db (an empty shard of a non-empty db is fine)
  QueryPlan whole{};
  return validate_query_plan(ctx, db->global_rows, db->L, Q, q_len, m, k, &whole);
}

extern "C" int smafa_query_sharded_dev(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc_dev, uint64_t Q, uint32_t q_len,
                                       int64_t m, int64_t k, smafa_hit *hits_dev, uint64_t hits_capacity, uint64_t *n_hits,
                                       void *stream, smafa_stats *stats) {
  if (!ctx || !db || !n_hits) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_query_sharded_dev: null argument");
  *n_hits = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  int rc = sharded_check(ctx, db, Q, q_len, m, k, "smafa_query_sharded_dev");
  if (rc || Q == 0) return rc;
  if (!q_enc_dev) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_query_sharded_dev: null query buffer");
  cudaSetDevice(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  cudaEventRecord(ctx->ev[2], s);
  uint64_t total = 0;
  for (uint64_t q0 = 0; q0 < Q; q0 += MAX_BATCH_QUERIES) {
    const uint64_t nq = std::min<uint64_t>(MAX_BATCH_QUERIES, Q - q0);
    const uint64_t room = hits_dev && total < hits_capacity ? hits_capacity - total : 0;
    uint64_t rows = 0;
    // without room the rows go to the exchange's own buffer and are only counted
    rc = sharded_slab_nccl(ctx, db, q_enc_dev + q0 * db->W, nq, q0, m, k, room ? hits_dev + total : nullptr, room, &rows, s, stats);
    if (rc) return rc;
    total += rows;
  }
  cudaEventRecord(ctx->ev[3], s);
  cudaEventSynchronize(ctx->ev[3]);
  if (stats) {
    cudaEventElapsedTime(&stats->total_ms, ctx->ev[2], ctx->ev[3]);
    stats->pairs = Q * db->D;
  }
  *n_hits = total;
  if (total > hits_capacity)
    return smafa_fail(ctx, SMAFA_E_OOM, "hits buffer too small: %llu rows needed, capacity %llu", (unsigned long long)total,
                      (unsigned long long)hits_capacity);
  return SMAFA_OK;
}

// Grows a malloc'd row buffer to hold n more rows.
static int grow_rows(smafa_ctx *ctx, smafa_hit *&all, uint64_t &cap_all, uint64_t n_all, uint64_t rows) {
  if (n_all + rows <= cap_all) return SMAFA_OK;
  cap_all = std::max<uint64_t>({n_all + rows, cap_all * 2, 1024});
  smafa_hit *grown = (smafa_hit *)realloc(all, cap_all * sizeof(smafa_hit));
  if (!grown) return smafa_fail(ctx, SMAFA_E_OOM, "realloc of %llu hits failed", (unsigned long long)cap_all);
  all = grown;
  return SMAFA_OK;
}

extern "C" int smafa_query_sharded(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len, int64_t m,
                                   int64_t k, smafa_hit **hits, uint64_t *n_hits, smafa_stats *stats) {
  if (!ctx || !db || !hits || !n_hits) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_query_sharded: null argument");
  *hits = nullptr;
  *n_hits = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  int rc = sharded_check(ctx, db, Q, q_len, m, k, "smafa_query_sharded");
  if (rc || Q == 0) return rc;
  if (!q_enc) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_query_sharded: null query buffer");
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  cudaEventRecord(ctx->ev[2], s);
  smafa_hit *all = nullptr;
  uint64_t n_all = 0, cap_all = 0;
  for (uint64_t q0 = 0; q0 < Q && rc == SMAFA_OK; q0 += MAX_BATCH_QUERIES) {
    const uint64_t nq = std::min<uint64_t>(MAX_BATCH_QUERIES, Q - q0);
    if ((rc = ensure_query_words(ctx, nq * db->W))) break;
    cudaError_t e = cudaMemcpyAsync(ctx->q_ref, q_enc + q0 * db->W, nq * db->W * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { rc = smafa_fail(ctx, SMAFA_E_CUDA, "H2D of queries: %s", cudaGetErrorString(e)); break; }
    uint64_t rows = 0;
    if ((rc = sharded_slab_nccl(ctx, db, ctx->q_ref, nq, q0, m, k, nullptr, 0, &rows, s, stats))) break;
    if ((rc = grow_rows(ctx, all, cap_all, n_all, rows))) break;
    if (rows) {
      e = cudaMemcpyAsync(all + n_all, ctx->xchg.hits, rows * sizeof(smafa_hit), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { rc = smafa_fail(ctx, SMAFA_E_CUDA, "D2H of hits: %s", cudaGetErrorString(e)); break; }
    }
    n_all += rows;
  }
  if (rc) { free(all); return rc; }
  cudaEventRecord(ctx->ev[3], s);
  cudaEventSynchronize(ctx->ev[3]);
  if (stats) {
    cudaEventElapsedTime(&stats->total_ms, ctx->ev[2], ctx->ev[3]);
    stats->pairs = Q * db->D;
  }
  if (!all && !(all = (smafa_hit *)malloc(sizeof(smafa_hit)))) return smafa_fail(ctx, SMAFA_E_OOM, "malloc failed");
  *hits = all;
  *n_hits = n_all;
  return SMAFA_OK;
}

// ------------------------------------------------------------------------------- one process, several GPUs

struct SmafaMulti {
  std::vector<smafa_ctx *> dev;      // one ordinary context per device; dev[0] merges
  std::vector<cudaEvent_t> done;     // local part of device r finished (recorded on dev[r]'s stream)
};

smafa_ctx *multi_first(smafa_ctx *ctx) { return ctx->multi->dev[0]; }

// Runs f(r) for every device on its own host thread (device 0 on the calling thread) and returns the first failure.
template <class F>
static int on_every_device(SmafaMulti *mu, F &&f) {
  const size_t n = mu->dev.size();
  std::vector<int> rc(n, SMAFA_OK);
  std::vector<std::thread> th;
  th.reserve(n);
  for (size_t r = 1; r < n; ++r) th.emplace_back([&, r] { rc[r] = f(r); });
  rc[0] = f(0);
  for (auto &t : th) t.join();
  for (size_t r = 0; r < n; ++r)
    if (rc[r]) return rc[r];
  return SMAFA_OK;
}

// first failing child's message becomes the parent's
static int adopt_error(smafa_ctx *ctx, int rc) {
  if (rc)
    for (smafa_ctx *c : ctx->multi->dev)
      if (!c->err.empty()) { smafa_fail(ctx, rc, "%s", c->err.c_str()); break; }
  return rc;
}

extern "C" int smafa_ctx_create_multi(smafa_ctx **out, const int *devices, int n_devices, int kernel) {
  if (!out || !devices || n_devices < 1) return smafa_fail(nullptr, SMAFA_E_INVALID, "smafa_ctx_create_multi: bad argument");
  *out = nullptr;
  if (n_devices == 1) return smafa_ctx_create(out, devices[0], kernel);
  for (int i = 0; i < n_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return smafa_fail(nullptr, SMAFA_E_INVALID, "smafa_ctx_create_multi: device %d listed twice", devices[i]);
  SmafaMulti *mu = new SmafaMulti();
  mu->dev.assign(n_devices, nullptr);
  mu->done.assign(n_devices, nullptr);
  // the primary contexts are created concurrently: each takes a good part of a second
  std::vector<std::string> errs(n_devices);
  int rc = on_every_device(mu, [&](size_t r) {
    int c = smafa_ctx_create(&mu->dev[r], devices[r], kernel);
    if (c) { errs[r] = smafa_last_error(nullptr); return c; }  // thread-local message of this thread
    cudaError_t e = cudaEventCreateWithFlags(&mu->done[r], cudaEventDisableTiming);
    if (e != cudaSuccess) { errs[r] = cudaGetErrorString(e); return (int)SMAFA_E_CUDA; }
    return (int)SMAFA_OK;
  });
  if (rc == SMAFA_OK) {
    // peer access lets the copies to the merging device go straight over NVLink (without it CUDA stages them through the host)
    cudaSetDevice(devices[0]);
    for (int r = 1; r < n_devices; ++r) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[0], devices[r]) == cudaSuccess && can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
        else (void)cudaGetLastError();
      }
    }
  }
  if (rc) {
    std::string msg;
    for (auto &m : errs)
      if (!m.empty()) { msg = m; break; }
    for (size_t r = 0; r < mu->dev.size(); ++r) {
      if (mu->done[r]) { cudaSetDevice(devices[r]); cudaEventDestroy(mu->done[r]); }
      smafa_ctx_destroy(mu->dev[r]);
    }
    delete mu;
    return smafa_fail(nullptr, rc, "smafa_ctx_create_multi: %s", msg.c_str());
  }
  smafa_ctx *ctx = new smafa_ctx();
  ctx->device = devices[0];
  ctx->kernel = kernel;
  ctx->multi = mu;
  *out = ctx;
  return SMAFA_OK;
}

extern "C" int smafa_ctx_device_count(const smafa_ctx *ctx) { return !ctx ? 0 : ctx->multi ? (int)ctx->multi->dev.size() : 1; }

void multi_destroy(smafa_ctx *ctx) {
  SmafaMulti *mu = ctx->multi;
  for (size_t r = 0; r < mu->dev.size(); ++r) {
    if (mu->done[r]) { cudaSetDevice(mu->dev[r]->device); cudaEventDestroy(mu->done[r]); }
    smafa_ctx_destroy(mu->dev[r]);
  }
  delete mu;
  ctx->multi = nullptr;
}

int multi_set(smafa_ctx *ctx, int what, int64_t value) {
  int rc = SMAFA_OK;
  for (smafa_ctx *c : ctx->multi->dev) {
    int r = what == 0 ? smafa_ctx_set_kernel(c, (int)value) : what == 1 ? smafa_ctx_set_alphabet(c, (int)value)
                                                                       : smafa_ctx_set_candidate_capacity(c, (uint64_t)value);
    if (r && !rc) rc = r;
  }
  if (what == 0 && !rc) ctx->kernel = (int)value;
  return adopt_error(ctx, rc);
}

// rows [lo, hi) of device r: contiguous ranges, the first D % n shards one row longer (like smafa_b200/dist.py shard_bounds)
static void shard_range(uint64_t D, size_t n, size_t r, uint64_t *lo, uint64_t *hi) {
  const uint64_t base = D / n, extra = D % n;
  *lo = r * base + std::min<uint64_t>(r, extra);
  *hi = *lo + base + (r < extra ? 1 : 0);
}

int multi_db_upload(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint64_t subject_offset, smafa_db **out) {
  SmafaMulti *mu = ctx->multi;
  const size_t R = mu->dev.size();
  if (D > 0 && (!enc || L == 0)) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_db_upload: D > 0 needs enc and L > 0");
  if (subject_offset + D >= (1ull << 32)) return smafa_fail(ctx, SMAFA_E_UNSUPPORTED, "db larger than 2^32-1 windows");
  smafa_db *db = new smafa_db();
  db->ctx = ctx;
  db->D = D;
  db->L = L;
  db->W = (L + 11) / 12;
  db->subject_offset = subject_offset;
  db->alphabet = ctx->alphabet;
  db->shards.assign(R, nullptr);
  // The db is grouped as a whole (api.cu group_order, on the first device) and the GROUPED order is cut into the
  // shards, so that every device holds whole families of similar windows; rows keep their subject numbers through the
  // shards' row -> subject tables, which is all the merge needs.
  int rc = SMAFA_OK;
  if (mu->dev[0]->db_group && D >= 65536) {
    db->perm_host.resize(D);
    uint64_t nc = 0;
    rc = smafa_group_order(mu->dev[0], enc, D, L, db->perm_host.data(), &nc);
    if (rc) { adopt_error(ctx, rc); multi_db_free(db); return rc; }
    if (nc == 0) db->perm_host.clear();  // no structure: plain contiguous shards
  }
  const bool mapped = !db->perm_host.empty();
  db->grouped = mapped;
  rc = on_every_device(mu, [&](size_t r) {
    uint64_t lo, hi;
    shard_range(D, R, r, &lo, &hi);
    if (!mapped) return db_upload_rows(mu->dev[r], enc ? enc + lo * db->W : nullptr, hi - lo, L, subject_offset + lo, nullptr, false, false, &db->shards[r]);
    std::vector<uint64_t> rows((hi - lo) * db->W);
    for (uint64_t i = lo; i < hi; ++i) memcpy(rows.data() + (i - lo) * db->W, enc + (uint64_t)db->perm_host[i] * db->W, db->W * sizeof(uint64_t));
    return db_upload_rows(mu->dev[r], rows.data(), hi - lo, L, subject_offset, db->perm_host.data() + lo, true, false, &db->shards[r]);
  });
  if (rc) { adopt_error(ctx, rc); multi_db_free(db); return rc; }
  *out = db;
  return SMAFA_OK;
}

int multi_db_append(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n) {
  // rows keep their global order when they join the last shard
  SmafaMulti *mu = ctx->multi;
  if (db->subject_offset + db->D + n >= (1ull << 32)) return smafa_fail(ctx, SMAFA_E_UNSUPPORTED, "db larger than 2^32-1 windows");
  int rc = db_append_rows(mu->dev.back(), db->shards.back(), enc, n, db->D);
  if (rc) return adopt_error(ctx, rc);
  if (!db->perm_host.empty())
    for (uint64_t i = 0; i < n; ++i) db->perm_host.push_back((uint32_t)(db->D + i));
  db->D += n;
  return SMAFA_OK;
}

void multi_db_free(smafa_db *db) {
  for (smafa_db *s : db->shards) smafa_db_free(s);
  delete db;
}

int multi_distances(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len, uint16_t *out) {
  SmafaMulti *mu = ctx->multi;
  if (Q == 0 || db->D == 0) return SMAFA_OK;
  if (!q_enc || !out) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_distances: null buffer");
  if (q_len != db->L)
    return smafa_fail(ctx, SMAFA_E_LENGTH_MISMATCH, "Cannot compute distances between seq of length %u and windows of lengths %u", q_len, db->L);
  int rc = on_every_device(mu, [&](size_t r) {
    const smafa_db *sh = db->shards[r];
    if (sh->D == 0) return (int)SMAFA_OK;
    std::vector<uint16_t> part(Q * sh->D);
    int c = smafa_distances(mu->dev[r], sh, q_enc, Q, q_len, part.data());
    if (c) return c;
    if (db->perm_host.empty()) {
      const uint64_t lo = sh->subject_offset - db->subject_offset;
      for (uint64_t q = 0; q < Q; ++q) memcpy(out + q * db->D + lo, part.data() + q * sh->D, sh->D * sizeof(uint16_t));
    } else {  // grouped db: column j of shard r is the window with subject number perm_host[first row of r + j]
      uint64_t first = 0;
      for (size_t o = 0; o < r; ++o) first += db->shards[o]->D;
      for (uint64_t q = 0; q < Q; ++q)
        for (uint64_t j = 0; j < sh->D; ++j) out[q * db->D + db->perm_host[first + j]] = part[q * sh->D + j];
    }
    return (int)SMAFA_OK;
  });
  return adopt_error(ctx, rc);
}

int multi_query(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q, uint32_t q_len, int64_t m, int64_t k,
                smafa_hit **hits, uint64_t *n_hits, smafa_stats *stats) {
  SmafaMulti *mu = ctx->multi;
  const size_t R = mu->dev.size();
  if (db->shards.size() != R) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_query: the db does not belong to this multi-device context");
  QueryPlan whole{};
  int rc = validate_query_plan(ctx, db->D, db->L, Q, q_len, m, k, &whole);
  if (rc || Q == 0) return rc;
  if (!q_enc) return smafa_fail(ctx, SMAFA_E_INVALID, "smafa_query: null query buffer");
  smafa_ctx *c0 = mu->dev[0];
  std::vector<smafa_stats> st(R);
  smafa_hit *all = nullptr;
  uint64_t n_all = 0, cap_all = 0;
  float total_ms = 0;
  for (uint64_t q0 = 0; q0 < Q && rc == SMAFA_OK; q0 += MAX_BATCH_QUERIES) {
    const uint64_t nq = std::min<uint64_t>(MAX_BATCH_QUERIES, Q - q0);
    uint64_t cap = c0->xchg.n_ranks == R && c0->xchg.cap_hint ? c0->xchg.cap_hint : initial_block_cap(nq);
    for (int attempt = 0; rc == SMAFA_OK; ++attempt) {
      // local part: every device on its own host thread (the scan's overflow handling synchronises its stream)
      std::vector<int> local_rc(R, SMAFA_OK);
      rc = on_every_device(mu, [&](size_t r) {
        smafa_ctx *c = mu->dev[r];
        cudaSetDevice(c->device);
        cudaStream_t s = c->stream;
        int e1 = exchange_reserve(c, r == 0 ? (uint32_t)R : 0u, cap, (uint32_t)nq, s);
        if (e1) return e1;
        if ((e1 = ensure_query_words(c, nq * db->W))) return e1;
        if (r == 0) cudaEventRecord(c->ev[2], s);
        cudaError_t e = cudaMemcpyAsync(c->q_ref, q_enc + q0 * db->W, nq * db->W * sizeof(uint64_t), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return smafa_fail(c, SMAFA_E_CUDA, "H2D of queries: %s", cudaGetErrorString(e));
        local_rc[r] = shard_local(c, db->shards[r], c->q_ref, nq, m, k, s, &st[r]);
        if (r == 0) cudaEventRecord(c->xchg.ev[0], s);
        e = cudaEventRecord(mu->done[r], s);
        return e == cudaSuccess ? (int)SMAFA_OK : smafa_fail(c, SMAFA_E_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
      });
      if (rc) { adopt_error(ctx, rc); break; }
      // exchange: the blocks travel to device 0 (peer copies on its stream, each behind its producer's event)
      cudaSetDevice(c0->device);
      cudaStream_t s0 = c0->stream;
      SmafaExchange &x = c0->xchg;
      const uint64_t stride = 2 + cap;
      cudaError_t e = cudaSuccess;
      for (size_t r = 0; r < R && e == cudaSuccess; ++r) {
        if (r) e = cudaStreamWaitEvent(s0, mu->done[r], 0);
        if (e == cudaSuccess)
          e = r == 0 ? cudaMemcpyAsync(x.gathered, x.block, stride * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s0)
                     : cudaMemcpyPeerAsync(x.gathered + r * stride, c0->device, mu->dev[r]->xchg.block, mu->dev[r]->device,
                                           stride * sizeof(uint64_t), s0);
      }
      if (e == cudaSuccess) {
        launch_merge_blocks(x.mw, x.gathered, (uint32_t)R, cap, (uint32_t)nq, merge_k(k), (uint32_t)q0, x.hits, x.mw.rows, x.info_dev, s0);
        cudaEventRecord(x.ev[1], s0);
        e = cudaMemcpyAsync(x.info_host, x.info_dev, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s0);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e != cudaSuccess) { rc = smafa_fail(ctx, SMAFA_E_CUDA, "multi-device query: %s", cudaGetErrorString(e)); break; }
      for (size_t r = 0; r < R && rc == SMAFA_OK; ++r)
        if (local_rc[r]) { rc = local_rc[r]; smafa_fail(ctx, rc, "%s", mu->dev[r]->err.c_str()); }
      if (rc) break;
      const uint64_t need = x.info_host[0];
      if (need > cap) {
        st[0].retries++;
        if (attempt > 8) { rc = smafa_fail(ctx, SMAFA_E_OOM, "multi-device query: block capacity keeps overflowing"); break; }
        cap = pow2_at_least(need + need / 4);
        continue;
      }
      const uint64_t rows = x.info_host[3];
      if ((rc = grow_rows(ctx, all, cap_all, n_all, rows))) break;
      if (rows) {
        e = cudaMemcpyAsync(all + n_all, x.hits, rows * sizeof(smafa_hit), cudaMemcpyDeviceToHost, s0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
        if (e != cudaSuccess) { rc = smafa_fail(ctx, SMAFA_E_CUDA, "D2H of hits: %s", cudaGetErrorString(e)); break; }
      }
      n_all += rows;
      x.cap_hint = adapted_block_cap(cap, need);
      cudaEventRecord(c0->ev[3], s0);
      cudaEventSynchronize(c0->ev[3]);
      float ms = 0;
      cudaEventElapsedTime(&ms, c0->ev[2], c0->ev[3]);
      total_ms += ms;
      cudaEventElapsedTime(&ms, x.ev[0], x.ev[1]);
      st[0].exchange_ms += ms;
      break;
    }
  }
  if (rc) { free(all); return rc; }
  if (stats) {
    *stats = st[0];  // kernel, guess: the first device's; sums and maxima over the devices below
    stats->scan_ms = 0;
    stats->candidates = stats->kernel_launches = 0;
    stats->retries = stats->rescanned = 0;
    for (size_t r = 0; r < R; ++r) {
      stats->scan_ms = std::max(stats->scan_ms, st[r].scan_ms);  // devices scan concurrently
      stats->candidates += st[r].candidates;
      stats->kernel_launches += st[r].kernel_launches;
      stats->retries += st[r].retries;
      stats->rescanned += st[r].rescanned;
    }
    stats->kernel_launches += 5;
    stats->total_ms = total_ms;
    stats->pairs = Q * db->D;
  }
  ctx->last_mma_k = c0->last_mma_k;
  if (!all && !(all = (smafa_hit *)malloc(sizeof(smafa_hit)))) return smafa_fail(ctx, SMAFA_E_OOM, "malloc failed");
  *hits = all;
  *n_hits = n_all;
  return SMAFA_OK;
}
