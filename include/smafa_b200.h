/*
 * smafa_b200.h -- C ABI of libsmafa_b200.so, the B200 (sm_100a) drop-in for the query/cluster
 * hot path of wwood/smafa v0.8.0.
 *
 * The reference has no FFI of its own: its hot path is the private method
 * WindowSet::get_distances (src/lib.rs:71-89) plus the selection code inlined in query()
 * (src/lib.rs:224-317) and cluster() (src/cluster.rs:45-74).  Each entry point below names the
 * reference code it replaces; INTEGRATION.md shows the Rust `extern "C"` block and the
 * call-site edits a maintainer would make.
 *
 * Conventions
 *  - Encoded windows use the reference's bit layout (src/lib.rs:29-52): ceil(L/12) u64 words
 *    per window, 5-bit one-hot code of symbol p at bit 5*(p%12) of word p/12.
 *  - "None" for the Option<u32> parameters of smafa::query is passed as -1.
 *  - All pointers are plain host pointers unless the name ends in _dev (then they are CUDA
 *    device pointers on the context's device and `stream` is a cudaStream_t passed as void*).
 *  - Input buffers are borrowed for the duration of the call.  Arrays returned through
 *    smafa_hit** are owned by the library until smafa_free().
 *  - Return value: SMAFA_OK (0) or a negative smafa_status.  No C++ exception crosses the ABI.
 *    smafa_last_error() gives a message; for the three statuses that model reference panics
 *    it is the reference's panic text.
 *  - A context is single-owner, not re-entrant (the reference is single-threaded).  A context made by
 *    smafa_ctx_create drives one GPU; smafa_ctx_create_multi makes one that row-shards every db over several GPUs
 *    of this process and is used through the same calls.  Launchers that start one process per GPU (torchrun, MPI)
 *    use smafa_ctx_comm_init + smafa_db_upload_shard + smafa_query_sharded instead.
 *  - There is no CPU fallback: without a usable sm_100 device every compute entry point
 *    returns SMAFA_E_CUDA.
 */
#ifndef SMAFA_B200_H
#define SMAFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMAFA_B200_ABI_VERSION 4 /* 2: smafa_stats gained guess_bound and rescanned; smafa_*_file_on_device.  3 (additive): smafa_ctx_last_mma_k.  4: multi-GPU entry points (smafa_ctx_create_multi, smafa_ctx_comm_init, smafa_db_upload_shard, smafa_query_sharded[_dev], smafa_*_file_on_devices); the round-1 probe hooks smafa_debug_mma_rate / smafa_debug_sparse_decode are gone */

typedef enum smafa_status {
  SMAFA_OK = 0,
  SMAFA_E_LENGTH_MISMATCH = -1, /* src/lib.rs:72-79 panic */
  SMAFA_E_EMPTY_DB = -2,        /* src/lib.rs:254,298 Option::unwrap() on an empty db */
  SMAFA_E_BAD_K = -3,           /* src/lib.rs:255 max_num_hits == 0 underflow */
  SMAFA_E_LIMIT_NEEDS_K = -4,   /* src/lib.rs:301-303 limit_per_sequence without max_num_hits>1 */
  SMAFA_E_CUDA = -10,
  SMAFA_E_OOM = -11,
  SMAFA_E_INVALID = -12,     /* bad argument (null pointer, L == 0 with D > 0, ...) */
  SMAFA_E_UNSUPPORTED = -13, /* L > 4095 or D >= 2^32 */
  SMAFA_E_NCCL = -14,        /* NCCL unavailable or a collective failed (one-process-per-GPU runs only) */
  SMAFA_E_PEER = -15,        /* another rank of a sharded query failed; its status is in the message */
  SMAFA_E_IO = -20,          /* host file API: Err(..) in the reference (exit code 1) */
  SMAFA_E_PANIC = -21        /* host file API: any other reference panic (exit code 101) */
} smafa_status;

/* Scan kernel formulations (north_star): both are kept, `AUTO` picks the measured default. */
typedef enum smafa_kernel {
  SMAFA_KERNEL_AUTO = 0,
  SMAFA_KERNEL_POPC = 1, /* CUDA-core XOR/OR + POPC over the bit-plane re-packing */
  SMAFA_KERNEL_MMA = 2   /* tcgen05 int8 MMA over one-hot operands, TMEM-drain epilogue */
} smafa_kernel;

/* Alphabets.  NUCLEOTIDE is the reference (src/lib.rs:167-184).  PROTEIN is an EXTENSION of this
 * build -- the reference panics on amino-acid bytes (src/lib.rs:35-42), so there is no reference
 * parity for it: same word geometry (12 five-bit groups per u64), but every group holds
 * a symbol NUMBER (1..20 = ACDEFGHIKLMNPQRSTVWY, 21 = X/B/Z/J/U/O, 22 = '-', 23 = '*') instead of a
 * one-hot code, and the distance is the number of positions whose symbols differ (X-X, gap-gap match,
 * mirroring the reference's N rule). */
typedef enum smafa_alphabet {
  SMAFA_ALPHABET_NUCLEOTIDE = 0,
  SMAFA_ALPHABET_PROTEIN = 1
} smafa_alphabet;

typedef struct smafa_ctx smafa_ctx;
typedef struct smafa_db smafa_db;

/* One output row of `smafa query`: the first three TSV columns (src/lib.rs:292,310). */
typedef struct smafa_hit {
  uint32_t query;    /* 0-based query number */
  uint32_t subject;  /* db window index (global: includes the shard's subject_offset) */
  uint32_t distance; /* number of differing positions */
} smafa_hit;

/* Per-call statistics (optional out-parameter, may be NULL). */
typedef struct smafa_stats {
  uint64_t pairs;          /* Q x D comparisons covered */
  uint64_t candidates;     /* rows the scan emitted before the final cutoff */
  uint64_t kernel_launches;/* kernels launched by this call */
  uint32_t retries;        /* candidate-buffer overflows that forced a smaller query batch */
  uint32_t kernel_used;    /* smafa_kernel actually run */
  float scan_ms;           /* device time of the scan kernels (CUDA events) */
  float total_ms;          /* device time of the whole call */
  int32_t guess_bound;     /* bound of the optimistic first pass of the last batch (-1: none; csrc/guess.cu) */
  uint32_t rescanned;      /* queries the first pass left unfinished (scanned again under the caller's bound) */
  float exchange_ms;       /* sharded queries: device time from the end of this shard's local part to the end of the
                              merge (block exchange + merge + the wait for the slowest shard); 0 otherwise */
  uint32_t union_degree;   /* db windows per accumulator of the last tcgen05 scan (1 = single-window operands; 0 = no such scan) */
} smafa_stats;

int smafa_abi_version(void);
const char *smafa_status_name(int status);

/* ---- context ------------------------------------------------------------------------- */
int smafa_ctx_create(smafa_ctx **ctx, int device, int kernel /* smafa_kernel */);
void smafa_ctx_destroy(smafa_ctx *ctx);
/* Message for the last failure on this context (or the last failure of ctx-less calls when
 * ctx == NULL).  Never NULL. */
const char *smafa_last_error(const smafa_ctx *ctx);
/* One context over several GPUs of this process: smafa_db_upload row-shards the db into contiguous
 * ranges, one per device (global subject numbers are kept, so print order survives), smafa_query / smafa_query_file
 * scan all shards concurrently (one host thread per device), move every shard's answer to the first device over
 * NVLink and merge there (csrc/merge.cu).  smafa_cluster's greedy is strictly sequential (src/cluster.rs:45-74) and
 * runs on the first device.  Device pointers belong to one device: the *_dev calls refuse such a context.
 * n_devices == 1 is exactly smafa_ctx_create. */
int smafa_ctx_create_multi(smafa_ctx **ctx, const int *devices, int n_devices, int kernel /* smafa_kernel */);
int smafa_ctx_device_count(const smafa_ctx *ctx);
int smafa_ctx_set_kernel(smafa_ctx *ctx, int kernel);
/* Alphabet of the dbs uploaded from now on (a db keeps the alphabet it was uploaded with; queries are
 * read in their db's alphabet), of smafa_cluster input and of the *_file calls on this context. */
int smafa_ctx_set_alphabet(smafa_ctx *ctx, int alphabet /* smafa_alphabet */);
/* Candidate-buffer capacity in rows (testing hook for the overflow/retry path; 0 = default). */
int smafa_ctx_set_candidate_capacity(smafa_ctx *ctx, uint64_t rows);

/* ---- db: replaces the in-memory WindowSet built by query() (src/lib.rs:208-218) ---------
 * Re-packs D windows of length L into the GPU-resident matrices (bit planes for the POPC
 * kernel, one-hot int8 tiles for the MMA kernel, reference words for exact verification and
 * output).  subject_offset is added to every reported subject index (row-sharding across
 * GPUs: SURVEY.md 8e).  D == 0 is allowed (queries then fail with SMAFA_E_EMPTY_DB, like
 * the reference). */
int smafa_db_upload(smafa_ctx *ctx, const uint64_t *enc /* [D][ceil(L/12)] */, uint64_t D,
                    uint32_t L, uint64_t subject_offset, smafa_db **db);
/* Similarity-grouped order of a db (csrc/api.cu group_order): perm_out[r] = index of the window to store at row r, such
 * that near-copies are neighbours and groups start at multiples of 16 rows where possible -- the order in which one
 * tcgen05 accumulator can filter up to 16 db windows at once ("union rows").  smafa_db_upload applies it by
 * itself (SMAFA_DB_GROUP=0 switches that off); callers that shard a db over processes call this once on the WHOLE db,
 * cut the grouped order into their shards and upload them with smafa_db_upload_mapped, so that every shard holds whole
 * groups.  *n_clusters = 0 and the identity when the db has too little structure (or is small, protein, longer than
 * 63).  Results never depend on the order: rows are reported under their subject numbers. */
int smafa_group_order(smafa_ctx *ctx, const uint64_t *enc, uint64_t D, uint32_t L, uint32_t *perm_out /* [D] */,
                      uint64_t *n_clusters);
/* The layout step of smafa_group_order alone (host only, no GPU): from a clustering -- centroid_of[i] = the window that
 * founded window i's cluster, itself for a founder, as smafa_cluster reports it -- to the row order: clusters contiguous,
 * members in input order, clusters arranged so that they start on multiples of 16 rows where their sizes allow. */
int smafa_group_layout(const uint32_t *centroid_of, uint64_t D, uint32_t *perm_out /* [D] */);
/* Rows in a caller-chosen order with explicit subject numbers: row r is reported as subject subjects[r] (< D_total, the
 * rows of the whole db this one is a part of).  grouped != 0 states that the order is a similarity-grouped one. */
int smafa_db_upload_mapped(smafa_ctx *ctx, const uint64_t *enc /* [D][ceil(L/12)] */, uint64_t D, uint32_t L,
                           const uint32_t *subjects /* [D] */, uint64_t D_total, int grouped, smafa_db **db);
/* Appends rows (used by cluster: the centroid set grows, src/cluster.rs:69-74). */
int smafa_db_append(smafa_ctx *ctx, smafa_db *db, const uint64_t *enc, uint64_t n);
uint64_t smafa_db_size(const smafa_db *db);
uint32_t smafa_db_window_len(const smafa_db *db);
void smafa_db_free(smafa_db *db);

/* ---- get_distances (src/lib.rs:71-89), for parity/debug: out[q*D + i] ------------------ */
int smafa_distances(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q,
                    uint32_t q_len, uint16_t *out /* [Q][D] */);

/* ---- query: get_distances + selection (src/lib.rs:238-314) -----------------------------
 * max_divergence / max_num_hits: -1 == None.  None or 1 => "Mode A" (all ties at the
 * minimum, ascending subject); any other k => "Mode B" (everything <= the k-th smallest
 * distance, ties included, in (distance, subject) order).  Hits are sorted in the
 * reference's print order.  --limit-per-sequence is a host-side run-length filter over this
 * list (smafa_apply_limit_per_sequence). */
int smafa_query(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q,
                uint32_t q_len, int64_t max_divergence, int64_t max_num_hits, smafa_hit **hits,
                uint64_t *n_hits, smafa_stats *stats);

/* Same, with queries already in device memory and hits left in a caller-provided device
 * buffer (capacity in rows).  *n_hits receives the number of rows the answer has; when it
 * exceeds hits_capacity the call returns SMAFA_E_OOM and nothing useful is in hits_dev.
 * Work is enqueued on `stream` (NULL = the legacy default stream, as everywhere in CUDA), so it is
 * ordered after whatever produced q_enc_dev on that stream; the call returns after the stream has
 * been synchronised. */
int smafa_query_dev(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc_dev, uint64_t Q,
                    uint32_t q_len, int64_t max_divergence, int64_t max_num_hits,
                    smafa_hit *hits_dev, uint64_t hits_capacity, uint64_t *n_hits,
                    void *stream, smafa_stats *stats);

/* Multi-GPU merge: applies the reference's selection to the union of the
 * per-shard candidate lists (after the NCCL all-gather).  In-place on a device buffer of n
 * rows; *n_out rows remain, sorted in print order.  *n_out is final on return; the rows themselves are
 * stream-ordered (valid for work queued on `stream` after the call, or after synchronising it). */
int smafa_merge_dev(smafa_ctx *ctx, smafa_hit *cands_dev, uint64_t n, int64_t max_divergence,
                    int64_t max_num_hits, uint64_t *n_out, void *stream);

/* ---- sharded query, one process per GPU ------------------------------------
 * Every rank holds a context on its GPU and one contiguous row range of the db; queries are replicated.  A call is
 * COLLECTIVE: every rank makes it with the same queries and parameters, every rank receives the complete answer
 * (identical rows, print order, global subject numbers).  Per <= 2^20 queries: local scan + selection, ONE
 * ncclAllGather of fixed-capacity blocks on the call's stream, a sort-free merge (csrc/merge.cu), one host
 * read-back.  A rank whose local part fails still joins the exchange, so all ranks return (the others with
 * SMAFA_E_PEER).  NCCL is loaded at run time (libnccl.so.2); nothing else in this header needs it.
 *   rank 0:  smafa_comm_unique_id(id);   broadcast id by whatever means the launcher has (a file, MPI, torch.distributed)
 *   all:     smafa_ctx_comm_init(ctx, id, rank, world_size);
 *            smafa_db_upload_shard(ctx, rows [lo, hi) of the db, hi - lo, L, lo, D_total, &db);
 *            smafa_query_sharded(ctx, db, queries, ...);            (or _dev: device pointers in and out) */
#define SMAFA_COMM_ID_BYTES 128
int smafa_comm_unique_id(uint8_t *id /* [SMAFA_COMM_ID_BYTES] */);
int smafa_ctx_comm_init(smafa_ctx *ctx, const uint8_t *id /* [SMAFA_COMM_ID_BYTES] */, int rank, int world_size);
void smafa_ctx_comm_free(smafa_ctx *ctx);
int smafa_db_upload_shard(smafa_ctx *ctx, const uint64_t *enc /* [D][ceil(L/12)]: this shard's rows */, uint64_t D,
                          uint32_t L, uint64_t subject_offset /* first global row of the shard */,
                          uint64_t D_total /* rows of the whole db */, smafa_db **db);
int smafa_query_sharded(smafa_ctx *ctx, const smafa_db *db_shard, const uint64_t *q_enc, uint64_t Q, uint32_t q_len,
                        int64_t max_divergence, int64_t max_num_hits, smafa_hit **hits, uint64_t *n_hits,
                        smafa_stats *stats);
int smafa_query_sharded_dev(smafa_ctx *ctx, const smafa_db *db_shard, const uint64_t *q_enc_dev, uint64_t Q,
                            uint32_t q_len, int64_t max_divergence, int64_t max_num_hits, smafa_hit *hits_dev,
                            uint64_t hits_capacity, uint64_t *n_hits, void *stream, smafa_stats *stats);

/* --limit-per-sequence (src/lib.rs:259-260,269-289): run-length cap over consecutive hits of
 * one query whose subjects have identical encodings.  In-place on a host array; returns the
 * new length.  db_enc is the host copy of the db words ([D][W]). */
uint64_t smafa_apply_limit_per_sequence(smafa_hit *hits, uint64_t n, const uint64_t *db_enc,
                                        uint32_t W, uint64_t subject_offset, uint32_t limit);

/* ---- cluster (src/cluster.rs:45-74) -----------------------------------------------------
 * enc: n encodings in input order, duplicates already removed (the HashSet step,
 * src/cluster.rs:46-48, stays on the host).  centroid_of[i] receives the input index of the
 * centroid sequence i joins (itself when it founds a new cluster).  The distance evaluation
 * is batched on the GPU; the order-dependent greedy assignment is replayed exactly.
 * n_comparisons (optional) receives sum_i |centroids before i|, the reference's work count. */
int smafa_cluster(smafa_ctx *ctx, const uint64_t *enc, uint64_t n, uint32_t L,
                  uint32_t max_divergence, uint32_t *centroid_of, uint64_t *n_centroids,
                  uint64_t *n_comparisons, smafa_stats *stats);

void smafa_free(void *p);

/* ---- host-side mirror of the reference's public functions (no GPU needed for makedb/count)
 * Same argument meaning as smafa::makedb / query / cluster / count (src/lib.rs:137,198,378;
 * src/cluster.rs:13).  Output goes to the given file descriptor.  On failure the return code
 * says whether the reference would have exited 1 (SMAFA_E_IO) or panicked (any other code);
 * smafa_last_error(NULL) holds the message. */
int smafa_makedb_file(const char *subject_fasta, const char *db_path);
int smafa_makedb_file_alphabet(const char *subject_fasta, const char *db_path, int alphabet);
int smafa_query_file(smafa_ctx *ctx, const char *db_path, const char *query_fasta,
                     int64_t max_divergence, int64_t max_num_hits, int64_t limit_per_sequence,
                     int out_fd);
/* Host helper of cluster: first[i] = 1 iff no earlier window has the encoding of window i -- the
 * HashSet<Vec<u64>> de-duplication of src/cluster.rs:24,46-48, hash-partitioned over the host threads
 * (the result does not depend on their number).  smafa_cluster expects exactly the windows with first[i] = 1. */
int smafa_mark_first_occurrences(const uint64_t *words, uint64_t n, uint32_t W, uint8_t *first);
int smafa_cluster_file(smafa_ctx *ctx, const char *input_fasta, uint32_t max_divergence, int out_fd);
/* The same two commands for a process that has no context yet (the CLI): the context for `device` is created on a
 * helper thread while the files are read, decoded and encoded -- CUDA initialisation takes 1-3 s and does not
 * depend on the inputs.  `kernel` is a smafa_kernel, `alphabet` a smafa_alphabet.  *ctx_out (may be NULL) receives
 * the context for error reporting and smafa_ctx_destroy; a device that cannot be initialised is an error even
 * when the inputs turned out to need no device work. */
int smafa_query_file_on_device(int device, int kernel, int alphabet, const char *db_path, const char *query_fasta,
                               int64_t max_divergence, int64_t max_num_hits, int64_t limit_per_sequence,
                               int out_fd, smafa_ctx **ctx_out);
int smafa_cluster_file_on_device(int device, int kernel, int alphabet, const char *input_fasta,
                                 uint32_t max_divergence, int out_fd, smafa_ctx **ctx_out);
/* The same over a list of GPUs (the CLI's --devices): `query` runs on a multi-device context (smafa_ctx_create_multi:
 * db row-sharded over the devices); `cluster` is sequential by definition and uses the first device only. */
int smafa_query_file_on_devices(const int *devices, int n_devices, int kernel, int alphabet, const char *db_path,
                                const char *query_fasta, int64_t max_divergence, int64_t max_num_hits,
                                int64_t limit_per_sequence, int out_fd, smafa_ctx **ctx_out);
int smafa_cluster_file_on_devices(const int *devices, int n_devices, int kernel, int alphabet, const char *input_fasta,
                                  uint32_t max_divergence, int out_fd, smafa_ctx **ctx_out);
int smafa_count_files(const char *const *paths, size_t n_paths, int out_fd);
/* Loads a db file into host words (src/lib.rs:206-218: File::open, version gate, postcard decode) -- the
 * LEB128 stream is decoded on all host threads.  *words ([n][W], reference bit layout) is released with
 * smafa_free(); window_len 0 == None (empty db).  No GPU needed. */
int smafa_db_file_load(const char *db_path, uint64_t **words, uint64_t *n, uint32_t *W, uint32_t *window_len);
/* Opens a db file and applies the version gate of src/lib.rs:208-217 (no GPU needed). */
int smafa_db_file_check(const char *db_path);

/* Encoding helpers (src/lib.rs:29-52, 113-135, 167-196). */
uint8_t smafa_encode_symbol(uint8_t byte); /* 0 == not a nucleotide */
int smafa_encode_window(const uint8_t *seq, size_t len, uint64_t *out_words, size_t *bad_pos);
int smafa_decode_window(const uint64_t *words, size_t len, char *out);
uint8_t smafa_encode_symbol_alphabet(uint8_t byte, int alphabet);
int smafa_encode_window_alphabet(const uint8_t *seq, size_t len, uint64_t *out_words, size_t *bad_pos, int alphabet);
int smafa_decode_window_alphabet(const uint64_t *words, size_t len, char *out, int alphabet);

/* ---- debug / parity hooks -------------------------------------------------------------
 * Raw int32 accumulators of the tcgen05 formulation for the first 128 db rows x (up to) 256
 * queries.  For the 5-symbol one-hot operands out[row*256 + col] = matches(row, col) - (L - bound);
 * the other operand encodings are documented in csrc/scan_mma.cu ("operand encodings").  Used by the
 * tests to pin the operand layout and descriptors of the MMA kernel independently of its epilogue. */
int smafa_debug_mma_dump(smafa_ctx *ctx, const smafa_db *db, const uint64_t *q_enc, uint64_t Q,
                         uint32_t bound, int32_t *out /* [128][256] */);
/* Measured dense int8 tcgen05 rate of this GPU in TOP/s (roofline denominator of the MMA kernel):
 * every SM issues mmas_per_cta back-to-back M128 x N256 x K32 kind::i8 MMAs on resident operands. */
int smafa_debug_mma_peak(smafa_ctx *ctx, uint32_t mmas_per_cta, double *tops);
/* Contraction depth K (int8 elements per window) of the tcgen05 operands of this db, 0 if the db is
 * not eligible for the MMA kernel: executed int8 ops per comparison = 2 * K. */
uint32_t smafa_db_mma_k(const smafa_db *db);
/* The same figure for the operands the LAST tcgen05 scan of this context actually used (union-row operands hold two
 * windows per row: K / 2 per window); 0 before the first such scan. */
uint32_t smafa_ctx_last_mma_k(const smafa_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SMAFA_B200_H */
