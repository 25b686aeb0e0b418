/*
 * oracle/smafa_oracle.h -- CPU restatement of wwood/smafa v0.8.0 (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity checker for the B200 path, not part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * Nothing under smafa_b200/ links, imports or executes it.
 *
 * Parity status: PINNED for nucleotide input -- every golden vector the reference's tests
 * hold for this path is reproduced (tests/test_oracle_golden.py): encoding KAT
 * (src/lib.rs:357-366), 9 query KATs + the v1-db error KAT (tests/test_cmdline.rs:9-247),
 * 3 cluster KATs (src/cluster.rs:101-143), count KATs (tests/test_cmdline.rs:183-201) and
 * the two v2 db fixtures byte-for-byte.  The reference itself (Rust) cannot be built here:
 * no cargo/rustc in the image, no vendored crates.
 *
 * All file:line citations are relative to the reference checkout (/root/reference).
 */
#ifndef SMAFA_ORACLE_H
#define SMAFA_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_DB_VERSION 2u /* src/lib.rs:18 */

/* Status codes.  Negative values model the reference's panics (process exit 101) or
 * Err-returns (exit 1); orc_last_error() holds the message the reference would print. */
enum {
  ORC_OK = 0,
  ORC_PANIC = -101, /* Rust panic!  -> stderr message, exit code 101 */
  ORC_ERR = -1      /* Err(..) from main -> "Error: ..." on stderr, exit code 1 */
};

typedef struct {
  uint32_t query, subject, distance;
} orc_hit;

/* A decoded WindowSet (src/lib.rs:54-60), stored flat: n windows x W words. */
typedef struct {
  uint32_t version;
  uint64_t *words; /* [n][W] */
  size_t n, W;
  size_t len; /* window length in symbols; 0 == None */
} orc_windowset;

/* A parsed FASTA/FASTQ file (needletail semantics, [unvendored]). */
typedef struct {
  char **ids;  /* header line without the leading '>' / '@' */
  char **seqs; /* sequence with line endings stripped, NOT upper-cased */
  size_t *lens;
  size_t n;
} orc_fastx;

const char *orc_last_error(void);

/* 0 = nucleotide (the reference, default); 1 = protein, the B200 build's extension -- PARITY UNPINNED
 * (the reference cannot process amino acids at all; see smafa_oracle.c "symbol rules"). */
void orc_set_alphabet(int alphabet);
int orc_get_alphabet(void);

/* src/lib.rs:167-196: byte -> 5-bit one-hot code, 0 = invalid. */
uint8_t orc_encode_single(uint8_t byte);
size_t orc_words_for_len(size_t len); /* ceil(len/12), src/lib.rs:32 */
/* src/lib.rs:29-52.  Returns ORC_OK or ORC_PANIC (invalid byte). */
int orc_encode(const char *id, const uint8_t *seq, size_t len, uint64_t *out_words);
/* src/lib.rs:113-135.  out must hold len bytes.  ORC_PANIC on an invalid 5-bit code. */
int orc_decode(const uint64_t *words, size_t len, char *out);
/* src/lib.rs:80-88: dist[i] = sum_w popcount(db[i][w] ^ q[w]) / 2 */
void orc_distances(const uint64_t *db, size_t n, size_t W, const uint64_t *q, size_t *dist);

/* FASTX input (needletail parse_fastx_file: FASTA or FASTQ, gz sniffed by magic). */
int orc_fastx_read(const char *path, orc_fastx *out);
void orc_fastx_free(orc_fastx *fx);

/* db bytes = postcard::to_allocvec(&WindowSet) (src/lib.rs:162). */
int orc_db_encode(const orc_windowset *ws, uint8_t **bytes, size_t *nbytes);
/* src/lib.rs:208-218 incl. the version gate. */
int orc_db_decode(const uint8_t *bytes, size_t nbytes, orc_windowset *out);
void orc_windowset_free(orc_windowset *ws);

/* src/lib.rs:137-165 */
int orc_makedb(const char *fasta_path, const char *db_path);

/* The selection part of query() on already-encoded input (src/lib.rs:224-317).
 * m, k, r: -1 == None.  Hits come back in the reference's print order with
 * --limit-per-sequence already applied.  q_len is the query window length in symbols
 * (checked against L like src/lib.rs:72-79).  threads>1 splits queries statically across
 * pthreads (a generous variant for the CPU baseline; the reference is single-threaded). */
int orc_query_encoded(const uint64_t *db, size_t D, size_t W, size_t L, const uint64_t *q,
                      size_t Q, size_t q_len, long m, long k, long r, int threads,
                      orc_hit **hits, size_t *n_hits);
/* Full query(): files in, TSV to `out` (src/lib.rs:198-325). */
int orc_query(const char *db_path, const char *query_path, long m, long k, long r, FILE *out);

/* cluster() on encoded, NOT de-duplicated input in file order (src/cluster.rs:22-84).
 * centroid_of[i] = input index of the centroid sequence i was assigned to, or
 * UINT32_MAX when i is a duplicate of an earlier encoding (no output line). */
int orc_cluster_encoded(const uint64_t *enc, size_t n, size_t W, size_t L, uint32_t t,
                        uint32_t *centroid_of, size_t *n_centroids, uint64_t *n_comparisons);
/* Full cluster(): file in, "raw\tdecoded centroid" lines to `out`. */
int orc_cluster(const char *fasta_path, uint32_t t, FILE *out);

/* count() (src/lib.rs:378-398): JSON to `out`. */
int orc_count(const char *const *paths, size_t n_paths, FILE *out);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
