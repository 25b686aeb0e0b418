/*
 * oracle/smafa_oracle.c -- CPU restatement of wwood/smafa v0.8.0 (TEST INFRASTRUCTURE ONLY).
 * See smafa_oracle.h for the parity status and the rules on who may call this.
 * Plain C11 + zlib + pthreads.  Every function cites the reference file:line it follows.
 */
#define _GNU_SOURCE
#include "smafa_oracle.h"

#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static __thread char g_err[1024];

const char *orc_last_error(void) { return g_err; }

static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------ symbol rules */

/* Alphabet switch.  0 = nucleotide = the reference.  1 = protein: an EXTENSION of the B200 build
 * (the reference panics on amino-acid bytes, src/lib.rs:35-42), PARITY UNPINNED -- there is no
 * reference behaviour to pin it to.  Definition: symbol numbers 1..20 = ACDEFGHIKLMNPQRSTVWY,
 * 21 = X/B/Z/J/U/O, 22 = '-', 23 = '*' in the same 5-bit groups; distance = number of positions whose
 * symbols differ.  Process-wide (set before any threads are started). */
static int g_alphabet = 0;
void orc_set_alphabet(int alphabet) { g_alphabet = alphabet; }
int orc_get_alphabet(void) { return g_alphabet; }

static uint8_t aa_encode_single(uint8_t b) {
  static const char aa[] = "ACDEFGHIKLMNPQRSTVWY";
  if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
  for (int i = 0; aa[i]; ++i)
    if (b == (uint8_t)aa[i]) return (uint8_t)(i + 1);
  switch (b) {
    case 'X': case 'B': case 'Z': case 'J': case 'U': case 'O': return 21;
    case '-': return 22;
    case '*': return 23;
    default: return 0;
  }
}

/* src/lib.rs:167-184 (create_lut) + src/lib.rs:190-196 (encode_single) */
uint8_t orc_encode_single(uint8_t b) {
  if (g_alphabet) return aa_encode_single(b);
  switch (b) {
    case 'A': case 'a': return 0x10;
    case 'C': case 'c': return 0x08;
    case 'G': case 'g': return 0x04;
    case 'T': case 't': case 'U': case 'u': return 0x02;
    case 'N': case 'W': case 'S': case 'M': case 'K': case 'R': case 'Y': case 'B':
    case 'D': case 'H': case 'V': case '-':
    case 'n': case 'w': case 's': case 'm': case 'k': case 'r': case 'y': case 'b':
    case 'd': case 'h': case 'v': return 0x01;
    default: return 0;
  }
}

size_t orc_words_for_len(size_t len) { return (len + 11) / 12; }

/* src/lib.rs:29-52: chunks of 12 symbols, code i of a chunk at bit 5*i. */
int orc_encode(const char *id, const uint8_t *seq, size_t len, uint64_t *out) {
  size_t W = orc_words_for_len(len);
  for (size_t w = 0; w < W; ++w) out[w] = 0;
  for (size_t p = 0; p < len; ++p) {
    uint8_t c = orc_encode_single(seq[p]);
    if (!c)
      return fail(ORC_PANIC,
                  "Byte %u cannot be interpreted as %s, in sequence \"%s\" at position %zu",
                  (unsigned)seq[p], g_alphabet ? "amino acid" : "nucleotide", id ? id : "", p);
    out[p / 12] |= (uint64_t)c << (5 * (p % 12));
  }
  return ORC_OK;
}

/* src/lib.rs:113-135 */
int orc_decode(const uint64_t *words, size_t len, char *out) {
  for (size_t i = 0; i < len; ++i) {
    unsigned b = (unsigned)((words[i / 12] >> (5 * (i % 12))) & 31u);
    if (g_alphabet) {
      static const char aa[] = "?ACDEFGHIKLMNPQRSTVWYX-*";
      if (b < 1 || b > 23) return fail(ORC_PANIC, "Invalid character in query sequence: %u", b);
      out[i] = aa[b];
      continue;
    }
    switch (b) {
      case 0x10: out[i] = 'A'; break;
      case 0x08: out[i] = 'C'; break;
      case 0x04: out[i] = 'G'; break;
      case 0x02: out[i] = 'T'; break;
      case 0x01: out[i] = 'N'; break;
      default: return fail(ORC_PANIC, "Invalid character in query sequence: %u", b);
    }
  }
  return ORC_OK;
}

/* src/lib.rs:80-88 */
void orc_distances(const uint64_t *db, size_t n, size_t W, const uint64_t *q, size_t *dist) {
  for (size_t i = 0; i < n; ++i) {
    const uint64_t *w = db + i * W;
    size_t s = 0;
    if (g_alphabet) { /* protein extension: positions whose 5-bit symbols differ */
      for (size_t j = 0; j < W; ++j)
        for (int g = 0; g < 12; ++g) s += (((w[j] ^ q[j]) >> (5 * g)) & 31u) != 0;
      dist[i] = s;
      continue;
    }
    for (size_t j = 0; j < W; ++j) s += (size_t)__builtin_popcountll(w[j] ^ q[j]);
    dist[i] = s / 2;
  }
}

/* ------------------------------------------------------------------ FASTX (needletail) */

static int slurp(const char *path, uint8_t **buf, size_t *n) {
  /* gzopen reads plain files transparently and inflates gzip ones (needletail sniffs the
   * magic bytes the same way; bz2/xz inputs are not supported by this oracle). */
  FILE *probe = fopen(path, "rb");
  if (!probe) return fail(ORC_ERR, "Os { code: %d, kind: NotFound, message: \"%s\" }", errno, strerror(errno));
  fclose(probe);
  gzFile f = gzopen(path, "rb");
  if (!f) return fail(ORC_ERR, "cannot open %s", path);
  size_t cap = 1 << 16, len = 0;
  uint8_t *b = malloc(cap);
  for (;;) {
    if (cap - len < (1 << 15)) b = realloc(b, cap *= 2);
    int r = gzread(f, b + len, (unsigned)(cap - len));
    if (r < 0) { gzclose(f); free(b); return fail(ORC_ERR, "read error on %s", path); }
    if (r == 0) break;
    len += (size_t)r;
  }
  gzclose(f);
  *buf = b;
  *n = len;
  return ORC_OK;
}

static void fx_push(orc_fastx *fx, size_t *cap, const uint8_t *id, size_t idn, const uint8_t *s, size_t sn) {
  if (fx->n == *cap) {
    *cap = *cap ? *cap * 2 : 64;
    fx->ids = realloc(fx->ids, *cap * sizeof *fx->ids);
    fx->seqs = realloc(fx->seqs, *cap * sizeof *fx->seqs);
    fx->lens = realloc(fx->lens, *cap * sizeof *fx->lens);
  }
  char *i = malloc(idn + 1);
  memcpy(i, id, idn);
  i[idn] = 0;
  char *q = malloc(sn + 1);
  size_t m = 0;
  for (size_t k = 0; k < sn; ++k)
    if (s[k] != '\n' && s[k] != '\r') q[m++] = (char)s[k];
  q[m] = 0;
  fx->ids[fx->n] = i;
  fx->seqs[fx->n] = q;
  fx->lens[fx->n] = m;
  fx->n++;
}

/* needletail::parse_fastx_file [unvendored]: '>' => FASTA (multi-line allowed, id = whole
 * header line), '@' => FASTQ (4-line records).  Pinned by every FASTA-driven reference test,
 * the record without trailing newline in tests/data/subjects.fa and the gz FASTQ in
 * tests/test_cmdline.rs:193-201. */
int orc_fastx_read(const char *path, orc_fastx *fx) {
  memset(fx, 0, sizeof *fx);
  uint8_t *b = NULL;
  size_t n = 0;
  int rc = slurp(path, &b, &n);
  if (rc) return rc;
  size_t cap = 0, p = 0;
  if (n == 0) { free(b); return fail(ORC_PANIC, "valid path/file: EmptyFile"); }
  if (b[0] == '>') {
    while (p < n) {
      if (b[p] != '>') { free(b); return fail(ORC_PANIC, "valid record: InvalidStart"); }
      size_t hs = p + 1, he = hs;
      while (he < n && b[he] != '\n') he++;
      size_t idn = he - hs;
      if (idn && b[hs + idn - 1] == '\r') idn--;
      size_t ss = he < n ? he + 1 : n, se = ss;
      /* the record ends at the next '>' that starts a line */
      while (se < n && !(b[se] == '>' && (se == ss || b[se - 1] == '\n'))) se++;
      fx_push(fx, &cap, b + hs, idn, b + ss, se - ss);
      p = se;
    }
  } else if (b[0] == '@') {
    while (p < n) {
      if (b[p] == '\n' || b[p] == '\r') { p++; continue; }
      if (b[p] != '@') { free(b); return fail(ORC_PANIC, "valid record: InvalidStart"); }
      size_t ls[4], le[4];
      for (int l = 0; l < 4; ++l) {
        ls[l] = p;
        while (p < n && b[p] != '\n') p++;
        le[l] = p;
        if (le[l] > ls[l] && b[le[l] - 1] == '\r') le[l]--;
        if (p < n) p++;
      }
      fx_push(fx, &cap, b + ls[0] + 1, le[0] - ls[0] - 1, b + ls[1], le[1] - ls[1]);
    }
  } else {
    free(b);
    return fail(ORC_PANIC, "valid path/file: InvalidStart");
  }
  free(b);
  return ORC_OK;
}

void orc_fastx_free(orc_fastx *fx) {
  for (size_t i = 0; i < fx->n; ++i) { free(fx->ids[i]); free(fx->seqs[i]); }
  free(fx->ids); free(fx->seqs); free(fx->lens);
  memset(fx, 0, sizeof *fx);
}

/* ------------------------------------------------------------------ db bytes (postcard) */

typedef struct { uint8_t *b; size_t n, cap; } bytebuf;

static void bb_put(bytebuf *bb, uint8_t v) {
  if (bb->n == bb->cap) bb->b = realloc(bb->b, bb->cap = bb->cap ? bb->cap * 2 : 256);
  bb->b[bb->n++] = v;
}

/* postcard varint = unsigned LEB128 (pinned by the .smafadb fixtures in tests/data) */
static void put_varint(bytebuf *bb, uint64_t v) {
  while (v >= 0x80) { bb_put(bb, (uint8_t)(v | 0x80)); v >>= 7; }
  bb_put(bb, (uint8_t)v);
}

static int get_varint(const uint8_t *b, size_t n, size_t *p, int max_bytes, uint64_t *out) {
  uint64_t v = 0;
  for (int i = 0; i < max_bytes; ++i) {
    if (*p >= n) return fail(ORC_ERR, "DeserializeUnexpectedEnd");
    uint8_t c = b[(*p)++];
    v |= (uint64_t)(c & 0x7f) << (7 * i);
    if (!(c & 0x80)) { *out = v; return ORC_OK; }
  }
  return fail(ORC_ERR, "DeserializeBadVarint");
}

/* src/lib.rs:162: version, windows (len + per window len + words), Option<NonZeroUsize> */
int orc_db_encode(const orc_windowset *ws, uint8_t **bytes, size_t *nbytes) {
  bytebuf bb = {0};
  put_varint(&bb, ws->version);
  put_varint(&bb, ws->n);
  for (size_t i = 0; i < ws->n; ++i) {
    put_varint(&bb, ws->W);
    for (size_t w = 0; w < ws->W; ++w) put_varint(&bb, ws->words[i * ws->W + w]);
  }
  if (ws->len) { bb_put(&bb, 1); put_varint(&bb, ws->len); }
  else bb_put(&bb, 0);
  *bytes = bb.b;
  *nbytes = bb.n;
  return ORC_OK;
}

/* src/lib.rs:212-218 */
int orc_db_decode(const uint8_t *b, size_t n, orc_windowset *ws) {
  memset(ws, 0, sizeof *ws);
  if (n < 4) return fail(ORC_PANIC, "range end index 4 out of range for slice of length %zu", n);
  size_t p = 0;
  uint64_t v = 0;
  int rc = get_varint(b, 4, &p, 5, &v); /* version parsed from buffer[0..4] only */
  if (rc) return rc;
  if (v != ORC_DB_VERSION)
    return fail(ORC_PANIC,
                "Unsupported db file version: %llu. This version of smafa only works with version %u "
                "databases. The last version to support version 1 databases was v0.7.1.",
                (unsigned long long)v, ORC_DB_VERSION);
  ws->version = (uint32_t)v;
  uint64_t cnt = 0;
  if ((rc = get_varint(b, n, &p, 10, &cnt))) return rc;
  ws->n = (size_t)cnt;
  size_t cap = 0;
  for (size_t i = 0; i < ws->n; ++i) {
    uint64_t w = 0;
    if ((rc = get_varint(b, n, &p, 10, &w))) { orc_windowset_free(ws); return rc; }
    if (i == 0) {
      ws->W = (size_t)w;
      cap = ws->n * ws->W;
      ws->words = malloc((cap ? cap : 1) * sizeof(uint64_t));
    } else if (w != ws->W) {
      orc_windowset_free(ws);
      return fail(ORC_ERR, "oracle: ragged db (window %zu has %llu words, expected %zu)", i,
                  (unsigned long long)w, ws->W);
    }
    for (size_t j = 0; j < ws->W; ++j)
      if ((rc = get_varint(b, n, &p, 10, &ws->words[i * ws->W + j]))) { orc_windowset_free(ws); return rc; }
  }
  if (p >= n) { orc_windowset_free(ws); return fail(ORC_ERR, "DeserializeUnexpectedEnd"); }
  uint8_t tag = b[p++];
  if (tag == 1) {
    uint64_t l = 0;
    if ((rc = get_varint(b, n, &p, 10, &l))) { orc_windowset_free(ws); return rc; }
    ws->len = (size_t)l;
  } else if (tag != 0) {
    orc_windowset_free(ws);
    return fail(ORC_ERR, "DeserializeBadOption");
  }
  return ORC_OK;
}

void orc_windowset_free(orc_windowset *ws) {
  free(ws->words);
  memset(ws, 0, sizeof *ws);
}

static int read_file(const char *path, uint8_t **buf, size_t *n) {
  FILE *f = fopen(path, "rb");
  if (!f) return fail(ORC_ERR, "Os { code: %d, kind: NotFound, message: \"%s\" }", errno, strerror(errno));
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  *buf = malloc(sz > 0 ? (size_t)sz : 1);
  *n = fread(*buf, 1, (size_t)sz, f);
  fclose(f);
  return ORC_OK;
}

/* Encode the records of a FASTX file into a WindowSet (src/lib.rs:147-152, push_encoding
 * :91-111).  On failure the WindowSet keeps the records that were pushed before it. */
static int encode_fastx(const orc_fastx *fx, uint32_t version, orc_windowset *ws) {
  memset(ws, 0, sizeof *ws);
  ws->version = version;
  for (size_t i = 0; i < fx->n; ++i) {
    size_t len = fx->lens[i];
    if (i == 0) {
      ws->W = orc_words_for_len(len);
      ws->words = malloc((fx->n * ws->W + 1) * sizeof(uint64_t));
    }
    uint64_t tmp[(len + 11) / 12 + 1];
    int rc = orc_encode(fx->ids[i], (const uint8_t *)fx->seqs[i], len, tmp);
    if (rc) return rc;
    if (ws->len) {
      if (ws->len != len)
        return fail(ORC_PANIC, "WindowSet seq length is %zu, got a new sequence of length %zu", ws->len, len);
    } else {
      if (len == 0) return fail(ORC_PANIC, "Cannot add empty sequence to WindowSet");
      ws->len = len;
    }
    memcpy(ws->words + ws->n * ws->W, tmp, ws->W * sizeof(uint64_t));
    ws->n++;
  }
  return ORC_OK;
}

/* src/lib.rs:137-165 */
int orc_makedb(const char *fasta_path, const char *db_path) {
  orc_fastx fx;
  int rc = orc_fastx_read(fasta_path, &fx);
  if (rc) return rc;
  orc_windowset ws;
  rc = encode_fastx(&fx, ORC_DB_VERSION, &ws);
  orc_fastx_free(&fx);
  if (rc) { orc_windowset_free(&ws); return rc; }
  uint8_t *bytes = NULL;
  size_t nb = 0;
  orc_db_encode(&ws, &bytes, &nb);
  orc_windowset_free(&ws);
  FILE *f = fopen(db_path, "wb");
  if (!f) { free(bytes); return fail(ORC_ERR, "cannot create %s", db_path); }
  fwrite(bytes, 1, nb, f);
  fclose(f);
  free(bytes);
  return ORC_OK;
}

/* ------------------------------------------------------------------ query selection */

typedef struct { size_t d, i; } pair_t;

static int pair_cmp(const void *a, const void *b) {
  const pair_t *x = a, *y = b; /* Vec<(usize,usize)>::sort(): lexicographic, src/lib.rs:250 */
  if (x->d != y->d) return x->d < y->d ? -1 : 1;
  if (x->i != y->i) return x->i < y->i ? -1 : 1;
  return 0;
}

typedef struct { orc_hit *h; size_t n, cap; } hitvec;

static void hv_push(hitvec *v, uint32_t q, uint32_t s, uint32_t d) {
  if (v->n == v->cap) v->h = realloc(v->h, (v->cap = v->cap ? v->cap * 2 : 1024) * sizeof(orc_hit));
  v->h[v->n].query = q;
  v->h[v->n].subject = s;
  v->h[v->n].distance = d;
  v->n++;
}

typedef struct {
  const uint64_t *db, *q;
  size_t D, W, q0, q1;
  long m, k, r;
  hitvec out;
  int rc;
  char err[1024];
} qjob;

/* One query record: src/lib.rs:238-315 */
static int query_one(const uint64_t *db, size_t D, size_t W, const uint64_t *qw, uint32_t qnum,
                     long m, long k, long r, size_t *dist, pair_t *pairs, hitvec *out) {
  orc_distances(db, D, W, qw, dist); /* :238 */
  int mode_b = (k >= 0 && k != 1);   /* :224  Some(k).filter(k != 1) */
  if (mode_b) {
    for (size_t i = 0; i < D; ++i) { pairs[i].d = dist[i]; pairs[i].i = i; } /* :243-247 */
    qsort(pairs, D, sizeof(pair_t), pair_cmp);                               /* :250 */
    size_t max_distance;
    if ((uint32_t)k > (uint32_t)D) { /* :253-254 */
      if (D == 0) return fail(ORC_PANIC, "called `Option::unwrap()` on a `None` value");
      max_distance = 0;
      for (size_t i = 0; i < D; ++i) if (dist[i] > max_distance) max_distance = dist[i];
    } else {
      if (k == 0) /* :255 (0u32 - 1) -> overflow panic (debug) / index out of bounds (release) */
        return fail(ORC_PANIC, "attempt to subtract with overflow");
      max_distance = pairs[k - 1].d;
    }
    /* :259-294 run-length limit keyed on the decoded subject == the encoding */
    const uint64_t *last = NULL;
    uint32_t last_count = 0;
    for (size_t t = 0; t < D; ++t) {
      size_t d = pairs[t].d, i = pairs[t].i;
      if (d <= max_distance && (m < 0 || d <= (size_t)m)) {
        if (r >= 0) {
          const uint64_t *s = db + i * W;
          if (last && memcmp(last, s, W * sizeof(uint64_t)) == 0) {
            if (last_count >= (uint32_t)r) continue; /* :274-275, run not reset */
            last_count += 1;
          } else {
            last = s;
            last_count = 1;
          }
        }
        hv_push(out, qnum, (uint32_t)i, (uint32_t)d); /* :292 */
      }
    }
  } else {
    if (D == 0) return fail(ORC_PANIC, "called `Option::unwrap()` on a `None` value"); /* :298 */
    size_t mn = dist[0];
    for (size_t i = 1; i < D; ++i) if (dist[i] < mn) mn = dist[i];
    if (r >= 0) /* :301-303 */
      return fail(ORC_PANIC, "limit_per_sequence is implemented unless max_num_hits > 1. It can be "
                             "implemented by analogy, just haven't gotten around to it.");
    if (m < 0 || mn <= (size_t)m) /* :306 */
      for (size_t i = 0; i < D; ++i)
        if (dist[i] == mn) hv_push(out, qnum, (uint32_t)i, (uint32_t)mn); /* :307-311 */
  }
  return ORC_OK;
}

static void *query_worker(void *arg) {
  qjob *j = arg;
  size_t *dist = malloc((j->D + 1) * sizeof(size_t)); /* :227 */
  pair_t *pairs = malloc((j->D + 1) * sizeof(pair_t));
  for (size_t qi = j->q0; qi < j->q1; ++qi) {
    int rc = query_one(j->db, j->D, j->W, j->q + qi * j->W, (uint32_t)qi, j->m, j->k, j->r, dist,
                       pairs, &j->out);
    if (rc) { j->rc = rc; snprintf(j->err, sizeof j->err, "%s", g_err); break; }
  }
  free(dist);
  free(pairs);
  return NULL;
}

int orc_query_encoded(const uint64_t *db, size_t D, size_t W, size_t L, const uint64_t *q, size_t Q,
                      size_t q_len, long m, long k, long r, int threads, orc_hit **hits,
                      size_t *n_hits) {
  *hits = NULL;
  *n_hits = 0;
  if (Q > 0 && L != 0 && q_len != L) /* src/lib.rs:72-79 */
    return fail(ORC_PANIC, "Cannot compute distances between seq of length %zu and windows of lengths %zu",
                q_len, L);
  if (threads < 1) threads = 1;
  if ((size_t)threads > Q) threads = Q ? (int)Q : 1;
  qjob *jobs = calloc((size_t)threads, sizeof *jobs);
  pthread_t *th = calloc((size_t)threads, sizeof *th);
  for (int t = 0; t < threads; ++t) {
    jobs[t].db = db; jobs[t].q = q; jobs[t].D = D; jobs[t].W = W;
    jobs[t].q0 = Q * (size_t)t / (size_t)threads;
    jobs[t].q1 = Q * (size_t)(t + 1) / (size_t)threads;
    jobs[t].m = m; jobs[t].k = k; jobs[t].r = r;
  }
  if (threads == 1) query_worker(&jobs[0]);
  else {
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, query_worker, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  }
  int rc = ORC_OK;
  size_t total = 0;
  for (int t = 0; t < threads; ++t) {
    if (jobs[t].rc && !rc) { rc = jobs[t].rc; snprintf(g_err, sizeof g_err, "%s", jobs[t].err); }
    total += jobs[t].out.n;
  }
  if (!rc) {
    orc_hit *all = malloc((total ? total : 1) * sizeof(orc_hit));
    size_t o = 0;
    for (int t = 0; t < threads; ++t) {
      memcpy(all + o, jobs[t].out.h, jobs[t].out.n * sizeof(orc_hit));
      o += jobs[t].out.n;
    }
    *hits = all;
    *n_hits = total;
  }
  for (int t = 0; t < threads; ++t) free(jobs[t].out.h);
  free(jobs);
  free(th);
  return rc;
}

/* src/lib.rs:198-325 */
int orc_query(const char *db_path, const char *query_path, long m, long k, long r, FILE *out) {
  uint8_t *bytes = NULL;
  size_t nb = 0;
  int rc = read_file(db_path, &bytes, &nb);
  if (rc) return rc;
  orc_windowset ws;
  rc = orc_db_decode(bytes, nb, &ws);
  free(bytes);
  if (rc) return rc;
  orc_fastx fx;
  rc = orc_fastx_read(query_path, &fx);
  if (rc) { orc_windowset_free(&ws); return rc; }
  size_t *dist = malloc((ws.n + 1) * sizeof(size_t));
  pair_t *pairs = malloc((ws.n + 1) * sizeof(pair_t));
  char *dec = malloc(ws.len + 1);
  for (size_t qi = 0; qi < fx.n && !rc; ++qi) {
    size_t len = fx.lens[qi];
    uint64_t qw[(len + 11) / 12 + ws.W + 1];
    memset(qw, 0, sizeof qw);
    rc = orc_encode(fx.ids[qi], (const uint8_t *)fx.seqs[qi], len, qw); /* :235 */
    if (rc) break;
    if (ws.len && ws.len != len) { /* :72-79 */
      rc = fail(ORC_PANIC, "Cannot compute distances between seq of length %zu and windows of lengths %zu",
                len, ws.len);
      break;
    }
    hitvec hv = {0};
    rc = query_one(ws.words, ws.n, ws.W, qw, (uint32_t)qi, m, k, r, dist, pairs, &hv);
    for (size_t h = 0; h < hv.n && !rc; ++h) {
      rc = orc_decode(ws.words + (size_t)hv.h[h].subject * ws.W, ws.len, dec); /* :266,:309 */
      dec[ws.len] = 0;
      if (!rc) fprintf(out, "%u\t%u\t%u\t%s\n", hv.h[h].query, hv.h[h].subject, hv.h[h].distance, dec);
    }
    free(hv.h);
  }
  free(dist); free(pairs); free(dec);
  orc_fastx_free(&fx);
  orc_windowset_free(&ws);
  return rc;
}

/* ------------------------------------------------------------------ cluster */

/* open-addressing set of encodings (HashSet<Vec<u64>>, src/cluster.rs:24,46) */
typedef struct { const uint64_t *base; size_t W; uint32_t *slot; size_t cap; } encset;

static uint64_t enc_hash(const uint64_t *w, size_t W) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < W; ++i) { h ^= w[i]; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29; }
  return h;
}

/* returns 1 when newly inserted, 0 when already present */
static int encset_insert(encset *s, uint32_t idx) {
  const uint64_t *w = s->base + (size_t)idx * s->W;
  size_t p = enc_hash(w, s->W) & (s->cap - 1);
  while (s->slot[p] != UINT32_MAX) {
    if (memcmp(s->base + (size_t)s->slot[p] * s->W, w, s->W * sizeof(uint64_t)) == 0) return 0;
    p = (p + 1) & (s->cap - 1);
  }
  s->slot[p] = idx;
  return 1;
}

/* src/cluster.rs:22-84 */
int orc_cluster_encoded(const uint64_t *enc, size_t n, size_t W, size_t L, uint32_t t,
                        uint32_t *centroid_of, size_t *n_centroids, uint64_t *n_comparisons) {
  (void)L;
  encset seen = {enc, W, NULL, 16};
  while (seen.cap < 2 * n + 1) seen.cap *= 2;
  seen.slot = malloc(seen.cap * sizeof(uint32_t));
  memset(seen.slot, 0xff, seen.cap * sizeof(uint32_t));
  uint64_t *cw = malloc((n * W + 1) * sizeof(uint64_t)); /* centroid encodings, contiguous */
  uint32_t *cidx = malloc((n + 1) * sizeof(uint32_t));   /* centroid -> input index */
  size_t *dist = malloc((n + 1) * sizeof(size_t));
  size_t C = 0;
  uint64_t cmp = 0;
  size_t max_div = t;
  for (size_t i = 0; i < n; ++i) {
    if (!encset_insert(&seen, (uint32_t)i)) { centroid_of[i] = UINT32_MAX; continue; } /* :46-48 */
    const uint64_t *q = enc + i * W;
    orc_distances(cw, C, W, q, dist); /* :51 */
    cmp += C;
    size_t mn = max_div * 2 + 2; /* :54-58 */
    if (C) { mn = dist[0]; for (size_t c = 1; c < C; ++c) if (dist[c] < mn) mn = dist[c]; }
    size_t assigned = 0;
    if (mn <= max_div) { /* :62-68 first index at the minimum */
      for (size_t c = 0; c < C; ++c) if (dist[c] == mn) { assigned = c; break; }
    } else {             /* :69-74 */
      assigned = C;
      memcpy(cw + C * W, q, W * sizeof(uint64_t));
      cidx[C] = (uint32_t)i;
      C++;
    }
    centroid_of[i] = cidx[assigned];
  }
  if (n_centroids) *n_centroids = C;
  if (n_comparisons) *n_comparisons = cmp;
  free(seen.slot); free(cw); free(cidx); free(dist);
  return ORC_OK;
}

int orc_cluster(const char *fasta_path, uint32_t t, FILE *out) {
  orc_fastx fx;
  int rc = orc_fastx_read(fasta_path, &fx);
  if (rc) return rc;
  orc_windowset ws;
  /* Records are handled one at a time by the reference, so everything before a bad record
   * is still clustered and printed; then the panic fires. */
  int bad = encode_fastx(&fx, 0, &ws);
  char saved[sizeof g_err];
  if (bad) {
    if (strncmp(g_err, "WindowSet", 9) == 0 && ws.n < fx.n)
      /* in cluster() a ragged record trips get_distances (src/lib.rs:72-79), not push_encoding */
      fail(ORC_PANIC, "Cannot compute distances between seq of length %zu and windows of lengths %zu",
           fx.lens[ws.n], ws.len);
    memcpy(saved, g_err, sizeof saved);
  }
  uint32_t *cof = malloc((fx.n + 1) * sizeof(uint32_t));
  rc = orc_cluster_encoded(ws.words, ws.n, ws.W, ws.len, t, cof, NULL, NULL);
  char *dec = malloc(ws.len + 1);
  for (size_t i = 0; i < ws.n && !rc; ++i) {
    if (cof[i] == UINT32_MAX) continue;
    rc = orc_decode(ws.words + (size_t)cof[i] * ws.W, ws.len, dec);
    dec[ws.len] = 0;
    if (!rc) fprintf(out, "%s\t%s\n", fx.seqs[i], dec); /* :79-84 raw input, decoded centroid */
  }
  free(dec); free(cof);
  orc_windowset_free(&ws);
  orc_fastx_free(&fx);
  if (!rc && bad) { memcpy(g_err, saved, sizeof saved); rc = bad; }
  return rc;
}

/* ------------------------------------------------------------------ count */

/* src/lib.rs:378-398; serde_json of Vec<CountResult{path,num_reads,num_bases}> */
int orc_count(const char *const *paths, size_t n_paths, FILE *out) {
  bytebuf bb = {0};
  bb_put(&bb, '[');
  for (size_t i = 0; i < n_paths; ++i) {
    orc_fastx fx;
    int rc = orc_fastx_read(paths[i], &fx);
    if (rc) { free(bb.b); return rc == ORC_PANIC ? fail(ORC_ERR, "%s", g_err) : rc; }
    size_t bases = 0;
    for (size_t k = 0; k < fx.n; ++k) bases += fx.lens[k];
    char tmp[64];
    const char *pre = "{\"path\":\"";
    if (i) bb_put(&bb, ',');
    for (const char *c = pre; *c; ++c) bb_put(&bb, (uint8_t)*c);
    for (const char *c = paths[i]; *c; ++c) {
      if (*c == '"' || *c == '\\') bb_put(&bb, '\\');
      bb_put(&bb, (uint8_t)*c);
    }
    snprintf(tmp, sizeof tmp, "\",\"num_reads\":%zu,\"num_bases\":%zu}", fx.n, bases);
    for (const char *c = tmp; *c; ++c) bb_put(&bb, (uint8_t)*c);
    orc_fastx_free(&fx);
  }
  bb_put(&bb, ']');
  bb_put(&bb, '\n');
  fwrite(bb.b, 1, bb.n, out);
  free(bb.b);
  return ORC_OK;
}
