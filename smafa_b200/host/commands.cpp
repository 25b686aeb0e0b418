// Host-side mirror of the reference's public functions: makedb / query / cluster / count
// (src/lib.rs:137-165, 198-325, 378-398; src/cluster.rs:13-94).  Everything except the distance
// scan + selection stays on the host exactly as in the reference; the scan goes through the
// C ABI (smafa_query / smafa_cluster) to the B200 kernels.  There is no CPU fallback for it.
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/smafa_b200.h"
#include "../csrc/internal.h"
#include "seqio.hpp"

using namespace smafa_host;

namespace {

// Rust's `{:?}` of a panic / error maps to: Panic -> exit 101, IoError -> exit 1.
template <class F>
int guarded(F &&f) {
  try {
    return f();
  } catch (const Panic &e) {
    smafa_set_global_error(e.what());
    return SMAFA_E_PANIC;
  } catch (const IoError &e) {
    smafa_set_global_error(e.what());
    return SMAFA_E_IO;
  } catch (const std::bad_alloc &) {
    smafa_set_global_error("out of host memory");
    return SMAFA_E_OOM;
  } catch (const std::exception &e) {
    smafa_set_global_error(e.what());
    return SMAFA_E_PANIC;
  }
}

// Buffered writer to a file descriptor (the reference's println! flushes per line; the bytes are
// the same).
struct FdWriter {
  int fd;
  std::string buf;
  explicit FdWriter(int f) : fd(f) { buf.reserve(1 << 20); }
  void flush() {
    size_t off = 0;
    while (off < buf.size()) {
      ssize_t w = ::write(fd, buf.data() + off, buf.size() - off);
      if (w <= 0) throw IoError("write failed");
      off += (size_t)w;
    }
    buf.clear();
  }
  void maybe_flush() { if (buf.size() >= (1 << 20) - 4096) flush(); }
  void put_u32(uint32_t v) {
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) buf.push_back(tmp[--n]);
  }
};

// SMAFA_TIMING=1: per-stage wall times of the file-level calls on stderr (measurement aid, not compared).
struct StageTimer {
  bool on = getenv("SMAFA_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char *what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[smafa timing] %-28s %9.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

struct EncodedInput {
  FastxFile file;               // owns what the records' views point into
  std::vector<Record> records;
  Words words;                  // [n_ok][W]
  uint32_t W = 0, L = 0;
  size_t n_ok = 0;              // records encoded before the first failure
  bool failed = false;
  std::string failure;          // panic text of the first bad record
};

// Encodes records in order until one cannot be handled.  `expect_len` (0 = take the first
// record's) is the window length every record must have; `mismatch` builds the panic text.
// `expect_text` is the reference's .expect(..) message for a record needletail cannot parse.
template <class MismatchMsg>
EncodedInput encode_all(FastxFile file, uint32_t expect_len, int alphabet, MismatchMsg mismatch, const char *expect_text) {
  EncodedInput in;
  in.records = std::move(file.recs);
  in.file = std::move(file);
  if (!in.file.parse_error.empty()) {  // reached only if every record before it was fine (overwritten below otherwise)
    in.failed = true;
    in.failure = std::string(expect_text) + ": " + in.file.parse_error;
  }
  if (in.records.empty()) return in;
  in.L = expect_len ? expect_len : (uint32_t)in.records[0].seq.size();
  in.W = words_for_len(in.L);
  in.words.resize(in.records.size() * (size_t)in.W);
  // Records are independent, so they are encoded on all host threads; the reference stops at
  // the first record it cannot handle, which is the lowest failing index over all chunks.
  struct Fail { size_t index = SIZE_MAX; std::string text; };
  const unsigned T = host_threads();
  std::vector<Fail> fails(T);
  parallel_chunks(in.records.size(), 16384, [&](unsigned t, size_t lo, size_t hi) {
    std::vector<uint64_t> tmp;
    for (size_t i = lo; i < hi; ++i) {
      const Record &r = in.records[i];
      tmp.assign(words_for_len(r.seq.size()) + 1, 0);
      try {
        encode_or_panic(r, tmp.data(), alphabet);  // encoding comes first in the reference (src/lib.rs:150,235)
      } catch (const Panic &e) {
        fails[t] = Fail{i, e.what()};
        return;
      }
      if (r.seq.size() != in.L) {
        fails[t] = Fail{i, mismatch(r.seq.size(), in.L)};
        return;
      }
      memcpy(in.words.data() + i * in.W, tmp.data(), in.W * sizeof(uint64_t));
    }
  });
  in.n_ok = in.records.size();
  for (const Fail &f : fails)
    if (f.index < in.n_ok) {
      in.n_ok = f.index;
      in.failed = true;
      in.failure = f.text;
    }
  return in;
}

}  // namespace

extern "C" uint8_t smafa_encode_symbol(uint8_t byte) { return SYMBOL_CODE[byte]; }

extern "C" int smafa_encode_window(const uint8_t *seq, size_t len, uint64_t *out_words, size_t *bad_pos) {
  return smafa_encode_window_alphabet(seq, len, out_words, bad_pos, SMAFA_ALPHABET_NUCLEOTIDE);
}

extern "C" int smafa_decode_window(const uint64_t *words, size_t len, char *out) {
  return smafa_decode_window_alphabet(words, len, out, SMAFA_ALPHABET_NUCLEOTIDE);
}

extern "C" uint8_t smafa_encode_symbol_alphabet(uint8_t byte, int alphabet) {
  return alphabet ? AA_SYMBOL_CODE[byte] : SYMBOL_CODE[byte];
}

extern "C" int smafa_encode_window_alphabet(const uint8_t *seq, size_t len, uint64_t *out_words, size_t *bad_pos, int alphabet) {
  if ((!seq && len) || !out_words) return SMAFA_E_INVALID;
  return encode_window(seq, len, out_words, bad_pos, alphabet) ? SMAFA_OK : SMAFA_E_PANIC;
}

extern "C" int smafa_decode_window_alphabet(const uint64_t *words, size_t len, char *out, int alphabet) {
  return guarded([&] { decode_window(words, len, out, alphabet); return (int)SMAFA_OK; });
}

extern "C" int smafa_makedb_file(const char *subject_fasta, const char *db_path) {
  return smafa_makedb_file_alphabet(subject_fasta, db_path, SMAFA_ALPHABET_NUCLEOTIDE);
}

// src/lib.rs:137-165
extern "C" int smafa_makedb_file_alphabet(const char *subject_fasta, const char *db_path, int alphabet) {
  return guarded([&]() -> int {
    StageTimer tm;
    FastxFile fx = read_fastx(subject_fasta);
    std::vector<Record> &recs = fx.recs;
    tm.lap("makedb: read FASTX");
    if (!recs.empty() && recs[0].seq.empty()) {
      std::vector<uint64_t> t(1);
      encode_or_panic(recs[0], t.data(), alphabet);
      throw Panic("Cannot add empty sequence to WindowSet: TryFromIntError(())");
    }
    EncodedInput in = encode_all(std::move(fx), 0, alphabet, [](size_t got, uint32_t want) {
      return "WindowSet seq length is " + std::to_string(want) + ", got a new sequence of length " + std::to_string(got);
    }, "valid record");
    tm.lap("makedb: encode");
    if (in.failed) throw Panic(in.failure);
    WindowDb db;
    db.n = in.n_ok;
    db.W = in.W;
    db.L = in.L;
    db.words = std::move(in.words);
    const Bytes bytes = serialize_db(db);
    tm.lap("makedb: serialize");
    FILE *f = fopen(db_path, "wb");
    if (!f) throw IoError(std::string("cannot create ") + db_path);
    const size_t w = fwrite(bytes.data(), 1, bytes.size(), f);
    fclose(f);
    if (w != bytes.size()) throw IoError("short write");
    tm.lap("makedb: write");
    return SMAFA_OK;
  });
}

// src/lib.rs:206-218
extern "C" int smafa_db_file_load(const char *db_path, uint64_t **words, uint64_t *n, uint32_t *W, uint32_t *window_len) {
  if (!db_path || !words || !n || !W || !window_len) return SMAFA_E_INVALID;
  *words = nullptr;
  return guarded([&]() -> int {
    WindowDb db = parse_db(read_file(db_path));
    uint64_t *out = (uint64_t *)malloc(std::max<size_t>(1, db.words.size()) * sizeof(uint64_t));
    if (!out) throw std::bad_alloc();
    memcpy(out, db.words.data(), db.words.size() * sizeof(uint64_t));
    *words = out;
    *n = db.n;
    *W = db.W;
    *window_len = db.L;
    return SMAFA_OK;
  });
}

// src/lib.rs:208-217: File::open(..)? then the version gate
extern "C" int smafa_db_file_check(const char *db_path) {
  return guarded([&]() -> int {
    // only the first bytes matter here; the whole file is read (once) by the command itself
    FILE *f = fopen(db_path, "rb");
    if (!f) throw IoError("Os { code: " + std::to_string(errno) + ", kind: NotFound, message: \"" + strerror(errno) + "\" }");
    uint8_t head[16];
    const size_t got = fread(head, 1, sizeof head, f);
    fclose(f);
    const std::vector<uint8_t> bytes(head, head + got);
    if (bytes.size() < 4) throw Panic("range end index 4 out of range for slice of length " + std::to_string(bytes.size()));
    uint64_t v = 0;
    for (int i = 0; i < 4; ++i) {
      v |= (uint64_t)(bytes[i] & 0x7f) << (7 * i);
      if (!(bytes[i] & 0x80)) break;
      if (i == 3) throw IoError("DeserializeUnexpectedEnd");
    }
    if (v != DB_VERSION)
      throw Panic("Unsupported db file version: " + std::to_string(v) + ". This version of smafa only works with version " +
                  std::to_string(DB_VERSION) + " databases. The last version to support version 1 databases was v0.7.1.");
    return SMAFA_OK;
  });
}

// first[i] = 1 iff no earlier record has the encoding of record i (the HashSet<Vec<u64>> of src/cluster.rs:24,46-48).
// Equal encodings hash alike, so the records are partitioned by hash and every host thread de-duplicates its own
// partitions in input order.  The result does not depend on the thread count.
static void mark_first_occurrences(const uint64_t *words, size_t n_rec, uint32_t W, uint8_t *first) {
  struct Key {
    const uint64_t *w;
    uint32_t W;
    bool operator==(const Key &o) const { return memcmp(w, o.w, W * sizeof(uint64_t)) == 0; }
  };
  struct KeyHash {
    size_t operator()(const Key &k) const {
      uint64_t h = 0x9E3779B97F4A7C15ull;
      for (uint32_t i = 0; i < k.W; ++i) { h ^= k.w[i]; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 31; }
      return (size_t)h;
    }
  };
  const unsigned P = n_rec >= 65536 ? host_threads() : 1;  // partitions
  if (P == 1) {
    std::unordered_set<Key, KeyHash> seen;
    seen.reserve(n_rec * 2 + 1);
    for (size_t i = 0; i < n_rec; ++i) first[i] = seen.insert(Key{words + i * W, W}).second;
    return;
  }
  // pass 1: partition number of every record; pass 2: per-partition index lists (input order); pass 3: dedup
  std::vector<uint8_t> part(n_rec);
  std::vector<std::vector<size_t>> counts(P, std::vector<size_t>(P, 0));  // [chunk][partition]
  parallel_chunks(P, 1, [&](unsigned, size_t c0, size_t c1) {
    for (size_t c = c0; c < c1; ++c)
      for (size_t i = n_rec * c / P; i < n_rec * (c + 1) / P; ++i) {
        const unsigned pt = (unsigned)((KeyHash()(Key{words + i * W, W}) >> 40) % P);
        part[i] = (uint8_t)pt;
        counts[c][pt]++;
      }
  });
  std::vector<std::vector<uint32_t>> members(P);
  std::vector<std::vector<size_t>> offset(P, std::vector<size_t>(P, 0));  // [chunk][partition] start inside members
  for (unsigned pt = 0; pt < P; ++pt) {
    size_t tot = 0;
    for (unsigned c = 0; c < P; ++c) { offset[c][pt] = tot; tot += counts[c][pt]; }
    members[pt].resize(tot);
  }
  parallel_chunks(P, 1, [&](unsigned, size_t c0, size_t c1) {
    for (size_t c = c0; c < c1; ++c) {
      std::vector<size_t> at = offset[c];
      for (size_t i = n_rec * c / P; i < n_rec * (c + 1) / P; ++i) members[part[i]][at[part[i]]++] = (uint32_t)i;
    }
  });
  parallel_chunks(P, 1, [&](unsigned, size_t p0, size_t p1) {
    for (size_t pt = p0; pt < p1; ++pt) {
      std::unordered_set<Key, KeyHash> seen;
      seen.reserve(members[pt].size() * 2 + 1);
      for (uint32_t i : members[pt]) first[i] = seen.insert(Key{words + (size_t)i * W, W}).second;
    }
  });
}

extern "C" int smafa_mark_first_occurrences(const uint64_t *words, uint64_t n, uint32_t W, uint8_t *first) {
  if ((n && (!words || !first)) || W == 0 || n >= (1ull << 32)) return SMAFA_E_INVALID;
  return guarded([&]() -> int { mark_first_occurrences(words, (size_t)n, W, first); return SMAFA_OK; });
}

// A context that may still be under construction on a helper thread: CUDA initialisation takes 1-3 s and does not
// depend on the inputs, so the CLI entry points (smafa_*_file_on_device) start it first and read, decode and encode
// the files meanwhile.  get() joins; a failed creation is reported through *rc and the global error text.
struct LazyCtx {
  smafa_ctx *ctx = nullptr;
  int rc = SMAFA_OK;
  std::string err;
  std::thread th;
  void start(const int *devices, int n_devices, int kernel, int alphabet) {
    const std::vector<int> devs(devices, devices + n_devices);
    th = std::thread([this, devs, kernel, alphabet] {
      rc = smafa_ctx_create_multi(&ctx, devs.data(), (int)devs.size(), kernel);
      if (rc) err = smafa_last_error(nullptr);  // the error text is thread-local: carry it over
      else smafa_ctx_set_alphabet(ctx, alphabet);
    });
  }
  smafa_ctx *get() {
    if (th.joinable()) th.join();
    if (rc) smafa_set_global_error(err);
    return rc ? nullptr : ctx;
  }
  ~LazyCtx() { if (th.joinable()) th.join(); }
};

static int query_file_impl(LazyCtx &lazy, int alphabet, const char *db_path, const char *query_fasta, int64_t max_divergence,
                           int64_t max_num_hits, int64_t limit_per_sequence, int out_fd);
static int cluster_file_impl(LazyCtx &lazy, int alphabet, const char *input_fasta, uint32_t max_divergence, int out_fd);

extern "C" int smafa_query_file(smafa_ctx *ctx, const char *db_path, const char *query_fasta, int64_t max_divergence,
                                int64_t max_num_hits, int64_t limit_per_sequence, int out_fd) {
  if (!ctx) { smafa_set_global_error("smafa_query_file needs a context (no CPU fallback)"); return SMAFA_E_PANIC; }
  LazyCtx lazy;
  lazy.ctx = ctx;
  return query_file_impl(lazy, ctx->alphabet, db_path, query_fasta, max_divergence, max_num_hits, limit_per_sequence, out_fd);
}

extern "C" int smafa_query_file_on_device(int device, int kernel, int alphabet, const char *db_path, const char *query_fasta,
                                          int64_t max_divergence, int64_t max_num_hits, int64_t limit_per_sequence,
                                          int out_fd, smafa_ctx **ctx_out) {
  return smafa_query_file_on_devices(&device, 1, kernel, alphabet, db_path, query_fasta, max_divergence, max_num_hits,
                                     limit_per_sequence, out_fd, ctx_out);
}

extern "C" int smafa_query_file_on_devices(const int *devices, int n_devices, int kernel, int alphabet, const char *db_path,
                                           const char *query_fasta, int64_t max_divergence, int64_t max_num_hits,
                                           int64_t limit_per_sequence, int out_fd, smafa_ctx **ctx_out) {
  if (ctx_out) *ctx_out = nullptr;
  if (!devices || n_devices < 1) { smafa_set_global_error("smafa_query_file_on_devices: no device given"); return SMAFA_E_INVALID; }
  LazyCtx lazy;
  lazy.start(devices, n_devices, kernel, alphabet);
  int rc = query_file_impl(lazy, alphabet, db_path, query_fasta, max_divergence, max_num_hits, limit_per_sequence, out_fd);
  smafa_ctx *ctx = lazy.get();  // also when the inputs needed no device work: a missing GPU is reported, not ignored
  if (ctx_out) *ctx_out = ctx; else if (ctx) smafa_ctx_destroy(ctx);
  return rc ? rc : lazy.rc;
}

extern "C" int smafa_cluster_file_on_device(int device, int kernel, int alphabet, const char *input_fasta,
                                            uint32_t max_divergence, int out_fd, smafa_ctx **ctx_out) {
  return smafa_cluster_file_on_devices(&device, 1, kernel, alphabet, input_fasta, max_divergence, out_fd, ctx_out);
}

extern "C" int smafa_cluster_file_on_devices(const int *devices, int n_devices, int kernel, int alphabet, const char *input_fasta,
                                             uint32_t max_divergence, int out_fd, smafa_ctx **ctx_out) {
  if (ctx_out) *ctx_out = nullptr;
  if (!devices || n_devices < 1) { smafa_set_global_error("smafa_cluster_file_on_devices: no device given"); return SMAFA_E_INVALID; }
  LazyCtx lazy;
  // the greedy is sequential (src/cluster.rs:45-74): one device does all of it, the others would only idle
  lazy.start(devices, 1, kernel, alphabet);
  int rc = cluster_file_impl(lazy, alphabet, input_fasta, max_divergence, out_fd);
  smafa_ctx *ctx = lazy.get();
  if (ctx_out) *ctx_out = ctx; else if (ctx) smafa_ctx_destroy(ctx);
  return rc ? rc : lazy.rc;
}

// src/lib.rs:198-325
static int query_file_impl(LazyCtx &lazy, int alphabet, const char *db_path, const char *query_fasta, int64_t max_divergence,
                           int64_t max_num_hits, int64_t limit_per_sequence, int out_fd) {
  smafa_db *dbh = nullptr;
  smafa_ctx *ctx = nullptr;
  int rc = guarded([&]() -> int {
    StageTimer tm;
    WindowDb db = parse_db(read_file(db_path));  // File::open(..)? -> Err, version gate -> panic
    tm.lap("query: read + decode db");
    FastxFile fx = read_fastx(query_fasta);
    tm.lap("query: read FASTX");
    // get_distances checks the length only when the db is non-empty (src/lib.rs:72)
    EncodedInput in = encode_all(std::move(fx), db.L, alphabet, [](size_t got, uint32_t want) {
      return "Cannot compute distances between seq of length " + std::to_string(got) + " and windows of lengths " +
             std::to_string(want);
    }, "Failed to parse query sequence");
    tm.lap("query: encode");
    const bool mode_b = max_num_hits >= 0 && max_num_hits != 1;  // src/lib.rs:224
    FdWriter out(out_fd);
    if (in.n_ok > 0) {
      if (!(ctx = lazy.get())) return lazy.rc;
      tm.lap("query: wait for the CUDA context");
      int r = smafa_db_upload(ctx, db.words.data(), db.n, db.L, 0, &dbh);
      if (r) return r;
      tm.lap("query: db upload + re-pack");
      // Mode A with --limit-per-sequence panics right after the first min() (src/lib.rs:298-303)
      const uint64_t nq = (!mode_b && limit_per_sequence >= 0 && db.n > 0) ? 0 : in.n_ok;
      smafa_hit *hits = nullptr;
      uint64_t n_hits = 0;
      r = smafa_query(ctx, dbh, in.words.data(), nq ? nq : 1, in.L, max_divergence, max_num_hits, &hits, &n_hits, nullptr);
      if (r) {
        if (r == SMAFA_E_EMPTY_DB || r == SMAFA_E_BAD_K || r == SMAFA_E_LENGTH_MISMATCH) throw Panic(smafa_last_error(ctx));
        smafa_set_global_error(smafa_last_error(ctx));
        return r;
      }
      tm.lap("query: scan + selection");
      if (nq == 0) {
        smafa_free(hits);
        throw Panic("limit_per_sequence is implemented unless max_num_hits > 1. It can be implemented by analogy, "
                    "just haven't gotten around to it.");
      }
      if (mode_b && limit_per_sequence >= 0)
        n_hits = smafa_apply_limit_per_sequence(hits, n_hits, db.words.data(), db.W, 0, (uint32_t)limit_per_sequence);
      // src/lib.rs:292,310: "query\tsubject\tdistance\tdecoded subject\n".  Lines are formatted on all host
      // threads into per-chunk buffers and written in order.
      const unsigned T = n_hits >= 65536 ? host_threads() : 1;
      std::vector<std::string> parts(T);
      parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
        for (size_t t = t0; t < t1; ++t) {
          FdWriter w(-1);
          w.buf.reserve((size_t)(n_hits / T + 1) * (db.L + 24));
          std::string dec(db.L, '\0');
          for (uint64_t i = n_hits * t / T; i < n_hits * (t + 1) / T; ++i) {
            decode_window(db.words.data() + (size_t)hits[i].subject * db.W, db.L, dec.data(), alphabet);
            w.put_u32(hits[i].query); w.buf.push_back('\t');
            w.put_u32(hits[i].subject); w.buf.push_back('\t');
            w.put_u32(hits[i].distance); w.buf.push_back('\t');
            w.buf.append(dec); w.buf.push_back('\n');
          }
          parts[t] = std::move(w.buf);
        }
      });
      for (std::string &part : parts) {
        out.buf = std::move(part);
        out.flush();
      }
      tm.lap("query: format + write TSV");
      smafa_free(hits);
      out.flush();
    }
    if (in.failed) throw Panic(in.failure);
    return SMAFA_OK;
  });
  if (dbh) smafa_db_free(dbh);
  return rc;
}

extern "C" int smafa_cluster_file(smafa_ctx *ctx, const char *input_fasta, uint32_t max_divergence, int out_fd) {
  if (!ctx) { smafa_set_global_error("smafa_cluster_file needs a context (no CPU fallback)"); return SMAFA_E_PANIC; }
  LazyCtx lazy;
  lazy.ctx = ctx;
  return cluster_file_impl(lazy, ctx->alphabet, input_fasta, max_divergence, out_fd);
}

// src/cluster.rs:13-94
static int cluster_file_impl(LazyCtx &lazy, int alphabet, const char *input_fasta, uint32_t max_divergence, int out_fd) {
  return guarded([&]() -> int {
    smafa_ctx *ctx = nullptr;
    StageTimer tm;
    FastxFile fx = read_fastx(input_fasta);
    std::vector<Record> &recs = fx.recs;
    tm.lap("cluster: read FASTX");
    if (!recs.empty() && recs[0].seq.empty()) {
      std::vector<uint64_t> t(1);
      encode_or_panic(recs[0], t.data(), alphabet);
      throw Panic("Cannot add empty sequence to WindowSet: TryFromIntError(())");
    }
    EncodedInput in = encode_all(std::move(fx), 0, alphabet, [](size_t got, uint32_t want) {
      return "Cannot compute distances between seq of length " + std::to_string(got) + " and windows of lengths " +
             std::to_string(want);
    }, "Failed to parse input sequence");
    // HashSet<Vec<u64>> de-duplication on encodings, first occurrence wins (src/cluster.rs:24,46-48)
    const size_t n_rec = in.n_ok;
    std::vector<uint8_t> first(n_rec, 0);
    mark_first_occurrences(in.words.data(), n_rec, in.W, first.data());
    std::vector<uint32_t> uniq;  // record index of each unique encoding, input order
    size_t n_uniq = 0;
    for (size_t i = 0; i < n_rec; ++i) n_uniq += first[i];
    uniq.reserve(n_uniq);
    for (size_t i = 0; i < n_rec; ++i)
      if (first[i]) uniq.push_back((uint32_t)i);
    std::vector<uint64_t> uwords(n_uniq * (size_t)in.W);
    parallel_chunks(n_uniq, 65536, [&](unsigned, size_t u0, size_t u1) {
      for (size_t u = u0; u < u1; ++u) memcpy(uwords.data() + u * in.W, in.words.data() + (size_t)uniq[u] * in.W, in.W * sizeof(uint64_t));
    });
    tm.lap("cluster: encode + de-duplicate");
    std::vector<uint32_t> cof(uniq.size());
    uint64_t n_centroids = 0;
    if (!uniq.empty()) {
      if (!(ctx = lazy.get())) return lazy.rc;
      tm.lap("cluster: wait for the CUDA context");
      int r = smafa_cluster(ctx, uwords.data(), uniq.size(), in.L, max_divergence, cof.data(), &n_centroids, nullptr, nullptr);
      if (r) { smafa_set_global_error(smafa_last_error(ctx)); return r; }
    }
    tm.lap("cluster: greedy (GPU distances)");
    // src/cluster.rs:79-84: raw input sequence, decoded centroid.  Formatted on all host threads into per-chunk
    // buffers and written in order, like the query TSV.
    FdWriter out(out_fd);
    const unsigned T = uniq.size() >= 65536 ? host_threads() : 1;
    std::vector<std::string> parts(T);
    parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
      for (size_t t = t0; t < t1; ++t) {
        std::string buf, dec(in.L, '\0');
        const size_t u0 = uniq.size() * t / T, u1 = uniq.size() * (t + 1) / T;
        buf.reserve((u1 - u0) * (2 * (size_t)in.L + 2));
        for (size_t u = u0; u < u1; ++u) {
          decode_window(uwords.data() + (size_t)cof[u] * in.W, in.L, dec.data(), alphabet);
          buf.append(in.records[uniq[u]].seq); buf.push_back('\t');
          buf.append(dec); buf.push_back('\n');
        }
        parts[t] = std::move(buf);
      }
    });
    for (std::string &part : parts) {
      out.buf = std::move(part);
      out.flush();
    }
    tm.lap("cluster: format + write");
    if (in.failed) throw Panic(in.failure);
    return SMAFA_OK;
  });
}

// src/lib.rs:378-398
extern "C" int smafa_count_files(const char *const *paths, size_t n_paths, int out_fd) {
  return guarded([&]() -> int {
    std::string js = "[";
    for (size_t i = 0; i < n_paths; ++i) {
      FastxFile fx;
      try {
        fx = read_fastx(paths[i], /*io_error_on_open=*/true);  // parse_fastx_file(&path)? -> Err
      } catch (const Panic &e) {
        throw IoError(e.what());  // count() propagates parse errors with `?`
      }
      if (!fx.parse_error.empty()) throw IoError(fx.parse_error);  // `let record = record?;` (src/lib.rs:385)
      size_t bases = 0;
      const std::vector<Record> &recs = fx.recs;
      for (const Record &r : recs) bases += r.seq.size();
      if (i) js += ',';
      js += "{\"path\":\"";
      for (const char *c = paths[i]; *c; ++c) {
        if (*c == '"' || *c == '\\') js += '\\';
        js += *c;
      }
      js += "\",\"num_reads\":" + std::to_string(recs.size()) + ",\"num_bases\":" + std::to_string(bases) + "}";
    }
    js += "]\n";
    FdWriter out(out_fd);
    out.buf = js;
    out.flush();
    return SMAFA_OK;
  });
}
