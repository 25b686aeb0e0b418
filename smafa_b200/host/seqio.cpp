#include "seqio.hpp"

#include <zlib.h>

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace smafa_host {

namespace {
struct Lut {
  uint8_t t[256];
  constexpr Lut() : t() {
    for (int i = 0; i < 256; ++i) t[i] = 0;
    const char *n = "NWSMKRYBDHV";
    t['A'] = t['a'] = 16;
    t['C'] = t['c'] = 8;
    t['G'] = t['g'] = 4;
    t['T'] = t['t'] = t['U'] = t['u'] = 2;
    for (int i = 0; n[i]; ++i) { t[(int)n[i]] = 1; t[(int)n[i] + 32] = 1; }
    t['-'] = 1;
  }
};
constexpr Lut kLut;

struct AaLut {
  uint8_t t[256];
  constexpr AaLut() : t() {
    for (int i = 0; i < 256; ++i) t[i] = 0;
    const char *aa = "ACDEFGHIKLMNPQRSTVWY";
    for (int i = 0; aa[i]; ++i) { t[(int)aa[i]] = (uint8_t)(i + 1); t[(int)aa[i] + 32] = (uint8_t)(i + 1); }
    const char *x = "XBZJUO";
    for (int i = 0; x[i]; ++i) { t[(int)x[i]] = 21; t[(int)x[i] + 32] = 21; }
    t['-'] = 22;
    t['*'] = 23;
  }
};
constexpr AaLut kAaLut;
}  // namespace

const uint8_t AA_SYMBOL_CODE[256] = {
#define X4(i) kAaLut.t[i], kAaLut.t[i + 1], kAaLut.t[i + 2], kAaLut.t[i + 3]
#define X16(i) X4(i), X4(i + 4), X4(i + 8), X4(i + 12)
#define X64(i) X16(i), X16(i + 16), X16(i + 32), X16(i + 48)
    X64(0), X64(64), X64(128), X64(192)
#undef X64
#undef X16
#undef X4
};

const uint8_t SYMBOL_CODE[256] = {
#define X4(i) kLut.t[i], kLut.t[i + 1], kLut.t[i + 2], kLut.t[i + 3]
#define X16(i) X4(i), X4(i + 4), X4(i + 8), X4(i + 12)
#define X64(i) X16(i), X16(i + 16), X16(i + 32), X16(i + 48)
    X64(0), X64(64), X64(128), X64(192)
#undef X64
#undef X16
#undef X4
};

bool encode_window(const uint8_t *seq, size_t len, uint64_t *out, size_t *bad_pos, int alphabet) {
  const uint32_t W = words_for_len(len);
  const uint8_t *lut = alphabet ? AA_SYMBOL_CODE : SYMBOL_CODE;
  for (uint32_t w = 0; w < W; ++w) {
    const size_t base = (size_t)w * 12, n = len - base < 12 ? len - base : 12;
    uint64_t acc = 0;
    for (size_t i = 0; i < n; ++i) {
      const uint8_t c = lut[seq[base + i]];
      if (!c) {
        if (bad_pos) *bad_pos = base + i;
        return false;
      }
      acc |= (uint64_t)c << (5 * i);
    }
    out[w] = acc;
  }
  return true;
}

void encode_or_panic(const Record &r, uint64_t *out, int alphabet) {
  size_t bad = 0;
  if (!encode_window(reinterpret_cast<const uint8_t *>(r.seq.data()), r.seq.size(), out, &bad, alphabet))
    throw Panic("Byte " + std::to_string((unsigned)(uint8_t)r.seq[bad]) + " cannot be interpreted as " +
                (alphabet ? "amino acid" : "nucleotide") + ", in sequence \"" + std::string(r.id) + "\" at position " +
                std::to_string(bad));
}

void decode_window(const uint64_t *words, size_t len, char *out, int alphabet) {
  static const char nuc[32] = {0, 'N', 'T', 0, 'G', 0, 0, 0, 'C', 0, 0, 0, 0, 0, 0, 0, 'A'};
  static const char aa[32] = {0,   'A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M',
                              'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'Y', 'X', '-', '*'};
  const char *sym = alphabet ? aa : nuc;
  for (size_t i = 0; i < len; ++i) {
    const unsigned b = (unsigned)(words[i / 12] >> (5 * (i % 12))) & 31u;
    const char c = sym[b];
    if (!c) throw Panic("Invalid character in query sequence: " + std::to_string(b));
    out[i] = c;
  }
}

unsigned host_threads() {
  static const unsigned n = [] {
    if (const char *e = getenv("SMAFA_HOST_THREADS")) {
      const int v = atoi(e);
      if (v >= 1) return (unsigned)std::min(v, 256);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return std::min(std::max(hc, 1u), 32u);
  }();
  return n;
}

void parallel_chunks(size_t n, size_t min_per_thread, const std::function<void(unsigned, size_t, size_t)> &fn) {
  unsigned T = (unsigned)std::min<size_t>(host_threads(), std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread)));
  if (T <= 1) { fn(0, 0, n); return; }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(T);
  for (unsigned t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      try { fn(t, n * t / T, n * (t + 1) / T); } catch (...) { err[t] = std::current_exception(); }
    });
  for (auto &x : th) x.join();
  for (auto &e : err)
    if (e) std::rethrow_exception(e);  // the lowest-numbered chunk's failure, i.e. the first in input order
}

Bytes read_file(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f)
    throw IoError("Os { code: " + std::to_string(errno) + ", kind: NotFound, message: \"" + strerror(errno) + "\" }");
  Bytes buf;
  struct stat st;
  if (fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {  // one allocation
    buf.resize((size_t)st.st_size);
    const size_t n = buf.size();
    const int fd = fileno(f);
    std::atomic<bool> short_read{false};
    // large files: every host thread preads its own range (the page faults of the fresh buffer and the copy out of
    // the page cache are the cost, and both parallelise)
    parallel_chunks(n, 32u << 20, [&](unsigned, size_t lo, size_t hi) {
      size_t got = lo;
      while (got < hi) {
        const ssize_t r = pread(fd, buf.data() + got, hi - got, (off_t)got);
        if (r <= 0) { short_read = true; return; }
        got += (size_t)r;
      }
    });
    if (short_read) {  // the file shrank under us, or an I/O error: fall back to one sequential read
      size_t got = 0, r;
      rewind(f);
      while (got < n && (r = fread(buf.data() + got, 1, n - got, f)) > 0) got += r;
      buf.resize(got);
    } else {
      fseek(f, (long)n, SEEK_SET);
    }
  }
  uint8_t tmp[1 << 16];  // pipes, or a file that grew
  size_t r;
  while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
  fclose(f);
  return buf;
}

static Bytes read_maybe_gz(const std::string &path, bool io_error_on_open) {
  FILE *probe = fopen(path.c_str(), "rb");
  if (!probe) {
    std::string msg = "Os { code: " + std::to_string(errno) + ", kind: NotFound, message: \"" + strerror(errno) + "\" }";
    if (io_error_on_open) throw IoError(msg);
    throw Panic("valid path/file of input: " + msg);
  }
  uint8_t magic[2] = {0, 0};
  const size_t got = fread(magic, 1, 2, probe);
  fclose(probe);
  if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) return read_file(path);  // not gzip: no zlib pass
  gzFile g = gzopen(path.c_str(), "rb");
  if (!g) throw IoError("cannot open " + path);
  gzbuffer(g, 1 << 20);
  Bytes buf;
  std::vector<uint8_t> tmp(1 << 20);
  for (;;) {
    int r = gzread(g, tmp.data(), (unsigned)tmp.size());
    if (r < 0) { gzclose(g); throw IoError("read error on " + path); }
    if (r == 0) break;
    buf.insert(buf.end(), tmp.begin(), tmp.begin() + r);
  }
  gzclose(g);
  return buf;
}

// Appends b[0, n) without line endings, one memchr + append per line.
static void append_stripped(std::string &dst, const uint8_t *b, size_t n) {
  size_t p = 0;
  while (p < n) {
    const void *nl = memchr(b + p, '\n', n - p);
    size_t e = nl ? (size_t)((const uint8_t *)nl - b) : n;
    size_t le = e;
    while (le > p && b[le - 1] == '\r') --le;
    if (memchr(b + p, '\r', le - p) == nullptr) {
      dst.append(reinterpret_cast<const char *>(b + p), le - p);
    } else {  // stray carriage returns inside a line: the slow, exact way
      for (size_t i = p; i < le; ++i)
        if (b[i] != '\r') dst.push_back((char)b[i]);
    }
    p = e + 1;
  }
}

FastxFile read_fastx(const std::string &path, bool io_error_on_open) {
  FastxFile f;
  f.buf = read_maybe_gz(path, io_error_on_open);
  const uint8_t *b = f.buf.data();
  const size_t n = f.buf.size();
  if (n == 0) throw Panic("valid path/file: EmptyFile");
  auto view = [&](size_t lo, size_t hi) { return std::string_view(reinterpret_cast<const char *>(b + lo), hi - lo); };
  if (b[0] == '>') {
    // the next record start ('>' at the start of a line) at or after q
    auto next_start = [&](size_t q) {
      while (q < n) {
        const void *gt = memchr(b + q, '>', n - q);
        if (!gt) return n;
        q = (size_t)((const uint8_t *)gt - b);
        if (q == 0 || b[q - 1] == '\n') return q;
        ++q;
      }
      return n;
    };
    // chunk boundaries = the first record start at or after the nominal split points
    const unsigned T = n >= (8u << 20) ? host_threads() : 1;
    std::vector<size_t> cut(T + 1, n);
    cut[0] = 0;
    for (unsigned t = 1; t < T; ++t) cut[t] = next_start(std::max(cut[t - 1], n * t / T));
    // pass 1: records per chunk, so that pass 2 can write them in place (no per-thread lists to merge)
    std::vector<size_t> first(T + 1, 0);
    parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
      for (size_t t = t0; t < t1; ++t) {
        size_t cnt = 0;
        for (size_t p = cut[t]; p < cut[t + 1]; p = next_start(p + 1)) {
          if (b[p] != '>') throw Panic("valid record: InvalidStart");
          ++cnt;
        }
        first[t + 1] = cnt;
      }
    });
    for (unsigned t = 0; t < T; ++t) first[t + 1] += first[t];
    f.recs.resize(first[T]);
    f.arenas.resize(T);
    // pass 2: records of [cut[t], cut[t+1])
    parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
      for (size_t t = t0; t < t1; ++t) {
        Record *dst = f.recs.data() + first[t];
        size_t p = cut[t];
        const size_t hi = cut[t + 1];
        while (p < hi) {
          size_t he = p + 1;
          const void *nl = memchr(b + he, '\n', n - he);
          he = nl ? (size_t)((const uint8_t *)nl - b) : n;
          size_t idn = he - (p + 1);
          if (idn && b[p + idn] == '\r') --idn;
          Record r;
          r.id = view(p + 1, p + 1 + idn);
          const size_t ss = he < n ? he + 1 : n, se = next_start(ss);
          // the usual record is one sequence line: a view into the file, trailing line end excluded
          size_t le = se;
          if (le > ss && b[le - 1] == '\n') --le;
          while (le > ss && b[le - 1] == '\r') --le;
          if (memchr(b + ss, '\n', le - ss) == nullptr && memchr(b + ss, '\r', le - ss) == nullptr) {
            r.seq = view(ss, le);
          } else {
            std::string joined;
            joined.reserve(se - ss);
            append_stripped(joined, b + ss, se - ss);
            f.arenas[t].push_back(std::move(joined));
            r.seq = f.arenas[t].back();
          }
          *dst++ = r;
          p = se;
        }
      }
    });
  } else if (b[0] == '@') {
    size_t p = 0;
    while (p < n) {
      if (b[p] == '\n' || b[p] == '\r') { ++p; continue; }
      const size_t rec_no = f.recs.size() + 1;
      if (b[p] != '@') {
        f.parse_error = "ParseError { kind: InvalidStart, record: " + std::to_string(rec_no) + " }";
        break;
      }
      size_t ls[4], le[4];
      int lines = 0;
      for (int l = 0; l < 4 && p < n; ++l, ++lines) {
        ls[l] = p;
        const void *nl = memchr(b + p, '\n', n - p);
        p = nl ? (size_t)((const uint8_t *)nl - b) : n;
        le[l] = p;
        if (le[l] > ls[l] && b[le[l] - 1] == '\r') --le[l];
        if (p < n) ++p;
      }
      // needletail rejects these (the reference then panics on the record): accept nothing it would not
      if (lines < 4) {
        f.parse_error = "ParseError { kind: UnexpectedEnd, record: " + std::to_string(rec_no) + " }";
        break;
      }
      if (le[2] == ls[2] || b[ls[2]] != '+') {
        f.parse_error = "ParseError { kind: InvalidSeparator, record: " + std::to_string(rec_no) + " }";
        break;
      }
      if (le[3] - ls[3] != le[1] - ls[1]) {
        f.parse_error = "ParseError { kind: UnequalLengths, record: " + std::to_string(rec_no) + " }";
        break;
      }
      f.recs.push_back(Record{view(ls[0] + 1, le[0]), view(ls[1], le[1])});
    }
  } else {
    throw Panic("valid path/file: InvalidStart");
  }
  return f;
}

// ---- db bytes: varint(version) varint(n) n x [varint(W) W x varint(u64)] option(len) ----

static inline void put_varint(Bytes &o, uint64_t v) {
  while (v >= 0x80) { o.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  o.push_back((uint8_t)v);
}

static inline size_t varint_len(uint64_t v) { return (size_t)(70 - __builtin_clzll(v | 1)) / 7; }
static inline uint8_t *put_varint_raw(uint8_t *o, uint64_t v) {
  while (v >= 0x80) { *o++ = (uint8_t)(v | 0x80); v >>= 7; }
  *o++ = (uint8_t)v;
  return o;
}

Bytes serialize_db(const WindowDb &db) {
  // Two passes on all host threads: the encoded size of every chunk of windows, then each chunk written in place
  // at its offset (no per-chunk buffers to concatenate, no capacity check per byte).
  const unsigned T = db.n >= 65536 ? host_threads() : 1;
  std::vector<size_t> off(T + 1, 0);
  const size_t wlen = varint_len(db.W);
  parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      const uint64_t lo = db.n * t / T, hi = db.n * (t + 1) / T;
      size_t bytes = (size_t)(hi - lo) * wlen;
      const uint64_t *w = db.words.data() + lo * db.W, *e = db.words.data() + hi * db.W;
      for (; w < e; ++w) bytes += varint_len(*w);
      off[t + 1] = bytes;
    }
  });
  Bytes o;
  put_varint(o, DB_VERSION);
  put_varint(o, db.n);
  const size_t head = o.size();
  off[0] = head;
  for (unsigned t = 0; t < T; ++t) off[t + 1] += off[t];
  o.resize(off[T]);
  parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      const uint64_t lo = db.n * t / T, hi = db.n * (t + 1) / T;
      uint8_t *p = o.data() + off[t];
      for (uint64_t i = lo; i < hi; ++i) {
        p = put_varint_raw(p, db.W);
        for (uint32_t w = 0; w < db.W; ++w) p = put_varint_raw(p, db.words[i * db.W + w]);
      }
    }
  });
  if (db.L) { o.push_back(1); put_varint(o, db.L); }
  else o.push_back(0);
  return o;
}

namespace {
struct Cursor {
  const uint8_t *b;
  size_t n, p = 0;
  uint64_t varint(int max_bytes) {
    uint64_t v = 0;
    for (int i = 0; i < max_bytes; ++i) {
      if (p >= n) throw IoError("DeserializeUnexpectedEnd");
      const uint8_t c = b[p++];
      v |= (uint64_t)(c & 0x7f) << (7 * i);
      if (!(c & 0x80)) return v;
    }
    throw IoError("DeserializeBadVarint");
  }
};
}  // namespace

// Sequential decode of the window list (also the arbiter of every malformed file: the parallel path below
// falls back to it whenever anything looks off, so error behaviour is that of one straight pass).
static void parse_body_sequential(Cursor &c, WindowDb &db) {
  for (uint64_t i = 0; i < db.n; ++i) {
    const uint64_t w = c.varint(10);
    if (i == 0) {
      // every window costs at least 1 + W bytes: a count the file cannot hold is a truncated file, not an allocation
      if (w > 512 || db.n > (c.n - c.p) / (w + 1) + 1) throw IoError("DeserializeUnexpectedEnd");
      db.W = (uint32_t)w;
      db.words.resize(db.n * db.W);
    } else if (w != db.W) {
      throw IoError("db file holds windows of different word counts");
    }
    uint64_t *dst = db.words.data() + i * db.W;
    for (uint32_t j = 0; j < db.W; ++j) dst[j] = c.varint(10);
  }
}

// Parallel decode.  A LEB128 stream can be entered at any byte that follows a terminator
// (MSB clear), so the body is cut into byte ranges that start on varint boundaries; a first pass counts the
// terminators of every range (= varints in it), the prefix sum gives each range the index of its first
// varint, and since every window is exactly 1 + W varints (inner Vec length, then the words) that index maps
// to (window, slot) directly.  Returns false when the stream does not have that regular shape.
static bool parse_body_parallel(Cursor &c, WindowDb &db) {
  const uint8_t *b = c.b;
  const size_t body = c.p, n = c.n;
  Cursor probe{b, n, body};
  const uint64_t W = probe.varint(10);
  if (W == 0 || W > 512) return false;
  const uint64_t V = db.n * (W + 1);  // varints in the body
  const unsigned T = host_threads();
  if (T <= 1 || n - body < (size_t)T * 4096) return false;
  std::vector<size_t> start(T + 1, n);
  start[0] = body;
  for (unsigned t = 1; t < T; ++t) {
    size_t q = std::max(start[t - 1], body + (n - body) * t / T);
    while (q < n && q > body && (b[q - 1] & 0x80)) ++q;  // advance to a byte that follows a terminator
    start[t] = q;
  }
  std::vector<uint64_t> count(T + 1, 0);
  parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      uint64_t cnt = 0;
      const uint8_t *p = b + start[t], *e = b + start[t + 1];
      for (; p + 8 <= e; p += 8) {
        uint64_t x;
        memcpy(&x, p, 8);
        cnt += 8 - (uint64_t)__builtin_popcountll(x & 0x8080808080808080ull);
      }
      for (; p < e; ++p) cnt += !(*p & 0x80);
      count[t + 1] = cnt;
    }
  });
  for (unsigned t = 0; t < T; ++t) count[t + 1] += count[t];
  if (count[T] < V + 1) return false;  // truncated (the Option tag after the body is one more terminator)
  db.W = (uint32_t)W;
  db.words.resize(db.n * W);
  std::atomic<bool> bad{false};
  std::vector<size_t> body_end(T, 0);
  parallel_chunks(T, 1, [&](unsigned, size_t t0, size_t t1) {
    for (size_t t = t0; t < t1; ++t) {
      uint64_t idx = count[t];
      if (idx >= V) continue;
      uint64_t win = idx / (W + 1), slot = idx % (W + 1);
      Cursor cur{b, start[t + 1], start[t]};
      while (cur.p < start[t + 1] && idx < V) {
        uint64_t v = 0;
        int i = 0;
        for (;; ++i) {
          if (i == 10 || cur.p >= n) { bad = true; return; }
          const uint8_t ch = b[cur.p++];
          v |= (uint64_t)(ch & 0x7f) << (7 * i);
          if (!(ch & 0x80)) break;
        }
        if (slot == 0) {
          if (v != W) { bad = true; return; }
        } else {
          db.words[win * W + slot - 1] = v;
        }
        if (++slot == W + 1) { slot = 0; ++win; }
        ++idx;
      }
      if (idx == V) body_end[t] = cur.p;
    }
  });
  if (bad) return false;
  size_t end = 0;
  for (unsigned t = 0; t < T; ++t) end = std::max(end, body_end[t]);
  if (end == 0) return false;
  c.p = end;
  return true;
}

WindowDb parse_db(const Bytes &bytes) {
  if (bytes.size() < 4)  // &buffer[0..4], src/lib.rs:214
    throw Panic("range end index 4 out of range for slice of length " + std::to_string(bytes.size()));
  Cursor head{bytes.data(), 4};
  const uint64_t version = head.varint(5);
  if (version != DB_VERSION)
    throw Panic("Unsupported db file version: " + std::to_string(version) +
                ". This version of smafa only works with version " + std::to_string(DB_VERSION) +
                " databases. The last version to support version 1 databases was v0.7.1.");
  Cursor c{bytes.data(), bytes.size()};
  c.varint(5);
  WindowDb db;
  db.n = c.varint(10);
  const size_t body = c.p;
  if (db.n < 65536 || db.n > bytes.size() || !parse_body_parallel(c, db)) {
    c.p = body;
    db.words.clear();
    parse_body_sequential(c, db);
  }
  if (c.p >= c.n) throw IoError("DeserializeUnexpectedEnd");
  const uint8_t tag = c.b[c.p++];
  if (tag == 1) db.L = (uint32_t)c.varint(10);
  else if (tag != 0) throw IoError("DeserializeBadOption");
  // `len` and the windows' word counts are separate fields of the file (src/lib.rs:54-60); makedb always writes them
  // consistently (ceil(len / 12) words per window, src/lib.rs:32).  Everything downstream -- the device upload, the TSV
  // decode, --limit-per-sequence -- derives the row stride from len, so a hand-made file that disagrees is refused
  // here instead of being read with the wrong stride.
  if (db.n > 0 && (db.L == 0 || db.W != words_for_len(db.L)))
    throw IoError("db file is inconsistent: windows of " + std::to_string(db.W) + " words, window length " +
                  (db.L ? std::to_string(db.L) : std::string("None")));
  return db;
}

}  // namespace smafa_host
