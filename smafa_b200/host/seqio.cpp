#include "seqio.hpp"

#include <zlib.h>

#include <cerrno>
#include <cstdio>
#include <cstring>

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
                (alphabet ? "amino acid" : "nucleotide") + ", in sequence \"" + r.id + "\" at position " +
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

std::vector<uint8_t> read_file(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f)
    throw IoError("Os { code: " + std::to_string(errno) + ", kind: NotFound, message: \"" + strerror(errno) + "\" }");
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t r;
  while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
  fclose(f);
  return buf;
}

static std::vector<uint8_t> read_maybe_gz(const std::string &path, bool io_error_on_open) {
  FILE *probe = fopen(path.c_str(), "rb");
  if (!probe) {
    std::string msg = "Os { code: " + std::to_string(errno) + ", kind: NotFound, message: \"" + strerror(errno) + "\" }";
    if (io_error_on_open) throw IoError(msg);
    throw Panic("valid path/file of input: " + msg);
  }
  fclose(probe);
  gzFile g = gzopen(path.c_str(), "rb");
  if (!g) throw IoError("cannot open " + path);
  gzbuffer(g, 1 << 20);
  std::vector<uint8_t> buf;
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

static void append_stripped(std::string &dst, const uint8_t *b, size_t n) {
  for (size_t i = 0; i < n; ++i)
    if (b[i] != '\n' && b[i] != '\r') dst.push_back((char)b[i]);
}

std::vector<Record> read_fastx(const std::string &path, bool io_error_on_open) {
  const std::vector<uint8_t> buf = read_maybe_gz(path, io_error_on_open);
  const uint8_t *b = buf.data();
  const size_t n = buf.size();
  std::vector<Record> out;
  if (n == 0) throw Panic("valid path/file: EmptyFile");
  size_t p = 0;
  if (b[0] == '>') {
    while (p < n) {
      if (b[p] != '>') throw Panic("valid record: InvalidStart");
      size_t he = p + 1;
      while (he < n && b[he] != '\n') ++he;
      size_t idn = he - (p + 1);
      if (idn && b[p + idn] == '\r') --idn;
      Record r;
      r.id.assign(reinterpret_cast<const char *>(b + p + 1), idn);
      size_t ss = he < n ? he + 1 : n, se = ss;
      while (se < n) {  // next '>' at the start of a line ends the record
        const void *gt = memchr(b + se, '>', n - se);
        if (!gt) { se = n; break; }
        se = (const uint8_t *)gt - b;
        if (se == ss || b[se - 1] == '\n') break;
        ++se;
      }
      r.seq.reserve(se - ss);
      append_stripped(r.seq, b + ss, se - ss);
      out.push_back(std::move(r));
      p = se;
    }
  } else if (b[0] == '@') {
    while (p < n) {
      if (b[p] == '\n' || b[p] == '\r') { ++p; continue; }
      if (b[p] != '@') throw Panic("valid record: InvalidStart");
      size_t ls[4], le[4];
      for (int l = 0; l < 4; ++l) {
        ls[l] = p;
        const void *nl = memchr(b + p, '\n', n - p);
        p = nl ? (size_t)((const uint8_t *)nl - b) : n;
        le[l] = p;
        if (le[l] > ls[l] && b[le[l] - 1] == '\r') --le[l];
        if (p < n) ++p;
      }
      Record r;
      r.id.assign(reinterpret_cast<const char *>(b + ls[0] + 1), le[0] - ls[0] - 1);
      r.seq.assign(reinterpret_cast<const char *>(b + ls[1]), le[1] - ls[1]);
      out.push_back(std::move(r));
    }
  } else {
    throw Panic("valid path/file: InvalidStart");
  }
  return out;
}

// ---- db bytes: varint(version) varint(n) n x [varint(W) W x varint(u64)] option(len) ----

static inline void put_varint(std::vector<uint8_t> &o, uint64_t v) {
  while (v >= 0x80) { o.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  o.push_back((uint8_t)v);
}

std::vector<uint8_t> serialize_db(const WindowDb &db) {
  std::vector<uint8_t> o;
  o.reserve(16 + db.n * (1 + 9 * (size_t)db.W));
  put_varint(o, DB_VERSION);
  put_varint(o, db.n);
  for (uint64_t i = 0; i < db.n; ++i) {
    put_varint(o, db.W);
    for (uint32_t w = 0; w < db.W; ++w) put_varint(o, db.words[i * db.W + w]);
  }
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

WindowDb parse_db(const std::vector<uint8_t> &bytes) {
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
  for (uint64_t i = 0; i < db.n; ++i) {
    const uint64_t w = c.varint(10);
    if (i == 0) {
      db.W = (uint32_t)w;
      db.words.resize(db.n * db.W);
    } else if (w != db.W) {
      throw IoError("db file holds windows of different word counts");
    }
    uint64_t *dst = db.words.data() + i * db.W;
    for (uint32_t j = 0; j < db.W; ++j) dst[j] = c.varint(10);
  }
  if (c.p >= c.n) throw IoError("DeserializeUnexpectedEnd");
  const uint8_t tag = c.b[c.p++];
  if (tag == 1) db.L = (uint32_t)c.varint(10);
  else if (tag != 0) throw IoError("DeserializeBadOption");
  return db;
}

}  // namespace smafa_host
