// Host-side sequence I/O of the B200 smafa drop-in: FASTA/FASTQ(+gz) reading, the 5-bit one-hot
// window encoding and the db file format.  Stands in for needletail + postcard + the encode /
// decode helpers of the reference (src/lib.rs:29-52,113-135,137-165,167-196,206-218).
#pragma once
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace smafa_host {

// Mirrors a Rust panic! (process exit code 101).
struct Panic : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// Mirrors an Err(..) bubbling out of main (exit code 1, "Error: ..." on stderr).
struct IoError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Host worker threads for the ingest paths (db decode, FASTA parse, window encoding): the hardware concurrency,
// capped at 32; SMAFA_HOST_THREADS overrides.  The reference is single-threaded; results do not depend on it.
unsigned host_threads();
// Runs fn(thread, begin, end) over [0, n) split into contiguous chunks; rethrows the failure of the lowest chunk.
void parallel_chunks(size_t n, size_t min_per_thread, const std::function<void(unsigned, size_t, size_t)> &fn);

constexpr uint32_t DB_VERSION = 2;  // CURRENT_DB_VERSION, src/lib.rs:18

// File contents: a byte vector whose resize() does not zero-fill (a whole-file read would otherwise touch every page
// twice), read with parallel pread()s when the file is large.
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
  template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
using Bytes = std::vector<uint8_t, DefaultInitAlloc<uint8_t>>;
using Words = std::vector<uint64_t, DefaultInitAlloc<uint64_t>>;  // every element is written by the (parallel) decoder / encoder

// One FASTX record as views into its FastxFile (no per-record allocation: 10 M records used to cost 20 M mallocs and
// a gigabyte of freshly faulted pages, which bounded the "multi-threaded" parse at one thread's speed).
struct Record {
  std::string_view id;   // header line without '>' / '@'
  std::string_view seq;  // line endings stripped, case preserved
};
struct FastxFile {
  Bytes buf;                                      // the (decompressed) file: single-line sequences are views into it
  std::vector<std::deque<std::string>> arenas;    // re-assembled sequences (multi-line, stray CR), one arena per parser thread
  std::vector<Record> recs;
  // Set when the records after recs.back() could not be parsed (FASTQ: no '+' line, quality and sequence of different
  // lengths, a truncated last record).  needletail hands out records one by one, so the reference handles every
  // record before the bad one and then panics on .expect(..) (src/lib.rs:149,234; src/cluster.rs:39) -- the callers
  // here do the same with this text.
  std::string parse_error;
  FastxFile() = default;
  FastxFile(FastxFile &&) = default;              // moving keeps every heap block, so the views stay valid
  FastxFile &operator=(FastxFile &&) = default;
  FastxFile(const FastxFile &) = delete;
  FastxFile &operator=(const FastxFile &) = delete;
};

// Whole-file FASTX parse (plain or gzip, sniffed by zlib).  Throws Panic on malformed input
// (the reference .expect()s needletail's result) and IoError when the file cannot be opened.
FastxFile read_fastx(const std::string &path, bool io_error_on_open = false);

extern const uint8_t SYMBOL_CODE[256];  // src/lib.rs:167-184; 0 = not a nucleotide
// Protein extension (not in the reference): symbol numbers 1..20 = ACDEFGHIKLMNPQRSTVWY,
// 21 = X (also B, Z, J, U, O), 22 = '-', 23 = '*'; lower case accepted; 0 = not an amino-acid symbol.
extern const uint8_t AA_SYMBOL_CODE[256];
inline uint32_t words_for_len(size_t len) { return (uint32_t)((len + 11) / 12); }
// alphabet: 0 = nucleotide (the reference's one-hot codes), 1 = protein (symbol numbers).
// Returns false and sets *bad_pos at the first byte without a code.
bool encode_window(const uint8_t *seq, size_t len, uint64_t *out, size_t *bad_pos, int alphabet = 0);
void encode_or_panic(const Record &r, uint64_t *out, int alphabet = 0);  // panic text of src/lib.rs:38-41
void decode_window(const uint64_t *words, size_t len, char *out, int alphabet = 0);  // src/lib.rs:113-135

struct WindowDb {
  Words words;  // [n][W]
  uint64_t n = 0;
  uint32_t W = 0;
  uint32_t L = 0;  // 0 == None (empty db)
};

Bytes serialize_db(const WindowDb &db);  // == postcard::to_allocvec(&WindowSet)
WindowDb parse_db(const Bytes &bytes);                  // version gate of src/lib.rs:212-217
Bytes read_file(const std::string &path);

}  // namespace smafa_host
