// `smafa` command line of the B200 drop-in.  Same subcommands and flags as the reference binary
// (src/main.rs:64-116); additive flags: --device N / --devices a,b,... (query: the db is row-sharded over the listed
// GPUs), --kernel {auto,popc,mma}, --protein (amino-acid windows, an extension: the reference only knows nucleotides).
// Exit codes follow Rust: 0 ok, 101 for a reference panic, 1 for an Err from main, 2 usage.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smafa_b200.h"

static const char *kVersion = "0.8.0-b200";

static int usage(const std::string &msg) {
  fprintf(stderr,
          "error: %s\n\nUsage: smafa [-v|-q] <COMMAND>\n\nCommands:\n"
          "  makedb   Generate a searchable database            -i <FILE> -d <FILE>\n"
          "  query    Search a database                         -d <FILE> -q <FILE> [--max-divergence <INT>]\n"
          "           [--max-num-hits <INT>] [--limit-per-sequence <INT>] [--device <INT> | --devices <INT,INT,...>] [--kernel auto|popc|mma]\n"
          "  cluster  Cluster sequences by similarity           -i <FILE> -d <INT> [--device <INT>] [--kernel ..]\n"
          "  (makedb/query/cluster: --protein treats the windows as amino acids -- an extension, not in the reference)\n"
          "  count    Print the number of reads/bases in a possibly gzipped FASTX file  -i <FILE>...\n",
          msg.c_str());
  return 2;
}

static int finish(int rc, smafa_ctx *ctx) {
  if (rc == SMAFA_OK) return 0;
  const char *msg = smafa_last_error(nullptr);
  if (rc == SMAFA_E_IO) {
    fprintf(stderr, "Error: %s\n", msg);
    return 1;
  }
  if (rc == SMAFA_E_PANIC) {
    fprintf(stderr, "thread 'main' panicked:\n%s\n", msg);
    return 101;
  }
  fprintf(stderr, "Error: %s: %s\n", smafa_status_name(rc), ctx ? smafa_last_error(ctx) : msg);
  return 1;
}

static bool parse_u32(const char *s, int64_t *out) {
  char *end = nullptr;
  unsigned long long v = strtoull(s, &end, 10);
  if (!*s || *end || s[0] == '-' || v > 0xFFFFFFFFull) return false;
  *out = (int64_t)v;
  return true;
}

// clap accepts `--name=value`, `-n=value` and `-nvalue` next to `--name value` (src/main.rs uses the derive defaults):
// split those forms so the parser below only sees separate tokens.  `value_shorts` = the short options of this
// subcommand that take a value.
static std::vector<std::string> normalize_args(int argc, char **argv) {
  std::vector<std::string> out;
  std::string sub;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (sub.empty() && !a.empty() && a[0] != '-') sub = a;
    const std::string value_shorts = sub == "query" ? "dq" : sub == "count" ? "i" : (sub.empty() ? "" : "id");
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      const size_t eq = a.find('=');
      if (eq != std::string::npos) { out.push_back(a.substr(0, eq)); out.push_back(a.substr(eq + 1)); continue; }
    } else if (a.size() > 2 && a[0] == '-' && a[1] != '-' && value_shorts.find(a[1]) != std::string::npos) {
      out.push_back(a.substr(0, 2));
      out.push_back(a.substr(a[2] == '=' ? 3 : 2));
      continue;
    }
    out.push_back(a);
  }
  return out;
}

int main(int argc_raw, char **argv_raw) {
  const std::vector<std::string> norm = normalize_args(argc_raw, argv_raw);
  std::vector<char *> argv_vec{argv_raw[0]};
  for (const std::string &a : norm) argv_vec.push_back(const_cast<char *>(a.c_str()));
  const int argc = (int)argv_vec.size();
  char **argv = argv_vec.data();
  int i = 1;
  auto is = [&](const char *a, const char *s, const char *l) { return (s && !strcmp(a, s)) || (l && !strcmp(a, l)); };
  while (i < argc && (is(argv[i], "-v", "--verbose") || is(argv[i], "-q", "--quiet"))) ++i;
  if (i < argc && is(argv[i], "-V", "--version")) { printf("smafa %s\n", kVersion); return 0; }
  if (i >= argc) { usage("a subcommand is required"); return 0; }  // reference prints help, exit 0
  const std::string cmd = argv[i++];
  const bool is_query = cmd == "query", is_cluster = cmd == "cluster", is_count = cmd == "count", is_makedb = cmd == "makedb";
  if (!is_query && !is_cluster && !is_count && !is_makedb) return usage("unrecognized subcommand '" + cmd + "'");
  const char *input = nullptr, *database = nullptr, *query = nullptr;
  std::vector<const char *> inputs;
  int64_t m = -1, k = -1, r = -1, device = 0;
  std::vector<int> devices;
  int kernel = SMAFA_KERNEL_AUTO;
  int alphabet = SMAFA_ALPHABET_NUCLEOTIDE;
  for (; i < argc; ++i) {
    const char *a = argv[i];
    auto need = [&](int64_t *dst) {
      if (i + 1 >= argc || !parse_u32(argv[i + 1], dst)) return false;
      ++i;
      return true;
    };
    if (is(a, "-v", "--verbose") || is(a, nullptr, "--quiet") || (!is_query && is(a, "-q", nullptr))) continue;
    if (is(a, "-i", "--input") && !is_query) {
      if (is_count) { while (i + 1 < argc && argv[i + 1][0] != '-') inputs.push_back(argv[++i]); }
      else if (i + 1 < argc) input = argv[++i];
      else return usage("a value is required for '--input <FILE>'");
    } else if (is_cluster && is(a, "-d", "--max-divergence")) { if (!need(&m)) return usage("invalid value for '--max-divergence <INT>'"); }
    else if ((is_query || is_makedb) && is(a, "-d", "--database") && i + 1 < argc) database = argv[++i];
    else if (is_query && is(a, "-q", "--query") && i + 1 < argc) query = argv[++i];
    else if (is_query && is(a, nullptr, "--max-divergence")) { if (!need(&m)) return usage("invalid value for '--max-divergence <INT>'"); }
    else if (is_query && is(a, nullptr, "--max-num-hits")) { if (!need(&k)) return usage("invalid value for '--max-num-hits <INT>'"); }
    else if (is_query && is(a, nullptr, "--limit-per-sequence")) { if (!need(&r)) return usage("invalid value for '--limit-per-sequence <INT>'"); }
    else if (!is_count && is(a, nullptr, "--protein")) alphabet = SMAFA_ALPHABET_PROTEIN;
    else if ((is_query || is_cluster) && is(a, nullptr, "--device")) { if (!need(&device)) return usage("invalid value for '--device <INT>'"); }
    else if ((is_query || is_cluster) && is(a, nullptr, "--devices") && i + 1 < argc) {
      // comma-separated GPU numbers: `query` row-shards the db over them; `cluster` uses the first
      std::string list = argv[++i];
      size_t p = 0;
      while (p <= list.size()) {
        const size_t c = std::min(list.find(',', p), list.size());
        int64_t d = 0;
        if (!parse_u32(list.substr(p, c - p).c_str(), &d) || d > 1023) return usage("invalid value for '--devices <INT,INT,...>'");
        devices.push_back((int)d);
        p = c + 1;
      }
    }
    else if ((is_query || is_cluster) && is(a, nullptr, "--kernel") && i + 1 < argc) {
      const std::string v = argv[++i];
      if (v == "auto") kernel = SMAFA_KERNEL_AUTO;
      else if (v == "popc") kernel = SMAFA_KERNEL_POPC;
      else if (v == "mma") kernel = SMAFA_KERNEL_MMA;
      else return usage("invalid value for '--kernel'");
    } else return usage(std::string("unexpected argument '") + a + "'");
  }
  if (is_makedb) {
    if (!input || !database) return usage("the following required arguments were not provided: --input <FILE> --database <FILE>");
    return finish(smafa_makedb_file_alphabet(input, database, alphabet), nullptr);
  }
  if (is_count) {
    if (inputs.empty()) return usage("the following required arguments were not provided: --input <FILE>");
    return finish(smafa_count_files(inputs.data(), inputs.size(), 1), nullptr);
  }
  if (is_query && (!database || !query)) return usage("the following required arguments were not provided: --database <FILE> --query <FILE>");
  if (is_cluster && !input) return usage("the following required arguments were not provided: --input <FILE>");
  if (is_cluster && m < 0) {  // src/main.rs:43 .unwrap() on the optional -d
    fprintf(stderr, "thread 'main' panicked:\ncalled `Option::unwrap()` on a `None` value\n");
    return 101;
  }
  if (is_query) {  // open + version gate come first in the reference, before any device work
    int pre = smafa_db_file_check(database);
    if (pre) return finish(pre, nullptr);
  }
  const bool timing = getenv("SMAFA_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[smafa timing] %-28s %9.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t0).count());
    t0 = now;
  };
  // the context is created on a helper thread while the inputs are read and encoded (SMAFA_TIMING shows the wait)
  smafa_ctx *ctx = nullptr;
  int rc;
  if (devices.empty()) devices.push_back((int)device);
  if (is_query) rc = smafa_query_file_on_devices(devices.data(), (int)devices.size(), kernel, alphabet, database, query, m, k, r, 1, &ctx);
  else rc = smafa_cluster_file_on_devices(devices.data(), (int)devices.size(), kernel, alphabet, input, (uint32_t)m, 1, &ctx);
  lap(is_query ? "query (all stages above)" : "cluster (all stages above)");
  int code = finish(rc, ctx);
  smafa_ctx_destroy(ctx);
  lap("context teardown");
  return code;
}
