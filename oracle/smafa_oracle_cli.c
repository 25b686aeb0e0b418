/*
 * oracle/smafa_oracle_cli.c -- command-line front end of the CPU oracle (TEST INFRASTRUCTURE
 * ONLY).  Mirrors the flags of the reference binary (src/main.rs:64-116) so the reference's
 * CLI golden vectors (tests/test_cmdline.rs) can be replayed against the restatement.
 * Exit codes follow Rust: 0 ok, 101 panic, 1 Err from main, 2 usage (clap).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "smafa_oracle.h"

static int finish(int rc) {
  if (rc == ORC_OK) return 0;
  if (rc == ORC_PANIC) {
    fprintf(stderr, "thread 'main' panicked:\n%s\n", orc_last_error());
    return 101;
  }
  fprintf(stderr, "Error: %s\n", orc_last_error());
  return 1;
}

static int usage(const char *msg) {
  fprintf(stderr, "error: %s\n\nUsage: smafa_oracle <makedb|query|cluster|count> [OPTIONS]\n", msg);
  return 2;
}

static int is_flag(const char *a, const char *s, const char *l) {
  return (s && strcmp(a, s) == 0) || (l && strcmp(a, l) == 0);
}

int main(int argc, char **argv) {
  int i = 1;
  while (i < argc && (is_flag(argv[i], "-v", "--verbose") || is_flag(argv[i], "-q", "--quiet"))) i++;
  if (i >= argc) return usage("a subcommand is required");
  const char *cmd = argv[i++];
  const char *input = NULL, *database = NULL, *query = NULL;
  const char *inputs[256];
  size_t n_inputs = 0;
  long m = -1, k = -1, r = -1;
  int is_cluster = strcmp(cmd, "cluster") == 0, is_query = strcmp(cmd, "query") == 0;
  for (; i < argc; ++i) {
    const char *a = argv[i];
    if (is_flag(a, "-v", "--verbose") || is_flag(a, NULL, "--quiet")) continue;
    if (is_flag(a, NULL, "--protein")) { orc_set_alphabet(1); continue; } /* extension, parity unpinned */
    if (!is_query && is_flag(a, "-q", NULL)) continue;
    if (is_flag(a, "-i", "--input")) {
      if (strcmp(cmd, "count") == 0) {
        while (i + 1 < argc && argv[i + 1][0] != '-' && n_inputs < 256) inputs[n_inputs++] = argv[++i];
      } else if (i + 1 < argc) input = argv[++i];
    } else if (is_cluster && is_flag(a, "-d", "--max-divergence") && i + 1 < argc) m = atol(argv[++i]);
    else if (is_flag(a, "-d", "--database") && i + 1 < argc) database = argv[++i];
    else if (is_query && is_flag(a, "-q", "--query") && i + 1 < argc) query = argv[++i];
    else if (is_flag(a, NULL, "--max-divergence") && i + 1 < argc) m = atol(argv[++i]);
    else if (is_flag(a, NULL, "--max-num-hits") && i + 1 < argc) k = atol(argv[++i]);
    else if (is_flag(a, NULL, "--limit-per-sequence") && i + 1 < argc) r = atol(argv[++i]);
    else return usage("unexpected argument");
  }
  if (strcmp(cmd, "makedb") == 0) {
    if (!input || !database) return usage("makedb needs -i and -d");
    return finish(orc_makedb(input, database));
  }
  if (is_query) {
    if (!query || !database) return usage("query needs -d and -q");
    int rc = orc_query(database, query, m, k, r, stdout);
    fflush(stdout);
    return finish(rc);
  }
  if (is_cluster) {
    if (!input) return usage("cluster needs -i");
    if (m < 0) { /* src/main.rs:43 .unwrap() on a missing -d */
      fprintf(stderr, "thread 'main' panicked:\ncalled `Option::unwrap()` on a `None` value\n");
      return 101;
    }
    int rc = orc_cluster(input, (uint32_t)m, stdout);
    fflush(stdout);
    return finish(rc);
  }
  if (strcmp(cmd, "count") == 0) return finish(orc_count(inputs, n_inputs, stdout));
  return usage("unknown subcommand");
}
