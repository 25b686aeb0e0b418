#!/usr/bin/env python3
"""bench.py -- pairwise window comparisons/sec of the smafa query hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--kernel auto|popc|mma] [--mode a|b]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's CPU algorithm on the host cores

One "step" = one pass of the hot path over one batch of synthetic input: every query of the
batch against every db window, max-divergence filter and best-hit selection included, producing
the final hit rows.  Workload at N=1 = BASELINE.json configs[1]: 100k queries x 1M windows, 60 nt,
--max-divergence 5 (Mode A).  At N>1 ONE db of N x 1M windows is generated (same generator, same seed), row-sharded
into contiguous ranges (1M-window shard per GPU: weak scaling) and the queries are drawn from the whole db, so every
shard holds hits of every query batch; the per-shard answers are exchanged with one ncclAllGather and merged inside
the library every step.  `--db-per-gpu 1250000 --queries 1000000` at N=8 is BASELINE.json configs[2] (1M x 10M).

`value`     : comparisons/s with queries and db resident in HBM (smafa_query_sharded_dev; smafa_query_dev at N=1).
`e2e`       : the same through the host-facing C-ABI call (smafa_query_sharded / smafa_query: pinned host queries
              in, host hit rows out; H2D, scan, exchange, merge and D2H inside the timed region).
At every N rank 0 checks the GPU rows of a query sample against the oracle on the WHOLE db (`cpu_baseline.matches_gpu_rows`).
`roofline`  : for the scan kernel, from CUDA events around it (smafa_stats.scan_ms).
`cpu_baseline`: the oracle (a port of the reference's algorithm; the Rust reference cannot be built
              in this image) timed on the host cores on a bounded query sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L = 60
Q_DEFAULT = 100_000
D_PER_GPU = 1_000_000
MAX_DIVERGENCE = 5
ALPHABET = "nucleotide"
# dram__bytes_read.sum + dram__bytes_write.sum of one scan_mma_kernel launch at the default workload (ncu --set full),
# keyed by the contraction depth per window of the operands that ran (192: +-1 feature operands, one window per row;
# 85: one-hot union rows, three windows per accumulator).  A CONSTANT from the named capture, not re-measured per run
# (ncu cannot run inside a timed bench); other operand choices have no capture: traffic = null.
MMA_TRAFFIC = {
    192: (230.0e6 + 4.9e6, "ncu dram__bytes_read+write, profiles/r01_ncu_mma_v8_summary.txt (algorithmic: 192 MB "
                           "db operand tiles + 19 MB query tiles, read once)"),
    85: (163.6e6 + 6.0e6, "ncu dram__bytes_read+write, profiles/r01_ncu_mma_v9_union3_summary.txt (algorithmic: 85 MB "
                          "union-row db image + 26 MB query tiles; the query tile is re-fetched per db chunk, mostly from L2)"),
    16: (76.7e6 + 3.9e6, "constant from one ncu --set full capture, profiles/r02_ncu_mma_u16_summary.txt (algorithmic: 16 MB union-row db "
                         "image of 16 windows per row + 26 MB query tiles + the reference words of verified windows)"),
}


def parse_args():
    global MAX_DIVERGENCE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "mma"])
    ap.add_argument("--mode", default="a", choices=["a", "b"], help="a: best hit + ties; b: --max-num-hits 10")
    ap.add_argument("--queries", type=int, default=Q_DEFAULT)
    ap.add_argument("--db-per-gpu", type=int, default=D_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--max-divergence", default=None, help="integer or 'none' (default 5 = configs[1]; protein: none)")
    ap.add_argument("--alphabet", default="nucleotide", choices=["nucleotide", "protein"],
                    help="protein = the configs[3] shape (20-aa windows, --max-num-hits 10), an extension without reference parity")
    ap.add_argument("--window-length", type=int, default=None,
                    help="ablation only: nucleotide windows of another length (the headline workload is 60 nt)")
    a = ap.parse_args()
    global L
    if a.window_length and a.alphabet == "nucleotide":
        L = a.window_length
    if a.alphabet == "protein":
        L, a.mode = 20, "b"
        a.max_divergence = a.max_divergence or "none"
    a.max_divergence = a.max_divergence or str(MAX_DIVERGENCE)
    MAX_DIVERGENCE = None if a.max_divergence.lower() == "none" else int(a.max_divergence)
    return a


def workload_name(a, n):
    k = "" if a.mode == "a" else " --max-num-hits 10"
    unit = "nt" if a.alphabet == "nucleotide" else "aa protein"
    return (f"synthetic {L}-{unit} SingleM windows: {a.queries} queries x {a.db_per_gpu * n} db "
            f"({a.db_per_gpu}/GPU row shard), " + (f"--max-divergence {MAX_DIVERGENCE}" if MAX_DIVERGENCE is not None else "no max-divergence") + k)


def make_inputs(a, rank, n):
    """ONE db of n x db_per_gpu windows (the SURVEY 8d generator, one seed), queries drawn from the whole of it.
    Returns (whole db as words, query words, whole db as symbols)."""
    from smafa_b200 import synth
    D = a.db_per_gpu * n
    if a.alphabet == "protein":
        db_sym = synth.make_db_aa(D, L=L, seed=synth.SEED_PROTEIN)
        q_sym = synth.make_queries_aa(db_sym, a.queries, seed=synth.SEED_PROTEIN + 1)
        pack = synth.pack_symbols_aa
    else:
        db_sym = synth.make_db(D, L=L, seed=synth.SEED_DB)
        q_sym = synth.make_queries(db_sym, a.queries, seed=synth.SEED_QUERY)
        pack = synth.pack_symbols
    # N > 1: every rank hands the WHOLE db to ShardedSearcher, which groups it as a whole (smafa_group_order) and keeps
    # this rank's range of the grouped order -- contiguous shards of the plain order when the db has no structure to group
    return pack(db_sym), pack(q_sym), db_sym


def pack_db(db_sym):
    from smafa_b200 import synth
    return (synth.pack_symbols_aa if ALPHABET == "protein" else synth.pack_symbols)(db_sym)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler process is
    started ahead of time (nvidia-smi needs ~0.5 s to come up); only rows whose arrival time falls
    between mark_begin() and mark_end() are used."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        rows = [r for ts, r in self.rows if len(r) >= 7 and self.t0 <= ts <= self.t1 + 0.03]
        if not rows:  # region shorter than one sampling period: take the nearest rows
            rows = [r for ts, r in self.rows if len(r) >= 7][-3:]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [v for v in (num(r[0]) for r in rows) if v is not None]
        mx = [v for v in (num(r[1]) for r in rows) if v is not None]
        pw = [v for v in (num(r[2]) for r in rows) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower() == "active" for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def cpu_baseline(db, q, seconds, mode_k):
    """Times the oracle's query (all host threads) on a bounded prefix of the same queries."""
    from oracle import c_oracle
    c_oracle.build()
    c_oracle.set_alphabet(1 if ALPHABET == "protein" else 0)
    threads = os.cpu_count() or 1
    D = db.shape[0]
    probe = min(q.shape[0], 4 * threads)
    t0 = time.perf_counter()
    c_oracle.query(db, L, q[:probe], L, MAX_DIVERGENCE, mode_k, None, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(q.shape[0], max(probe, probe * seconds / dt)))
    n = max(threads, n // threads * threads)
    t0 = time.perf_counter()
    hits = c_oracle.query(db, L, q[:n], L, MAX_DIVERGENCE, mode_k, None, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n * D / dt, "unit": "comparisons/s", "cores": threads, "kind": "port",
            "sample": f"first {n} of {q.shape[0]} queries x full {D}-window db, {dt:.1f} s, "
                      f"-O3 -march=native, queries split statically over {threads} threads "
                      f"(the reference itself is single-threaded)",
            "seconds": dt, "rows": int(hits.shape[0])}, hits, n


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def roofline(kernel_used, pairs_per_launch, scan_ms, peaks, peaks_kind, clocks, int8_peak, mma_k):
    """Roofline of the dominant (scan) kernel.  This path is compute-bound: operands
    are reused Q x D times, compulsory HBM traffic is a few hundred MB per step."""
    secs = scan_ms / 1e3
    if kernel_used == 2:
        # `frac` = EXECUTED int8 ops / s over the int8 tensor peak: the tensor-pipe utilisation (<= 1).  The kernel
        # executes 2 * mma_k ops per comparison (mma_k = contraction depth per window of the operands that ran: +-1
        # character features K = 192, or one-hot union rows K = 256 shared by 2 or 3 windows).  SURVEY.md 8d's
        # ALGORITHMIC figure -- one-hot(query) . one-hot(db)^T over 5 symbols x L = 2*5*L ops per comparison -- is
        # reported next to it as `frac_algorithmic`; it exceeds 1 because the union-row filter does 3.5x less tensor
        # work per comparison than that formulation, which is an algorithmic saving, not a utilisation.
        ops = 2 * (5 if ALPHABET == "nucleotide" else 23) * L  # one-hot symbols x positions
        algorithmic = pairs_per_launch * ops / secs / 1e12
        executed = pairs_per_launch * 2 * mma_k / secs / 1e12
        # MEASURED_PEAKS.json only has bf16; the int8 dense rate is measured live by the library's issue-only tcgen05
        # kind::i8 probe (smafa_debug_mma_peak) on this GPU at the same clocks, and cross-checked once against the
        # theoretical rate (148 SMs x 16384 int8 ops/clk x SM clock) and cuBLASLt IGEMM (profiles/r02_int8_peak_crosscheck.txt).
        return {"bound": "tensor", "achieved": executed, "peak": int8_peak, "unit": "TFLOP/s",
                "unit_note": "integer path: 1 'FLOP' here = one int8 multiply or add on the tensor pipe (TOP/s)",
                "frac": executed / int8_peak, "traffic": MMA_TRAFFIC.get(mma_k, (None, None))[0],
                "peak_source": "measured in this run: tcgen05.mma kind::i8 M128xN256xK32 issue-only probe "
                               f"on all SMs; for reference 2 x {peaks_kind} bf16_tflops = {2 * peaks['bf16_tflops']:.0f}",
                "executed_ops_per_comparison": 2 * mma_k, "algorithmic_ops_per_comparison": ops,
                "achieved_algorithmic": algorithmic, "frac_algorithmic": algorithmic / int8_peak,
                "traffic_source": MMA_TRAFFIC.get(mma_k, (None, "no ncu capture for these operands"))[1],
                "operands": {192: "+-1 character features, one window per accumulator (K = 192)",
                             128: "one-hot union rows, two windows per accumulator (K = 256 per row)",
                             85: "one-hot union rows, three windows per accumulator (K = 256 per row)",
                             64: "one-hot union rows, four windows per accumulator (grouped db)",
                             32: "one-hot union rows, eight windows per accumulator (grouped db)",
                             16: "one-hot union rows, sixteen windows per accumulator (grouped db)"}.get(
                                 mma_k, f"K = {mma_k} per window")}
    # POPC formulation: the binding unit is the POPC pipe.  Reference layout = 10 x (XOR32+POPC32)
    # per comparison (5 u64 words, src/lib.rs:85).  The bit-plane packing needs 2 (1 with early exit).
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    popc_peak = 148 * 16 * sm_mhz * 1e6 / 1e12  # XU pipe: 16 POPC lanes/clk/SM (ncu: xu 89-94% busy)
    popc_per_cmp = 1  # bit-plane fast path with early exit (--max-divergence <= L/4): 1 POPC per pair
    achieved = pairs_per_launch * popc_per_cmp / secs / 1e12
    return {"bound": "alu", "achieved": achieved, "peak": popc_peak, "unit": "TPOPC/s",
            "frac": achieved / popc_peak, "traffic": 35.6e6 + 0.1e6,
            "peak_source": f"148 SM x 16 POPC/clk x {sm_mhz:.0f} MHz (sampled SM clock); the XU (POPC) pipe is the "
                           "binding unit of the bit-plane formulation (1 POPC + 2 LOP3 per pair); on the reference "
                           "word layout the same comparison costs 10 POPC",
            "ops_per_comparison": popc_per_cmp, "reference_layout_popc_per_comparison": 10,
            "traffic_source": "constant from ncu dram__bytes, profiles/r01_ncu_popc_early_summary.txt"}


def run_reference(a):
    """The reference arm: the oracle port of the reference's algorithm on all host threads (the Rust reference cannot
    be built in this image), same workload as the B200 arm at this N, each step a bounded query sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(a.gpus))), a.gpus)
    db, q, _ = make_inputs(a, 0, world)
    mode_k = None if a.mode == "a" else 10
    times, cb = [], None
    per_step = max(2.0, min(a.cpu_seconds, 60.0 / max(1, a.steps + a.warmup)))
    for i in range(a.warmup + a.steps):
        cb, _, n = cpu_baseline(db, q, per_step, mode_k)
        if i >= a.warmup:
            times.append((cb["seconds"], n))
    pairs = sum(n for _, n in times) * db.shape[0]
    secs = sum(t for t, _ in times)
    value = pairs / secs
    cb["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": "pairwise window comparisons/sec", "value": value, "unit": "comparisons/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": secs / max(1, len(times)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64-popcount",
        "data": "synthetic", "config": {"workload": workload_name(a, world)},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": "comparisons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    global ALPHABET
    a = parse_args()
    ALPHABET = a.alphabet
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist

    import smafa_b200
    from smafa_b200.dist import ShardedSearcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # torch.distributed carries the barrier, the max-over-ranks of the timings and the one-time broadcast of the
        # library's communicator id; the per-step exchange is the library's own ncclAllGather (csrc/sharded.cu)
        dist.init_process_group("nccl", device_id=dev)

    db_words, q, db_sym = make_inputs(a, rank, world)
    D_total = db_words.shape[0]
    del db_sym
    ctx = smafa_b200.Context(local_rank, a.kernel)
    ctx.set_alphabet(a.alphabet)
    t_up = time.perf_counter()
    searcher = ShardedSearcher(ctx, db_words, L, world_size=world, rank=rank)
    t_up = time.perf_counter() - t_up
    if rank != 0 or a.no_cpu_baseline:
        db_words = None  # only rank 0 needs the whole db again (oracle check)
    mode_k = None if a.mode == "a" else 10
    q_pinned = torch.from_numpy(q.view(np.int64)).pin_memory()
    q_dev = q_pinned.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    pairs_per_step = a.queries * D_total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing (value) ----
    # rank 0 samples its own GPU only: one nvidia-smi poller per rank (8 of them at 50 Hz) contends for the
    # driver lock and showed up as multi-ms launch stalls in the 8-GPU run
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    int8_peak = ctx.mma_peak_tops(50000) if rank == 0 else None
    for _ in range(a.warmup):
        rows = searcher.query_dev(q_dev, MAX_DIVERGENCE, mode_k)
    n_rows = rows.shape[0]
    barrier()
    sampler.mark_begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    scan_ms, launches = 0.0, 0
    t_wall = time.perf_counter()
    for i in range(a.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        rows = searcher.query_dev(q_dev, MAX_DIVERGENCE, mode_k)
        ev[i][1].record()
        scan_ms += searcher.last_stats["scan_ms"]
        launches += searcher.last_launches
    barrier()
    t_wall = time.perf_counter() - t_wall
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = pairs_per_step * a.steps / (ms_total / 1e3)

    # ---- end to end through the host-facing C-ABI call ----
    for _ in range(2):
        host_rows = searcher.query_host(q_pinned, MAX_DIVERGENCE, mode_k)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        host_rows = searcher.query_host(q_pinned, MAX_DIVERGENCE, mode_k)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = pairs_per_step * a.steps / float(te.item())
    assert host_rows.shape[0] == n_rows
    # every rank holds the complete answer: they must agree bit for bit
    same_everywhere = True
    if world > 1:
        digest = torch.tensor([int(host_rows.astype(np.uint64).sum() % (1 << 62)), host_rows.shape[0]], dtype=torch.int64, device=dev)
        lo_d, hi_d = digest.clone(), digest.clone()
        dist.all_reduce(lo_d, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_d, op=dist.ReduceOp.MAX)
        same_everywhere = bool((lo_d == hi_d).all().item())

    if rank == 0:
        peaks, peaks_kind = load_peaks()
        st = searcher.last_stats
        # contraction depth per window of the operands the scan really used (union rows: K / 2 per window)
        mma_k = ctx.last_mma_k or searcher.db.mma_k
        roof = roofline(st["kernel_used"], a.queries * a.db_per_gpu, scan_ms / a.steps, peaks, peaks_kind, clocks, int8_peak,
                        mma_k)
        out = {
            "metric": "pairwise window comparisons/sec", "value": value, "unit": "comparisons/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s8" if st["kernel_used"] == 2 else "u32-popcount", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "kernel": {1: "popc", 2: "mma", 0: "generic"}[st["kernel_used"]],
                       "l2": "flushed between timed steps (256 MiB write)", "hit_rows": int(n_rows),
                       "candidates_per_step": int(st["candidates"]), "parallelism": f"db-row-shard x{world}",
                       "db": f"one {D_total}-window db (seed {0x5AFA0001:#x}), queries drawn from the whole db; stored on the device in "
                             "similarity-grouped order (smafa_group_order), " + ("the grouped order cut into one contiguous range per GPU"
                                                                                 if world > 1 else "one GPU"),
                       "db_upload_s": round(t_up, 3), "union_degree": int(st.get("union_degree", 0)),
                       "rows_identical_on_all_ranks": same_everywhere,
                       # selection without a useful bound scans under a guessed bound first (csrc/guess.cu)
                       "guess_bound": int(st["guess_bound"]), "rescanned_queries": int(st["rescanned"])},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "comparisons/s",
                    "h2d_bytes_per_step": int(q.nbytes) * world, "d2h_bytes_per_step": int(host_rows.nbytes) * world,
                    "entry_point": "smafa_query_sharded (C ABI, host buffers)" if world > 1 else "smafa_query (C ABI, host buffers)"},
            # kernels of this rank per timed region as the library counts them (smafa_stats.kernel_launches: its own
            # launches plus CUB's radix-sort / select kernels; NCCL's all-gather kernel counts as one)
            "gpu_launches": int(launches),
            "roofline": roof,
            "scan_ms_per_step": scan_ms / a.steps, "wall_s_timed_region": t_wall,
        }
        if not a.no_cpu_baseline:
            # rank 0: the oracle on a bounded prefix of the same queries against the WHOLE db -- the timing is the CPU
            # baseline, the rows are the parity check of the sharded scan + exchange + merge at this N
            cb, cpu_hits, n = cpu_baseline(db_words, q, a.cpu_seconds if world == 1 else min(a.cpu_seconds, 8.0), mode_k)
            got = host_rows[host_rows[:, 0] < n]
            cb["matches_gpu_rows"] = bool(got.shape == cpu_hits.shape and (got == cpu_hits).all()) and same_everywhere
            cb["checked_queries"] = int(n)
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.barrier()  # the other ranks wait for rank 0's oracle check before the process group goes away
    searcher.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
