"""Cross-check of the int8 tensor peak that bench.py's roofline divides by: the library's own issue-only tcgen05 probe
(smafa_debug_mma_peak) next to cuBLASLt's int8 GEMM (torch._int_mm, 8192^3 and 16384 x 8192 x 8192) and the theoretical
rate 148 SMs x 16384 int8 ops/clk x SM clock."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smafa_b200

dev = torch.device("cuda", 0)
ctx = smafa_b200.Context(0)
probe = [ctx.mma_peak_tops(50000) for _ in range(3)]
print("library probe (tcgen05.mma kind::i8 M128xN256xK32, issue only, all SMs): %s TOP/s" % ", ".join("%.0f" % x for x in probe))
for M, N, K in [(8192, 8192, 8192), (16384, 8192, 8192)]:
    a = torch.randint(-8, 8, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(-8, 8, (K, N), dtype=torch.int8, device=dev).t().contiguous().t()
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # sustained: back to back for ~2 s
    n = max(4, int(2000 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n
    print("cuBLASLt int8 GEMM (torch._int_mm) %d x %d x %d: best %.3f ms = %.0f TOP/s, sustained %.3f ms = %.0f TOP/s"
          % (M, N, K, best, 2.0 * M * N * K / best / 1e9, sus, 2.0 * M * N * K / sus / 1e9))
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.split()[0]
print("theoretical: 148 SMs x 16384 int8 ops/clk x %s MHz = %.0f TOP/s" % (clk, 148 * 16384 * float(clk) / 1e6))
ctx.close()
