"""Compile-time guard (CPU, nvcc cross-compiles): the hot loops of both scan kernels must stay free of local-memory
traffic.  The POPC kernel once lost half its rate (4.3e12 -> 2.2e12 comparisons/s) without any change to its own
code: a heavier out-of-line slow path made ptxas spill the query planes and reload them inside the inner loop."""
import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "smafa_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
         "-Xptxas", "-v"]


def _compile(src, out):
    r = subprocess.run([NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    stats = {}
    text = r.stderr + r.stdout
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes "
                         r"spill stores, (\d+) bytes spill loads\s*\n[^\n]*Used (\d+) registers", text):
        stats[m.group(1)] = dict(stack=int(m.group(2)), st=int(m.group(3)), ld=int(m.group(4)), regs=int(m.group(5)))
    return stats


def _loop_local_accesses(obj, kernel):
    """LDL/STL instructions between the first and the last POPC of a kernel's SASS (= inside its scan loop)."""
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    body, on = [], False
    for line in sass.splitlines():
        if "Function : " in line:
            on = kernel in line
            continue
        if on:
            body.append(line)
    pop = [i for i, l in enumerate(body) if " POPC " in l]
    assert pop, kernel
    return sum(1 for i, l in enumerate(body) if pop[0] < i < pop[-1] and ("LDL" in l or "STL" in l))


@pytest.fixture(scope="module")
def compiled(tmp_path_factory):
    d = tmp_path_factory.mktemp("ptxas")
    with ThreadPoolExecutor(2) as ex:
        a = ex.submit(_compile, "scan_popc.cu", str(d / "popc.o"))
        b = ex.submit(_compile, "scan_mma.cu", str(d / "mma.o"))
        return {"popc": a.result(), "mma": b.result(), "popc_obj": str(d / "popc.o")}


def test_popc_scan_loop_has_no_local_memory_traffic(compiled):
    kernels = {k: v for k, v in compiled["popc"].items() if "scan_popc_kernel" in k}
    assert len(kernels) == 12  # PW x R x EARLY x alphabet
    for name, st in kernels.items():
        assert st["regs"] <= 80, (name, st)           # 3 CTAs of 256 threads per SM
        if name.endswith("ELb0EEEvNS_10ScanParamsEjj"):  # nucleotide instantiations (AA = false): the measured path
            assert st["ld"] <= 64 and st["st"] <= 96, (name, st)
            assert _loop_local_accesses(compiled["popc_obj"], name) == 0, name


def test_mma_scan_kernels_do_not_spill(compiled):
    kernels = {k: v for k, v in compiled["mma"].items() if "scan_mma_kernel" in k}
    assert kernels
    for name, st in kernels.items():
        assert st["st"] == 0 and st["ld"] == 0, (name, st)
        assert st["regs"] <= 144, (name, st)          # 448 threads, one CTA per SM
