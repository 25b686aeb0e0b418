"""Builds libsmafa_b200.so (CUDA kernels + C ABI + C++ host) and the `smafa` CLI for sm_100a.

Run as `python -m smafa_b200.build` or through __graft_entry__.build().  nvcc cross-compiles
without a GPU; the outputs stay in-tree (git-ignored) so they travel to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsmafa_b200.so")
BIN_DIR = os.path.join(HERE, "bin")
CLI = os.path.join(BIN_DIR, "smafa")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
                     "--expt-relaxed-constexpr"]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-I/usr/local/cuda/include"]

CU_SOURCES = ["pack.cu", "scan_popc.cu", "scan_mma.cu", "guess.cu", "finalize.cu", "merge.cu", "sharded.cu", "api.cu"]
CXX_SOURCES = ["seqio.cpp", "commands.cpp"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(BIN_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    headers += [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith(".hpp")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "smafa_b200.h"))
    objs = []
    procs = []
    for src in CU_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _newer([s] + headers, o):
            procs.append((subprocess.Popen([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                           stderr=subprocess.STDOUT, text=True), s))
    for src in CXX_SOURCES + ["main.cpp"]:
        s, o = os.path.join(HOST, src), os.path.join(OBJ, src + ".o")
        if src != "main.cpp":
            objs.append(o)
        if force or _newer([s] + headers, o):
            procs.append((subprocess.Popen(["g++"] + CXX_FLAGS + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                           stderr=subprocess.STDOUT, text=True), s))
    for p, s in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("compile failed: " + s)
        if verbose and out.strip():
            print(out)
    if force or procs or not os.path.exists(LIB):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lz", "-cudart", "static"])
    main_o = os.path.join(OBJ, "main.cpp.o")
    if force or procs or not os.path.exists(CLI):
        _run(["g++", "-o", CLI, main_o, "-L" + HERE, "-lsmafa_b200", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB, "and", CLI)
