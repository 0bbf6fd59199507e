"""canonicalvoting_b200/build.py -- compile csrc/*.cu into the in-tree C-ABI library.

    python -m canonicalvoting_b200.build [--force] [--verbose]

Produces canonicalvoting_b200/_C/libcvb200.so for sm_100a ONLY
(`-gencode arch=compute_100a,code=sm_100a -lineinfo`).  nvcc cross-compiles without
a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
No --use_fast_math: the vote op's voxel indices depend on accurate cosf/sinf/div.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "_C")
OBJ_DIR = os.path.join(CSRC, "build")
LIB = os.path.join(OUT_DIR, "libcvb200.so")

NVCC = os.environ.get("CVB200_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "cvb200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile (if stale) and return the path of libcvb200.so."""
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs, hdrs = _sources(), _headers()
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs + [__file__]):
            jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            print(r.stdout, r.stderr, flush=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOST_CXX, "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
