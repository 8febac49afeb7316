"""Builds libipcb200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python ipc-toolkit_b200/build.py [--force]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3; -fmad=false so that
the FP64 classification arithmetic rounds exactly like the CPU oracle (explicit
fma() calls in the eigen-solver are still fused).  One object per .cu, compiled
in parallel, then linked into one shared library with no torch dependency.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libipcb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["api.cu", "broad.cu", "collisions.cu", "potential.cu", "friction.cu", "ccd.cu"]
# -fmad=false everywhere: decisions taken in floating point (boxes, distance types, the dhat filter, CCD) must round
# exactly like the CPU oracle, and in potential.cu an implicit contraction a * b - c * d -> fma(a, b, -(c * d)) would turn
# the EXACT zeros of symmetric / axis-aligned configurations into rounding residue — entries the reference's sparse
# matrix does not have (local_to_global.hpp:290-291).  Explicit fma() calls (Jacobi rotations) are still fused.
FMAD_OK = set()
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr",
    "-Xcudafe", "--diag_suppress=177",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ipcb200.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer(o, [s] + headers):
            cmd = [NVCC] + FLAGS + ([] if src in FMAD_OK else ["-fmad=false"]) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=6) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            print(l)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or force or _newer(OUT, objs):
        run([NVCC, "--shared", "-o", OUT] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
