"""Builds libbos_b200.so (CUDA kernels + C ABI, sm_100a only) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box too.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbos_b200.so")

CU_SOURCES = ["linearize.cu", "solve_pcg.cu", "solve_dense.cu", "misc.cu", "setup.cu", "ctx.cu"]
CPP_SOURCES = ["pattern.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))) + [os.path.join("..", "..", "include", "bos_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_info=False, extra_flags=None, out=None, objdir=None):
    """extra_flags / out / objdir build an experimental variant (e.g. -DBOS_EPT=4) beside the default library."""
    global OBJ, LIB
    saved = (OBJ, LIB)
    if objdir:
        OBJ = objdir
    if out:
        LIB = out
    try:
        return _build(force, verbose, ptxas_info, list(extra_flags or []))
    finally:
        OBJ, LIB = saved


def _build(force, verbose, ptxas_info, extra_flags):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in CU_SOURCES + CPP_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc()] + NVCC_FLAGS + extra_flags + (["-Xptxas", "-v"] if ptxas_info else []) + ["-x", "cu", "-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(run, jobs):
                if verbose or ptxas_info:
                    sys.stderr.write(out)
    if jobs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
