"""In-tree build of libb200hist.so (sm_100a only) and of the test oracle.

nvcc cross-compiles without a GPU.  Objects go to build/, the library to xgboost_b200/lib/ so that
it travels with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "xgboost_b200", "csrc")
LIBDIR = os.path.join(ROOT, "xgboost_b200", "lib")
OBJDIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(LIBDIR, "libb200hist.so")
ORACLE_SRC = os.path.join(ROOT, "oracle", "hist_oracle.c")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]
SOURCES = ["ellpack.cu", "quantise.cu", "histogram.cu", "evaluate.cu", "partition.cu", "updater.cu",
           "components.cu", "comm.cu", "sketch.cu", "predict.cu", "sampling.cu", "colsample.cu"]


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _headers():
    out = [os.path.join(ROOT, "include", "b200hist.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_library(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = _headers()
    objs, jobs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src + ".o")
        stamp = obj + ".sha"
        dig = _digest([path] + hdrs)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((path, obj, stamp, dig))

    def compile_one(job):
        path, obj, stamp, dig = job
        extra = ["-Xptxas", "-v"] if verbose else []
        out = _run([NVCC] + NVCC_FLAGS + extra + ["-c", path, "-o", obj])
        with open(stamp, "w") as f:
            f.write(dig)
        return out

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(compile_one, jobs):
                if verbose and out.strip():
                    print(out)
    if jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


def build_oracle() -> str:
    """Compile the CPU restatement (test infrastructure only)."""
    if (not os.path.exists(ORACLE_LIB)) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(ORACLE_SRC):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", ORACLE_LIB,
              ORACLE_SRC, "-lm"])
    return ORACLE_LIB


if __name__ == "__main__":
    print(build_library(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build_oracle())
