"""Builds libnblic_b200.so in-tree (nvcc cross-compiles sm_100a without a GPU).

The shared library is the product: CUDA kernels + batch C ABI (csrc/stream_kernels.cu) and the
reference codec's five drop-in entry points in C (csrc/nblic_dropin.c).  It is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnblic_b200.so")
CLI = os.path.join(HERE, "nblic_batch")  # batch command-line front end (csrc/nblic_batch_cli.c)
SOURCES_CU = ["stream_kernels.cu"]
SOURCES_C = ["nblic_dropin.c"]
HEADERS = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith((".cuh", ".h", ".c"))] + \
          [os.path.join(HERE, "..", "include", "nblic_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


LIB_SEQ = os.path.join(HERE, "libnblic_b200_seq.so")  # test build: the product + the sequential one-agent-per-stream kernels


def build_library(force: bool = False, verbose: bool = False, sequential: bool = False) -> str:
    """sequential=True builds libnblic_b200_seq.so (-DNBLIC_B200_SEQUENTIAL): the same library plus coder_kernel, the plain
    sequential formulation the parity tests use as a second implementation (MAP_LANE).  Not part of the product."""
    lib = LIB_SEQ if sequential else LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES_CU + SOURCES_C]
    if not force and not _stale(lib, srcs + HEADERS + [os.path.abspath(__file__)]):
        return lib
    objdir = os.path.join(HERE, "build", "seq" if sequential else "product")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for s in SOURCES_CU:
        o = os.path.join(objdir, s + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DNBLIC_B200_SEQUENTIAL"] if sequential else []), *os.environ.get("NBLIC_NVCC_EXTRA", "").split(),
               "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
        objs.append(o)
    for s in SOURCES_C:
        o = os.path.join(objdir, s + ".o")
        subprocess.run(["gcc", "-O2", "-fPIC", "-Wall", "-Wextra", "-std=gnu99", "-c", os.path.join(CSRC, s), "-o", o], check=True)
        objs.append(o)
    subprocess.run([_nvcc(), "-shared", "-o", lib, *objs, "-lpthread", "-cudart", "shared"], check=True)
    if sequential:
        return lib
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-std=gnu99", "-o", CLI, os.path.join(CSRC, "nblic_batch_cli.c"),
                    "-L" + HERE, "-lnblic_b200", "-lpthread", "-Wl,-rpath,$ORIGIN"], check=True)
    return LIB


def source_fingerprint() -> str:
    """16 hex digits over the kernel sources and compile flags: stamps ncu captures so that bench.py only quotes
    profile figures taken from the code it is running (the GPU box has no .git to ask for HEAD)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh")):
            h.update(name.encode())
            h.update(open(os.path.join(CSRC, name), "rb").read())
    return h.hexdigest()[:16]


def build_microbench() -> str:
    """tools/microbench/int_issue_peak: the integer-issue peak measurement the roofline is quoted against."""
    src = os.path.join(HERE, "..", "tools", "microbench", "int_issue_peak.cu")
    out = os.path.join(HERE, "..", "tools", "microbench", "int_issue_peak")
    if _stale(out, [src]):
        subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-o", out, src], check=True)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_library(force="--force" in sys.argv, sequential=True))
    print(build_microbench())
