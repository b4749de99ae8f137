"""Builds libssym.so (sm_100a CUDA + C-ABI) and the verify-batch CLI, in-tree.

    python stark-symphony_b200/build.py [--force]

nvcc cross-compiles without a GPU; the artefacts are git-ignored but travel with the tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libssym.so")
CLI = os.path.join(HERE, "bin", "verify-batch")

CU_SOURCES = ["stwo_kernels.cu", "prover_kernels.cu", "s101_kernels.cu", "jets_kernels.cu", "wit_kernels.cu", "compact_kernels.cu", "api.cu"]
CPP_SOURCES = ["witness.cpp", "cost.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libssym cannot be built (there is no CPU fallback)")


def _deps_mtime() -> float:
    m = 0.0
    for d in (CSRC, INCLUDE):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return max(m, os.path.getmtime(__file__))


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False, tuning: bool = False, synccheck: bool = False) -> str:
    """tuning=True adds -DSSYM_TUNING: every kernel variant of the experiments DESIGN.md section 4 reports (SSYM_ADDMODE, SSYM_ROLLED, SSYM_CHANNEL_NP,
    SSYM_FRONT, SSYM_M31_INV_K environment switches).  The release library has none of them."""
    if tuning and "-DSSYM_TUNING" not in NVCC_FLAGS:
        NVCC_FLAGS.append("-DSSYM_TUNING")
        force = True
    if synccheck and "-DSSYM_SYNCCHECK" not in NVCC_FLAGS:  # named barriers behind one program location (csrc/stwo_kernels.cu: k1_bar_rs), for tools/sanitize.sh
        NVCC_FLAGS.append("-DSSYM_SYNCCHECK")
        force = True
    extra = os.environ.get("SSYM_NVCC_EXTRA", "").split()  # experiment builds: extra -D switches (e.g. -DK1_R_ADDMODE=0)
    if extra:
        NVCC_FLAGS.extend(f for f in extra if f not in NVCC_FLAGS)
        force = True
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    stale = force or not os.path.exists(LIB) or not os.path.exists(CLI) or min(os.path.getmtime(LIB), os.path.getmtime(CLI)) < _deps_mtime()
    if not stale:
        return LIB
    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(_compile, CU_SOURCES + CPP_SOURCES))
    # default visibility only for the extern "C" API (marked in the sources via SSYM_API? no: export everything extern "C")
    cmd = [_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
    subprocess.run(cmd, check=True)
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-o", CLI, os.path.join(CSRC, "verify_batch.cpp"), "-L" + HERE, "-lssym", "-lpthread",
           "-Wl,-rpath,$ORIGIN/.."]
    subprocess.run(cmd, check=True)
    if verbose:
        for o in objs:
            print(open(o + ".log").read())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv, synccheck="--synccheck" in sys.argv))
