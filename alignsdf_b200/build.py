"""Build libalignsdf_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object loaded with ctypes).

    python -m alignsdf_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libalignsdf_b200.so")
SOURCES = ["api.cu", "k1_simt.cu", "k1_tc.cu", "bind.cu", "mc.cu", "cc.cu", "nn.cu"]
# test / profiling build of the tensor-core kernel with its cycle counters and stage knock-outs
# (asdf_tc_eval_debug); never loaded by the product path
DEBUG_LIB = os.path.join(HERE, "libalignsdf_b200_debug.so")
DEBUG_SOURCES = ["api.cu", "k1_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), "include", "alignsdf_b200.h")
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [inc]
    for f in files:
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    from . import mc_tables
    mc_tables.write_header()
    stamp = LIB + ".sha256"
    dig = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read().strip() == dig):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


def build_debug(force: bool = False) -> str:
    stamp = DEBUG_LIB + ".sha256"
    dig = _digest()
    if (not force and os.path.exists(DEBUG_LIB) and os.path.exists(stamp)
            and open(stamp).read().strip() == dig):
        return DEBUG_LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DASDF_TC_DEBUG"] + [os.path.join(CSRC, s) for s in DEBUG_SOURCES] + ["-o", DEBUG_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return DEBUG_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--debug" in sys.argv:
        print(build_debug(force="--force" in sys.argv))
