"""Build libadmm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadmm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".cpp", ".hpp"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "admm_b200.h"))
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    if not force and _newer(LIB, srcs):
        return LIB
    obj_mesh = os.path.join(CSRC, "mesh.o")
    cmds = [
        ["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-c", os.path.join(CSRC, "mesh.cpp"), "-o", obj_mesh],
        [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", "-shared",
         *(["-Xptxas", "-v"] if verbose else []),
         os.path.join(CSRC, "lib.cu"), obj_mesh, "-o", LIB, "-lgomp"],
    ]
    for c in cmds:
        r = subprocess.run(c, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(c), r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
