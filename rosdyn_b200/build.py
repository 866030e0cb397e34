"""Build the C-ABI shared library IN-TREE: rosdyn_b200/librosdyn_b200.so (nvcc, sm_100a only).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the tree.
`python -m rosdyn_b200.build [--force] [--verbose]`
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJDIR = os.path.join(ROOT, "build")
LIB = os.path.join(HERE, "librosdyn_b200.so")
SOURCES = ["kernels.cu", "gram.cu", "gram_fused.cu", "components.cu", "aux.cu", "ik.cu", "group.cu", "capi.cu", "urdf.cpp", "solve.cpp", "fold.cpp"]
HEADERS = ["chain_dev.h", "spatial.cuh", "launch.h", "gram_common.cuh", os.path.join("..", "..", "include", "rosdyn_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-ccbin", "/usr/bin/g++"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=5) as ex:
        list(ex.map(run, jobs))
    if force or jobs or not _newer(LIB, objs):
        run([NVCC, "-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
