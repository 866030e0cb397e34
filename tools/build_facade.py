"""Builds build/facade_check from examples/facade_check.cpp with plain g++ against the in-tree C-ABI library
(no CUDA headers needed above the ABI)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build() -> str:
    out = os.path.join(ROOT, "build", "facade_check")
    src = os.path.join(ROOT, "examples", "facade_check.cpp")
    deps = [src, os.path.join(ROOT, "include", "rosdyn_b200", "chain.hpp"), os.path.join(ROOT, "include", "rosdyn_b200.h"),
            os.path.join(ROOT, "rosdyn_b200", "librosdyn_b200.so")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        build_group_check()
        return out
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                           "-L", os.path.join(ROOT, "rosdyn_b200"), "-lrosdyn_b200", "-Wl,-rpath,$ORIGIN/../rosdyn_b200"])
    build_group_check()
    return out


def build_group_check() -> str:
    """examples/group_check.cpp: the NCCL group of the C-ABI from plain C++ (g++, libcudart for the device buffers)."""
    out = os.path.join(ROOT, "build", "group_check")
    src = os.path.join(ROOT, "examples", "group_check.cpp")
    deps = [src, os.path.join(ROOT, "include", "rosdyn_b200.h"), os.path.join(ROOT, "rosdyn_b200", "librosdyn_b200.so")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", src, "-o", out,
                           "-L", os.path.join(ROOT, "rosdyn_b200"), "-lrosdyn_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath,$ORIGIN/../rosdyn_b200", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return out


if __name__ == "__main__":
    print(build())
