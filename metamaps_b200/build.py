"""Builds metamaps_b200/libmetamaps_b200.so (the C-ABI library) with nvcc for sm_100a, in-tree."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "mm_lib.cu")
OUT = os.path.join(HERE, "libmetamaps_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared", "-I" + os.path.join(ROOT, "include"), "-lgomp"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(ROOT, "include", "metamaps_b200.h"))
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [SRC, "-o", OUT]
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)       # the image exports a wrapper g++ without libgomp.spec
    subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], check=True, env=env)
    return OUT


HOST_SRC = os.path.join(HERE, "csrc", "host", "metamaps_main.cpp")
HOST_BIN = os.path.join(HERE, "metamaps")


def build_host(lib_dir: str = HERE, lib_name: str = "metamaps_b200", out: str = HOST_BIN, force: bool = False) -> str:
    """The C++ host (`metamaps mapDirectly|classify`) linked against the C-ABI library."""
    deps = [HOST_SRC, os.path.join(HERE, "csrc", "host", "mm_fastx.hpp"), os.path.join(ROOT, "include", "metamaps_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    subprocess.run(["/usr/bin/g++", "-O3", "-std=c++17", "-fopenmp", "-pthread", HOST_SRC, "-o", out, "-L" + lib_dir, "-l" + lib_name,
                    "-Wl,-rpath," + lib_dir + ":$ORIGIN", "-lz"], check=True)
    return out


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
    print(build_host(force=True))
