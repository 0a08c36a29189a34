"""Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m pyglm_b200.build [--force]

Output: pyglm_b200/_lib/libpyglm_b200.so (git-ignored, but shipped to the GPU box by gpurun).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIBPATH = os.path.join(LIBDIR, "libpyglm_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh"))
    return any(os.path.getmtime(p) > t for p in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-o", LIBPATH] + sources()
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stdout.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libpyglm_b200.so (exit %d)" % proc.returncode)
    return LIBPATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
