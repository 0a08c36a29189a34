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


def _compile(nvcc, src, obj):
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-ccbin", "/usr/bin/g++", "-c", src, "-o", obj]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, proc.returncode, proc.stdout


def build_library(force=False, verbose=False):
    """One object per .cu file (compiled in parallel, only when stale), then one link."""
    if not force and not _stale():
        return LIBPATH
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(os.path.dirname(HERE), "build", "obj")     # git- and gpurun-ignored
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    headers = glob.glob(os.path.join(CSRC, "*.cuh"))
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        newest = max(os.path.getmtime(p) for p in [src] + headers)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(lambda j: _compile(nvcc, *j), jobs))
    for src, rc, out in results:
        if verbose or rc != 0:
            sys.stdout.write(out)
        if rc != 0:
            raise RuntimeError("nvcc failed compiling %s (exit %d)" % (os.path.basename(src), rc))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", LIBPATH] + objs
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stdout.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed linking libpyglm_b200.so (exit %d)" % proc.returncode)
    return LIBPATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
