"""Build libkgeo_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libkgeo_b200.so")
SOURCES = [os.path.join(CSRC, "kgeo_b200.cu")]
HEADERS = [os.path.join(CSRC, h) for h in ("kg_special.cuh", "kg_math.cuh", "kg_geodesic.cuh", "kg_image.cuh", "kg_fieldline.cuh", "kg_offeq.cuh", "kg_theta.cuh", "kg_theta_layout.cuh", "kg_hot.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "kgeo_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=true",
              "-diag-suppress", "177", "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libkgeo_b200.so")


def is_stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources.  Returns the path."""
    if not force and not is_stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIBPATH] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
