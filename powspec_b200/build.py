"""Build libpowspec_b200.so in-tree with nvcc for sm_100a.

    python -m powspec_b200.build [--force]

The library is the product: hand-written CUDA kernels (csrc/assign.cu,
csrc/binning.cu, csrc/fft_strided.cu, csrc/generate.cu), the host orchestration + C ABI
(csrc/context.cu) and the reference-ABI seam (csrc/refabi.cpp).  It links
cuFFT dynamically and the CUDA runtime statically.  No GPU is needed to build.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libpowspec_b200.so")

SOURCES = ["assign.cu", "assign_tiles.cu", "binning.cu", "fft_strided.cu", "cnvt.cu", "ingest.cu", "generate.cu", "context.cu", "dist.cu", "hostcopy.cpp", "refabi.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _cuda_lib_dir(nvcc: str) -> str:
    root = os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
    for sub in ("lib64", os.path.join("targets", "x86_64-linux", "lib")):
        d = os.path.join(root, sub)
        if os.path.exists(os.path.join(d, "libcufft.so")):
            return d
    return os.path.join(root, "lib64")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "psb_internal.h"), os.path.join(CSRC, "psb_context.h"), os.path.join(CSRC, "assign_common.cuh"),
               os.path.join(ROOT, "include", "powspec_b200.h"),
               os.path.join(ROOT, "include", "powspec_refabi.h")]
    # the image exports CC/CXX pointing at a wrapper; pin the system compiler
    host = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    common = [nvcc, *ARCH, *host, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(ROOT, "include")]

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        if force or _stale(o, [s, *headers]):
            cmd = common + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        libdir = _cuda_lib_dir(nvcc)
        cmd = [nvcc, *ARCH, *host, "-shared", "-o", LIB, *objs, "-L", libdir, "-lcufft",
               "-Xlinker", "-rpath", "-Xlinker", libdir, "-Xlinker", "-Bsymbolic"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
