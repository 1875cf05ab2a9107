"""Builds the native pieces in-tree (the .so files travel to the GPU box with the snapshot).

  titancna_b200/csrc/libtitancna_b200.so   CUDA kernels + C ABI (nvcc, sm_100a only)
  titancna_b200/csrc/TitanCNA_shim.so      the R `.Call` glue compiled against oracle/shim/R.h
                                           (test vehicle for the drop-in boundary; with a real R
                                           the same r_glue.c is compiled by R CMD SHLIB)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libtitancna_b200.so")
GLUE = os.path.join(CSRC, "TitanCNA_shim.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # host code must not contract a*b+c: the host-side exact tables mirror the reference's plain C
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-shared",
]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    """Every .cu is its own translation unit (compiled in parallel, nvcc cross-compiles without a GPU), linked into one
    shared object."""
    units = ["titan_api.cu", "titan_fb2.cu", "titan_util.cu"]
    headers = [os.path.join(CSRC, f) for f in ("titan_kernels.cuh", "titan_common.cuh", "titan_fb2.h", "titan_host.hpp")]
    headers.append(os.path.join(ROOT, "include", "titancna_b200.h"))
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("TITAN_NVCC_DEFS", "").split()   # e.g. -DTITAN_PROBE_PIECES=4
    jobs, objs = [], []
    for u in units:
        src, obj = os.path.join(CSRC, u), os.path.join(objdir, u.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd))
            jobs.append((cmd, subprocess.Popen(cmd)))
    for cmd, j in jobs:
        if j.wait() != 0:
            raise subprocess.CalledProcessError(j.returncode, cmd)
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    glue_src = [os.path.join(CSRC, "r_glue.c"), os.path.join(CSRC, "r_glue_fused.c")]
    shim = os.path.join(ROOT, "oracle", "shim")
    shim_src = [os.path.join(shim, f) for f in ("shim.c", "R.h")]
    if all(os.path.exists(g) for g in glue_src) and os.path.isdir(shim) and (force or _stale(GLUE, glue_src + shim_src + [LIB])):
        cc = shutil.which("gcc") or "gcc"
        cmd = [cc, "-O2", "-Wall", "-fPIC", "-shared", "-I", shim, "-I", os.path.join(ROOT, "include"), "-o", GLUE,
               *glue_src, os.path.join(shim, "shim.c"), "-L", CSRC, "-ltitancna_b200", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
