"""Builds libquadsim.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.  No JIT cache: the .so sits next to
this file so that it travels to the GPU box with the repository snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libquadsim.so")
SOURCES = [os.path.join(CSRC, "quadsim_capi.cu")]
HEADERS = [os.path.join(CSRC, h) for h in ("quadsim_kernels.cuh", "quadsim_policy.cuh", "quadsim_rollout.cuh", "quadsim_train.cuh")] + \
          [os.path.join(ROOT, "include", "quadsim.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fopenmp", "-shared", "-I" + os.path.join(ROOT, "include"), "-lgomp"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(p):
        raise RuntimeError("nvcc not found: cannot build libquadsim.so")
    return p


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False, out=None, extra_flags=()):
    """Compile if missing or older than its sources; returns the path of the shared library.
    ``out`` / ``extra_flags`` build an experimental variant (e.g. -DQS_EXP_NOMLP) next to the product library."""
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags, "-o", out, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
