"""Build libcmpc_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcmpc_b200.so")
HOSTLIB = os.path.join(HERE, "libcmpc_host.so")
SOURCES = [os.path.join(HERE, "csrc", "cmpc_kernels.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in ("cmpc_kernels.cu", "cmpc_core.cuh", "cmpc_warp.cuh", "cmpc_ipm.cuh", "cmpc_layout.cuh", "cmpc_sparse.cuh", "cmpc_populate.cuh")]
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "cmpc_b200.h"))
HOST_SOURCES = [os.path.join(HERE, "host", f) for f in ("CentroidalMPC.cpp", "IniParametersHandler.cpp", "Contacts.cpp", "BlockUtilities.cpp", "capi.cpp")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550,177"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcmpc_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, DEPS):
        cmd = [_nvcc(), *NVCC_FLAGS, *SOURCES, "-o", LIB, "-lcudart"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    host_src = [s for s in HOST_SOURCES if os.path.exists(s)]
    hdir = os.path.join(HERE, "host", "BipedalLocomotion")
    hdir2 = os.path.join(HERE, "host", "CentroidalMPCWalking")
    hdeps = host_src + [os.path.join(d, f) for d in (hdir, hdir2) if os.path.isdir(d) for f in os.listdir(d)] + [LIB]
    if host_src and (force or _stale(HOSTLIB, hdeps)):
        inc = os.path.join(os.path.dirname(HERE), "include")
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", inc, "-I", os.path.join(HERE, "host"), *host_src, "-o", HOSTLIB,
               "-L", HERE, "-lcmpc_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
