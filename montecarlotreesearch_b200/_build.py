"""In-tree build of the CUDA backend: nvcc -> montecarlotreesearch_b200/libmctsb.so (sm_100a only).

The shared object is the C-ABI drop-in boundary (include/mctsb.h).  It is git-ignored but travels
to the GPU box with gpurun.  nvcc cross-compiles without a GPU, so this also runs on CPU-only hosts.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmctsb.so")
HOSTLIB = os.path.join(PKG, "libmcts_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",           # the score arithmetic must round like the host reference (no contraction)
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def all_inputs():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "mctsb.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def build_cuda(force=False, verbose=False):
    if not force and not _stale(LIB, all_inputs()):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + cuda_sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(PKG, "ptxas_report.txt"), "w") as f:
        f.write(res.stdout)
    return LIB


def build_host(force=False, verbose=False):
    """C++ host side (tree, game classes, GPU rollout backend glue) -> libmcts_host.so."""
    hdir = os.path.join(PKG, "host")
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir)) if f.endswith(".cpp")]
    if not srcs:
        return None
    deps = [os.path.join(hdir, f) for f in os.listdir(hdir)] + all_inputs()
    if not force and not _stale(HOSTLIB, deps):
        return HOSTLIB
    cmd = ["g++", "-O2", "-ftree-vectorize", "-fno-math-errno", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-I", os.path.join(ROOT, "include"),
           "-o", HOSTLIB] + srcs + ["-L", PKG, "-lmctsb", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("g++ failed: " + " ".join(cmd))
    return HOSTLIB


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv, verbose=True)
