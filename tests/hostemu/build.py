"""Build the TEST-ONLY host compile of the device core (see hostemu.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostemu.so")
SO32 = os.path.join(HERE, "_hostemu32.so")      # the same with 32-lane warps (simt.h fibers)
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = [os.path.join(HERE, "hostemu.cpp")] + [
    os.path.join(ROOT, "montecarlotreesearch_b200", "csrc", f) for f in ("qcore.cuh", "ttt_core.cuh", "philox.cuh", "quoridor_lockstep_body.cuh", "quoridor_uni_lockstep_body.cuh", "rollout_kernels.cuh", "tree_body.cuh")
] + [os.path.join(ROOT, "include", "mctsb.h")]


def build(force=False, warp32=False):
    so = SO32 if warp32 else SO
    deps = SRC + [os.path.join(HERE, "simt.h"), os.path.abspath(__file__)]
    if force or not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DMCTSB_HOST_EMU"] + (["-DMCTSB_EMU_WARP32"] if warp32 else []) +
            ["-I/usr/local/cuda/include", "-x", "c++", SRC[0], "-o", so]
        )
    return so


if __name__ == "__main__":
    print(build(force=True)); print(build(force=True, warp32=True))
