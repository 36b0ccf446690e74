"""Build the TEST-ONLY host compile of the device core (see hostemu.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostemu.so")
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = [os.path.join(HERE, "hostemu.cpp")] + [
    os.path.join(ROOT, "montecarlotreesearch_b200", "csrc", f) for f in ("qcore.cuh", "ttt_core.cuh", "philox.cuh", "quoridor_lockstep_body.cuh", "rollout_kernels.cuh")
] + [os.path.join(ROOT, "include", "mctsb.h")]


def build(force=False):
    if force or not os.path.exists(SO) or any(os.path.getmtime(f) > os.path.getmtime(SO) for f in SRC):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DMCTSB_HOST_EMU", "-I/usr/local/cuda/include",
             "-x", "c++", SRC[0], "-o", SO]
        )
    return SO


if __name__ == "__main__":
    print(build(force=True))
