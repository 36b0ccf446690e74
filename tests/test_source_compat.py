"""CPU, where /root/reference exists: programs written against the reference's headers — its own two
main.cpp files, unmodified — compile against this repo's host headers, link against the two shared
libraries, and (without a GPU) fail loudly instead of falling back to a CPU rollout."""
import os
import subprocess

import pytest

import montecarlotreesearch_b200 as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HOST = os.path.join(ROOT, "montecarlotreesearch_b200", "host")


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    if not os.path.isdir(REF):
        pytest.skip("reference sources not present")
    m.build()
    d = tmp_path_factory.mktemp("compat")
    for game in ("TicTacToe", "Quoridor"):
        os.makedirs(d / "examples" / game)
        os.symlink(os.path.join(REF, "examples", game, "main.cpp"), d / "examples" / game / "main.cpp")
        os.symlink(os.path.join(HOST, game + ".h"), d / "examples" / game / (game + ".h"))
        os.symlink(os.path.join(HOST, "state.h"), d / "examples" / game / "state.h")
    os.makedirs(d / "mcts" / "include")
    for f in ("mcts.h", "state.h", "rollout_backend.h"):
        os.symlink(os.path.join(HOST, f), d / "mcts" / "include" / f)
    os.symlink(os.path.join(ROOT, "montecarlotreesearch_b200", "csrc"), d / "examples" / "csrc")
    return d


@pytest.mark.parametrize("game", ["TicTacToe", "Quoridor"])
def test_reference_mains_compile_and_link_unmodified(tree, game):
    exe = tree / (game + "_demo")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "examples/%s/main.cpp" % game,
           "-L", m.PKG_DIR, "-lmcts_host", "-lmctsb", "-Wl,-rpath," + m.PKG_DIR, "-o", str(exe)]
    res = subprocess.run(cmd, cwd=tree, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert os.path.exists(exe)


@pytest.mark.skipif(m.device_count() > 0, reason="needs a CPU-only host")
def test_demo_refuses_to_roll_out_without_a_gpu(tree):
    exe = tree / "TicTacToe_demo"
    if not os.path.exists(exe):
        pytest.skip("demo not built")
    res = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert res.returncode != 0
    assert "no CUDA device" in res.stdout or "BackendError" in res.stdout or "mctsb_create" in res.stdout
