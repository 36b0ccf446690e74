"""The reference's OWN, unmodified tree and game code with its pthread JobScheduler swapped at link time for
integration/gpu_jobscheduler.cpp (oracle/_ref/libref_gpu_dropin.so, built by `make -C oracle dropin`):
RolloutJobs scheduled by the reference's MCTS_node::rollout() (mcts.cpp:57-81) are played on the B200 through
the C ABI.  CPU part: the library links, the state packer is faithful on the whole corpus, and it refuses to
run without a device.  GPU part: the reference tree on the GPU scheduler builds exactly the tree this repo's
host tree builds, node for node."""
import ctypes as C
import os

import numpy as np
import pytest

import montecarlotreesearch_b200 as m
from oracle import reflib
from tests.util import golden

DROPIN = os.path.join(reflib.REF_DIR, "libref_gpu_dropin.so")
pytestmark = pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/libref_gpu_dropin.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def dropin():
    m.build()
    return reflib.RefLib("gpu_dropin")


def test_dropin_replaces_the_scheduler_symbols(dropin):
    # the reference's JobScheduler member functions are defined by the drop-in object, and the library
    # depends on the CUDA backend — not on a second copy of the pthread pool
    for sym in ("_ZN12JobScheduler8scheduleEP3Job", "_ZN12JobScheduler25waitUntilJobsHaveFinishedEi",
                "_ZN12JobSchedulerC1Ej", "_ZN12JobSchedulerD1Ev", "ref_grow_tree", "mctsb_dropin_reset"):
        assert hasattr(dropin.lib, sym), sym
    import subprocess
    needed = subprocess.run(["readelf", "-d", DROPIN], capture_output=True, text=True).stdout
    assert "libmctsb.so" in needed


def test_dropin_packer_is_faithful_on_the_corpus(dropin):
    """reference Quoridor_state (walls[][] / wall_connections[][] only) -> packed record: the orientation of
    every wall is recovered; compared through the raw arrays the reference's rules read"""
    corpus = golden("corpus_q.npy")
    for i in range(0, len(corpus), 3):
        ok, back = dropin.q_pack_roundtrip(corpus[i])
        assert ok, i
        assert np.array_equal(reflib.raw_from_packed(back), reflib.raw_from_packed(corpus[i])), i
        for f in ("wx", "wy", "bx", "by", "wwalls", "bwalls", "turn", "move_counter"):
            assert back[f] == corpus[i][f], (i, f)
    ok, back = dropin.q_pack_roundtrip(reflib.opening_state())
    assert ok and back["h_anchor"] == 0 and back["v_anchor"] == 0


def _gpu_sequential_backend(be, log):
    """RolloutBackend callback for this repo's host tree: GPU rollouts, summed in job order in double precision
    exactly as the reference's aggregation loop does (mcts.cpp:68-75)."""
    def cb(game, policy, states, n_leaves, R, first_id, out, user):
        for i in range(n_leaves):
            if game == m.GAME_QUORIDOR:
                s = np.frombuffer(C.string_at(states + 32 * i, 32), dtype=m.QSTATE_DTYPE)
            else:
                s = np.frombuffer(C.string_at(states + 4 * i, 4), dtype=m.TSTATE_DTYPE)
            o, _ = be.rollout_outcomes(game, policy, s, R, first_id + i * R, finals=False)
            acc = 0.0
            for v in o["score"]:
                acc += float(v)
            out[i] = acc
        log.append(n_leaves * R)
    return cb


@pytest.mark.gpu
@pytest.mark.parametrize("game,iters", [(m.GAME_TICTACTOE, 3000), (m.GAME_QUORIDOR, 400)])
def test_reference_tree_on_the_gpu_scheduler_equals_host_tree(dropin, game, iters):
    from tests.hostlib import Session, load
    assert dropin.dropin_device_ok()
    host = load()
    dropin.dropin_reset(host.mcts_host_peek_rollout_id())                    # both trees draw the same Philox stream ids
    ref = dropin.grow_tree(0 if game == m.GAME_QUORIDOR else 1, iters)      # the reference's grow_tree, rollouts on the GPU
    flushes, rollouts = dropin.dropin_counters()
    assert flushes == iters and rollouts == 4 * iters                        # NUMBER_OF_THREADS = 4 jobs per iteration

    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    log = []
    root = m.quoridor_opening() if game == m.GAME_QUORIDOR else np.zeros((), dtype=m.TSTATE_DTYPE)
    mine = Session(host, game, root, rollouts_per_leaf=4, leaves_per_flush=1, callback=_gpu_sequential_backend(be, log))
    mine.grow(iters)
    st = mine.root_stats()
    mine.close()
    be.close()
    assert sum(log) == 4 * iters
    assert st["sims"] == ref["root_sims"] and st["score"] == ref["root_score"] and st["size"] == ref["size"]
    assert st["children"].shape == ref["children"].shape
    assert np.array_equal(st["children"], ref["children"])                   # move id, simulations, score: bit for bit


@pytest.mark.gpu
def test_rollout_jobs_on_the_gpu_scheduler(dropin):
    """JobScheduler::schedule(new RolloutJob(...)) x N + waitUntilJobsHaveFinished(), the reference's own calls"""
    dropin.dropin_reset(1 << 20)
    secs, mean = dropin.jobscheduler_bench(0, 4, 4096)
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    out, _ = be.rollout_outcomes(m.GAME_QUORIDOR, m.POLICY_REF, m.quoridor_opening(), 4096, 1 << 20, finals=False)
    be.close()
    assert mean == float(np.sum(out["score"].astype(np.float64)) / 4096) or abs(mean - out["score"].mean()) < 1e-12
    assert dropin.dropin_counters() == (1, 4096)


def test_dropin_refuses_to_run_without_a_device(dropin):
    if dropin.dropin_device_ok():
        pytest.skip("a CUDA device is present")
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from oracle import reflib; "
            "reflib.RefLib('gpu_dropin').grow_tree(1, 10)" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources not present")
def test_reference_example_main_links_against_the_gpu_scheduler(tmp_path):
    """the recipe of INTEGRATION.md (C): the reference's own Quoridor REPL, its tree and its game, with the
    drop-in object in place of JobScheduler.cpp — compiles and links without touching a reference file"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = "/root/reference"
    obj = str(tmp_path / "gpu_jobscheduler.o")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-w", "-DMCTSB_DROPIN_ONLY_QUORIDOR", "-I" + ref, "-I" + os.path.join(root, "include"),
                           "-include", os.path.join(root, "integration", "open_classes.h"),
                           "-c", os.path.join(root, "integration", "gpu_jobscheduler.cpp"), "-o", obj])
    exe = str(tmp_path / "quoridor_gpu")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-w", "-pthread", "-I" + ref,
                           ref + "/mcts/src/mcts.cpp", ref + "/examples/Quoridor/Quoridor.cpp",
                           ref + "/examples/Quoridor/main.cpp", obj,
                           "-L" + m.PKG_DIR, "-lmctsb", "-Wl,-rpath," + m.PKG_DIR, "-o", exe])
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True).stdout
    assert "JobScheduler::waitUntilJobsHaveFinished" in syms and "mctsb_rollout_outcomes" in syms
    assert "thread_code" in syms and "pthread_create" not in syms          # no pool threads are ever created


@pytest.mark.gpu
def test_reference_repl_binary_plays_on_the_gpu():
    """oracle/_ref/quoridor_gpu = the reference's examples/Quoridor/main.cpp + mcts.cpp + Quoridor.cpp, unmodified,
    linked with the drop-in scheduler.  `genmove` grows the reference's tree (MAXITER 20 000 iterations x 4
    RolloutJobs, main.cpp:6,145) with every rollout played by the CUDA kernels, then prints its move."""
    import re
    import subprocess
    exe = os.path.join(reflib.REF_DIR, "quoridor_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/quoridor_gpu not built")
    r = subprocess.run([exe], input="genmove\nwinner\nq\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-400:]
    sims = re.search(r"Number of simulations: (\d+)", r.stdout)
    assert sims and int(sims.group(1)) >= 4 * 1000          # the tree was grown through the GPU scheduler
    assert "Best moves:" in r.stdout and "Illegal" not in r.stdout and "illegal" not in r.stderr
