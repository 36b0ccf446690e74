"""The C++ host side (host/*.cpp): game classes against the reference goldens, and the tree logic
(selection / expansion / backpropagation / advance / batching / root merge).  On CPU the tree is
driven with a test-only RolloutBackend whose rollouts come from the oracle; the -m gpu tests use the
real CudaRolloutBackend."""
import ctypes as C

import numpy as np
import pytest

import montecarlotreesearch_b200 as m
from tests.hostlib import Session, load
from tests.util import golden


@pytest.fixture(scope="module")
def host():
    return load()


@pytest.fixture(scope="module")
def corpus():
    return golden("corpus_q.npy")


def oracle_backend(port, seed=m.DEFAULT_SEED, log=None):
    """RolloutBackend callback: exact per-leaf sums computed by the oracle (test infrastructure)."""
    def cb(game, policy, states, n_leaves, R, first_id, out, user):
        for i in range(n_leaves):
            if game == m.GAME_QUORIDOR:
                s = np.frombuffer(C.string_at(states + 32 * i, 32), dtype=m.QSTATE_DTYPE)[0]
                o, _, _ = port.q_rollout_philox(s, seed, first_id + i * R, R, policy=policy)
            else:
                t = np.frombuffer(C.string_at(states + 4 * i, 4), dtype=m.TSTATE_DTYPE)[0]
                o = port.t_rollout_philox(int(t["x_mask"]), int(t["o_mask"]), seed, first_id + i * R, R)
            _, as_double = port.leaf_sum(o["score"])
            out[i] = as_double
        if log is not None:
            log.append((game, n_leaves, R, first_id))
    return cb


def test_actions_order_and_info_vs_reference_golden(host, corpus):
    g = golden("golden_q_primitives.npz")
    ids = np.zeros(256, dtype=np.int32)
    info = np.zeros(6, dtype=np.int32)
    ev = C.c_double(0)
    for j, i in enumerate(g["index"]):
        s = np.ascontiguousarray(corpus[i])
        n = host.mcts_host_q_actions(s.ctypes.data, ids.ctypes.data, 256)
        assert n == g["n_actions"][j] and (ids[:n] == g["actions"][j, :n]).all(), i
        best = host.mcts_host_q_info(s.ctypes.data, info.ctypes.data, C.addressof(ev))
        assert (info[0], info[1], info[2], info[3]) == (g["winner"][j], g["sp_white"][j], g["sp_black"][j], g["force_playwall"][j])
        assert ev.value == g["eval"][j]
        assert best == g["best_step"][j]


def test_good_moves_of_the_host_class_vs_reference_golden(host, corpus):
    """Quoridor_state::generate_good_moves (the selectable actions_to_try, Quoridor.cpp:518-592,618-627)"""
    g = golden("golden_q_goodmoves.npz")
    ids = np.zeros(256, dtype=np.int32)
    for j, i in enumerate(g["index"]):
        s = np.ascontiguousarray(m.quoridor_opening() if i < 0 else corpus[i])
        n = host.mcts_host_q_good_actions(s.ctypes.data, ids.ctypes.data, 256)
        assert n == g["count"][j] and (ids[:n] == g["ids"][j, :n]).all(), i


def test_tree_over_good_moves_only(host, port):
    """the tree expands generate_good_moves() when asked to: root children = that list, in that order"""
    g = golden("golden_q_goodmoves.npz")
    t = Session(host, m.GAME_QUORIDOR, m.quoridor_opening(), rollouts_per_leaf=1, callback=oracle_backend(port))
    t.set_good_moves_only(True)
    t.grow(12)
    r = t.root_stats()
    n = int(g["count"][0])
    assert len(r["children"]) == n and (r["children"][:, 0] == g["ids"][0, :n]).all()
    t.close()


def test_next_state_vs_oracle(host, port, corpus):
    rng = np.random.default_rng(21)
    out = np.zeros((), dtype=m.QSTATE_DTYPE)
    for i in rng.choice(len(corpus), 300, replace=False):
        s = np.ascontiguousarray(corpus[i])
        for mid in rng.integers(0, 209, size=10):
            ok = host.mcts_host_q_next(s.ctypes.data, int(mid), out.ctypes.data)
            e_ok, e = port.q_play(corpus[i], int(mid))
            assert bool(ok) == e_ok and out.tobytes() == e.tobytes()


def test_move_notation(host):
    s = np.ascontiguousarray(m.quoridor_opening())
    buf = C.create_string_buffer(128)
    host.mcts_host_q_sprint(s.ctypes.data, 13, buf, 128)
    assert buf.value == b"White moves to E2"                       # Quoridor.h:24-28 wording
    host.mcts_host_q_sprint(s.ctypes.data, 81 + 8 * 2 + 3, buf, 128)
    assert buf.value == b"White places horizontal wall at D3"


def test_tictactoe_tree_semantics(host, port):
    log = []
    t = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=4, callback=oracle_backend(port, log=log))
    t.grow(2000)
    r = t.root_stats()
    # reference bookkeeping: every iteration back-propagates 4 simulations and bumps `size` once per ancestor
    assert r["sims"] == 8000 and r["size"] == 2000 and len(r["children"]) == 9
    assert r["children"][:, 1].sum() == 8000                       # root sims = sum over children
    assert (r["children"][:, 0] == np.arange(9)).all()             # children in actions_to_try order
    assert len(log) == 2000 and all(e[1:3] == (1, 4) for e in log)
    ids = sorted(e[3] for e in log)
    assert all(b - a == 4 for a, b in zip(ids, ids[1:]))           # disjoint Philox id ranges
    p1 = r["score"] / r["sims"]
    assert 0.55 < p1 < 0.85
    best = t.best_move()
    assert best in (0, 2, 4, 6, 8)                                 # corner or centre
    # advance keeps the chosen subtree
    child = r["children"][r["children"][:, 0] == best][0]
    assert not t.advance(best)
    r2 = t.root_stats()
    assert r2["sims"] == child[1] and r2["score"] == child[2]
    # advancing by a move that was never expanded starts over
    t.grow(50)
    t.close()


def test_tictactoe_matches_stock_reference_tree_statistics(host, port, ref_native):
    """Config 1 plumbing: same iteration count and rollouts per leaf as the stock reference tree."""
    ref = ref_native.grow_tree(1, 4000)
    t = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=4, callback=oracle_backend(port))
    t.grow(4000)
    r = t.root_stats()
    assert r["sims"] == ref["root_sims"] and r["size"] == ref["size"]
    assert abs(r["score"] / r["sims"] - ref["root_score"] / ref["root_sims"]) < 0.06
    t.close()


def test_leaf_batching_with_virtual_loss(host, port):
    log = []
    t = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=8, leaves_per_flush=16,
                callback=oracle_backend(port, log=log))
    t.grow(960)
    r = t.root_stats()
    assert r["sims"] == 960 * 8 and r["size"] == 960
    assert len(log) == 60 and all(e[1] == 16 for e in log)          # 16 leaves per backend call
    assert t.counters() == (60, 960 * 8)
    assert 0.5 < r["score"] / r["sims"] < 0.9
    t.close()


@pytest.mark.parametrize("in_flight", [1, 2, 3])
def test_overlapped_flushes(host, port, in_flight):
    """Selection of the next batch overlaps the previous flushes: same bookkeeping with one, two or three
    flushes in flight (results are collected oldest first; every leaf keeps its virtual loss until then)."""
    log = []
    t = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=8, leaves_per_flush=16,
                callback=oracle_backend(port, log=log))
    t.set_overlap(in_flight)
    t.grow(1000)                       # 62 full flushes + one of 8 leaves
    r = t.root_stats()
    assert r["sims"] == 1000 * 8 and r["size"] == 1000
    assert len(log) == 63 and log[-1][1] == 8 and t.counters() == (63, 8000)
    assert r["children"][:, 1].sum() == 8000
    assert 0.5 < r["score"] / r["sims"] < 0.9
    t.close()


def test_quoridor_tree_small(host, port, corpus):
    t = Session(host, m.GAME_QUORIDOR, corpus[4200], rollouts_per_leaf=2, callback=oracle_backend(port))
    t.grow(40)
    r = t.root_stats()
    assert r["sims"] == 80 and r["size"] == 40
    g = golden("golden_q_primitives.npz")
    j = int(np.nonzero(g["index"] == 4200)[0][0])
    n = min(40, int(g["n_actions"][j]))
    assert (r["children"][:n, 0] == g["actions"][j, :n]).all()      # expansion follows generate_all_moves order
    mv = t.best_move()
    assert mv in g["actions"][j, : g["n_actions"][j]]
    t.close()


def test_root_parallel_merge(host, port):
    a = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=4, callback=oracle_backend(port, seed=1))
    b = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=4, callback=oracle_backend(port, seed=2))
    a.grow(300)
    b.grow(5)                      # b has expanded only 5 of the 9 root actions
    ea, eb = a.export_root(), b.export_root()
    assert len(ea) == 18 and len(eb) == 18 and (eb[10:] == 0).all()
    merged = ea + eb               # what the all-reduce produces on every rank
    a.import_root(merged)
    b.import_root(merged)
    ra, rb = a.root_stats(), b.root_stats()
    assert len(rb["children"]) == 9
    assert (ra["children"] == rb["children"]).all() and ra["sims"] == rb["sims"] == merged[0::2].sum()
    assert a.best_move() == b.best_move()
    a.close(); b.close()


def _tree_signature(t):
    r = t.root_stats()
    return r["sims"], r["score"], r["size"], r["children"].tobytes(), t.best_move()


@pytest.mark.parametrize("K,in_flight", [(1, 0), (8, 0), (8, 2)])
def test_expansion_from_device_masks_builds_the_same_tree(host, port, corpus, K, in_flight):
    """MCTS_batching::gpu_expand: a leaf's move list is read off the legal-move mask that came back with its flush
    (here the masks come from the shared rule code compiled for the host) — same list, same order, so the tree is
    the same tree, statistic for statistic, as with actions_to_try() on the CPU (mcts.cpp:19)."""
    sigs = []
    for expand in (False, True):
        base = host.mcts_host_peek_rollout_id()
        inner = oracle_backend(port)

        def cb(game, policy, states, n_leaves, R, first_id, out, user, base=base, inner=inner):
            inner(game, policy, states, n_leaves, R, first_id - base, out, user)
        t = Session(host, m.GAME_QUORIDOR, corpus[4200], rollouts_per_leaf=1, leaves_per_flush=K, callback=cb)
        t.set_overlap(in_flight)
        t.set_gpu_expand(expand)
        t.grow(260)                    # more iterations than root actions: second-level nodes get expanded too
        sigs.append(_tree_signature(t))
        mv = t.best_move()
        t.advance(mv)
        t.grow(60)                     # the reused subtree keeps working (children expanded from kept masks)
        sigs[-1] += _tree_signature(t)
        t.close()
    assert sigs[0] == sigs[1]


@pytest.mark.gpu
@pytest.mark.parametrize("K,R", [(16, 64), (256, 16)])
def test_gpu_expanded_tree_equals_host_expanded_tree(host, K, R):
    """The same on the B200: masks from quoridor_movegen_kernel riding on every flush (mctsb_submit_x with
    MCTSB_SUBMIT_LEGAL) against move lists generated on the CPU; rollout ids aligned, so every statistic agrees."""
    sigs = []
    for expand in (False, True):
        base = host.mcts_host_peek_rollout_id()
        t = Session(host, m.GAME_QUORIDOR, m.quoridor_opening(), rollouts_per_leaf=R, leaves_per_flush=K, seed=77)
        host.mcts_host_set_next_rollout_id(0)
        t.set_gpu_expand(expand)
        t.grow(4 * K + 300)
        sigs.append(_tree_signature(t))
        t.advance(t.best_move())
        t.grow(2 * K)
        sigs[-1] += _tree_signature(t)
        t.close()
        host.mcts_host_set_next_rollout_id(base + 10 ** 9)
    assert sigs[0] == sigs[1]


@pytest.mark.gpu
def test_config1_tictactoe_on_gpu_backend(host):
    """BASELINE.json configs[0]: Tic-Tac-Toe from the empty board, 10k iterations, rollouts on the B200."""
    t = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=4)
    t.grow(10000)
    r = t.root_stats()
    assert r["sims"] == 40000 and r["size"] == 10000
    assert 0.58 < r["score"] / r["sims"] < 0.76                     # stock reference: ~0.657-0.668
    assert t.best_move() in (0, 2, 4, 6, 8)
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("in_flight", [0, 1, 2])
def test_config4_quoridor_selfplay_flush_on_gpu(host, in_flight):
    """BASELINE.json configs[3] shape: leaf-parallel flushes of 64k rollouts (K=16 leaves x R=4096), synchronous
    or with one / two flushes queued on the GPU (two C-ABI handles = two streams)."""
    t = Session(host, m.GAME_QUORIDOR, m.quoridor_opening(), rollouts_per_leaf=4096, leaves_per_flush=16)
    t.set_overlap(in_flight)
    t.grow(160)
    r = t.root_stats()
    assert r["sims"] == 160 * 4096 and t.counters() == (10, 160 * 4096)
    assert 0.4 < r["score"] / r["sims"] < 0.9
    mv = t.best_move()
    assert 0 <= mv < 209
    t.advance(mv)
    t.close()


@pytest.mark.gpu
def test_cpp_root_parallel_selfplay_program(tmp_path):
    """integration/selfplay_root_parallel.cpp: configs[3]/[4] as a plain C++ program over libmcts_host.so — one tree
    and one host thread per GPU of the box, NCCL merge of the root statistics behind the C ABI, no Python in the loop."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "montecarlotreesearch_b200")
    exe = str(tmp_path / "selfplay_cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(root, "integration", "selfplay_root_parallel.cpp"), "-L", pkg, "-lmcts_host", "-lmctsb",
                           "-Wl,-rpath," + pkg, "-o", exe])
    r = subprocess.run([exe, "0", "96", "16", "256", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == m.device_count() and line["moves"] == 8 and line["same_move_on_every_gpu"]
    assert line["rollouts_total"] == line["n_gpus"] * 8 * 96 * 256
