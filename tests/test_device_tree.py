"""The device-resident tree (csrc/tree_body.cuh, C ABI mctsb_tree_*): Selection / Expansion / Backpropagation on the
GPU must build the SAME tree as the host tree (host/mcts.cpp) under the same batching — same nodes in the same
order, same simulation counts, same score doubles.  On CPU the warp-level code runs on emulated 32-lane warps
(tests/hostemu) and both trees are fed the same playout results; the -m gpu tests run the kernels."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import montecarlotreesearch_b200 as m
from oracle.reflib import LEAFSUM_DTYPE
from tests.hostlib import Session, load


@pytest.fixture(scope="module")
def host():
    return load()


@pytest.fixture(scope="module")
def emu32():
    from tests.hostemu.emulib import EmuLib
    return EmuLib(warp32=True)


def fake_sum(state_bytes, first_id, R):
    """a made-up exact playout sum of one leaf, a function of (position, Philox ids): {lo, hi} limbs of sum(score * 2^57)"""
    h = hashlib.blake2b(state_bytes + int(first_id).to_bytes(8, "little"), digest_size=16).digest()
    a, b = int.from_bytes(h[:8], "little"), int.from_bytes(h[8:], "little")
    return a % (R << 29), b % ((R << 28) + 1)


def drive_both(host, emu32, root, R, K, flushes, port=None, advance_every=0, chunk=64, good_moves=False):
    """grow the host tree and the emulated device tree side by side; compare after every flush"""
    from tests.hostemu.emulib import EmuTree
    to_double = emu32.lib.emu_leaf_sum_to_double
    log = []

    def cb(game, policy, states, n_leaves, Rr, first_id, out, user):
        for i in range(n_leaves):
            sb = C.string_at(states + 32 * i, 32)
            if port is None:
                lo, hi = fake_sum(sb, first_id + i * Rr, Rr)
            else:
                s = np.frombuffer(sb, dtype=m.QSTATE_DTYPE)[0]
                o, _, _ = port.q_rollout_philox(s, m.DEFAULT_SEED, first_id + i * Rr, Rr)
                acc, _ = port.leaf_sum(o["score"])
                lo, hi = int(acc["lo"]), int(acc["hi"])
            out[i] = to_double(lo, hi)
        log.append((n_leaves, first_id))

    base_id = 1 << 40
    host.mcts_host_set_next_rollout_id(base_id)
    ht = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=R, leaves_per_flush=K, callback=cb)
    dt = EmuTree(emu32, root, R, K, chunk=max(chunk, (K + 63) // 64 // 32 * 32 + 32))
    if good_moves:
        ht.set_good_moves_only(True); dt.set_good_moves_only(True)
    done = 0
    for f in range(flushes):
        n = K if not isinstance(flushes, list) else K
        ht.grow(n)
        leaves, err = dt.select(n)
        assert err == 0, err
        sums = np.zeros(n, dtype=LEAFSUM_DTYPE)
        for d in range(n):
            sb = leaves[d].tobytes()
            if port is None:
                lo, hi = fake_sum(sb, base_id + (done + d) * R, R)
            else:
                o, _, _ = port.q_rollout_philox(leaves[d], m.DEFAULT_SEED, base_id + (done + d) * R, R)
                acc, _ = port.leaf_sum(o["score"])
                lo, hi = int(acc["lo"]), int(acc["hi"])
            sums[d] = (lo, hi, R)
        assert dt.backprop(sums) == 0
        done += n
        a, b = ht.dump(), dt.dump()
        assert len(a["depth"]) == len(b["depth"]), (f, len(a["depth"]), len(b["depth"]))
        for k in ("depth", "move", "n_children", "sims", "size"):
            assert (a[k] == b[k]).all(), (f, k, np.flatnonzero(a[k] != b[k])[:8])
        assert a["score"].tobytes() == b["score"].tobytes(), (f, "score", np.flatnonzero(a["score"] != b["score"])[:8])
        assert (b["virt"] == 0).all()
        if advance_every and (f + 1) % advance_every == 0:
            mv = ht.best_move()
            ht.advance(mv)
            assert dt.advance(mv)
    st = dt.stats()
    ht.close(); dt.close()
    return st, log


def test_emulated_device_tree_with_one_flush_in_flight(host, emu32):
    """overlap mode: S(1) S(2) B(1) S(3) B(2) ... B(n) — the host tree's order with MCTS_batching::overlap; every flush is
    selected while the one before still carries its virtual losses"""
    from tests.hostemu.emulib import EmuTree
    to_double = emu32.lib.emu_leaf_sum_to_double
    for root, R, K, flushes in ((m.quoridor_opening(), 4, 64, 9), (m.quoridor_opening(), 1, 512, 5)):
        def cb(game, policy, states, n_leaves, Rr, first_id, out, user):
            for i in range(n_leaves):
                lo, hi = fake_sum(C.string_at(states + 32 * i, 32), first_id + i * Rr, Rr)
                out[i] = to_double(lo, hi)
        base_id = 1 << 40
        host.mcts_host_set_next_rollout_id(base_id)
        ht = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=R, leaves_per_flush=K, callback=cb)
        ht.set_overlap(1)
        tail = K // 2 + 1
        ht.grow(K * (flushes - 1) + tail)
        dt = EmuTree(emu32, root, R, K, chunk=32)
        pending = None
        for f in range(flushes):
            n = K if f < flushes - 1 else tail
            leaves, err = dt.select(n, slot=f & 1)
            assert err == 0
            sums = np.zeros(n, dtype=LEAFSUM_DTYPE)
            for d in range(n):
                lo, hi = fake_sum(leaves[d].tobytes(), base_id + (f * K + d) * R, R)
                sums[d] = (lo, hi, R)
            if pending is not None:
                assert dt.backprop(pending[0], slot=pending[1]) == 0
            pending = (sums, f & 1)
        assert dt.backprop(pending[0], slot=pending[1]) == 0
        a, b = ht.dump(), dt.dump()
        assert len(a["depth"]) == len(b["depth"])
        for k in ("depth", "move", "n_children", "sims", "size"):
            assert (a[k] == b[k]).all(), (k, np.flatnonzero(a[k] != b[k])[:8])
        assert a["score"].tobytes() == b["score"].tobytes()
        assert (b["virt"] == 0).all()
        ht.close(); dt.close()


def test_leaf_sum_conversion_matches_the_abi_helper(emu32, port):
    """tr_leaf_sum_to_double (device code) = mctsb_leaf_sum_to_double's correctly rounded exact sum"""
    rng = np.random.default_rng(5)
    to_double = emu32.lib.emu_leaf_sum_to_double
    for R in (1, 4, 16, 127, 128, 4096, 1 << 20):
        for _ in range(200):
            lo, hi = int(rng.integers(0, R << 29)), int(rng.integers(0, (R << 28) + 1))
            exact = (hi << 29) + lo
            # correctly rounded exact / 2^57 with Python integers
            from fractions import Fraction
            assert to_double(lo, hi) == float(Fraction(exact, 1 << 57)), (R, lo, hi)
    assert to_double(0, 0) == 0.0


@pytest.mark.parametrize("R,K,flushes,chunk", [(4, 1, 40, 32), (4, 16, 12, 32), (16, 256, 6, 64), (1, 1024, 4, 32), (1, 1024, 3, 512)])
def test_emulated_device_tree_equals_host_tree(host, emu32, R, K, flushes, chunk):
    """chunk = descents per chunk of the wavefront (the flush is selected in chunks staggered by one level)"""
    st, _ = drive_both(host, emu32, m.quoridor_opening(), R, K, flushes, chunk=chunk)
    assert st["error"] == 0


def test_emulated_device_tree_midgame_with_advance(host, emu32):
    """a mid-game root (few walls left: deeper, narrower trees, terminal nodes inside a flush) and tree reuse across moves"""
    from tests.util import golden
    corpus = golden("corpus_q.npy")
    for i in (7, 1234, 4000):
        st, _ = drive_both(host, emu32, corpus[i], 2, 128, 8, advance_every=2)
        assert st["error"] == 0


def test_emulated_device_tree_with_oracle_playouts(host, emu32, port):
    """the same with real playouts (oracle): exact sums through both conversions"""
    st, _ = drive_both(host, emu32, m.quoridor_opening(), 2, 32, 4, port=port)
    assert st["error"] == 0


# ---- the kernels, on the GPU ---------------------------------------------------------------------------------------
def compare_gpu_overlapped(host, root, R, K, flushes, policy=0):
    """one flush in flight on both sides: host tree with MCTS_batching::overlap vs mctsb_tree_set_overlap"""
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    base_id = 1 << 43
    iters = K * (flushes - 1) + K // 2 + 1
    host.mcts_host_set_next_rollout_id(base_id)
    ht = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=R, leaves_per_flush=K, policy=policy)
    ht.set_overlap(1)
    dt = m.DeviceTree(be, root, R, K, policy=policy)
    dt.set_overlap(True)
    for rep in range(2):                     # two grows: the second starts from a tree whose pipeline has drained
        ht.grow(iters)
        dt.grow(iters, base_id + rep * iters * R)
        a, b = ht.dump(), dt.dump()
        assert len(a["depth"]) == len(b["depth"]), (len(a["depth"]), len(b["depth"]))
        for k in ("depth", "move", "n_children", "sims", "size"):
            assert (a[k] == b[k]).all(), (k, np.flatnonzero(a[k] != b[k])[:8])
        assert a["score"].tobytes() == b["score"].tobytes()
        assert ht.best_move() == dt.best_move()
    info = dt.info()
    ht.close(); dt.close(); be.close()
    return info


@pytest.mark.gpu
@pytest.mark.parametrize("partition", [True, False])
@pytest.mark.parametrize("R,K,flushes", [(4, 16, 12), (16, 1024, 6), (16, 4096, 5), (256, 256, 5)])
def test_device_tree_with_one_flush_in_flight_equals_host_tree_on_gpu(host, R, K, flushes, partition, monkeypatch):
    """partition: the tree's kernels on SMs of their own (green contexts) / both kinds of kernels on all SMs, two plain streams"""
    if not partition:
        monkeypatch.setenv("MCTSB_TREE_NO_PARTITION", "1")
    info = compare_gpu_overlapped(host, m.quoridor_opening(), R, K, flushes)
    assert info["flags"] == 0


def compare_gpu(host, root, R, K, flushes, policy=0, advance_every=0, tail=0, max_nodes=1 << 23, good_moves=False):
    """device tree vs host tree (both playing their rollouts on the GPU, same Philox ids): identical after every grow"""
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    base_id = 1 << 41
    host.mcts_host_set_next_rollout_id(base_id)
    ht = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=R, leaves_per_flush=K, policy=policy)
    dt = m.DeviceTree(be, root, R, K, policy=policy, max_nodes=max_nodes)
    if good_moves:
        ht.set_good_moves_only(True); dt.set_good_moves_only(True)
    done = 0
    for f in range(flushes):
        n = K if not (tail and f == flushes - 1) else tail
        ht.grow(n)
        dt.grow(n, base_id + done * R)
        done += n
        a, b = ht.dump(), dt.dump()
        assert len(a["depth"]) == len(b["depth"]), (f, len(a["depth"]), len(b["depth"]))
        for k in ("depth", "move", "n_children", "sims", "size"):
            assert (a[k] == b[k]).all(), (f, k, np.flatnonzero(a[k] != b[k])[:8])
        assert a["score"].tobytes() == b["score"].tobytes(), (f, "score", np.flatnonzero(a["score"] != b["score"])[:8])
        assert ht.best_move() == dt.best_move()
        if advance_every and (f + 1) % advance_every == 0:
            mv = ht.best_move()
            ht.advance(mv); dt.advance(mv)
            assert ht.current_state().tobytes() == dt.root_state().tobytes()
    info = dt.info()
    r = ht.root_stats()
    assert info["root_sims"] == r["sims"] and info["root_score"] == r["score"] and info["root_size"] == r["size"]
    ht.close(); dt.close(); be.close()
    return info


@pytest.mark.gpu
@pytest.mark.parametrize("R,K,flushes", [(4, 1, 30), (4, 16, 20), (16, 256, 10), (16, 4096, 6), (1, 8192, 4), (1, 40000, 3)])
def test_device_tree_equals_host_tree_on_gpu(host, R, K, flushes):
    info = compare_gpu(host, m.quoridor_opening(), R, K, flushes, tail=max(1, K // 3))
    assert info["flags"] == 0 and info["flushes"] == flushes


@pytest.mark.gpu
def test_device_tree_midgame_advance_and_uni_policy_on_gpu(host):
    from tests.util import golden
    corpus = golden("corpus_q.npy")
    deepest = 0
    for i in (7, 1234, 4000):
        info = compare_gpu(host, corpus[i], 8, 512, 8, advance_every=2)
        assert info["flags"] == 0
        deepest = max(deepest, info["deepest_level"])
    assert deepest >= 2
    info = compare_gpu(host, m.quoridor_opening(), 4, 256, 4, policy=m.POLICY_UNI)
    assert info["flags"] == 0


@pytest.mark.gpu
def test_device_tree_full_pool_and_bad_arguments():
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    with pytest.raises(m.BackendError) as e:
        t = m.DeviceTree(be, m.quoridor_opening(), 4, 64, max_nodes=600)      # root + 131 children, then 4 x 131 do not fit
        t.grow(64 * 40, 0)
    assert e.value.status == 9
    bad = m.quoridor_opening().copy(); bad["wx"] = 9
    with pytest.raises(m.BackendError) as e:
        m.DeviceTree(be, bad, 4, 64)
    assert e.value.status == 4
    t = m.DeviceTree(be, m.quoridor_opening(), 4, 64)
    with pytest.raises(m.BackendError):
        t.grow(10, 0, leaves_per_flush=65)                                     # more than the tree was created for
    with pytest.raises(m.BackendError):
        t.advance(0)                                                           # not a legal move of the opening
    t.advance(13)                                                              # never expanded: fresh root (mcts.cpp:149-152)
    assert t.info()["root_sims"] == 0 and int(t.root_state()["turn"]) == 1
    t.close(); be.close()


@pytest.mark.gpu
def test_cpp_device_tree_class_plays_the_host_trees_game(host):
    """MCTS_device_tree (host/device_tree.h: MCTS_tree's interface over the C ABI) against MCTS_tree: same moves, same sizes"""
    root = np.ascontiguousarray(m.quoridor_opening())
    err = C.create_string_buffer(512)
    games = []
    for device_tree in (False, True):
        host.mcts_host_set_next_rollout_id(1 << 44)
        if device_tree:
            h = host.mcts_host_dtree_create(root.ctypes.data, 0, 0, m.DEFAULT_SEED, 8, 512, 0, 1 << 22, err, 512)
            assert h, err.value
            grow = lambda n: host.mcts_host_dtree_grow(h, n, 1e9, err, 512)
            best, advance, size = (lambda: host.mcts_host_dtree_best_move(h)), (lambda mv: host.mcts_host_dtree_advance(h, mv)), (lambda: host.mcts_host_dtree_size(h))
        else:
            t = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=8, leaves_per_flush=512)
            grow = lambda n: (t.grow(n), 0)[1]
            best, advance, size = t.best_move, (lambda mv: int(t.advance(mv))), (lambda: int(t.root_stats()["size"]))
        moves = []
        for ply in range(6):
            assert grow(2048) == 0, err.value
            moves.append((best(), size()))
            assert advance(moves[-1][0]) == 0
        games.append(moves)
        if device_tree:
            host.mcts_host_dtree_destroy(h)
        else:
            t.close()
    assert games[0] == games[1], games


@pytest.mark.gpu
def test_cpp_selfplay_program_device_tree_vs_host_tree(tmp_path):
    """integration/selfplay_device_tree.cpp: the reference's self-play loop (examples/Quoridor/main.cpp:129-169) as a plain C++
    program over MCTS_device_tree and, with `host`, over MCTS_tree — no Python in the loop; the two play the same game"""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "montecarlotreesearch_b200")
    exe = str(tmp_path / "selfplay_device_tree")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(root, "integration", "selfplay_device_tree.cpp"), "-L", pkg, "-lmcts_host", "-lmctsb",
                           "-Wl,-rpath," + pkg, "-o", exe])
    lines = []
    for where in ("device", "host"):
        r = subprocess.run([exe, "2048", "512", "8", "10", where], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        lines.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert lines[0]["moves"] == 10 and lines[0]["move_ids"] == lines[1]["move_ids"], lines
    assert lines[0]["rollouts_total"] == lines[1]["rollouts_total"] == 10 * 2048 * 8


@pytest.mark.gpu
def test_root_merge_of_device_trees_equals_the_host_trees_merge(host):
    """mctsb_tree_merge_root is MCTS_tree::merge_root_stats for device trees: a no-op for a lone backend; over every GPU of
    the box (different Philox keys) all ranks end with the same root-child statistics — the sums — and the same best move"""
    root = m.quoridor_opening()
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    host.mcts_host_set_next_rollout_id(1 << 46)
    ht = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=8, leaves_per_flush=128)
    dt = m.DeviceTree(be, root, 8, 128)
    ht.grow(512); dt.grow(512, 1 << 46)
    ht.merge_root(); dt.merge_root()                            # lone backends: both leave the tree as it is
    a, b = ht.dump(), dt.dump()
    for k in ("depth", "move", "n_children", "sims", "size"):
        assert (a[k] == b[k]).all(), k
    assert a["score"].tobytes() == b["score"].tobytes() and ht.best_move() == dt.best_move()
    ht.close(); dt.close(); be.close()
    n = m.device_count()
    if n >= 2:
        bes = [m.RolloutBackend(g, m.DEFAULT_SEED + g) for g in range(n)]
        m.comm_init_all(bes)
        trees = [m.DeviceTree(b_, root, 8, 128) for b_ in bes]
        trees[0].grow(64, 0)
        with pytest.raises(m.BackendError):
            trees[0].merge_root()                               # root not fully expanded yet: its slots are not all there
        trees[0].grow(448, 64 * 8)
        for t in trees[1:]:
            t.grow(512, 0)
        import threading
        th = [threading.Thread(target=t.merge_root) for t in trees]     # one blocking all-reduce per rank: issue them together
        [x.start() for x in th]; [x.join() for x in th]
        kids = [t.root_children() for t in trees]
        for k in kids[1:]:
            assert all((x == y).all() for x, y in zip(k, kids[0]))
        assert len({t.best_move() for t in trees}) == 1
        assert int(kids[0][1].sum()) == n * 512 * 8
        [t.close() for t in trees]; [b_.close() for b_ in bes]


@pytest.mark.gpu
def test_advance_tree_reclaims_the_dropped_subtrees(host, monkeypatch):
    """advance_tree on the device copies the kept subtree into the spare pool (the dropped ones are gone, mcts.cpp:132-157):
    the tree stays the host tree's after every collection, and a pool far too small for the whole game is enough"""
    monkeypatch.setenv("MCTSB_TREE_GC", "always")
    info = compare_gpu(host, m.quoridor_opening(), 4, 256, 12, advance_every=1)
    assert info["flags"] == 0 and info["nodes_used"] < 60000         # one move's subtree, not twelve moves' worth of ids
    monkeypatch.delenv("MCTSB_TREE_GC")
    # default policy (collect once a quarter of the pool is used): 30 moves x 1 024 iterations need ~1 M ids, the pool has 150 000
    info = compare_gpu(host, m.quoridor_opening(), 4, 256, 120, advance_every=4, max_nodes=150000)
    assert info["flags"] == 0 and info["nodes_used"] <= 150000
    monkeypatch.setenv("MCTSB_TREE_GC", "never")
    with pytest.raises(m.BackendError) as e:
        compare_gpu(host, m.quoridor_opening(), 4, 256, 120, advance_every=4, max_nodes=150000)
    assert e.value.status == 9


def test_emulated_device_tree_over_good_moves(host, emu32):
    """generate_good_moves (Quoridor.cpp:518-592) as the tree's actions_to_try: Quoridor_state::set_good_moves_only on the
    host, mctsb_tree_set_good_moves on the device — the ordered list is rebuilt by the warp whenever a node is expanded"""
    from tests.util import golden
    corpus = golden("corpus_q.npy")
    for root in (m.quoridor_opening(), corpus[7], corpus[1234]):
        st, _ = drive_both(host, emu32, root, 2, 64, 6, advance_every=3, chunk=32, good_moves=True)
        assert st["error"] == 0


@pytest.mark.gpu
def test_device_tree_over_good_moves_on_gpu(host):
    from tests.util import golden
    corpus = golden("corpus_q.npy")
    for root in (m.quoridor_opening(), corpus[7], corpus[4000]):
        info = compare_gpu(host, root, 8, 512, 8, advance_every=2, good_moves=True)
        assert info["flags"] == 0
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    t = m.DeviceTree(be, m.quoridor_opening(), 4, 64)
    t.grow(64, 0)
    with pytest.raises(m.BackendError):
        t.set_good_moves_only(True)                              # not on a grown tree
    t.close(); be.close()
