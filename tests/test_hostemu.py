"""CPU: the DEVICE rule/rollout core (csrc/qcore.cuh, ttt_core.cuh) compiled for the host by
tests/hostemu (test-only, never shipped) against the oracle and the reference goldens.  Lets the
kernel logic be checked bit-for-bit on a box without a GPU; the -m gpu tests repeat this through
the C ABI on the real kernels."""
import numpy as np
import pytest

from oracle.reflib import opening_state
from tests.util import golden, philox_stream


@pytest.fixture(scope="module")
def emu():
    from tests.hostemu.emulib import EmuLib
    return EmuLib()


@pytest.fixture(scope="module")
def corpus():
    return golden("corpus_q.npy")


def test_core_primitives_vs_golden(emu, corpus):
    g = golden("golden_q_primitives.npz")
    for j, i in enumerate(g["index"]):
        p = emu.q_primitives(corpus[i])
        for k in ("winner", "sp_white", "sp_black", "force_playwall"):
            assert p[k] == g[k][j], (k, i)
        assert p["eval"] == g["eval"][j]
        assert (np.packbits(p["legal"], bitorder="little") == g["legal"][j]).all()
        assert (np.packbits(p["cheap"], bitorder="little") == g["cheap"][j]).all()
        s2 = g["steps2"][j]
        assert (p["steps2"] == s2[s2 >= 0]).all()
        if g["winner"][j] == 0:
            assert emu.q_best_step(corpus[i]) == g["best_step"][j]


def test_core_primitives_vs_oracle_on_corpus(emu, port, corpus):
    for i in range(300, 4608, 9):
        a, b = port.q_primitives(corpus[i]), emu.q_primitives(corpus[i])
        for k in ("winner", "sp_white", "sp_black", "force_playwall", "eval"):
            assert a[k] == b[k]
        assert (a["legal"] == b["legal"]).all() and (a["cheap"] == b["cheap"]).all()


def test_pruning_masks_are_exact_on_corpus(emu, corpus):
    """every cheap-legal wall outside the shortest-path pruning mask leaves the path length unchanged
    (brute force: one flood fill per wall, both players, the whole mid-game corpus + rollout finals)"""
    g = golden("golden_q_replay.npz")
    states = np.concatenate([corpus, g["final"].astype(corpus.dtype)])
    for player in (0, 1):
        bad, n, kept = emu.q_cut_check(states, player)
        assert bad == 0
        assert n > 100000 and kept < 0.45 * n        # the masks prune: most walls need no search


def test_detour_rule_never_passes_an_illegal_wall(emu, corpus):
    """q_detour_safe_masks (legal without a search: at most one of the wall's three lattice points touches anything):
    on the whole corpus plus dense late-game positions every wall it settles keeps both paths, the one-wall form agrees
    with the mask form, and it settles most candidates."""
    from tests.util import serpentine_state
    bad, cheap, settled, illegal = emu.q_detour_check(corpus)
    assert bad == 0 and illegal > 0
    assert settled > 0.6 * cheap, (settled, cheap)
    bad, cheap, settled, illegal = emu.q_detour_check(np.array([serpentine_state(turn=0), serpentine_state(rows=(0, 2, 4, 7), turn=1)]))
    assert bad == 0 and illegal > 0 and settled < cheap


def test_core_replay_vs_golden(emu, corpus):
    g = golden("golden_q_replay.npz")
    seed = int(g["seed"])
    streams = np.stack([philox_stream(seed, int(i), 16384) for i in g["id"]])
    out, fin, _ = emu.q_rollout_streams(corpus[g["index"]], streams)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    assert fin.tobytes() == g["final"].tobytes()


def test_core_philox_vs_golden(emu):
    g = golden("golden_q_opening.npz")
    out, _, _ = emu.q_rollout_philox(opening_state(), int(g["seed"]), 0, 512)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()


def test_core_uniform_vs_golden(emu, corpus):
    g = golden("golden_q_uniform.npz")
    for j, i in enumerate(g["index"]):
        out, _, _ = emu.q_rollout_philox(corpus[i], int(g["seed"]), int(g["id"][j]), 1, policy=1)
        assert out["score"][0] == g["score"][j] and out["draws"][0] == g["draws"][j]
        assert out["kind"][0] == g["kind"][j] and out["plies"][0] == g["plies"][j]


def test_core_play_vs_oracle(emu, port, corpus):
    rng = np.random.default_rng(3)
    for i in rng.choice(4608, 200, replace=False):
        for mid in rng.integers(0, 209, size=12):
            a_ok, a = port.q_play(corpus[i], int(mid))
            b_ok, b = emu.q_play(corpus[i], int(mid))
            assert a_ok == b_ok and a.tobytes() == b.tobytes()


def test_core_tictactoe_vs_golden(emu):
    g = golden("golden_t.npz")
    out = emu.t_rollout_philox(0, 0, int(g["seed"]), 0, 20000)
    assert (out["score"] == g["empty_scores"]).all() and (out["draws"] == g["empty_draws"]).all()
    for j in range(0, len(g["states"]), 3):
        x, o = int(g["states"][j][0]), int(g["states"][j][1])
        out = emu.t_rollout_philox(x, o, int(g["seed"]), 7000 + j, 1)
        assert out["score"][0] == g["score"][j] and out["draws"][0] == g["draws"][j]


# ---- the production lockstep kernel body, emulated with a one-lane warp --------------------------------
def test_lockstep_body_vs_reference_golden(emu, corpus):
    g = golden("golden_q_opening.npz")
    out, _, sums, _ = emu.lockstep(opening_state(), 512, int(g["seed"]), 0)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    assert int(sums["n"][0]) == 512
    g = golden("golden_q_replay.npz")
    streams = np.stack([philox_stream(int(g["seed"]), int(i), 16384) for i in g["id"]])
    out, fin, _, _ = emu.lockstep(corpus[g["index"]], 1, 0, 0, streams=streams)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    assert fin.tobytes() == g["final"].tobytes()


def test_lockstep_body_vs_oracle_many_leaves(emu, port, corpus):
    rng = np.random.default_rng(17)
    idx = rng.choice(len(corpus), 120, replace=False)
    out, fin, sums, _ = emu.lockstep(corpus[idx], 5, 4242, 900)
    for j, i in enumerate(idx):
        exp, efin, _ = port.q_rollout_philox(corpus[i], 4242, 900 + 5 * j, 5, finals=True)
        assert exp.tobytes() == out[5 * j: 5 * j + 5].tobytes() and efin.tobytes() == fin[5 * j: 5 * j + 5].tobytes()
        acc, _ = port.leaf_sum(exp["score"])
        assert (int(acc["lo"]), int(acc["hi"]), int(acc["n"])) == (int(sums["lo"][j]), int(sums["hi"][j]), int(sums["n"][j]))


def test_long_paths_overflow_the_layer_budget(emu, port):
    """Shortest paths longer than MCTSB_MAX_LAYERS: the kernel falls back to measuring every wall."""
    from tests.util import serpentine_state
    for turn in (0, 1):
        s = serpentine_state(turn=turn)
        p = port.q_primitives(s)
        assert p["sp_white"] > 24 and p["sp_black"] > 24
        exp, efin, _ = port.q_rollout_philox(s, 5, 0, 64, finals=True)
        out, fin, _, _ = emu.lockstep(s, 64, 5, 0)
        assert exp.tobytes() == out.tobytes() and efin.tobytes() == fin.tobytes()
        e = emu.q_primitives(s)
        assert (e["legal"] == p["legal"]).all() and e["sp_white"] == p["sp_white"] and e["sp_black"] == p["sp_black"]


# ---- the same body on emulated 32-lane warps (tests/hostemu/simt.h): votes, shuffles, the shared task queue,
# ---- the atomicMax fold and the global work counter are real here ----------------------------------------
@pytest.fixture(scope="module")
def emu32():
    from tests.hostemu.emulib import EmuLib
    e = EmuLib(warp32=True)
    assert e.lib.emu_warp_width() == 32
    return e


def test_lockstep_warp32_vs_reference_golden(emu32, corpus):
    g = golden("golden_q_opening.npz")
    out, _, sums, _ = emu32.lockstep(opening_state(), 512, int(g["seed"]), 0, blocks=2)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    assert int(sums["n"][0]) == 512
    g = golden("golden_q_replay.npz")
    streams = np.stack([philox_stream(int(g["seed"]), int(i), 16384) for i in g["id"]])
    out, fin, _, _ = emu32.lockstep(corpus[g["index"]], 1, 0, 0, streams=streams, blocks=1)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    assert fin.tobytes() == g["final"].tobytes()


def test_lockstep_warp32_ragged_mixed_leaves(emu32, port, corpus):
    """Leaves change inside a warp (7 rollouts per leaf, 33+ distinct leaves per warp boundary), mid-game positions
    next to terminal and wall-less ones, three blocks racing for the work counter: every playout and every
    exact leaf sum equals the oracle's."""
    rng = np.random.default_rng(23)
    idx = rng.choice(len(corpus), 150, replace=False)
    st = corpus[idx].copy()
    st[5]["wx"] = 8                      # a terminal leaf in the middle of a warp
    st[40]["wwalls"] = 0; st[40]["bwalls"] = 0
    out, fin, sums, _ = emu32.lockstep(st, 7, 777, 31, blocks=3)
    for j in range(len(st)):
        exp, efin, _ = port.q_rollout_philox(st[j], 777, 31 + 7 * j, 7, finals=True)
        assert exp.tobytes() == out[7 * j: 7 * j + 7].tobytes() and efin.tobytes() == fin[7 * j: 7 * j + 7].tobytes()
        acc, _ = port.leaf_sum(exp["score"])
        assert (int(acc["lo"]), int(acc["hi"]), int(acc["n"])) == (int(sums["lo"][j]), int(sums["hi"][j]), int(sums["n"][j]))


def test_lockstep_warp32_long_paths(emu32, port):
    from tests.util import serpentine_state
    s = serpentine_state(turn=0)
    exp, efin, _ = port.q_rollout_philox(s, 5, 0, 96, finals=True)
    out, fin, _, _ = emu32.lockstep(s, 96, 5, 0, blocks=1)
    assert exp.tobytes() == out.tobytes() and efin.tobytes() == fin.tobytes()


def test_lockstep_warp32_forced_rare_branches(emu32, port, corpus):
    """The branches a free-running playout only meets by chance, forced by replay streams (tests/util.py
    forced_branch_cases): first-legal walls whose shuffled head is ILLEGAL (the rest of the pool follows in key
    order, Quoridor.cpp:773-788) and best-wall / guided searches with more candidates than one batch of the task
    queue takes per lane.  108 different leaves, one rollout each, so every warp mixes leaves and branches."""
    from tests.util import forced_branch_cases
    states, streams, notes = forced_branch_cases(port, corpus)
    assert sum(1 for n in notes if n[0] == "head_miss") >= 33 and any(n[0] == "many_candidates" and n[1] > 64 for n in notes)
    exp, efin = port.q_rollout_streams(states, streams)
    out, fin, _, _ = emu32.lockstep(states, 1, 0, 0, streams=streams, blocks=2)
    assert out.tobytes() == exp.tobytes() and fin.tobytes() == efin.tobytes()


# ---- the UNI policy on the lockstep machinery (quoridor_uni_lockstep_body.cuh + the walk of phase B) ----------
@pytest.mark.parametrize("warp32", [False, True])
def test_uni_lockstep_vs_reference_golden_and_oracle(warp32, port, corpus):
    from tests.hostemu.emulib import EmuLib
    e = EmuLib(warp32=warp32)
    g = golden("golden_q_uniform.npz")
    streams = np.stack([philox_stream(int(g["seed"]), int(i), 8192) for i in g["id"]])
    out, _, _, _, _ = e.uni_lockstep(corpus[g["index"]], 1, 0, 0, streams=streams, blocks=2)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    # ragged mixed leaves: 5 rollouts per leaf, leaves change inside a warp; terminal, wall-less and one-sided leaves
    rng = np.random.default_rng(5)
    st = corpus[rng.choice(len(corpus), 60, replace=False)].copy()
    st[3]["bx"] = 0; st[3]["by"] = (int(st[3]["wy"]) + 1) % 9      # Black has already won (and does not stand on White)
    st[7]["wwalls"] = 0; st[7]["bwalls"] = 0          # nobody has a wall: handed over at once
    st[11]["wwalls"] = 0                              # only Black may place walls
    out, fin, sums, _, handed = e.uni_lockstep(st, 5, 4242, 17, blocks=2)
    assert handed > 200
    for j in range(len(st)):
        exp, efin, _ = port.q_rollout_philox(st[j], 4242, 17 + 5 * j, 5, policy=1, finals=True)
        assert exp.tobytes() == out[5 * j: 5 * j + 5].tobytes() and efin.tobytes() == fin[5 * j: 5 * j + 5].tobytes(), j
        acc, _ = port.leaf_sum(exp["score"])
        assert (int(acc["lo"]), int(acc["hi"]), int(acc["n"])) == (int(sums["lo"][j]), int(sums["hi"][j]), int(sums["n"][j]))


def test_uni_lockstep_warp32_long_paths_and_small_batches(emu32, port):
    """mazes whose paths overflow the stored layers (nothing is pruned: 80 candidates per ply, several queue batches),
    and a launch shape with fewer than 32 lanes per batch"""
    from tests.util import serpentine_state
    for s in (serpentine_state(turn=0, wwalls=3, bwalls=3), serpentine_state(rows=(0, 2, 4, 7), turn=1, wwalls=2, bwalls=4)):
        exp, efin, _ = port.q_rollout_philox(s, 5, 0, 40, policy=1, finals=True)
        out, fin, _, _, _ = emu32.uni_lockstep(s, 40, 5, 0, blocks=1, lanes=11)
        assert exp.tobytes() == out.tobytes() and efin.tobytes() == fin.tobytes()
