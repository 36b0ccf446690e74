"""GPU (-m gpu): the sm_100a kernels, called through the C ABI, against the oracle and the
reference goldens.  Bit-exact for every integer / index / score value."""
from fractions import Fraction

import numpy as np
import pytest

import montecarlotreesearch_b200 as m
from tests.util import golden, philox_stream

pytestmark = pytest.mark.gpu
G, T = m.GAME_QUORIDOR, m.GAME_TICTACTOE


@pytest.fixture(scope="module")
def be():
    m.build()          # no-op when libmctsb.so is newer than its sources
    b = m.RolloutBackend(0, m.DEFAULT_SEED)
    yield b
    b.close()


@pytest.fixture(scope="module")
def corpus():
    return golden("corpus_q.npy")


def test_library_is_the_cuda_one(be):
    assert be.counters()["kernel_launches"] == 0
    be.rollout_batch(T, 0, np.zeros(1, dtype=m.TSTATE_DTYPE), 32)
    assert be.counters()["kernel_launches"] == 1


# ---- rule primitives on the whole corpus -----------------------------------------------------------
def test_movegen_vs_golden(be, corpus):
    g = golden("golden_q_primitives.npz")
    legal, info, ev = be.q_movegen(corpus[g["index"]])
    bits = m.unpack_legal(legal)
    assert (np.packbits(bits, axis=1, bitorder="little") == g["legal"]).all()
    assert (info["winner"] == g["winner"]).all() and (info["sp_white"] == g["sp_white"]).all()
    assert (info["sp_black"] == g["sp_black"]).all() and (info["force_playwall"] == g["force_playwall"]).all()
    assert (ev == g["eval"]).all()
    # the legal set is exactly what generate_all_moves enumerates
    for j in range(len(g["index"])):
        n = int(g["n_actions"][j])
        assert sorted(np.nonzero(bits[j])[0].tolist()) == sorted(g["actions"][j, :n].tolist())


def test_movegen_vs_oracle_full_corpus(be, port, corpus):
    legal, info, ev = be.q_movegen(corpus)
    bits = m.unpack_legal(legal)
    for i in range(len(corpus)):
        p = port.q_primitives(corpus[i])
        assert (bits[i] == p["legal"]).all(), i
        assert (info["winner"][i], info["sp_white"][i], info["sp_black"][i], info["force_playwall"][i]) == (
            p["winner"], p["sp_white"], p["sp_black"], p["force_playwall"])
        assert ev[i] == p["eval"]


def test_play_vs_oracle(be, port, corpus):
    rng = np.random.default_rng(4)
    idx = rng.choice(len(corpus), 2000)
    ids = rng.integers(0, 209, size=2000).astype(np.int32)
    out, ok = be.q_play(corpus[idx], ids)
    for j in range(2000):
        a_ok, a = port.q_play(corpus[idx[j]], int(ids[j]))
        assert bool(ok[j]) == a_ok and out[j].tobytes() == a.tobytes()


# ---- replayed playouts -----------------------------------------------------------------------------
def test_replay_vs_reference_golden(be, corpus):
    g = golden("golden_q_replay.npz")
    streams = np.stack([philox_stream(int(g["seed"]), int(i), 16384) for i in g["id"]])
    out, fin = be.rollout_replay(G, m.POLICY_REF, corpus[g["index"]], streams)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()
    assert fin.tobytes() == g["final"].tobytes()


def test_replay_vs_oracle_random_streams(be, port, corpus):
    rng = np.random.default_rng(12)
    idx = rng.choice(len(corpus), 1024)
    streams = rng.integers(0, 2**32, size=(1024, 8192), dtype=np.uint32)
    out, fin = be.rollout_replay(G, m.POLICY_REF, corpus[idx], streams)
    exp, efin = port.q_rollout_streams(corpus[idx], streams)
    assert out.tobytes() == exp.tobytes() and fin.tobytes() == efin.tobytes()


def test_forced_rare_branches(be, port, corpus):
    """Replay streams that force what free-running playouts meet only by chance (tests/util.py forced_branch_cases):
    96 first-legal picks whose shuffled head is an ILLEGAL wall (the rest of the pool follows in key order,
    Quoridor.cpp:773-788) and best-wall / guided searches over 80 unpruned candidates — more than one batch of the
    warp's task queue takes per lane.  108 distinct leaves in one launch: every warp mixes leaves and branches."""
    from tests.util import forced_branch_cases
    states, streams, notes = forced_branch_cases(port, corpus)
    exp, efin = port.q_rollout_streams(states, streams)
    for j, (kind, k) in enumerate(notes):                # the oracle did take the branch: a wall other than the head
        if kind == "head_miss":
            first = int(port.q_rollout_stream(states[j], streams[j])["moves"][0])
            assert first >= 81 and first != (145 if k & 1 else 81) + (k >> 1)
    out, fin = be.rollout_replay(G, m.POLICY_REF, states, streams)
    assert out.tobytes() == exp.tobytes() and fin.tobytes() == efin.tobytes()


def test_known_answer_streams(be):
    g = golden("golden_kat.npz")
    streams = np.stack([
        np.zeros(65536, np.uint32),
        np.full(65536, 0x80000000, np.uint32),
        ((np.arange(65536, dtype=np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        (((np.arange(65536, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B9)) ^ (np.arange(65536, dtype=np.uint64) >> np.uint64(3))).astype(np.uint32),
    ])
    out, _ = be.rollout_replay(G, m.POLICY_REF, np.repeat(m.quoridor_opening(), 4), streams)
    tout, _ = be.rollout_replay(T, 0, np.zeros(4, dtype=m.TSTATE_DTYPE), streams)
    for j, name in enumerate(("S0", "S1", "S2", "S3")):
        assert [out["score"][j], out["draws"][j], out["kind"][j], out["plies"][j], tout["score"][j], tout["draws"][j]] == list(g[name])


# ---- Philox-keyed playouts ---------------------------------------------------------------------------
def test_opening_rollouts_vs_reference_golden(be):
    g = golden("golden_q_opening.npz")
    out, _ = be.rollout_outcomes(G, m.POLICY_REF, m.quoridor_opening(), 512, 0)
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()
    assert (out["kind"] == g["kind"]).all() and (out["plies"] == g["plies"]).all()


@pytest.mark.parametrize("name", ["golden_q_distribution.npz", "golden_q_distribution64.npz"])
def test_free_running_scores_vs_reference_golden(be, corpus, name):
    g = golden(name)
    n = int(g["n"])
    states = np.array([m.quoridor_opening() if i < 0 else corpus[i] for i in g["index"]], dtype=m.QSTATE_DTYPE)
    # leaf j, rollout k uses id first + j*n + k; the golden used first_id for every state, so run one leaf at a time
    for j in range(len(states)):
        out, _ = be.rollout_outcomes(G, m.POLICY_REF, states[j], n, int(g["first_id"]), finals=False)
        assert (out["score"] == g["scores"][j]).all(), j


def test_three_sigma_agreement_on_disjoint_streams(be, corpus):
    """Free-running win rates with streams the reference never saw: sample means within 3 sigma
    (sample variance; scores are not Bernoulli)."""
    g = golden("golden_q_distribution.npz")
    states = np.array([m.quoridor_opening() if i < 0 else corpus[i] for i in g["index"]], dtype=m.QSTATE_DTYPE)
    R = 20000
    out, _ = be.rollout_outcomes(G, m.POLICY_REF, states, R, 1 << 40, finals=False)
    sc = out["score"].reshape(len(states), R)
    for j in range(len(states)):
        ref = g["scores"][j]
        se = np.sqrt(ref.var(ddof=1) / len(ref) + sc[j].var(ddof=1) / R)
        assert abs(ref.mean() - sc[j].mean()) <= 3.0 * se + 1e-12, (j, ref.mean(), sc[j].mean(), se)


def test_sixty_four_state_distribution_on_disjoint_streams(be, corpus):
    """64 states (opening + 63 mid-game), 20 000 rollouts each on streams the reference never
    saw.  Every state's mean within 3.5 sigma of the reference's, and the 64 deviations together chi-square
    plausible (sum of z^2 below the 99.9 % quantile of chi2(64) = 112.3)."""
    g = golden("golden_q_distribution64.npz")
    states = np.array([m.quoridor_opening() if i < 0 else corpus[i] for i in g["index"]], dtype=m.QSTATE_DTYPE)
    R = 20000
    out, _ = be.rollout_outcomes(G, m.POLICY_REF, states, R, 1 << 41, finals=False)
    sc = out["score"].reshape(len(states), R)
    z = []
    for j in range(len(states)):
        ref = g["scores"][j]
        se = np.sqrt(ref.var(ddof=1) / len(ref) + sc[j].var(ddof=1) / R)
        if se == 0:                                  # decided positions: every rollout returns the same value
            assert ref.mean() == sc[j].mean(), j
            continue
        z.append((ref.mean() - sc[j].mean()) / se)
    z = np.array(z)
    assert np.abs(z).max() <= 3.5, z
    assert (z ** 2).sum() <= 112.3 * len(z) / 64 + 1e-9, (z ** 2).sum()


def test_good_moves_vs_reference_golden(be, corpus):
    """mctsb_q_goodmoves (one warp per position, one lane per candidate wall) against the reference's public
    generate_good_moves() on the opening and the whole corpus: same moves, same queue order."""
    g = golden("golden_q_goodmoves.npz")
    states = np.array([m.quoridor_opening() if i < 0 else corpus[i] for i in g["index"]], dtype=m.QSTATE_DTYPE)
    got = be.q_goodmoves(states)
    for j in range(len(states)):
        n = int(g["count"][j])
        assert len(got[j]) == n and (got[j] == g["ids"][j, :n]).all(), j


def test_uniform_policy_vs_golden(be, corpus):
    g = golden("golden_q_uniform.npz")
    for j, i in enumerate(g["index"]):
        out, _ = be.rollout_outcomes(G, m.POLICY_UNI, corpus[i], 1, int(g["id"][j]))
        assert out["score"][0] == g["score"][j] and out["draws"][0] == g["draws"][j]
        assert out["kind"][0] == g["kind"][j] and out["plies"][0] == g["plies"][j]


def test_uniform_policy_vs_oracle(be, port, corpus):
    rng = np.random.default_rng(2)
    idx = rng.choice(4096, 64)
    out, fin = be.rollout_outcomes(G, m.POLICY_UNI, corpus[idx], 2, 77)
    for j, i in enumerate(idx):
        exp, efin, _ = port.q_rollout_philox(corpus[i], m.DEFAULT_SEED, 77 + 2 * j, 2, policy=1, finals=True)
        assert out[2 * j: 2 * j + 2].tobytes() == exp.tobytes() and fin[2 * j: 2 * j + 2].tobytes() == efin.tobytes()


# ---- Tic-Tac-Toe ---------------------------------------------------------------------------------------
def test_uniform_policy_lockstep_path_vs_simple_kernel_and_oracle(be, port, corpus):
    """The production UNI path (lockstep wall phase + walk kernel) against the one-thread-per-rollout kernel
    (MCTSB_SIMPLE_KERNEL=1) on 256 mid-game leaves x 8 playouts, and against the oracle on ragged special leaves
    (terminal, nobody has walls, one side has walls, a maze that overflows the stored layers); exact leaf sums too."""
    import os
    from tests.util import serpentine_state
    os.environ["MCTSB_SIMPLE_KERNEL"] = "1"
    try:
        simple = m.RolloutBackend(0, m.DEFAULT_SEED)
    finally:
        del os.environ["MCTSB_SIMPLE_KERNEL"]
    rng = np.random.default_rng(35)
    states = corpus[rng.choice(len(corpus), 256, replace=False)]
    a, fa = be.rollout_outcomes(G, m.POLICY_UNI, states, 8, 4711)
    b, fb = simple.rollout_outcomes(G, m.POLICY_UNI, states, 8, 4711)
    simple.close()
    assert a.tobytes() == b.tobytes() and fa.tobytes() == fb.tobytes()
    st = corpus[rng.choice(len(corpus), 40, replace=False)].copy()
    st[3]["bx"] = 0; st[3]["by"] = (int(st[3]["wy"]) + 1) % 9      # Black has already won (and does not stand on White)
    st[7]["wwalls"] = 0; st[7]["bwalls"] = 0
    st[11]["wwalls"] = 0
    st[13] = serpentine_state(turn=0, wwalls=3, bwalls=3)
    out, fin = be.rollout_outcomes(G, m.POLICY_UNI, st, 5, 99)
    _, sums = be.rollout_batch(G, m.POLICY_UNI, st, 5, 99)
    for j in range(len(st)):
        exp, efin, _ = port.q_rollout_philox(st[j], m.DEFAULT_SEED, 99 + 5 * j, 5, policy=1, finals=True)
        assert exp.tobytes() == out[5 * j: 5 * j + 5].tobytes() and efin.tobytes() == fin[5 * j: 5 * j + 5].tobytes(), j
        acc, _ = port.leaf_sum(exp["score"])
        assert (int(acc["lo"]), int(acc["hi"]), int(acc["n"])) == (int(sums["lo"][j]), int(sums["hi"][j]), int(sums["n"][j]))


def test_tictactoe_vs_golden(be):
    g = golden("golden_t.npz")
    out, _ = be.rollout_outcomes(T, 0, np.zeros(1, dtype=m.TSTATE_DTYPE), 20000, 0)
    assert (out["score"] == g["empty_scores"]).all() and (out["draws"] == g["empty_draws"]).all()
    st = np.zeros(len(g["states"]), dtype=m.TSTATE_DTYPE)
    st["x_mask"], st["o_mask"] = g["states"][:, 0], g["states"][:, 1]
    out, _ = be.rollout_outcomes(T, 0, st, 1, 7000)       # leaf j uses id 7000 + j
    assert (out["score"] == g["score"]).all() and (out["draws"] == g["draws"]).all()


# ---- aggregation, partition independence, edge cases ------------------------------------------------------
def test_leaf_sums_are_exact_and_order_free(be, corpus):
    states = corpus[:37]
    R = 257                                   # ragged: warps straddle leaves
    out, _ = be.rollout_outcomes(G, m.POLICY_REF, states, R, 123, finals=False)
    score, sums = be.rollout_batch(G, m.POLICY_REF, states, R, 123)
    sc = out["score"].reshape(37, R)
    for i in range(37):
        exact = sum(Fraction(float(v)) for v in sc[i])
        assert Fraction(int(sums["hi"][i]) * (1 << 29) + int(sums["lo"][i]), 1 << 57) == exact
        assert int(sums["n"][i]) == R and score[i] == float(exact)


def test_results_do_not_depend_on_batch_shape(be, corpus):
    """Rollout id r always yields the same playout, however the ids are cut into leaves/launches."""
    s = corpus[5]
    a, _ = be.rollout_outcomes(G, m.POLICY_REF, s, 600, 1000, finals=False)
    b1, _ = be.rollout_outcomes(G, m.POLICY_REF, s, 173, 1000, finals=False)
    b2, _ = be.rollout_outcomes(G, m.POLICY_REF, s, 427, 1173, finals=False)
    assert a.tobytes() == b1.tobytes() + b2.tobytes()
    c, _ = be.rollout_outcomes(G, m.POLICY_REF, np.repeat(s, 3), 200, 1000, finals=False)
    assert a.tobytes() == c.tobytes()


def test_terminal_and_single_inputs(be):
    s = m.quoridor_opening().copy()
    s["wx"], s["by"] = 8, 3                    # White already on its goal row: no draws, score 1
    out, fin = be.rollout_outcomes(G, m.POLICY_REF, s, 3, 0)
    assert (out["score"] == 1.0).all() and (out["draws"] == 0).all() and (out["plies"] == 0).all() and (out["kind"] == 0).all()
    s["wx"], s["bx"] = 3, 0                    # Black on row 0
    out, _ = be.rollout_outcomes(G, m.POLICY_REF, s, 1, 0)
    assert out["score"][0] == 0.0 and out["kind"][0] == 1
    t = np.zeros(1, dtype=m.TSTATE_DTYPE)
    t["x_mask"] = 0b000000111                  # x already won
    out, _ = be.rollout_outcomes(T, 0, t, 2, 0)
    assert (out["score"] == 1.0).all() and (out["draws"] == 0).all()


def test_bad_arguments_are_refused(be):
    s = m.quoridor_opening().copy()
    s["bx"], s["by"] = 0, 4                    # both pawns on one square
    with pytest.raises(m.BackendError) as e:
        be.rollout_batch(G, m.POLICY_REF, s, 4)
    assert e.value.status == 4
    with pytest.raises(m.BackendError) as e:
        be.rollout_batch(G, 7, m.quoridor_opening(), 4)
    assert e.value.status == 3
    be.submit(G, m.POLICY_REF, m.quoridor_opening(), 32)
    with pytest.raises(m.BackendError) as e:
        be.submit(G, m.POLICY_REF, m.quoridor_opening(), 32)
    assert e.value.status == 6
    be.wait()
    with pytest.raises(m.BackendError) as e:
        be.wait()
    assert e.value.status == 5


def test_full_size_properties(be):
    """Config 2 size (1M rollouts from the opening): properties that need no CPU replay."""
    n = 1 << 20
    score, sums = be.rollout_batch(G, m.POLICY_REF, m.quoridor_opening(), n, 0)
    assert int(sums["n"][0]) == n
    mean = score[0] / n
    g = golden("golden_q_distribution.npz")
    ref = g["scores"][0]
    assert abs(mean - ref.mean()) <= 3.0 * np.sqrt(ref.var(ddof=1) / len(ref) + ref.var(ddof=1) / n)
    # same ids in two halves -> the exact integer sums add up
    _, s1 = be.rollout_batch(G, m.POLICY_REF, m.quoridor_opening(), n // 2, 0)
    _, s2 = be.rollout_batch(G, m.POLICY_REF, m.quoridor_opening(), n // 2, n // 2)
    assert int(s1["lo"][0]) + int(s2["lo"][0]) == int(sums["lo"][0])
    assert int(s1["hi"][0]) + int(s2["hi"][0]) == int(sums["hi"][0])


def test_two_independent_kernels_agree(be, corpus):
    """The production lockstep kernel against the one-thread-per-rollout kernel (MCTSB_SIMPLE_KERNEL=1):
    two device implementations of the REF policy, same outcomes and final positions."""
    import os
    os.environ["MCTSB_SIMPLE_KERNEL"] = "1"
    try:
        simple = m.RolloutBackend(0, m.DEFAULT_SEED)
    finally:
        del os.environ["MCTSB_SIMPLE_KERNEL"]
    rng = np.random.default_rng(33)
    states = corpus[rng.choice(len(corpus), 256, replace=False)]
    a, fa = be.rollout_outcomes(G, m.POLICY_REF, states, 8, 31337)
    b, fb = simple.rollout_outcomes(G, m.POLICY_REF, states, 8, 31337)
    simple.close()
    assert a.tobytes() == b.tobytes() and fa.tobytes() == fb.tobytes()
    a, fa = be.rollout_outcomes(G, m.POLICY_REF, m.quoridor_opening(), 4096, 99, finals=False)
    assert abs(a["score"].mean() - 0.65) < 0.03


def test_long_paths_overflow_the_layer_budget(be, port):
    """Maze positions whose shortest paths exceed the stored-layer budget (MCTSB_MAX_LAYERS = 24)."""
    from tests.util import serpentine_state
    states = np.array([serpentine_state(turn=0), serpentine_state(turn=1), serpentine_state(rows=(0, 2, 4, 7), turn=0)],
                      dtype=m.QSTATE_DTYPE)
    legal, info, ev = be.q_movegen(states)
    bits = m.unpack_legal(legal)
    out, fin = be.rollout_outcomes(G, m.POLICY_REF, states, 96, 4321)
    for i, s in enumerate(states):
        p = port.q_primitives(s)
        assert p["sp_white"] > 24 and p["sp_black"] > 24
        assert (bits[i] == p["legal"]).all() and info["sp_white"][i] == p["sp_white"] and info["sp_black"][i] == p["sp_black"]
        exp, efin, _ = port.q_rollout_philox(s, m.DEFAULT_SEED, 4321 + 96 * i, 96, finals=True)
        assert exp.tobytes() == out[96 * i: 96 * (i + 1)].tobytes() and efin.tobytes() == fin[96 * i: 96 * (i + 1)].tobytes()


def test_every_launch_shape_plays_the_same_rollouts(be):
    """Launch sizes that take different grid shapes (a few lanes per warp, one balanced wave, two equal batches,
    the persistent grid): rollout id r is the same playout in all of them, and the exact sums add up."""
    s = m.quoridor_opening()
    big, _ = be.rollout_outcomes(G, m.POLICY_REF, s, 200000, 5000, finals=False)          # persistent grid, 32 lanes
    for n in (1, 4, 33, 1000, 70001, 120000):
        out, _ = be.rollout_outcomes(G, m.POLICY_REF, s, n, 5000, finals=False)
        assert out.tobytes() == big[:n].tobytes(), n
    _, whole = be.rollout_batch(G, m.POLICY_REF, s, 70001, 5000)
    lo = hi = cnt = 0
    for first, n in ((5000, 1), (5001, 4), (5005, 995), (6000, 69001)):
        _, part = be.rollout_batch(G, m.POLICY_REF, s, n, first)
        lo += int(part["lo"][0]); hi += int(part["hi"][0]); cnt += int(part["n"][0])
    assert (lo + (hi << 29), cnt) == (int(whole["lo"][0]) + (int(whole["hi"][0]) << 29), 70001)


def test_resident_states_are_invalidated_by_other_uploads(be, corpus):
    """mctsb_run_resident plays what mctsb_upload_states made resident — any other call that re-uploads leaf
    states drops the marker instead of silently mixing new and stale positions."""
    be.upload_states(G, corpus[:64])
    be.run_resident(G, m.POLICY_REF, 8, 0)
    _, a = be.wait()
    assert int(a["n"].sum()) == 64 * 8
    be.rollout_batch(G, m.POLICY_REF, corpus[100:110], 4, 0)          # overwrites the device copy of the states
    with pytest.raises(m.BackendError) as e:
        be.run_resident(G, m.POLICY_REF, 8, 0)
    assert e.value.status == 5
    be.upload_states(G, corpus[:64])
    be.run_resident(G, m.POLICY_REF, 8, 0)
    _, b = be.wait()
    assert a.tobytes() == b.tobytes()


def test_two_host_threads_two_handles(be, corpus):
    """Distinct handles are independent: two host threads drive their own handle at the same time (the launch
    configuration is per device and built under a lock), results equal the single-threaded ones."""
    import threading
    states = corpus[:200]
    want, _ = be.rollout_outcomes(G, m.POLICY_REF, states, 6, 77, finals=False)
    got, errs = {}, []

    def work(tag):
        try:
            h = m.RolloutBackend(0, m.DEFAULT_SEED)
            for _ in range(3):
                got[tag], _ = h.rollout_outcomes(G, m.POLICY_REF, states, 6, 77, finals=False)
            h.close()
        except Exception as ex:  # pragma: no cover
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert got[0].tobytes() == want.tobytes() and got[1].tobytes() == want.tobytes()


# ---- expansion prepared on the device, the exchange behind the C ABI, the stock-RNG distribution ------------
def test_submit_x_returns_the_movegen_masks_next_to_the_sums(be, corpus):
    """mctsb_submit_x(MCTSB_SUBMIT_LEGAL): same sums as the plain pair, plus the masks mctsb_q_movegen returns."""
    st = corpus[100:100 + 333]                       # ragged: not a multiple of anything
    legal_ref, _, _ = be.q_movegen(st)
    score0, sums0 = be.rollout_batch(G, 0, st, 8, 12345)
    be.submit_x(G, 0, st, 8, 12345, flags=m.SUBMIT_LEGAL)
    score1, sums1, legal = be.wait_x()
    assert (legal == legal_ref).all() and (sums0 == sums1).all() and (score0 == score1).all()
    be.submit(G, 0, st[:4], 8, 0)                    # masks were not asked for: asking for them at wait time is an error
    out = np.zeros((4, 4), dtype=np.uint64)
    assert be.lib.mctsb_wait_x(be.handle, None, None, out.ctypes.data) == 3     # MCTSB_ERR_BAD_ARG, the batch stays pending
    be.wait()
    with pytest.raises(m.BackendError):              # Tic-Tac-Toe has no move-generation kernel
        be.submit_x(T, 0, np.zeros(1, dtype=m.TSTATE_DTYPE), 8, 0, flags=m.SUBMIT_LEGAL)


def test_comm_single_rank_is_the_identity():
    """A communicator of one (mctsb_comm_unique_id -> mctsb_comm_init_rank) brings NCCL up behind the C ABI: sums over
    one rank return their input, bit for bit, and the error paths answer."""
    b = m.RolloutBackend(0, 5)
    v = np.array([1.5, -2.25, 3e300, 0.1])
    with pytest.raises(m.BackendError) as e:
        b.allreduce_root_stats(v)
    assert e.value.status == 5                       # no communicator yet
    assert b.comm_info() == (1, 0)
    b.comm_init_rank(m.comm_unique_id(), 1, 0)
    assert b.comm_info() == (1, 0)
    assert (b.allreduce_root_stats(v) == v).all()
    s = np.zeros(3, dtype=m.LEAFSUM_DTYPE)
    s["lo"], s["hi"], s["n"] = [1, 2 ** 63 + 5, 7], [9, 8, 2 ** 40], [3, 3, 3]
    assert (b.allreduce_leaf_sums(s) == s).all()
    with pytest.raises(m.BackendError):
        b.comm_init_rank(m.comm_unique_id(), 1, 0)   # one communicator per handle
    b.submit(T, 0, np.zeros(1, dtype=m.TSTATE_DTYPE), 4, 0)
    with pytest.raises(m.BackendError) as e:
        b.allreduce_root_stats(v)
    assert e.value.status == 6                       # busy: a submit is pending
    b.wait()
    b.close()


def test_comm_over_every_device_of_the_box():
    """One process, one handle per GPU (mctsb_comm_init_all), reductions issued as one group from one thread; the
    merged leaf sums of rollouts sharded by id equal the single-GPU sums over the whole id range."""
    n = m.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    bs = [m.RolloutBackend(d, m.DEFAULT_SEED) for d in range(n)]
    m.comm_init_all(bs)
    assert [b.comm_info() for b in bs] == [(n, r) for r in range(n)]
    import ctypes as C
    vecs = [np.arange(6, dtype=np.float64) * (r + 1) for r in range(n)]
    hs = (C.c_void_p * n)(*[b.handle for b in bs])
    ps = (C.c_void_p * n)(*[v.ctypes.data for v in vecs])
    assert bs[0].lib.mctsb_allreduce_root_stats_group(hs, ps, n, 6) == 0
    want = np.arange(6, dtype=np.float64) * (n * (n + 1) // 2)
    assert all((v == want).all() for v in vecs)
    for b in bs:
        b.close()


def test_mean_score_agrees_with_the_stock_rng_reference(be, ref_native):
    """The reference with its OWN generator and std::shuffle (no shim): its mean result over RolloutJobs from the
    opening position against the GPU's mean over 2^20 Philox-keyed rollouts.  4 sigma of the CPU sample (the GPU
    sample's own error is 30x smaller); bench.py reports the same comparison at 3 sigma in cpu_baseline."""
    import os
    threads = len(os.sched_getaffinity(0))
    jobs = 3000 * threads
    _, mean, std = ref_native.jobscheduler_bench2(0, threads, jobs)
    score, _ = be.rollout_batch(G, 0, np.array([m.quoridor_opening()]), 1 << 20, 1 << 40)
    gpu_mean = float(score[0]) / (1 << 20)
    sem = std / jobs ** 0.5
    assert abs(gpu_mean - mean) <= 4 * sem, (gpu_mean, mean, sem)
