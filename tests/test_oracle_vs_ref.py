"""CPU, only where the compiled reference exists (oracle/_ref/*.so, built from /root/reference by
`make -C oracle ref`): live differential tests of the C restatement against the reference, on
fresh random games — the goldens cover the same ground with fixed inputs."""
import numpy as np

from oracle.reflib import apply_move_packed, opening_state, raw_from_packed


def _random_games(ref, seed, games, step_bias):
    rng = np.random.default_rng(seed)
    for _ in range(games):
        s = opening_state()
        for _ply in range(250):
            p = ref.q_primitives(s)
            yield s, p
            if p["terminal"] or len(p["actions"]) == 0:
                break
            acts = p["actions"]
            steps = acts[acts < 81]
            if len(steps) and rng.random() < step_bias:
                m = int(rng.choice(steps))
            else:
                m = int(rng.choice(acts))
            ok, raw = ref.q_play(s, m)
            assert ok
            s = apply_move_packed(s, m)
            assert (raw == raw_from_packed(s)).all()


def test_primitives_live(ref, port):
    n = 0
    for s, a in _random_games(ref, 5, 25, 0.5):
        b = port.q_primitives(s)
        for k in a:
            if isinstance(a[k], np.ndarray):
                assert a[k].shape == b[k].shape and (a[k] == b[k]).all(), k
            else:
                assert a[k] == b[k], k
        assert (ref.q_raw(s) == port.q_raw(s)).all()
        if not a["terminal"]:
            assert ref.q_best_step(s) == port.q_best_step(s)
        n += 1
    assert n > 500


def test_shortest_path_queries_live(ref, port):
    rng = np.random.default_rng(6)
    for s, a in _random_games(ref, 6, 6, 0.4):
        for wid in np.nonzero(a["cheap"])[0][:: 7]:
            mid = 81 + wid // 2 if wid % 2 == 0 else 145 + wid // 2
            for black in (0, 1):
                assert ref.q_sp_with_wall(s, black, int(mid)) == port.q_sp_with_wall(s, black, int(mid))
        x, y = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        for black in (0, 1):
            assert ref.q_sp_from(s, black, x, y) == port.q_sp_from(s, black, x, y)


def test_replay_live(ref, port):
    rng = np.random.default_rng(8)
    states = [s.copy() for s, a in _random_games(ref, 9, 12, 0.5) if not a["terminal"]]
    for i in rng.choice(len(states), 120, replace=False):
        st = rng.integers(0, 2**32, size=16384, dtype=np.uint32)
        a = ref.q_rollout_stream(states[i], st)
        b = port.q_rollout_stream(states[i], st)
        assert (a["score"], a["draws"], a["kind"], a["plies"]) == (b["score"], b["draws"], b["kind"], b["plies"])
        assert (a["moves"] == b["moves"]).all()
        assert (a["raw_final"] == raw_from_packed(b["final"])).all()
    for i in rng.choice(len(states), 12, replace=False):
        st = rng.integers(0, 2**32, size=16384, dtype=np.uint32)
        a = ref.q_uniform_stream(states[i], st)
        b = port.q_rollout_stream(states[i], st, policy=1)
        assert (a["score"], a["draws"], a["kind"], a["plies"]) == (b["score"], b["draws"], b["kind"], b["plies"])
        assert (a["raw_final"] == raw_from_packed(b["final"])).all()


def test_tictactoe_live(ref, port):
    rng = np.random.default_rng(10)
    for _ in range(300):
        cells = rng.permutation(9)[: int(rng.integers(0, 9))]
        x = o = 0
        for k, c in enumerate(cells):
            if k % 2 == 0:
                x |= 1 << int(c)
            else:
                o |= 1 << int(c)
        if len(cells) % 2 == 1:
            o |= 0x8000
        a, b = ref.t_primitives(x, o), port.t_primitives(x, o)
        assert (a["winner"], a["terminal"], a["player1_turn"]) == (b["winner"], b["terminal"], b["player1_turn"])
        st = rng.integers(0, 2**32, size=16, dtype=np.uint32)
        assert ref.t_rollout_stream(x, o, st) == port.t_rollout_stream(x, o, st)


def test_stock_tree_tictactoe_statistics(ref_native):
    """Config 1 on the stock reference: 10k iterations, 4 rollouts each (mcts.cpp:57-81)."""
    r = ref_native.grow_tree(1, 10000)
    assert r["root_sims"] == 40000 and r["size"] == 10000
    assert 0.55 < r["root_score"] / r["root_sims"] < 0.80
