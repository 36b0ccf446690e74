"""Generate the golden fixtures under tests/golden/ FROM THE COMPILED, UNMODIFIED REFERENCE
(oracle/_ref/libref_replay.so, built by `make -C oracle ref` where /root/reference exists).

The reference ships no golden vectors of its own, so these files are the pin:
every value in them was produced by the reference's own functions.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reflib import QSTATE_DTYPE, RefLib, apply_move_packed, opening_state, raw_from_packed  # noqa: E402
from tests.util import philox_stream  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
R = RefLib("replay")
SEED = 0x2026101900000001


def walls_on_board(s):
    return bin(int(s["h_anchor"])).count("1") + bin(int(s["v_anchor"])).count("1")


def midgame_state(i):
    """uniformly random LEGAL play through the reference API from the opening,
    stopped when a per-state target of 10..20 walls on board is reached; games that end first are
    discarded."""
    rng = np.random.default_rng(42 + i)
    while True:
        target = int(rng.integers(10, 21))
        s = opening_state()
        ok = True
        while walls_on_board(s) < target:
            p = R.q_primitives(s)
            if p["terminal"] or len(p["actions"]) == 0:
                ok = False
                break
            m = int(p["actions"][int(rng.integers(0, len(p["actions"])))])
            s = apply_move_packed(s, m)
        if ok and not R.q_primitives(s)["terminal"]:
            return s


def varied_states(n, seed):
    """Positions from every phase (step-heavy play so games progress to the end), terminal ones included."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        s = opening_state()
        step_bias = rng.uniform(0.3, 0.9)
        for ply in range(300):
            p = R.q_primitives(s)
            if rng.random() < 0.25:
                out.append(s.copy())
            if p["terminal"] or len(p["actions"]) == 0:
                out.append(s.copy())
                break
            acts = p["actions"]
            steps = acts[acts < 81]
            if len(steps) and rng.random() < step_bias:
                m = int(steps[int(rng.integers(0, len(steps)))])
            else:
                m = int(acts[int(rng.integers(0, len(acts)))])
            ok, raw = R.q_play(s, m)
            assert ok
            s = apply_move_packed(s, m)
            assert (raw == raw_from_packed(s)).all()
    return out[:n]


def primitives_table(states):
    n = len(states)
    t = {
        "winner": np.zeros(n, np.int8), "sp_white": np.zeros(n, np.int8), "sp_black": np.zeros(n, np.int8),
        "force_playwall": np.zeros(n, np.int8), "eval": np.zeros(n, np.float64),
        "legal": np.zeros((n, 27), np.uint8), "cheap": np.zeros((n, 16), np.uint8),
        "n_actions": np.zeros(n, np.int16), "actions": np.full((n, 140), -1, np.int16),
        "steps2": np.full((n, 12), -1, np.int16), "best_step": np.zeros(n, np.int16),
    }
    for i, s in enumerate(states):
        p = R.q_primitives(s)
        for k in ("winner", "sp_white", "sp_black", "force_playwall", "eval"):
            t[k][i] = p[k]
        t["legal"][i] = np.packbits(p["legal"], bitorder="little")
        t["cheap"][i] = np.packbits(p["cheap"], bitorder="little")
        t["n_actions"][i] = len(p["actions"])
        t["actions"][i, : len(p["actions"])] = p["actions"]
        t["steps2"][i, : len(p["steps2"])] = p["steps2"]
        t["best_step"][i] = R.q_best_step(s) if not p["terminal"] else -1
    return t


def main():
    t0 = time.time()
    mid = [midgame_state(i) for i in range(4096)]
    var = varied_states(512, 7)
    corpus = np.array(mid + var, dtype=QSTATE_DTYPE)
    np.save(os.path.join(OUT, "corpus_q.npy"), corpus)
    print("corpus", corpus.shape, "%.0fs" % (time.time() - t0))

    sel = np.concatenate([np.arange(256), np.arange(4096, 4608)])
    prim = primitives_table([corpus[i] for i in sel])
    np.savez_compressed(os.path.join(OUT, "golden_q_primitives.npz"), index=sel, **prim)
    print("primitives", len(sel), "%.0fs" % (time.time() - t0))

    # replayed REF playouts: stream = Philox(SEED, id)
    rng = np.random.default_rng(99)
    ridx = np.concatenate([rng.choice(4096, 256, replace=False), 4096 + rng.choice(512, 128, replace=False)])
    n = len(ridx)
    rep = {"index": ridx, "id": np.arange(n, dtype=np.uint64) + 1000, "score": np.zeros(n), "draws": np.zeros(n, np.uint32),
           "kind": np.zeros(n, np.uint8), "plies": np.zeros(n, np.uint16), "final": np.zeros(n, QSTATE_DTYPE),
           "moves": np.full((n, 50), -1, np.int16)}
    for j, i in enumerate(ridx):
        st = philox_stream(SEED, int(rep["id"][j]), 16384)
        r = R.q_rollout_stream(corpus[i], st)
        assert r["draws"] <= 16384
        s = corpus[i].copy()
        for m in r["moves"]:
            s = apply_move_packed(s, int(m))
        assert (raw_from_packed(s) == r["raw_final"]).all()
        rep["score"][j], rep["draws"][j], rep["kind"][j], rep["plies"][j] = r["score"], r["draws"], r["kind"], r["plies"]
        rep["final"][j] = s
        rep["moves"][j, : len(r["moves"])] = r["moves"]
    np.savez_compressed(os.path.join(OUT, "golden_q_replay.npz"), seed=np.uint64(SEED), **rep)
    print("replay", n, "%.0fs" % (time.time() - t0))

    # opening position: 512 Philox rollouts through the reference's own rollout()
    o = R.q_rollout_philox(opening_state(), SEED, 0, 512, traced=True, finals=False)
    np.savez_compressed(os.path.join(OUT, "golden_q_opening.npz"), seed=np.uint64(SEED), score=o["score"],
                        draws=o["draws"], kind=o["kind"].astype(np.uint8), plies=o["plies"].astype(np.uint16))

    # known-answer streams on the opening position (S0 reproduces the survey's value)
    kat = {}
    streams = {
        "S0": np.zeros(65536, np.uint32),
        "S1": np.full(65536, 0x80000000, np.uint32),
        "S2": ((np.arange(65536, dtype=np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        "S3": (((np.arange(65536, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B9)) ^ (np.arange(65536, dtype=np.uint64) >> np.uint64(3))).astype(np.uint32),
    }
    for name, st in streams.items():
        r = R.q_rollout_stream(opening_state(), st)
        tt = R.t_rollout_stream(0, 0, st)
        kat[name] = np.array([r["score"], r["draws"], r["kind"], r["plies"], tt[0], tt[1]])
        print(name, repr(r["score"]), r["draws"], tt)
    np.savez_compressed(os.path.join(OUT, "golden_kat.npz"), **kat)

    # UNI playouts (driver over the reference's public API)
    uidx = rng.choice(4096, 48, replace=False)
    uni = {"index": uidx, "id": np.arange(48, dtype=np.uint64) + 5000, "score": np.zeros(48), "draws": np.zeros(48, np.uint32),
           "kind": np.zeros(48, np.uint8), "plies": np.zeros(48, np.uint16)}
    for j, i in enumerate(uidx):
        r = R.q_uniform_philox(corpus[i], SEED, int(uni["id"][j]), 1)
        uni["score"][j], uni["draws"][j], uni["kind"][j], uni["plies"][j] = r["score"][0], r["draws"][0], r["kind"][0], r["plies"][0]
    np.savez_compressed(os.path.join(OUT, "golden_q_uniform.npz"), seed=np.uint64(SEED), **uni)
    print("uniform", "%.0fs" % (time.time() - t0))

    # free-running REF score distributions (for the 3-sigma tests): opening + 15 corpus states
    didx = np.concatenate([[-1], rng.choice(4096, 15, replace=False)])
    dist = {"index": didx, "n": 3000, "first_id": np.uint64(1 << 32), "scores": np.zeros((len(didx), 3000))}
    for j, i in enumerate(didx):
        s = opening_state() if i < 0 else corpus[i]
        dist["scores"][j] = R.q_rollout_philox(s, SEED, 1 << 32, 3000, traced=False)["score"]
    np.savez_compressed(os.path.join(OUT, "golden_q_distribution.npz"), seed=np.uint64(SEED), **dist)
    print("distribution", "%.0fs" % (time.time() - t0))

    # Tic-Tac-Toe: every reachable state
    seen, stack, rows = set(), [(0, 0)], []
    while stack:
        x, o = stack.pop()
        if (x, o) in seen:
            continue
        seen.add((x, o))
        p = R.t_primitives(x, o)
        amask = 0
        for c in p["actions"]:
            amask |= 1 << int(c)
        rows.append((x, o, p["winner"], p["terminal"], p["player1_turn"], amask))
        if p["terminal"]:
            continue
        for c in p["actions"]:
            if o & 0x8000:
                stack.append((x, (o | (1 << int(c))) & 0x7FFF))
            else:
                stack.append((x | (1 << int(c)), o | 0x8000))
    rows = np.array(sorted(rows), dtype=np.int32)
    tsc = np.zeros(len(rows))
    tdr = np.zeros(len(rows), np.uint32)
    for j, (x, o) in enumerate(rows[:, :2]):
        sc, dr = R.t_rollout_philox(int(x), int(o), SEED, 7000 + j, 1)
        tsc[j], tdr[j] = sc[0], dr[0]
    e = R.t_rollout_philox(0, 0, SEED, 0, 20000)
    np.savez_compressed(os.path.join(OUT, "golden_t.npz"), seed=np.uint64(SEED), states=rows, score=tsc, draws=tdr,
                        empty_scores=e[0], empty_draws=e[1])
    print("ttt", len(rows), "%.0fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
