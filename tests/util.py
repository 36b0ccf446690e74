"""Shared helpers for the tests: a numpy Philox4x32-10 (independent third implementation of the
draw convention in oracle/philox_ref.h) and fixture loading."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = np.uint64((int(k0) + W0) & 0xFFFFFFFF)
        k1 = np.uint64((int(k1) + W1) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_stream(seed, rollout_id, n):
    """First n draws of rollout `rollout_id` under `seed` (draw k = block k>>2, word k&3)."""
    nb = (n + 3) // 4
    blk = np.arange(nb, dtype=np.uint64)
    lo = np.full(nb, rollout_id & 0xFFFFFFFF, dtype=np.uint64)
    hi = np.full(nb, (rollout_id >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    z = np.zeros(nb, dtype=np.uint64)
    o = philox4x32_10(blk, lo, hi, z, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(o, axis=1).reshape(-1)[:n].astype(np.uint32)


def golden(name):
    path = os.path.join(GOLDEN_DIR, name)
    if name.endswith(".npy"):
        return np.load(path)
    return np.load(path)


def serpentine_state(rows=(1, 3, 5, 6), wwalls=2, bwalls=2, turn=0):
    """A maze: full-width horizontal barriers under the given rows with the gap alternating between column 0 and
    column 8, so both shortest paths are far longer than the kernel's stored-layer budget (MCTSB_MAX_LAYERS)."""
    from oracle.reflib import QSTATE_DTYPE
    s = np.zeros((), dtype=QSTATE_DTYPE)
    h = 0
    for n, x in enumerate(rows):
        for y in ((1, 3, 5, 7) if n % 2 == 0 else (0, 2, 4, 6)):
            h |= 1 << (8 * x + y)
    s["h_anchor"] = np.uint64(h)
    s["wx"], s["wy"], s["bx"], s["by"] = 0, 4, 8, 4
    s["wwalls"], s["bwalls"], s["turn"], s["move_counter"] = wwalls, bwalls, turn, 16
    return s


THR_0_10, THR_0_75 = 429496730, 3221225472      # qcore.cuh: dist(gen) < 0.10 / < 0.75 as u32 thresholds


def forced_first_ply_stream(port, state, mode, head_rank=None, length=8192, seed=0):
    """A replay stream whose first ply takes a chosen branch of pick_semirandom_move (Quoridor.cpp:693-788):
    mode 1 = best wall (first roll < 0.10), 2 = guided (second roll < 0.75), 3 = first legal wall.  head_rank = pool
    rank (reference loop order: row, column, h before v) the shuffle puts first.  Draw layout of the ply (qcore.cuh
    q_pick_semirandom): [one ignored draw unless force_playwall] [head of the shuffle, then n-1 keys, when the pool has
    n >= 2 walls] [roll 1] [roll 2 unless best wall] ...  Returns (stream, n_pool)."""
    import numpy as np
    p = port.q_primitives(state)
    n_pool = int(p["cheap"].sum())
    rng = np.random.default_rng(seed)
    st = rng.integers(0, 2 ** 32, size=length, dtype=np.uint32)
    pos = 0 if p["force_playwall"] else 1
    if n_pool >= 2:
        if head_rank is not None:
            st[pos] = min(2 ** 32 - 1, -(-head_rank * 2 ** 32 // n_pool) + 1)      # mulhi(head, n_pool) == head_rank
        pos += n_pool
    st[pos] = 0 if mode == 1 else 2 ** 32 - 1
    if mode != 1:
        st[pos + 1] = 0 if mode == 2 else 2 ** 32 - 1
    return st, n_pool


def forced_branch_cases(port, corpus, want=96):
    """(states, streams, notes) driving the rare branches of the candidate-wall phase on purpose:
      * first-legal with an ILLEGAL head (cheap-legal wall that closes a pawn in): the rest of the pool follows in
        key order (Quoridor.cpp:773-788);
      * best-wall on positions with more candidates than one batch of the kernel's task queue takes per lane
        (mazes whose paths overflow the layer budget, so nothing is pruned, and the corpus states with the most
        path-lengthening walls)."""
    import numpy as np
    states, streams, notes = [], [], []
    # -- illegal heads
    for i in range(len(corpus)):
        s = corpus[i]
        if int(s["bwalls"]) == 0 or int(s["wwalls" if s["turn"] == 0 else "bwalls"]) == 0:
            continue
        p = port.q_primitives(s)
        if p["winner"]:
            continue
        cheap = p["cheap"].astype(bool)
        legal_wall = np.zeros(128, dtype=bool)
        for k in range(128):
            legal_wall[k] = p["legal"][(145 if k & 1 else 81) + (k >> 1)]
        bad = np.nonzero(cheap & ~legal_wall)[0]
        if len(bad) == 0 or cheap.sum() < 2:
            continue
        k = int(bad[len(states) % len(bad)])
        rank = int(cheap[:k].sum())
        st, _ = forced_first_ply_stream(port, s, 3, head_rank=rank, seed=1000 + i)
        states.append(s); streams.append(st); notes.append(("head_miss", k))
        if len(states) >= want:
            break
    # -- more candidates than a batch
    heavy = [serpentine_state(turn=0, wwalls=5, bwalls=5), serpentine_state(turn=1, wwalls=5, bwalls=5),
             serpentine_state(rows=(0, 2, 4, 7), turn=0, wwalls=3, bwalls=4)]
    for j, s in enumerate(heavy * 4):
        st, n_pool = forced_first_ply_stream(port, s, 1 if j % 2 == 0 else 2, seed=2000 + j)
        states.append(s); streams.append(st); notes.append(("many_candidates", n_pool))
    return np.array(states, dtype=corpus.dtype), np.stack(streams), notes
