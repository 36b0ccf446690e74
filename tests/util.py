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
