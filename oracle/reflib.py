"""TEST INFRASTRUCTURE — ctypes bindings for the compiled, unmodified reference
(oracle/_ref/libref_replay.so, libref_native.so; built by oracle/Makefile from /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It is the checker, never the thing measured as the product.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

QSTATE_DTYPE = np.dtype(
    [
        ("h_anchor", "<u8"),
        ("v_anchor", "<u8"),
        ("wx", "u1"),
        ("wy", "u1"),
        ("bx", "u1"),
        ("by", "u1"),
        ("wwalls", "u1"),
        ("bwalls", "u1"),
        ("turn", "u1"),
        ("flags", "u1"),
        ("move_counter", "<u4"),
        ("reserved", "<u4"),
    ]
)
assert QSTATE_DTYPE.itemsize == 32

OUTCOME_DTYPE = np.dtype(
    [("score", "<f8"), ("plies", "<u2"), ("kind", "u1"), ("reserved", "u1"), ("draws", "<u4")]
)
assert OUTCOME_DTYPE.itemsize == 16

LEAFSUM_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("n", "<u8")])

NUM_MOVES = 209


def opening_state():
    """Quoridor_state() default constructor, Quoridor.cpp:17-23."""
    s = np.zeros((), dtype=QSTATE_DTYPE)
    s["wx"], s["wy"], s["bx"], s["by"] = 0, 4, 8, 4
    s["wwalls"] = s["bwalls"] = 10
    return s


def raw_from_packed(s):
    """Packed 32-byte record -> the reference's raw members as the harness dumps them
    (walls[81] chars, wall_connections[64], wx, wy, bx, by, wwallsno, bwallsno, turn char,
    move_counter) — 153 int32.  Pure Python/numpy; mirrors include/mctsb.h's wire format."""
    out = np.zeros(153, dtype=np.int32)
    walls = np.full((9, 9), ord(" "), dtype=np.int32)
    h = int(s["h_anchor"])
    v = int(s["v_anchor"])
    hseg = np.zeros((9, 9), dtype=bool)
    vseg = np.zeros((9, 9), dtype=bool)
    for a in range(64):
        x, y = divmod(a, 8)
        if (h >> a) & 1:
            hseg[x, y] = hseg[x, y + 1] = True
            out[81 + a] = 1
        if (v >> a) & 1:
            vseg[x, y] = vseg[x + 1, y] = True
            out[81 + a] = 1
    walls[hseg & ~vseg] = ord("h")
    walls[vseg & ~hseg] = ord("v")
    walls[hseg & vseg] = ord("b")
    out[:81] = walls.reshape(-1)
    out[145:151] = [s["wx"], s["wy"], s["bx"], s["by"], s["wwalls"], s["bwalls"]]
    out[151] = ord("B") if int(s["turn"]) else ord("W")
    out[152] = int(s["move_counter"])
    return out


def apply_move_packed(s, move_id):
    """Packed-state successor (no legality check) — host-side bookkeeping for corpus building."""
    t = s.copy()
    white = int(s["turn"]) == 0
    if move_id < 81:
        x, y = divmod(move_id, 9)
        if white:
            t["wx"], t["wy"] = x, y
        else:
            t["bx"], t["by"] = x, y
    else:
        if move_id < 145:
            t["h_anchor"] = np.uint64(int(s["h_anchor"]) | (1 << (move_id - 81)))
        else:
            t["v_anchor"] = np.uint64(int(s["v_anchor"]) | (1 << (move_id - 145)))
        if white:
            t["wwalls"] = int(s["wwalls"]) - 1
        else:
            t["bwalls"] = int(s["bwalls"]) - 1
    t["turn"] = 0 if not white else 1
    t["move_counter"] = int(s["move_counter"]) + 1
    return t


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class RefLib:
    """One of the two builds of the reference.  kind = 'replay' (stream RNG) or 'native'."""

    def __init__(self, kind="replay"):
        """kind = 'replay' (stream RNG), 'native' (stock RNG; the timed CPU baseline) or 'gpu_dropin' (the
        reference's tree + games linked with integration/gpu_jobscheduler.cpp instead of its JobScheduler.cpp:
        its RolloutJobs run on the B200 through libmctsb.so)."""
        path = os.path.join(REF_DIR, "libref_%s.so" % kind)
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s missing: run `make -C oracle ref dropin` where /root/reference exists" % path
            )
        self.kind = kind
        self.lib = L = C.CDLL(path)
        vp = C.c_void_p
        L.ref_has_stream_rng.restype = C.c_int
        L.ref_q_raw.argtypes = [vp, vp]
        L.ref_q_primitives.argtypes = [vp] * 9
        L.ref_q_sp_with_wall.argtypes = [vp, C.c_int, C.c_int]
        L.ref_q_sp_from.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.ref_q_play.argtypes = [vp, C.c_int, vp]
        L.ref_q_best_step.argtypes = [vp]
        L.ref_q_good_moves.argtypes = [vp, vp]
        L.ref_t_primitives.argtypes = [C.c_uint32, C.c_uint32, vp, vp, vp]
        L.ref_grow_tree.restype = C.c_double
        L.ref_grow_tree.argtypes = [C.c_int, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, C.c_int]
        L.ref_jobscheduler_bench.restype = C.c_double
        L.ref_jobscheduler_bench.argtypes = [C.c_int, vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp]
        L.ref_uniform_bench.restype = C.c_double
        L.ref_uniform_bench.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.ref_jobscheduler_bench2.restype = C.c_double
        L.ref_jobscheduler_bench2.argtypes = [C.c_int, vp, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, vp]
        if kind == "gpu_dropin":
            L.mctsb_dropin_reset.argtypes = [C.c_uint64]
            L.mctsb_dropin_counters.argtypes = [vp, vp]
            L.ref_q_pack_roundtrip.argtypes = [vp, vp]
        if kind == "replay":
            L.ref_q_rollout_stream.restype = C.c_double
            L.ref_q_rollout_stream.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, vp, vp, vp]
            L.ref_q_rollout_philox.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, C.c_int]
            L.ref_q_uniform_stream.restype = C.c_double
            L.ref_q_uniform_stream.argtypes = [vp, vp, C.c_uint64, C.c_int, vp, vp, vp, vp]
            L.ref_q_uniform_philox.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vp, vp, vp, vp, vp]
            L.ref_t_rollout_stream.restype = C.c_double
            L.ref_t_rollout_stream.argtypes = [C.c_uint32, C.c_uint32, vp, C.c_uint64, vp]
            L.ref_t_rollout_philox.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, vp, vp]

    # ---- Quoridor primitives -------------------------------------------------------------
    @staticmethod
    def _s(s):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        return a, a.ctypes.data

    def q_raw(self, s):
        a, p = self._s(s)
        out = np.zeros(153, dtype=np.int32)
        self.lib.ref_q_raw(p, out.ctypes.data)
        return out

    def q_primitives(self, s):
        a, p = self._s(s)
        info = np.zeros(8, dtype=np.int32)
        legal = np.zeros(NUM_MOVES, dtype=np.uint8)
        actions = np.zeros(NUM_MOVES, dtype=np.int32)
        n_actions = C.c_int32(0)
        steps2 = np.zeros(12, dtype=np.int32)
        n_steps2 = C.c_int32(0)
        ev = C.c_double(0)
        cheap = np.zeros(128, dtype=np.uint8)
        self.lib.ref_q_primitives(
            p, info.ctypes.data, legal.ctypes.data, actions.ctypes.data, C.addressof(n_actions),
            steps2.ctypes.data, C.addressof(n_steps2), C.addressof(ev), cheap.ctypes.data,
        )
        return {
            "winner": int(info[0]), "sp_white": int(info[1]), "sp_black": int(info[2]),
            "force_playwall": int(info[3]), "terminal": int(info[4]), "player1_turn": int(info[5]),
            "legal": legal, "actions": actions[: n_actions.value].copy(),
            "steps2": steps2[: n_steps2.value].copy(), "eval": ev.value, "cheap": cheap,
        }

    def q_sp_with_wall(self, s, black, wall_id):
        a, p = self._s(s)
        return self.lib.ref_q_sp_with_wall(p, int(black), int(wall_id))

    def q_sp_from(self, s, black, x, y):
        a, p = self._s(s)
        return self.lib.ref_q_sp_from(p, int(black), int(x), int(y))

    def q_play(self, s, move_id):
        a, p = self._s(s)
        raw = np.zeros(153, dtype=np.int32)
        ok = self.lib.ref_q_play(p, int(move_id), raw.ctypes.data)
        return bool(ok), raw

    def q_good_moves(self, s):
        """the reference's generate_good_moves(): canonical move ids in queue order"""
        a, p = self._s(s)
        ids = np.zeros(256, dtype=np.int32)
        n = self.lib.ref_q_good_moves(p, ids.ctypes.data)
        return ids[:n].copy()

    def q_best_step(self, s):
        a, p = self._s(s)
        return self.lib.ref_q_best_step(p)

    # ---- Quoridor rollouts ------------------------------------------------------------------
    def q_rollout_stream(self, s, stream):
        a, p = self._s(s)
        st = np.ascontiguousarray(stream, dtype=np.uint32)
        draws = C.c_uint64(0)
        raw = np.zeros(153, dtype=np.int32)
        moves = np.zeros(64, dtype=np.int32)
        n_moves = C.c_int32(0)
        kind = C.c_int32(0)
        mism = C.c_int32(0)
        score = self.lib.ref_q_rollout_stream(
            p, st.ctypes.data, st.size, C.addressof(draws), raw.ctypes.data, moves.ctypes.data,
            C.addressof(n_moves), C.addressof(kind), C.addressof(mism),
        )
        assert mism.value == 0, "harness-traced loop disagrees with reference rollout()"
        return {"score": score, "draws": draws.value, "raw_final": raw,
                "moves": moves[: n_moves.value].copy(), "kind": kind.value, "plies": n_moves.value}

    def q_rollout_philox(self, s, seed, first_id, n, traced=True, finals=False):
        a, p = self._s(s)
        scores = np.zeros(n, dtype=np.float64)
        draws = np.zeros(n, dtype=np.uint32)
        kinds = np.zeros(n, dtype=np.int32)
        plies = np.zeros(n, dtype=np.int32)
        raws = np.zeros((n, 153), dtype=np.int32) if finals else None
        bad = self.lib.ref_q_rollout_philox(
            p, seed, first_id, n, scores.ctypes.data, draws.ctypes.data,
            raws.ctypes.data if finals else None, kinds.ctypes.data, plies.ctypes.data, int(traced),
        )
        assert bad == 0, "harness-traced loop disagrees with reference rollout()"
        return {"score": scores, "draws": draws, "kind": kinds, "plies": plies, "raw_final": raws}

    def q_uniform_stream(self, s, stream, max_plies=4096):
        a, p = self._s(s)
        st = np.ascontiguousarray(stream, dtype=np.uint32)
        draws = C.c_uint64(0)
        raw = np.zeros(153, dtype=np.int32)
        plies = C.c_int32(0)
        kind = C.c_int32(0)
        score = self.lib.ref_q_uniform_stream(
            p, st.ctypes.data, st.size, max_plies, C.addressof(draws), raw.ctypes.data,
            C.addressof(plies), C.addressof(kind),
        )
        return {"score": score, "draws": draws.value, "raw_final": raw, "plies": plies.value, "kind": kind.value}

    def q_uniform_philox(self, s, seed, first_id, n, max_plies=4096, finals=False):
        a, p = self._s(s)
        scores = np.zeros(n, dtype=np.float64)
        draws = np.zeros(n, dtype=np.uint32)
        kinds = np.zeros(n, dtype=np.int32)
        plies = np.zeros(n, dtype=np.int32)
        raws = np.zeros((n, 153), dtype=np.int32) if finals else None
        self.lib.ref_q_uniform_philox(
            p, seed, first_id, n, max_plies, scores.ctypes.data, draws.ctypes.data,
            raws.ctypes.data if finals else None, kinds.ctypes.data, plies.ctypes.data,
        )
        return {"score": scores, "draws": draws, "kind": kinds, "plies": plies, "raw_final": raws}

    # ---- Tic-Tac-Toe ----------------------------------------------------------------------------
    def t_primitives(self, x_mask, o_mask):
        info = np.zeros(4, dtype=np.int32)
        actions = np.zeros(9, dtype=np.int32)
        n = C.c_int32(0)
        self.lib.ref_t_primitives(x_mask, o_mask, info.ctypes.data, actions.ctypes.data, C.addressof(n))
        return {"winner": int(info[0]), "terminal": int(info[1]), "player1_turn": int(info[2]),
                "actions": actions[: n.value].copy()}

    def t_rollout_stream(self, x_mask, o_mask, stream):
        st = np.ascontiguousarray(stream, dtype=np.uint32)
        draws = C.c_uint64(0)
        score = self.lib.ref_t_rollout_stream(x_mask, o_mask, st.ctypes.data, st.size, C.addressof(draws))
        return score, draws.value

    def t_rollout_philox(self, x_mask, o_mask, seed, first_id, n):
        scores = np.zeros(n, dtype=np.float64)
        draws = np.zeros(n, dtype=np.uint32)
        self.lib.ref_t_rollout_philox(x_mask, o_mask, seed, first_id, n, scores.ctypes.data, draws.ctypes.data)
        return scores, draws

    # ---- stock host tree and the timed CPU baseline ------------------------------------------------
    def grow_tree(self, game, iters, qstate=None, x_mask=0, o_mask=0, max_children=256):
        qs = np.ascontiguousarray(qstate if qstate is not None else opening_state(), dtype=QSTATE_DTYPE)
        out4 = np.zeros(4, dtype=np.float64)
        ch = np.zeros((max_children, 3), dtype=np.float64)
        secs = self.lib.ref_grow_tree(game, qs.ctypes.data, x_mask, o_mask, iters, out4.ctypes.data,
                                      ch.ctypes.data, max_children)
        n = int(min(out4[3], max_children))
        return {"seconds": secs, "root_sims": out4[0], "root_score": out4[1], "size": out4[2],
                "children": ch[:n].copy()}

    # ---- gpu_dropin build only ------------------------------------------------------------------
    def dropin_reset(self, first_rollout_id=0):
        self.lib.mctsb_dropin_reset(first_rollout_id)

    def dropin_counters(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.lib.mctsb_dropin_counters(C.byref(a), C.byref(b))
        return a.value, b.value

    def dropin_device_ok(self):
        return bool(self.lib.mctsb_dropin_device_ok())

    def q_pack_roundtrip(self, s):
        """packed -> reference Quoridor_state -> the drop-in's packer -> packed"""
        a, p = self._s(s)
        out = np.zeros((), dtype=QSTATE_DTYPE)
        ok = self.lib.ref_q_pack_roundtrip(p, out.ctypes.data)
        return bool(ok), out

    def jobscheduler_bench(self, game, n_threads, n_jobs, qstates=None, x_mask=0, o_mask=0):
        qs = np.ascontiguousarray(qstates if qstates is not None else opening_state(), dtype=QSTATE_DTYPE).reshape(-1)
        mean = C.c_double(0)
        secs = self.lib.ref_jobscheduler_bench(game, qs.ctypes.data, qs.size, x_mask, o_mask, n_threads,
                                               n_jobs, C.addressof(mean))
        return secs, mean.value

    def uniform_bench(self, n_threads, n_playouts, qstates=None, max_plies=4096):
        """UNI playouts through the reference's public API on n_threads threads -> (seconds, mean score, mean plies)"""
        qs = np.ascontiguousarray(qstates if qstates is not None else opening_state(), dtype=QSTATE_DTYPE).reshape(-1)
        mean, plies = C.c_double(0), C.c_double(0)
        secs = self.lib.ref_uniform_bench(qs.ctypes.data, qs.size, n_threads, n_playouts, max_plies, C.addressof(mean), C.addressof(plies))
        return secs, mean.value, plies.value

    def jobscheduler_bench2(self, game, n_threads, n_jobs, qstates=None, x_mask=0, o_mask=0):
        """-> (wall seconds, mean result, standard deviation of one result)"""
        qs = np.ascontiguousarray(qstates if qstates is not None else opening_state(), dtype=QSTATE_DTYPE).reshape(-1)
        mean, sq = C.c_double(0), C.c_double(0)
        secs = self.lib.ref_jobscheduler_bench2(game, qs.ctypes.data, qs.size, x_mask, o_mask, n_threads,
                                                n_jobs, C.addressof(mean), C.addressof(sq))
        return secs, mean.value, max(0.0, sq.value - mean.value ** 2) ** 0.5
