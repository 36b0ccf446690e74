"""TEST INFRASTRUCTURE — ctypes bindings for oracle/liboracle.so, the plain-C restatement of the
reference's rollout path (oracle/quoridor_oracle.c, oracle/ttt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .reflib import LEAFSUM_DTYPE, NUM_MOVES, OUTCOME_DTYPE, QSTATE_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

WORK_DTYPE = np.dtype(
    [("bfs_calls", "<u8"), ("bfs_pops", "<u8"), ("flood_steps", "<u8"), ("cheap_wall_tests", "<u8"), ("plies", "<u8")]
)
TSTATE_DTYPE = np.dtype([("x_mask", "<u2"), ("o_mask", "<u2")])


class _Draws(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint64), ("seed", C.c_uint64),
                ("rollout_id", C.c_uint64), ("pos", C.c_uint64)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("quoridor_oracle.c", "ttt_oracle.c", "oracle.h", "philox_ref.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class PortLib:
    def __init__(self):
        build()
        self.lib = L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oracle_q_primitives.argtypes = [vp] * 9
        L.oracle_q_raw.argtypes = [vp, vp]
        L.oracle_q_sp_with_wall.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_q_sp_from.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.oracle_q_play.argtypes = [vp, C.c_int, vp]
        L.oracle_q_best_step.argtypes = [vp]
        L.oracle_q_good_moves.argtypes = [vp, vp]
        L.oracle_q_rollout.restype = C.c_double
        L.oracle_q_rollout.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
        L.oracle_q_rollout_philox.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp]
        L.oracle_q_rollout_streams.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, C.c_uint32, vp, vp]
        L.oracle_t_primitives.argtypes = [C.c_uint32, C.c_uint32, vp, vp, vp]
        L.oracle_t_rollout.restype = C.c_double
        L.oracle_t_rollout.argtypes = [C.c_uint32, C.c_uint32, vp, vp, vp]
        L.oracle_t_rollout_philox.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, vp]
        L.oracle_t_rollout_streams.argtypes = [vp, C.c_int, vp, C.c_uint32, C.c_uint32, vp]
        L.oracle_leaf_sum_add.argtypes = [vp, C.c_double]
        L.oracle_leaf_sum_to_double.restype = C.c_double
        L.oracle_leaf_sum_to_double.argtypes = [vp]

    @staticmethod
    def _s(s):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        return a, a.ctypes.data

    def q_raw(self, s):
        a, p = self._s(s)
        out = np.zeros(153, dtype=np.int32)
        self.lib.oracle_q_raw(p, out.ctypes.data)
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
        self.lib.oracle_q_primitives(
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
        return self.lib.oracle_q_sp_with_wall(p, int(black), int(wall_id))

    def q_sp_from(self, s, black, x, y):
        a, p = self._s(s)
        return self.lib.oracle_q_sp_from(p, int(black), int(x), int(y))

    def q_play(self, s, move_id):
        a, p = self._s(s)
        out = np.zeros((), dtype=QSTATE_DTYPE)
        ok = self.lib.oracle_q_play(p, int(move_id), out.ctypes.data)
        return bool(ok), out

    def q_best_step(self, s):
        a, p = self._s(s)
        return self.lib.oracle_q_best_step(p)

    def q_good_moves(self, s):
        """generate_good_moves (Quoridor.cpp:518-592): canonical move ids in queue order"""
        a, p = self._s(s)
        ids = np.zeros(209, dtype=np.int32)
        n = self.lib.oracle_q_good_moves(p, ids.ctypes.data)
        return ids[:n].copy()

    def q_rollout_stream(self, s, stream, policy=0, work=False):
        a, p = self._s(s)
        st = np.ascontiguousarray(stream, dtype=np.uint32)
        d = _Draws(st.ctypes.data, st.size, 0, 0, 0)
        fin = np.zeros((), dtype=QSTATE_DTYPE)
        moves = np.zeros(64, dtype=np.int32)
        nm = C.c_int32(0)
        kind = C.c_int32(0)
        w = np.zeros((), dtype=WORK_DTYPE)
        score = self.lib.oracle_q_rollout(p, policy, C.addressof(d), fin.ctypes.data, moves.ctypes.data,
                                          C.addressof(nm), C.addressof(kind), w.ctypes.data if work else None)
        return {"score": score, "draws": d.pos, "final": fin, "moves": moves[: min(nm.value, 64)].copy(),
                "kind": kind.value, "plies": nm.value, "work": w}

    def q_rollout_philox(self, s, seed, first_id, n, policy=0, finals=False, work=False):
        a, p = self._s(s)
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        fin = np.zeros(n, dtype=QSTATE_DTYPE) if finals else None
        w = np.zeros((), dtype=WORK_DTYPE)
        self.lib.oracle_q_rollout_philox(p, policy, seed, first_id, n, out.ctypes.data,
                                         fin.ctypes.data if finals else None, w.ctypes.data if work else None)
        return out, fin, w

    def q_rollout_streams(self, states, streams, policy=0, finals=True):
        a = np.ascontiguousarray(states, dtype=QSTATE_DTYPE).reshape(-1)
        st = np.ascontiguousarray(streams, dtype=np.uint32)
        assert st.ndim == 2 and st.shape[0] == a.size
        out = np.zeros(a.size, dtype=OUTCOME_DTYPE)
        fin = np.zeros(a.size, dtype=QSTATE_DTYPE) if finals else None
        self.lib.oracle_q_rollout_streams(a.ctypes.data, policy, a.size, st.ctypes.data, st.shape[1], st.shape[1],
                                          out.ctypes.data, fin.ctypes.data if finals else None)
        return out, fin

    def t_primitives(self, x_mask, o_mask):
        info = np.zeros(4, dtype=np.int32)
        actions = np.zeros(9, dtype=np.int32)
        n = C.c_int32(0)
        self.lib.oracle_t_primitives(x_mask, o_mask, info.ctypes.data, actions.ctypes.data, C.addressof(n))
        return {"winner": int(info[0]), "terminal": int(info[1]), "player1_turn": int(info[2]),
                "actions": actions[: n.value].copy()}

    def t_rollout_stream(self, x_mask, o_mask, stream):
        st = np.ascontiguousarray(stream, dtype=np.uint32)
        d = _Draws(st.ctypes.data, st.size, 0, 0, 0)
        score = self.lib.oracle_t_rollout(x_mask, o_mask, C.addressof(d), None, None)
        return score, d.pos

    def t_rollout_philox(self, x_mask, o_mask, seed, first_id, n):
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        self.lib.oracle_t_rollout_philox(x_mask, o_mask, seed, first_id, n, out.ctypes.data)
        return out

    def t_rollout_streams(self, states, streams):
        a = np.ascontiguousarray(states, dtype=TSTATE_DTYPE).reshape(-1)
        st = np.ascontiguousarray(streams, dtype=np.uint32)
        out = np.zeros(a.size, dtype=OUTCOME_DTYPE)
        self.lib.oracle_t_rollout_streams(a.ctypes.data, a.size, st.ctypes.data, st.shape[1], st.shape[1], out.ctypes.data)
        return out

    def leaf_sum(self, scores):
        acc = np.zeros((), dtype=LEAFSUM_DTYPE)
        for s in np.asarray(scores, dtype=np.float64):
            self.lib.oracle_leaf_sum_add(acc.ctypes.data, float(s))
        return acc, self.lib.oracle_leaf_sum_to_double(acc.ctypes.data)
