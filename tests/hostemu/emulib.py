"""ctypes bindings of the TEST-ONLY host compile of the device core (tests/hostemu/hostemu.cpp)."""
import ctypes as C

import numpy as np

from oracle.reflib import OUTCOME_DTYPE, QSTATE_DTYPE
from oracle.portlib import TSTATE_DTYPE
from .build import build


class EmuLib:
    def __init__(self, warp32=False):
        """warp32=True: the build whose lockstep body runs on emulated 32-lane warps (simt.h fibers)."""
        self.lib = L = C.CDLL(build(warp32=warp32))
        vp = C.c_void_p
        L.emu_q_primitives.argtypes = [vp] * 6
        L.emu_q_sp_with_wall.argtypes = [vp, C.c_int, C.c_int]
        L.emu_q_sp_from.argtypes = [vp, C.c_int, C.c_int]
        L.emu_q_best_step.argtypes = [vp]
        L.emu_q_play.argtypes = [vp, C.c_int, vp]
        L.emu_q_cut_check.argtypes = [vp, C.c_int, vp]
        L.emu_q_detour_check.argtypes = [vp, vp]
        L.emu_q_rollout_streams.argtypes = [vp, C.c_int, C.c_int, vp, C.c_uint32, C.c_uint32, vp, vp, vp]
        L.emu_q_rollout_philox.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp]
        L.emu_t_rollout_philox.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, vp]
        L.emu_t_rollout_streams.argtypes = [vp, C.c_int, vp, C.c_uint32, C.c_uint32, vp]
        L.emu_t_winner.argtypes = [C.c_uint32, C.c_uint32]
        L.emu_lockstep.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, C.c_int, C.c_int]
        L.emu_uni_lockstep.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        if warp32:
            L.emu_tree_create.restype = vp
            L.emu_tree_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32]
            L.emu_tree_destroy.argtypes = [vp]
            L.emu_tree_set_good_moves.argtypes = [vp, C.c_int]
            L.emu_tree_select.argtypes = [vp, C.c_int, C.c_uint32, vp]
            L.emu_tree_backprop.argtypes = [vp, C.c_int, vp]
            L.emu_tree_dump.restype = C.c_uint64
            L.emu_tree_dump.argtypes = [vp, C.c_uint64] + [vp] * 7
            L.emu_tree_advance.argtypes = [vp, C.c_int]
            L.emu_tree_stats.argtypes = [vp, vp]
            L.emu_leaf_sum_to_double.restype = C.c_double
            L.emu_leaf_sum_to_double.argtypes = [C.c_uint64, C.c_uint64]

    def q_primitives(self, s):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        info = np.zeros(4, dtype=np.int32)
        legal4 = np.zeros(4, dtype=np.uint64)
        ev = C.c_double(0)
        cheap2 = np.zeros(2, dtype=np.uint64)
        steps = np.zeros(12, dtype=np.int32)
        self.lib.emu_q_primitives(a.ctypes.data, info.ctypes.data, legal4.ctypes.data, C.addressof(ev),
                                  cheap2.ctypes.data, steps.ctypes.data)
        legal = np.array([(int(legal4[i >> 6]) >> (i & 63)) & 1 for i in range(209)], dtype=np.uint8)
        cheap = np.zeros(128, dtype=np.uint8)
        for a_ in range(64):
            cheap[2 * a_] = (int(cheap2[0]) >> a_) & 1
            cheap[2 * a_ + 1] = (int(cheap2[1]) >> a_) & 1
        return {"winner": int(info[0]), "sp_white": int(info[1]), "sp_black": int(info[2]),
                "force_playwall": int(info[3]), "eval": ev.value, "legal": legal, "cheap": cheap,
                "steps2": steps[steps >= 0].copy()}

    def q_sp_with_wall(self, s, black, wall_id):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        k = 2 * (wall_id - 81) if wall_id < 145 else 2 * (wall_id - 145) + 1
        return self.lib.emu_q_sp_with_wall(a.ctypes.data, int(black), k)

    def q_detour_check(self, states):
        """(violations, cheap-legal walls, walls settled without a search, cheap-legal but illegal walls) over `states`"""
        a = np.ascontiguousarray(states, dtype=QSTATE_DTYPE).reshape(-1)
        out = np.zeros(3, dtype=np.int64)
        bad = 0
        for i in range(a.size):
            bad += self.lib.emu_q_detour_check(a[i:i + 1].ctypes.data, out.ctypes.data)
        return bad, int(out[0]), int(out[1]), int(out[2])

    def q_cut_check(self, states, player):
        """(violations, candidates, candidates the pruning masks keep) summed over `states`"""
        a = np.ascontiguousarray(states, dtype=QSTATE_DTYPE).reshape(-1)
        out = np.zeros(2, dtype=np.int64)
        bad = 0
        for i in range(a.size):
            bad += self.lib.emu_q_cut_check(a[i:i + 1].ctypes.data, int(player), out.ctypes.data)
        return bad, int(out[0]), int(out[1])

    def q_sp_from(self, s, black, x, y):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        return self.lib.emu_q_sp_from(a.ctypes.data, int(black), 9 * x + y)

    def q_best_step(self, s):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        return self.lib.emu_q_best_step(a.ctypes.data)

    def q_play(self, s, move_id):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        out = np.zeros((), dtype=QSTATE_DTYPE)
        ok = self.lib.emu_q_play(a.ctypes.data, int(move_id), out.ctypes.data)
        return bool(ok), out

    def q_rollout_streams(self, states, streams, policy=0):
        a = np.ascontiguousarray(states, dtype=QSTATE_DTYPE).reshape(-1)
        st = np.ascontiguousarray(streams, dtype=np.uint32)
        out = np.zeros(a.size, dtype=OUTCOME_DTYPE)
        fin = np.zeros(a.size, dtype=QSTATE_DTYPE)
        work = np.zeros(3, dtype=np.uint64)
        self.lib.emu_q_rollout_streams(a.ctypes.data, policy, a.size, st.ctypes.data, st.shape[1], st.shape[1],
                                       out.ctypes.data, fin.ctypes.data, work.ctypes.data)
        return out, fin, work

    def q_rollout_philox(self, s, seed, first_id, n, policy=0):
        a = np.ascontiguousarray(s, dtype=QSTATE_DTYPE)
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        fin = np.zeros(n, dtype=QSTATE_DTYPE)
        work = np.zeros(3, dtype=np.uint64)
        self.lib.emu_q_rollout_philox(a.ctypes.data, policy, seed, first_id, n, out.ctypes.data, fin.ctypes.data,
                                      work.ctypes.data)
        return out, fin, work

    def t_rollout_philox(self, x, o, seed, first_id, n):
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        self.lib.emu_t_rollout_philox(x, o, seed, first_id, n, out.ctypes.data)
        return out

    def t_rollout_streams(self, states, streams):
        a = np.ascontiguousarray(states, dtype=TSTATE_DTYPE).reshape(-1)
        st = np.ascontiguousarray(streams, dtype=np.uint32)
        out = np.zeros(a.size, dtype=OUTCOME_DTYPE)
        self.lib.emu_t_rollout_streams(a.ctypes.data, a.size, st.ctypes.data, st.shape[1], st.shape[1], out.ctypes.data)
        return out

    def lockstep(self, states, R, seed, first_id, streams=None, blocks=1, lanes=32):
        """The production kernel body on the host (one-lane warp, or `blocks` emulated thread blocks of 32-lane
        warps in the warp32 build): Philox mode (streams=None) or replay."""
        from oracle.reflib import LEAFSUM_DTYPE
        a = np.ascontiguousarray(states, dtype=QSTATE_DTYPE).reshape(-1)
        if streams is not None:
            st = np.ascontiguousarray(streams, dtype=np.uint32)
            assert R == 1 and st.shape[0] == a.size
            sp, sl, ss = st.ctypes.data, st.shape[1], st.shape[1]
        else:
            sp, sl, ss = None, 0, 0
        n = a.size * R
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        fin = np.zeros(n, dtype=QSTATE_DTYPE)
        sums = np.zeros(a.size, dtype=LEAFSUM_DTYPE)
        ctr = np.zeros(4, dtype=np.uint64)
        self.lib.emu_lockstep(a.ctypes.data, a.size, R, seed, first_id, sp, sl, ss, out.ctypes.data, fin.ctypes.data,
                              sums.ctypes.data, ctr.ctypes.data, blocks, lanes)
        return out, fin, sums, ctr

    def uni_lockstep(self, states, R, seed, first_id, streams=None, blocks=1, lanes=32):
        """The UNI path on the host: phase A (uni_lockstep_body) + phase B (q_walk_uni over the hand-over records).
        -> (outcomes, finals, sums, counters, rollouts handed over to phase B)"""
        import ctypes as C
        from oracle.reflib import LEAFSUM_DTYPE
        a = np.ascontiguousarray(states, dtype=QSTATE_DTYPE).reshape(-1)
        if streams is not None:
            st = np.ascontiguousarray(streams, dtype=np.uint32)
            assert R == 1 and st.shape[0] == a.size
            sp, sl, ss = st.ctypes.data, st.shape[1], st.shape[1]
        else:
            sp, sl, ss = None, 0, 0
        n = a.size * R
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        fin = np.zeros(n, dtype=QSTATE_DTYPE)
        sums = np.zeros(a.size, dtype=LEAFSUM_DTYPE)
        ctr = np.zeros(4, dtype=np.uint64)
        handed = C.c_uint64(0)
        rc = self.lib.emu_uni_lockstep(a.ctypes.data, a.size, R, seed, first_id, sp, sl, ss, out.ctypes.data, fin.ctypes.data,
                                       sums.ctypes.data, ctr.ctypes.data, blocks, lanes, C.byref(handed))
        assert rc == 0
        return out, fin, sums, ctr, handed.value


class EmuTree:
    """The device-resident tree (csrc/tree_body.cuh) on emulated 32-lane warps; the playouts of a flush are the caller's."""

    def __init__(self, emu, root, R, K, pool_cap=1 << 18, c=1.41, logtab_n=1 << 16, chunk=512):
        assert emu.lib.emu_warp_width() == 32
        self.lib, self.R, self.K = emu.lib, R, K
        r = np.ascontiguousarray(root, dtype=QSTATE_DTYPE)
        self.h = self.lib.emu_tree_create(r.ctypes.data, R, K, pool_cap, c, logtab_n, chunk)
        assert self.h, "bad chunk size"

    def close(self):
        if self.h:
            self.lib.emu_tree_destroy(self.h)
            self.h = None

    def set_good_moves_only(self, on=True):
        self.lib.emu_tree_set_good_moves(self.h, int(bool(on)))

    def select(self, n, slot=0):
        """Selection + Expansion of a flush of n descents into flush slot 0 / 1 -> (leaf positions, error bits)"""
        leaves = np.zeros(n, dtype=QSTATE_DTYPE)
        err = self.lib.emu_tree_select(self.h, slot, n, leaves.ctypes.data)
        assert err >= 0
        return leaves, err

    def backprop(self, sums, slot=0):
        from oracle.reflib import LEAFSUM_DTYPE
        s = np.ascontiguousarray(sums, dtype=LEAFSUM_DTYPE)
        return self.lib.emu_tree_backprop(self.h, slot, s.ctypes.data)

    def advance(self, move_id):
        return self.lib.emu_tree_advance(self.h, int(move_id)) == 0

    def stats(self):
        o = np.zeros(4, dtype=np.uint32)
        self.lib.emu_tree_stats(self.h, o.ctypes.data)
        return {"pool_top": int(o[0]), "items": int(o[1]), "error": int(o[2]), "exact_evals": int(o[3])}

    def dump(self, cap=1 << 20):
        a = {"depth": np.zeros(cap, np.int32), "move": np.zeros(cap, np.int32), "sims": np.zeros(cap, np.uint64),
             "score": np.zeros(cap, np.float64), "size": np.zeros(cap, np.uint32), "n_children": np.zeros(cap, np.uint32),
             "virt": np.zeros(cap, np.uint32)}
        n = self.lib.emu_tree_dump(self.h, cap, *[a[k].ctypes.data for k in ("depth", "move", "sims", "score", "size", "n_children", "virt")])
        assert n <= cap
        return {k: v[:n].copy() for k, v in a.items()}
