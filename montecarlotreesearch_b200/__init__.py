"""montecarlotreesearch_b200 — B200-native rollout backend for the MCTS Simulation phase.

Python is only the test/bench harness here: the product is the C-ABI shared library
`libmctsb.so` (include/mctsb.h; CUDA kernels for sm_100a under csrc/) and the C++ host side that
mirrors the reference's MCTS_state / MCTS_tree API (host/).  This module binds the C ABI with
ctypes, one method per entry point.  There is no CPU fallback: if the library is missing, or no
CUDA device is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmctsb.so")

GAME_QUORIDOR, GAME_TICTACTOE = 0, 1
POLICY_REF, POLICY_UNI = 0, 1
NUM_MOVES = 209
DEFAULT_SEED = 0x2026101900000001

QSTATE_DTYPE = np.dtype(
    [("h_anchor", "<u8"), ("v_anchor", "<u8"), ("wx", "u1"), ("wy", "u1"), ("bx", "u1"), ("by", "u1"),
     ("wwalls", "u1"), ("bwalls", "u1"), ("turn", "u1"), ("flags", "u1"), ("move_counter", "<u4"), ("reserved", "<u4")]
)
TSTATE_DTYPE = np.dtype([("x_mask", "<u2"), ("o_mask", "<u2")])
OUTCOME_DTYPE = np.dtype([("score", "<f8"), ("plies", "<u2"), ("kind", "u1"), ("reserved", "u1"), ("draws", "<u4")])
LEAFSUM_DTYPE = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("n", "<u8")])
QINFO_DTYPE = np.dtype([("winner", "i1"), ("sp_white", "i1"), ("sp_black", "i1"), ("force_playwall", "i1")])
COUNTERS_DTYPE = np.dtype([("kernel_launches", "<u8"), ("rollouts", "<u8"), ("plies", "<u8"),
                           ("flood_steps", "<u8"), ("flood_queries", "<u8")])
assert QSTATE_DTYPE.itemsize == 32 and OUTCOME_DTYPE.itemsize == 16 and LEAFSUM_DTYPE.itemsize == 24

EXPORTS = [
    "mctsb_create", "mctsb_destroy", "mctsb_abi_version", "mctsb_strerror", "mctsb_last_cuda_error",
    "mctsb_device_count", "mctsb_submit", "mctsb_wait", "mctsb_rollout_batch", "mctsb_rollout_outcomes",
    "mctsb_rollout_replay", "mctsb_q_movegen", "mctsb_q_play", "mctsb_upload_states", "mctsb_run_resident",
    "mctsb_int_issue_peak", "mctsb_lop3_peak", "mctsb_flush_l2", "mctsb_get_counters", "mctsb_leaf_sum_to_double",
    "mctsb_submit_x", "mctsb_wait_x", "mctsb_q_goodmoves",
    "mctsb_comm_unique_id", "mctsb_comm_init_rank", "mctsb_comm_init_all", "mctsb_comm_destroy", "mctsb_comm_info",
    "mctsb_allreduce_root_stats", "mctsb_allreduce_leaf_sums", "mctsb_allreduce_root_stats_group",
    "mctsb_tree_create", "mctsb_tree_destroy", "mctsb_tree_grow", "mctsb_tree_get_info", "mctsb_tree_root_children",
    "mctsb_tree_best_move", "mctsb_tree_advance", "mctsb_tree_root_state", "mctsb_tree_dump", "mctsb_tree_set_overlap", "mctsb_tree_merge_root", "mctsb_tree_set_good_moves",
]
TREE_INFO_DTYPE = np.dtype([("root_sims", "<u8"), ("root_score", "<f8"), ("root_size", "<u4"), ("root_children", "<u4"),
                            ("nodes_used", "<u4"), ("deepest_level", "<u4"), ("exact_evals", "<u4"), ("flags", "<u4"),
                            ("flushes", "<u8"), ("rollouts", "<u8"), ("ms_select", "<f4"), ("ms_rollout", "<f4"),
                            ("ms_backprop", "<f4"), ("ms_total", "<f4")])
SUBMIT_LEGAL = 1
Q_GOOD_MAX = 144
COMM_ID_BYTES = 128


class BackendError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("mctsb status %d: %s" % (status, message))
        self.status = status


def quoridor_opening():
    """Quoridor_state() (reference Quoridor.cpp:17-23) as a packed record."""
    s = np.zeros((), dtype=QSTATE_DTYPE)
    s["wx"], s["wy"], s["bx"], s["by"] = 0, 4, 8, 4
    s["wwalls"] = s["bwalls"] = 10
    return s


def build(force=False, verbose=False):
    from ._build import build_all
    build_all(force=force, verbose=verbose)


def load_library():
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            "%s not built: run `python -m montecarlotreesearch_b200._build` (needs nvcc). "
            "The rollout backend has no CPU fallback." % LIB_PATH
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
    L.mctsb_create.argtypes = [C.POINTER(vp), i32, u64]
    L.mctsb_destroy.argtypes = [vp]
    L.mctsb_strerror.restype = C.c_char_p
    L.mctsb_strerror.argtypes = [i32]
    L.mctsb_last_cuda_error.restype = C.c_char_p
    L.mctsb_last_cuda_error.argtypes = [vp]
    L.mctsb_device_count.argtypes = [C.POINTER(i32)]
    L.mctsb_submit.argtypes = [vp, i32, i32, vp, i32, u32, u64]
    L.mctsb_wait.argtypes = [vp, vp, vp]
    L.mctsb_submit_x.argtypes = [vp, i32, i32, vp, i32, u32, u64, u32]
    L.mctsb_wait_x.argtypes = [vp, vp, vp, vp]
    L.mctsb_comm_unique_id.argtypes = [vp]
    L.mctsb_comm_init_rank.argtypes = [vp, vp, i32, i32]
    L.mctsb_comm_init_all.argtypes = [C.POINTER(vp), i32]
    L.mctsb_comm_destroy.argtypes = [vp]
    L.mctsb_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.mctsb_allreduce_root_stats.argtypes = [vp, vp, i32]
    L.mctsb_allreduce_leaf_sums.argtypes = [vp, vp, i32]
    L.mctsb_allreduce_root_stats_group.argtypes = [C.POINTER(vp), C.POINTER(vp), i32, i32]
    L.mctsb_rollout_batch.argtypes = [vp, i32, i32, vp, i32, u32, u64, vp, vp]
    L.mctsb_rollout_outcomes.argtypes = [vp, i32, i32, vp, i32, u32, u64, vp, vp]
    L.mctsb_rollout_replay.argtypes = [vp, i32, i32, vp, i32, vp, u32, u32, vp, vp]
    L.mctsb_q_movegen.argtypes = [vp, vp, i32, vp, vp, vp]
    L.mctsb_q_play.argtypes = [vp, vp, vp, i32, vp, vp]
    L.mctsb_q_goodmoves.argtypes = [vp, vp, i32, vp, vp]
    L.mctsb_upload_states.argtypes = [vp, i32, vp, i32]
    L.mctsb_run_resident.argtypes = [vp, i32, i32, u32, u64, C.POINTER(C.c_float)]
    L.mctsb_int_issue_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.mctsb_lop3_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.mctsb_flush_l2.argtypes = [vp]
    L.mctsb_get_counters.argtypes = [vp, vp]
    L.mctsb_leaf_sum_to_double.restype = C.c_double
    L.mctsb_leaf_sum_to_double.argtypes = [vp]
    L.mctsb_tree_create.argtypes = [vp, C.POINTER(vp), vp, i32, u32, u32, u64, C.c_double]
    L.mctsb_tree_destroy.argtypes = [vp]
    L.mctsb_tree_grow.argtypes = [vp, u32, u32, u64]
    L.mctsb_tree_set_overlap.argtypes = [vp, i32]
    L.mctsb_tree_merge_root.argtypes = [vp]
    L.mctsb_tree_set_good_moves.argtypes = [vp, i32]
    L.mctsb_tree_get_info.argtypes = [vp, vp]
    L.mctsb_tree_root_children.argtypes = [vp, vp, vp, vp, i32, C.POINTER(i32)]
    L.mctsb_tree_best_move.argtypes = [vp, C.POINTER(C.c_int32)]
    L.mctsb_tree_advance.argtypes = [vp, C.c_int32]
    L.mctsb_tree_root_state.argtypes = [vp, vp]
    L.mctsb_tree_dump.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, C.POINTER(u64)]
    return L


def device_count():
    L = load_library()
    n = C.c_int(0)
    L.mctsb_device_count(C.byref(n))
    return n.value


class RolloutBackend:
    """Thin object wrapper over the opaque `mctsb_backend*` handle."""

    def __init__(self, device=0, seed=DEFAULT_SEED):
        self.lib = load_library()
        self.handle = C.c_void_p()
        st = self.lib.mctsb_create(C.byref(self.handle), device, seed)
        if st != 0:
            self.handle = None
            raise BackendError(st, self.lib.mctsb_strerror(st).decode())
        self.device, self.seed = device, seed

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mctsb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            msg = self.lib.mctsb_strerror(st).decode()
            if st in (2, 8, 9):
                msg += " / " + self.lib.mctsb_last_cuda_error(self.handle).decode()
            raise BackendError(st, msg)

    @staticmethod
    def _states(game, states):
        dt = QSTATE_DTYPE if game == GAME_QUORIDOR else TSTATE_DTYPE
        return np.ascontiguousarray(states, dtype=dt).reshape(-1)

    # -- hot path -----------------------------------------------------------------------------
    def submit(self, game, policy, states, rollouts_per_leaf, first_rollout_id=0):
        a = self._states(game, states)
        self._ck(self.lib.mctsb_submit(self.handle, game, policy, a.ctypes.data, a.size, rollouts_per_leaf, first_rollout_id))
        self._pending_n = a.size

    def wait(self):
        n = self._pending_n
        score = np.zeros(n, dtype=np.float64)
        sums = np.zeros(n, dtype=LEAFSUM_DTYPE)
        self._ck(self.lib.mctsb_wait(self.handle, score.ctypes.data, sums.ctypes.data))
        return score, sums

    def rollout_batch(self, game, policy, states, rollouts_per_leaf, first_rollout_id=0):
        self.submit(game, policy, states, rollouts_per_leaf, first_rollout_id)
        return self.wait()

    def submit_x(self, game, policy, states, rollouts_per_leaf, first_rollout_id=0, flags=SUBMIT_LEGAL):
        a = self._states(game, states)
        self._ck(self.lib.mctsb_submit_x(self.handle, game, policy, a.ctypes.data, a.size, rollouts_per_leaf, first_rollout_id, flags))
        self._pending_n = a.size

    def wait_x(self):
        """-> (score sums, exact sums, legal-move masks (n, 4) of the submitted leaves)"""
        n = self._pending_n
        score = np.zeros(n, dtype=np.float64)
        sums = np.zeros(n, dtype=LEAFSUM_DTYPE)
        legal = np.zeros((n, 4), dtype=np.uint64)
        self._ck(self.lib.mctsb_wait_x(self.handle, score.ctypes.data, sums.ctypes.data, legal.ctypes.data))
        return score, sums, legal

    # -- multi-GPU exchange (NCCL behind the C ABI) -------------------------------------------------
    def comm_init_rank(self, unique_id, world, rank):
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.mctsb_comm_init_rank(self.handle, buf, world, rank))

    def comm_info(self):
        w, r = C.c_int(0), C.c_int(0)
        self._ck(self.lib.mctsb_comm_info(self.handle, C.byref(w), C.byref(r)))
        return w.value, r.value

    def allreduce_root_stats(self, vec):
        v = np.ascontiguousarray(vec, dtype=np.float64).copy()
        self._ck(self.lib.mctsb_allreduce_root_stats(self.handle, v.ctypes.data, v.size))
        return v

    def allreduce_leaf_sums(self, sums):
        s = np.ascontiguousarray(sums, dtype=LEAFSUM_DTYPE).copy()
        self._ck(self.lib.mctsb_allreduce_leaf_sums(self.handle, s.ctypes.data, s.size))
        return s

    def rollout_outcomes(self, game, policy, states, rollouts_per_leaf, first_rollout_id=0, finals=True):
        a = self._states(game, states)
        n = a.size * rollouts_per_leaf
        out = np.zeros(n, dtype=OUTCOME_DTYPE)
        fin = np.zeros(n, dtype=QSTATE_DTYPE) if (finals and game == GAME_QUORIDOR) else None
        self._ck(self.lib.mctsb_rollout_outcomes(self.handle, game, policy, a.ctypes.data, a.size, rollouts_per_leaf,
                                                 first_rollout_id, out.ctypes.data, fin.ctypes.data if fin is not None else None))
        return out, fin

    def rollout_replay(self, game, policy, states, streams, finals=True):
        a = self._states(game, states)
        st = np.ascontiguousarray(streams, dtype=np.uint32)
        assert st.ndim == 2 and st.shape[0] == a.size
        out = np.zeros(a.size, dtype=OUTCOME_DTYPE)
        fin = np.zeros(a.size, dtype=QSTATE_DTYPE) if (finals and game == GAME_QUORIDOR) else None
        self._ck(self.lib.mctsb_rollout_replay(self.handle, game, policy, a.ctypes.data, a.size, st.ctypes.data,
                                               st.shape[1], st.shape[1], out.ctypes.data,
                                               fin.ctypes.data if fin is not None else None))
        return out, fin

    # -- rule primitives ------------------------------------------------------------------------
    def q_movegen(self, states):
        a = self._states(GAME_QUORIDOR, states)
        legal = np.zeros((a.size, 4), dtype=np.uint64)
        info = np.zeros(a.size, dtype=QINFO_DTYPE)
        ev = np.zeros(a.size, dtype=np.float64)
        self._ck(self.lib.mctsb_q_movegen(self.handle, a.ctypes.data, a.size, legal.ctypes.data, info.ctypes.data, ev.ctypes.data))
        return legal, info, ev

    def q_goodmoves(self, states):
        """generate_good_moves of every state -> list of int arrays (canonical move ids in queue order)"""
        a = self._states(GAME_QUORIDOR, states)
        moves = np.zeros((a.size, Q_GOOD_MAX), dtype=np.uint8)
        counts = np.zeros(a.size, dtype=np.int32)
        self._ck(self.lib.mctsb_q_goodmoves(self.handle, a.ctypes.data, a.size, moves.ctypes.data, counts.ctypes.data))
        return [moves[i, :counts[i]].astype(np.int32) for i in range(a.size)]

    def q_play(self, states, move_ids):
        a = self._states(GAME_QUORIDOR, states)
        ids = np.ascontiguousarray(move_ids, dtype=np.int32)
        out = np.zeros(a.size, dtype=QSTATE_DTYPE)
        ok = np.zeros(a.size, dtype=np.uint8)
        self._ck(self.lib.mctsb_q_play(self.handle, a.ctypes.data, ids.ctypes.data, a.size, out.ctypes.data, ok.ctypes.data))
        return out, ok

    # -- device-resident throughput path ------------------------------------------------------------
    def upload_states(self, game, states):
        a = self._states(game, states)
        self._ck(self.lib.mctsb_upload_states(self.handle, game, a.ctypes.data, a.size))
        self._pending_n = a.size

    def run_resident(self, game, policy, rollouts_per_leaf, first_rollout_id=0, timed=True):
        ms = C.c_float(0)
        self._ck(self.lib.mctsb_run_resident(self.handle, game, policy, rollouts_per_leaf, first_rollout_id,
                                             C.byref(ms) if timed else None))
        return ms.value

    def int_issue_peak(self):
        ops = C.c_double(0)
        ms = C.c_float(0)
        self._ck(self.lib.mctsb_int_issue_peak(self.handle, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value

    def lop3_peak(self):
        ops = C.c_double(0)
        ms = C.c_float(0)
        self._ck(self.lib.mctsb_lop3_peak(self.handle, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value

    def flush_l2(self):
        self._ck(self.lib.mctsb_flush_l2(self.handle))

    def counters(self):
        c = np.zeros((), dtype=COUNTERS_DTYPE)
        self._ck(self.lib.mctsb_get_counters(self.handle, c.ctypes.data))
        return {k: int(c[k]) for k in COUNTERS_DTYPE.names}


def comm_unique_id():
    """128 opaque bytes identifying a new communicator: made by one rank, shipped to the others by any means."""
    L = load_library()
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    st = L.mctsb_comm_unique_id(buf)
    if st != 0:
        raise BackendError(st, L.mctsb_strerror(st).decode())
    return bytes(buf)


def comm_init_all(backends):
    """One process, one RolloutBackend per device: a communicator over all of them."""
    L = backends[0].lib
    arr = (C.c_void_p * len(backends))(*[b.handle for b in backends])
    st = L.mctsb_comm_init_all(arr, len(backends))
    if st != 0:
        raise BackendError(st, L.mctsb_strerror(st).decode() + " / " + L.mctsb_last_cuda_error(backends[0].handle).decode())


def unpack_legal(mask4):
    """(n,4) uint64 -> (n,209) uint8"""
    m = np.asarray(mask4, dtype=np.uint64).reshape(-1, 4)
    bits = np.unpackbits(m.view(np.uint8).reshape(m.shape[0], 32), axis=1, bitorder="little")
    return bits[:, :NUM_MOVES]


class DeviceTree:
    """The search tree resident in HBM (mctsb_tree_*): Selection, Expansion, Simulation and Backpropagation of a
    Quoridor tree on the device, `leaves_per_flush` leaves x `rollouts_per_leaf` playouts per flush."""

    def __init__(self, backend, root, rollouts_per_leaf, leaves_per_flush, policy=POLICY_REF, max_nodes=1 << 22, c=1.41):
        self.be, self.lib = backend, backend.lib
        self.R, self.K = rollouts_per_leaf, leaves_per_flush
        r = np.ascontiguousarray(root, dtype=QSTATE_DTYPE)
        self.h = C.c_void_p()
        backend._ck(self.lib.mctsb_tree_create(backend.handle, C.byref(self.h), r.ctypes.data, policy, rollouts_per_leaf,
                                               leaves_per_flush, max_nodes, c))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mctsb_tree_destroy(self.h)
            self.h = None

    def set_good_moves_only(self, on=True):
        """before the first grow: expand generate_good_moves (Quoridor.cpp:518-592) instead of all legal moves"""
        self.be._ck(self.lib.mctsb_tree_set_good_moves(self.h, int(bool(on))))

    def set_overlap(self, on=True):
        """select flush f+1 while the playouts of flush f run (one flush in flight, as MCTS_batching::overlap)"""
        self.be._ck(self.lib.mctsb_tree_set_overlap(self.h, int(bool(on))))

    def grow(self, iterations, first_rollout_id, leaves_per_flush=None):
        self.be._ck(self.lib.mctsb_tree_grow(self.h, iterations, leaves_per_flush or self.K, first_rollout_id))

    def info(self):
        o = np.zeros((), dtype=TREE_INFO_DTYPE)
        self.be._ck(self.lib.mctsb_tree_get_info(self.h, o.ctypes.data))
        return {k: o[k].item() for k in TREE_INFO_DTYPE.names}

    def root_children(self, max_children=256):
        mv, sims, sc = np.zeros(max_children, np.int32), np.zeros(max_children, np.uint64), np.zeros(max_children, np.float64)
        n = C.c_int(0)
        self.be._ck(self.lib.mctsb_tree_root_children(self.h, mv.ctypes.data, sims.ctypes.data, sc.ctypes.data, max_children, C.byref(n)))
        k = min(n.value, max_children)
        return mv[:k].copy(), sims[:k].copy(), sc[:k].copy()

    def merge_root(self):
        """root-parallel trees: root-child statistics summed over the GPUs the backend is joined with (NCCL behind the C ABI)"""
        self.be._ck(self.lib.mctsb_tree_merge_root(self.h))

    def best_move(self):
        mv = C.c_int32(-1)
        self.be._ck(self.lib.mctsb_tree_best_move(self.h, C.byref(mv)))
        return mv.value

    def advance(self, move_id):
        self.be._ck(self.lib.mctsb_tree_advance(self.h, int(move_id)))

    def root_state(self):
        s = np.zeros((), dtype=QSTATE_DTYPE)
        self.be._ck(self.lib.mctsb_tree_root_state(self.h, s.ctypes.data))
        return s

    def dump(self, cap=1 << 20):
        a = {"depth": np.zeros(cap, np.int32), "move": np.zeros(cap, np.int32), "sims": np.zeros(cap, np.uint64),
             "score": np.zeros(cap, np.float64), "size": np.zeros(cap, np.uint32), "n_children": np.zeros(cap, np.uint32)}
        n = C.c_uint64(0)
        self.be._ck(self.lib.mctsb_tree_dump(self.h, cap, *[a[k].ctypes.data for k in ("depth", "move", "sims", "score", "size", "n_children")], C.byref(n)))
        assert n.value <= cap, "tree larger than the dump buffer"
        return {k: v[:n.value].copy() for k, v in a.items()}
