"""ctypes bindings of the C++ host side (montecarlotreesearch_b200/libmcts_host.so, host/capi.cpp): the host
tree (MCTS_tree over the B200 rollout backend) as the bench / self-play scripts and the tests drive it."""
import ctypes as C
import os

import numpy as np

import montecarlotreesearch_b200 as m

CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(C.c_double), C.c_void_p)


def load():
    m.build()
    L = C.CDLL(os.path.join(m.PKG_DIR, "libmcts_host.so"))
    vp = C.c_void_p
    L.mcts_host_q_actions.argtypes = [vp, vp, C.c_int]
    L.mcts_host_q_good_actions.argtypes = [vp, vp, C.c_int]
    L.mcts_host_set_good_moves_only.argtypes = [vp, C.c_int]
    L.mcts_host_q_next.argtypes = [vp, C.c_int, vp]
    L.mcts_host_q_info.argtypes = [vp, vp, vp]
    L.mcts_host_q_sprint.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.mcts_host_session_create.restype = vp
    L.mcts_host_session_create.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, C.c_char_p, C.c_int]
    L.mcts_host_session_destroy.argtypes = [vp]
    L.mcts_host_set_overlap.argtypes = [vp, C.c_int]
    L.mcts_host_grow.argtypes = [vp, C.c_int, C.c_double, C.c_char_p, C.c_int]
    L.mcts_host_root_stats.argtypes = [vp, vp, vp, C.c_int]
    L.mcts_host_best_move.argtypes = [vp]
    L.mcts_host_dump_tree.restype = C.c_uint64
    L.mcts_host_dump_tree.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp, vp]
    L.mcts_host_advance.argtypes = [vp, C.c_int]
    L.mcts_host_current_state.argtypes = [vp, vp]
    L.mcts_host_counters.argtypes = [vp, vp, vp]
    L.mcts_host_peek_rollout_id.restype = C.c_uint64
    L.mcts_host_peek_rollout_id.argtypes = []
    L.mcts_host_set_next_rollout_id.argtypes = [C.c_uint64]
    L.mcts_host_export_root.argtypes = [vp, vp, C.c_int]
    L.mcts_host_import_root.argtypes = [vp, vp, C.c_int]
    L.mcts_host_set_gpu_expand.argtypes = [vp, C.c_int]
    L.mcts_host_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mcts_host_comm_unique_id.argtypes = [vp]
    L.mcts_host_comm_init_rank.argtypes = [vp, vp, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.mcts_host_merge_root.argtypes = [vp, C.c_char_p, C.c_int]
    L.mcts_host_allreduce.argtypes = [vp, vp, C.c_int, C.c_char_p, C.c_int]
    L.mcts_host_dtree_create.restype = vp
    L.mcts_host_dtree_create.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_char_p, C.c_int]
    L.mcts_host_dtree_destroy.argtypes = [vp]
    L.mcts_host_dtree_grow.argtypes = [vp, C.c_int, C.c_double, C.c_char_p, C.c_int]
    L.mcts_host_dtree_best_move.argtypes = [vp]
    L.mcts_host_dtree_advance.argtypes = [vp, C.c_int]
    L.mcts_host_dtree_size.argtypes = [vp]
    L.mcts_host_dtree_size.restype = C.c_uint
    L.mcts_host_dtree_print_stats.argtypes = [vp]
    return L


class Session:
    """One MCTS_tree.  callback=None -> the CUDA backend (needs a GPU)."""

    def __init__(self, lib, game, root, rollouts_per_leaf=4, leaves_per_flush=1, policy=0, device=0,
                 seed=m.DEFAULT_SEED, callback=None):
        self.lib, self.game = lib, game
        dt = m.QSTATE_DTYPE if game == m.GAME_QUORIDOR else m.TSTATE_DTYPE
        r = np.ascontiguousarray(root, dtype=dt)
        self._cb = CB(callback) if callback else None
        err = C.create_string_buffer(512)
        self.h = lib.mcts_host_session_create(game, r.ctypes.data, policy, device, seed, rollouts_per_leaf, leaves_per_flush,
                                              C.cast(self._cb, C.c_void_p) if self._cb else None, None, err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())

    def close(self):
        if self.h:
            self.lib.mcts_host_session_destroy(self.h)
            self.h = None

    def set_overlap(self, on=True):
        """on = False / 0: synchronous flushes; True / 1: one flush in flight; 2: two flushes queued on the GPU"""
        self.lib.mcts_host_set_overlap(self.h, int(on))

    def set_good_moves_only(self, on=True):
        """Quoridor, before the first grow: expand generate_good_moves() (Quoridor.cpp:518-592) instead of all legal moves"""
        assert self.lib.mcts_host_set_good_moves_only(self.h, int(bool(on))) == 0

    def set_gpu_expand(self, on=True):
        """every flush also returns its leaves' legal-move masks from the move-generation kernel (MCTS_batching::gpu_expand)"""
        self.lib.mcts_host_set_gpu_expand(self.h, int(bool(on)))

    def times(self):
        """(seconds inside grow_tree so far, seconds of them inside backend calls)"""
        a, b = C.c_double(0), C.c_double(0)
        self.lib.mcts_host_times(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    # -- multi-GPU exchange (NCCL behind the C ABI) --
    def comm_init_rank(self, unique_id, world, rank):
        err = C.create_string_buffer(512)
        buf = (C.c_ubyte * m.COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        if self.lib.mcts_host_comm_init_rank(self.h, buf, world, rank, err, 512) != 0:
            raise RuntimeError(err.value.decode())

    def merge_root(self):
        """MCTS_tree::merge_root_stats: root-child statistics summed over the ranks, in place"""
        err = C.create_string_buffer(512)
        if self.lib.mcts_host_merge_root(self.h, err, 512) != 0:
            raise RuntimeError(err.value.decode())

    def allreduce(self, vec):
        v = np.ascontiguousarray(vec, dtype=np.float64).copy()
        err = C.create_string_buffer(512)
        if self.lib.mcts_host_allreduce(self.h, v.ctypes.data, v.size, err, 512) != 0:
            raise RuntimeError(err.value.decode())
        return v

    def grow(self, iters, seconds=1e9):
        err = C.create_string_buffer(512)
        if self.lib.mcts_host_grow(self.h, iters, seconds, err, 512) != 0:
            raise RuntimeError(err.value.decode())

    def root_stats(self, max_children=256):
        out4 = np.zeros(4)
        ch = np.zeros((max_children, 3))
        n = self.lib.mcts_host_root_stats(self.h, out4.ctypes.data, ch.ctypes.data, max_children)
        return {"sims": out4[0], "score": out4[1], "size": out4[2], "children": ch[:n].copy()}

    def best_move(self):
        return self.lib.mcts_host_best_move(self.h)

    def dump(self, cap=1 << 20):
        """the whole tree in pre-order: dict of arrays depth, move, sims, score, size, n_children (one row per node)"""
        a = {"depth": np.zeros(cap, np.int32), "move": np.zeros(cap, np.int32), "sims": np.zeros(cap, np.uint64),
             "score": np.zeros(cap, np.float64), "size": np.zeros(cap, np.uint32), "n_children": np.zeros(cap, np.uint32)}
        n = self.lib.mcts_host_dump_tree(self.h, cap, *[a[k].ctypes.data for k in ("depth", "move", "sims", "score", "size", "n_children")])
        assert n <= cap, "tree larger than the dump buffer"
        return {k: v[:n].copy() for k, v in a.items()}

    def advance(self, move_id):
        return bool(self.lib.mcts_host_advance(self.h, move_id))

    def current_state(self):
        dt = m.QSTATE_DTYPE if self.game == m.GAME_QUORIDOR else m.TSTATE_DTYPE
        s = np.zeros((), dtype=dt)
        self.lib.mcts_host_current_state(self.h, s.ctypes.data)
        return s

    def counters(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.lib.mcts_host_counters(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def export_root(self):
        buf = np.zeros(600)
        n = self.lib.mcts_host_export_root(self.h, buf.ctypes.data, 600)
        return buf[:n].copy()

    def import_root(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert self.lib.mcts_host_import_root(self.h, v.ctypes.data, v.size) == 0
