#!/usr/bin/env python
"""selfplay.py — BASELINE.json configs[3] and [4]: Quoridor full-game MCTS self-play.

  1 GPU  (configs[3]): host tree (C++, montecarlotreesearch_b200/host) + B200 rollout backend, every tree
          flush is K leaves x R rollouts (default 16 x 4096 = 65 536 playouts) in one kernel launch.
  N GPUs (configs[4], launch under torchrun): root-parallel — one tree per rank, same root, different
          Philox keys; ONE all-reduce of the root-child statistics per move (2 doubles per root action,
          slot = position in the root's actions_to_try() order) done by the C++ host through the C ABI
          (MCTS_tree::merge_root_stats -> mctsb_allreduce_root_stats, NCCL over NVLink), after which every rank
          applies the reference's final choice rule (win rate of the side to move, first maximum,
          mcts.cpp:104-129 with c = 0, :268-270) to identical numbers and plays the same move.
          torch.distributed only ships the 128-byte communicator id and provides the start barrier.

Prints one JSON line (rank 0): rollouts/s over the whole game, flush latency, the host tree's share of the
search's wall time (selection + expansion + backpropagation, i.e. everything outside backend calls), game
length, winner and the move list.  --agreement replays the game on ONE GPU with rank 0's seed (same positions,
tree reuse as in the game) and reports how often its best child is the move the N-GPU search chose.
The reference's loop of the same shape: examples/Quoridor/main.cpp:129-169.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class DeviceTreeSession:
    """The tree resident in HBM (mctsb_tree_*) behind the few calls play_game makes of a host-tree Session."""

    def __init__(self, m, args, seed, device):
        self.m = m
        self.be = m.RolloutBackend(device, seed)
        self.tree = m.DeviceTree(self.be, m.quoridor_opening(), args.rollouts_per_leaf, args.leaves_per_flush, max_nodes=args.device_tree_nodes)
        if args.overlap:
            self.tree.set_overlap(True)
        self.R, self.next_id = args.rollouts_per_leaf, 0
        self.wall, self.gpu_ms = 0.0, 0.0
        self.phase_ms = {"select": 0.0, "rollout": 0.0, "backprop": 0.0}

    def grow(self, iters):
        t0 = time.perf_counter()
        self.tree.grow(iters, self.next_id)
        self.wall += time.perf_counter() - t0
        self.next_id += iters * self.R
        i = self.tree.info()
        self.gpu_ms += i["ms_total"]
        for k in self.phase_ms:                      # CUDA-event time of the three phases (measured in synchronous mode only)
            self.phase_ms[k] += i["ms_" + k]

    def best_move(self):
        return self.tree.best_move()

    # root-parallel trees: the same calls a host-tree Session offers
    def comm_init_rank(self, unique_id, world, rank):
        self.be.comm_init_rank(unique_id, world, rank)

    def merge_root(self):
        self.tree.merge_root()

    def allreduce(self, vec):
        return self.be.allreduce_root_stats(vec)

    def advance(self, mv):
        self.tree.advance(mv)
        s = self.tree.root_state()
        return int(s["wx"]) == 8 or int(s["bx"]) == 0

    def counters(self):
        i = self.tree.info()
        return i["flushes"], i["rollouts"]

    def times(self):
        """(seconds inside grow, seconds of them the GPU was busy with the tree's launches)"""
        return self.wall, 1e-3 * self.gpu_ms

    def current_state(self):
        return self.tree.root_state()

    def info(self):
        return self.tree.info()

    def close(self):
        self.tree.close(); self.be.close()


def play_game(tree, args, dist_ready, follow=None):
    """One game on `tree`.  follow = list of move ids to play instead of the tree's own choice (agreement rerun).
    Returns (moves, own_choices, terminal, seconds in grow, seconds in merge)."""
    moves, own, t_grow, t_merge = [], [], 0.0, 0.0
    terminal = False
    while not terminal and len(moves) < args.max_moves and (follow is None or len(moves) < len(follow)):
        a = time.perf_counter()
        tree.grow(args.iters)
        b = time.perf_counter()
        t_grow += b - a
        if dist_ready:
            tree.merge_root()                          # C++ host -> C ABI -> NCCL; blocks until every rank arrived
        c = time.perf_counter()
        t_merge += c - b
        mv = tree.best_move()
        if dist_ready and args.check_agreement_across_ranks:
            # every rank must have picked the same move from the merged statistics: sum(mv) == world * mv and
            # sum(mv^2) == world * mv^2 only when all values are equal
            w = tree.allreduce(np.array([1.0, float(mv), float(mv) * mv]))
            assert w[1] == w[0] * mv and w[2] == w[0] * mv * mv, "ranks disagree on the move after the all-reduce"
        if mv < 0:
            break
        own.append(mv)
        if follow is not None:
            mv = follow[len(moves)]
        moves.append(mv)
        terminal = tree.advance(mv)
    return moves, own, terminal, t_grow, t_merge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=256, help="tree iterations (leaf expansions) per move")
    ap.add_argument("--rollouts-per-leaf", type=int, default=4096)
    ap.add_argument("--leaves-per-flush", type=int, default=16)
    ap.add_argument("--max-moves", type=int, default=400)
    ap.add_argument("--overlap", type=int, nargs="?", const=1, default=0,
                    help="flushes in flight: 1 = select the next batch of leaves while the previous flush runs on the GPU, 2 = two flushes queued")
    ap.add_argument("--gpu-expand", type=int, default=1,
                    help="1 = every flush also returns its leaves' legal-move masks from the move-generation kernel (no host movegen); 0 = actions_to_try() on the CPU")
    ap.add_argument("--agreement", action="store_true", help="replay the game on one GPU at rank 0's seed and report per-move best-child agreement")
    ap.add_argument("--check-agreement-across-ranks", type=int, default=1)
    ap.add_argument("--device-tree", type=int, default=0,
                    help="1 = the tree itself lives on the GPU (mctsb_tree_*: Selection, Expansion and Backpropagation as kernels); under torchrun one tree per GPU, merged per move by mctsb_tree_merge_root")
    ap.add_argument("--device-tree-nodes", type=int, default=1 << 26, help="node pool of the device tree (94 bytes per node)")
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x2026101900000001)
    args = ap.parse_args()

    # stdout carries exactly one JSON line: native libraries' banners (NCCL version) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import montecarlotreesearch_b200 as m
    from montecarlotreesearch_b200.hosttree import Session, load

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    host = load()

    def new_tree(seed, device):
        if args.device_tree:
            return DeviceTreeSession(m, args, seed, device)
        t = Session(host, m.GAME_QUORIDOR, m.quoridor_opening(), rollouts_per_leaf=args.rollouts_per_leaf,
                    leaves_per_flush=args.leaves_per_flush, device=device, seed=seed)
        if args.overlap:
            t.set_overlap(args.overlap)
        t.set_gpu_expand(bool(args.gpu_expand))
        return t

    # different Philox key per rank (root-parallel trees must be independent), same root
    tree = new_tree(args.seed + rank, local_rank)
    if dist is not None:
        # the communicator of the C ABI: rank 0 makes the id, torch.distributed ships it (plumbing only)
        import torch
        uid = torch.zeros(m.COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(m.comm_unique_id()), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        tree.comm_init_rank(uid.cpu().numpy().tobytes(), world, rank)
        tree.allreduce(np.zeros(8))                    # bring the communicator up before the clock starts
        dist.barrier()
    t0 = time.perf_counter()
    moves, _, terminal, t_grow, t_merge = play_game(tree, args, dist is not None)
    wall = time.perf_counter() - t0
    flushes, rollouts = tree.counters()
    sec_total, sec_backend = tree.times()
    device_tree_info = None
    if args.device_tree:
        i = tree.info()
        device_tree_info = {"ms_per_flush_by_phase": {k: v / max(1, flushes) for k, v in tree.phase_ms.items()} if not args.overlap else None,
                            "nodes_used": i["nodes_used"], "deepest_level": i["deepest_level"], "exact_evals": i["exact_evals"], "flags": i["flags"]}
    final = tree.current_state()
    rollouts_all = float(tree.allreduce(np.array([float(rollouts)]))[0]) if dist is not None else float(rollouts)
    tree.close()

    agreement = None
    if args.agreement and rank == 0:
        solo = new_tree(args.seed, local_rank)
        _, own, _, _, _ = play_game(solo, args, False, follow=moves)
        solo.close()
        same = sum(1 for a, b in zip(own, moves) if a == b)
        agreement = {"moves_compared": len(own), "same_best_child": same, "fraction": same / max(1, len(own)),
                     "what": "1-GPU search at rank 0's seed over the game's positions (tree reused along the game's moves) vs the move the %d-GPU search played" % world}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "config": "configs[%d]: Quoridor self-play, %d iterations/move, flush = %d leaves x %d rollouts = %d playouts"
                      % (4 if world > 1 else 3, args.iters, args.leaves_per_flush, args.rollouts_per_leaf,
                         args.leaves_per_flush * args.rollouts_per_leaf),
            "n_gpus": world, "moves": len(moves), "terminal": bool(terminal),
            "winner": "White" if int(final["wx"]) == 8 else "Black" if int(final["bx"]) == 0 else None,
            "rollouts_total": rollouts_all, "rollouts_per_sec": rollouts_all / wall, "wall_s": wall,
            "leaves_per_flush": args.leaves_per_flush, "rollouts_per_leaf": args.rollouts_per_leaf,
            "flushes_in_flight": args.overlap, "gpu_expand": bool(args.gpu_expand),
            "tree": "device (mctsb_tree_*: host share = time of grow the GPU was NOT busy with the tree's launches)" if args.device_tree else "host (MCTS_tree)",
            "flushes_rank0": flushes, "ms_per_flush": 1e3 * t_grow / max(1, flushes),
            "host_tree_share_of_search_time": (sec_total - sec_backend) / sec_total if sec_total > 0 else None,
            "search_seconds": sec_total, "backend_seconds": sec_backend,
            "merge_ms_per_move": 1e3 * t_merge / max(1, len(moves)) if dist is not None else 0.0,
            "merge": "MCTS_tree::merge_root_stats -> mctsb_allreduce_root_stats (NCCL behind the C ABI); includes waiting for the slowest rank" if dist is not None else None,
            "device_tree": device_tree_info,
            "agreement_with_1gpu_rerun": agreement,
            "move_ids": moves,
        }) + "\n").encode())


if __name__ == "__main__":
    main()
