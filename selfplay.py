#!/usr/bin/env python
"""selfplay.py — BASELINE.json configs[3] and [4]: Quoridor full-game MCTS self-play.

  1 GPU  (configs[3]): host tree (C++, montecarlotreesearch_b200/host) + B200 rollout backend, every tree
          flush is K leaves x R rollouts = 65 536 playouts in one kernel launch.
  N GPUs (configs[4], launch under torchrun): root-parallel — one tree per rank, same root, different
          Philox keys; ONE NCCL all-reduce of the root-child statistics per move (2 doubles per root
          action, slot = position in the root's actions_to_try() order), after which every rank applies
          the reference's final choice rule (win rate of the side to move, first maximum,
          mcts.cpp:104-129 with c = 0) to identical numbers and plays the same move.

Prints one JSON line (rank 0): rollouts/s over the whole game, flush latency, host-tree share of wall
time, game length and the move list.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=256, help="tree iterations (leaf expansions) per move")
    ap.add_argument("--rollouts-per-leaf", type=int, default=4096)
    ap.add_argument("--leaves-per-flush", type=int, default=16)
    ap.add_argument("--max-moves", type=int, default=120)
    ap.add_argument("--overlap", type=int, nargs="?", const=1, default=0,
                    help="flushes in flight: 1 = select the next batch of leaves while the previous flush runs on the GPU, 2 = two flushes queued")
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x2026101900000001)
    args = ap.parse_args()

    # stdout carries exactly one JSON line: native libraries' banners (NCCL version) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import montecarlotreesearch_b200 as m
    from montecarlotreesearch_b200 import multi
    from tests.hostlib import Session, load

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    import torch
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    host = load()
    # different Philox key per rank (root-parallel trees must be independent), same root
    tree = Session(host, m.GAME_QUORIDOR, m.quoridor_opening(), rollouts_per_leaf=args.rollouts_per_leaf,
                   leaves_per_flush=args.leaves_per_flush, device=local_rank, seed=args.seed + rank)
    if args.overlap:
        tree.set_overlap(args.overlap)
    moves, t_grow, t_merge, t_skew = [], 0.0, 0.0, 0.0
    if dist is not None:                               # bring the communicator up before the clock starts
        dist.all_reduce(torch.zeros(8, dtype=torch.float64, device="cuda"))
        torch.cuda.synchronize()
        dist.barrier()
    t0 = time.perf_counter()
    terminal = False
    while not terminal and len(moves) < args.max_moves:
        a = time.perf_counter()
        tree.grow(args.iters)
        b = time.perf_counter()
        t_grow += b - a
        if dist is not None:
            dist.barrier()                             # waiting for the slowest tree of this move (skew), kept apart
            b2 = time.perf_counter()
            t_skew += b2 - b
            merged = multi.allreduce_root_stats(tree.export_root(), dist, device="cuda")
            tree.import_root(merged)
            b = b2
        c = time.perf_counter()
        mv = tree.best_move()
        if dist is not None:
            # every rank must have picked the same move from the merged statistics: max(mv) == -max(-mv)
            chk = torch.tensor([mv, -mv], dtype=torch.int64, device="cuda")
            dist.all_reduce(chk, op=dist.ReduceOp.MAX)
            assert int(chk[0]) == -int(chk[1]) == mv, "ranks disagree on the move after the all-reduce"
        if mv < 0:
            break
        moves.append(mv)
        terminal = tree.advance(mv)
        t_merge += c - b
    wall = time.perf_counter() - t0
    flushes, rollouts = tree.counters()
    final = tree.current_state()
    tree.close()
    if dist is not None:
        tot = torch.tensor([float(rollouts)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        rollouts_all = float(tot[0])
        dist.barrier()
        dist.destroy_process_group()
    else:
        rollouts_all = float(rollouts)
    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "config": "configs[%d]: Quoridor self-play, %d iterations/move, flush = %d leaves x %d rollouts = %d playouts"
                      % (4 if world > 1 else 3, args.iters, args.leaves_per_flush, args.rollouts_per_leaf,
                         args.leaves_per_flush * args.rollouts_per_leaf),
            "n_gpus": world, "moves": len(moves), "terminal": bool(terminal),
            "winner": "White" if int(final["wx"]) == 8 else "Black" if int(final["bx"]) == 0 else None,
            "rollouts_total": rollouts_all, "rollouts_per_sec": rollouts_all / wall, "wall_s": wall,
            "flushes_rank0": flushes, "ms_per_flush": 1e3 * t_grow / max(1, flushes),
            "allreduce_ms_per_move": 1e3 * t_merge / max(1, len(moves)),
            "wait_for_slowest_rank_ms_per_move": 1e3 * t_skew / max(1, len(moves)),
            "move_ids": moves,
        }) + "\n").encode())


if __name__ == "__main__":
    main()
