"""Worker of tests/test_multi_gloo.py: launched twice by torch.distributed.run with the gloo backend.
Each rank plays its shard of one flush with the oracle standing in for the GPU (CPU test infrastructure),
merges with the product's multi-GPU helpers, and rank 0 writes what it saw."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

import montecarlotreesearch_b200 as m  # noqa: E402
from montecarlotreesearch_b200 import multi  # noqa: E402
from oracle.portlib import PortLib  # noqa: E402
from tests.hostlib import Session, load  # noqa: E402
from tests.test_host_tree import oracle_backend  # noqa: E402
from tests.util import golden  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    port = PortLib()
    corpus = golden("corpus_q.npy")

    # (i) throughput / leaf-parallel sharding: 3 leaves x 40 rollouts, ids cut into contiguous blocks
    leaves, R, seed, first = corpus[[7, 4100, 4400]], 40, 77, 5000
    total = len(leaves) * R
    start, count = multi.shard_range(total, world, rank)
    sums = np.zeros(len(leaves), dtype=m.LEAFSUM_DTYPE)
    for g in range(start, start + count):
        leaf = g // R
        o, _, _ = port.q_rollout_philox(leaves[leaf], seed, first + g, 1)
        acc, _ = port.leaf_sum(o["score"])
        sums["lo"][leaf] += acc["lo"]; sums["hi"][leaf] += acc["hi"]; sums["n"][leaf] += acc["n"]
    pending = multi.allreduce_leaf_sums_async(sums, dist)        # in flight while the rank does other work
    merged_sync = multi.allreduce_leaf_sums(sums, dist)
    merged = pending.result()
    assert merged.tobytes() == merged_sync.tobytes()

    # (ii) root-parallel trees: same root, different Philox keys per rank, one all-reduce of the root statistics
    host = load()
    t = Session(host, m.GAME_TICTACTOE, np.zeros((), dtype=m.TSTATE_DTYPE), rollouts_per_leaf=4,
                callback=oracle_backend(port, seed=1000 + rank))
    t.grow(150 + 50 * rank)
    mine = t.export_root()
    allv = multi.allreduce_root_stats(mine, dist)
    t.import_root(allv)
    best = t.best_move()
    bests = [None] * world
    dist.all_gather_object(bests, best)
    stats = t.root_stats()
    t.close()

    if rank == 0:
        json.dump({"world": world, "lo": [int(x) for x in merged["lo"]], "hi": [int(x) for x in merged["hi"]],
                   "n": [int(x) for x in merged["n"]], "bests": bests, "root_sims": stats["sims"],
                   "merged_vec": allv.tolist(), "mine": mine.tolist()}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
