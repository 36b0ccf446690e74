"""CPU, world_size 2 over gloo: the N>1 host-side logic (id sharding, exact integer merge of leaf
statistics, root-parallel merge of root-child statistics)."""
import json
import os
import subprocess
import sys

import numpy as np

import montecarlotreesearch_b200 as m
from montecarlotreesearch_b200 import multi
from tests.util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [multi.shard_range(total, world, r) for r in range(world)]
            assert sum(c for _, c in blocks) == total
            pos = 0
            for s, c in blocks:
                assert s == pos
                pos += c
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_two_ranks_over_gloo(tmp_path, port):
    out = tmp_path / "gloo.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "_gloo_worker.py"), str(out)]
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    got = json.load(open(out))
    assert got["world"] == 2
    # (i) the sharded, all-reduced accumulators equal one process playing the whole id range
    corpus = golden("corpus_q.npy")
    leaves, R, seed, first = corpus[[7, 4100, 4400]], 40, 77, 5000
    for i in range(3):
        o, _, _ = port.q_rollout_philox(leaves[i], seed, first + i * R, R)
        acc, as_double = port.leaf_sum(o["score"])
        assert (got["lo"][i], got["hi"][i], got["n"][i]) == (int(acc["lo"]), int(acc["hi"]), R)
        assert multi.leaf_sum_to_double(got["lo"][i], got["hi"][i]) == as_double
    # (ii) both ranks hold the same merged root statistics and choose the same move
    assert got["bests"][0] == got["bests"][1]
    v = np.array(got["merged_vec"])
    assert v[0::2].sum() == got["root_sims"] == (150 + 200) * 4
    assert (v >= np.array(got["mine"])).all()
