"""Host tree vs device-resident tree on one GPU: the same search (opening position, REF policy, K leaves x R playouts
per flush, K*R = 65 536 unless told otherwise), wall time per flush and where it goes.
    python tools/tree_bench.py [--shapes 16x4096,256x256,4096x16] [--flushes 16] [--policy 0] > out.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montecarlotreesearch_b200 as m  # noqa: E402
from montecarlotreesearch_b200.hosttree import Session, load  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="16x4096,256x256,4096x16")
    ap.add_argument("--flushes", type=int, default=16)
    ap.add_argument("--policy", type=int, default=0)
    ap.add_argument("--no-host", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="skip the one-flush-in-flight run of the device tree (profiling runs: one stream)")
    a = ap.parse_args()
    host = load()
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    root = m.quoridor_opening()
    out = []
    for shape in a.shapes.split(","):
        K, R = (int(x) for x in shape.split("x"))
        iters = K * a.flushes
        row = {"leaves_per_flush": K, "rollouts_per_leaf": R, "flushes": a.flushes, "policy": "REF" if a.policy == 0 else "UNI"}
        # device-resident tree: one warm-up grow (first-touch of the pool, kernel load), then the timed search from scratch
        for rep in range(0 if a.no_overlap else 2):
            dt = m.DeviceTree(be, root, R, K, policy=a.policy)
            dt.set_overlap(True)
            t0 = time.perf_counter()
            dt.grow(iters, 1 << 42)
            wall = time.perf_counter() - t0
            info = dt.info()
            dt.close()
        if not a.no_overlap:
            row["device_tree_1_in_flight"] = {"wall_ms_per_flush": 1e3 * wall / a.flushes, "rollouts_per_sec": iters * R / wall, "iterations_per_sec": iters / wall,
                                          "nodes_used": info["nodes_used"]}
        for rep in range(2):
            dt = m.DeviceTree(be, root, R, K, policy=a.policy)
            t0 = time.perf_counter()
            dt.grow(iters, 1 << 42)
            wall = time.perf_counter() - t0
            info = dt.info()
            dt.close()
        row["device_tree"] = {"wall_ms_per_flush": 1e3 * wall / a.flushes, "rollouts_per_sec": iters * R / wall, "iterations_per_sec": iters / wall,
                              "ms_select_per_flush": info["ms_select"] / a.flushes, "ms_rollout_per_flush": info["ms_rollout"] / a.flushes,
                              "ms_backprop_per_flush": info["ms_backprop"] / a.flushes, "nodes_used": info["nodes_used"],
                              "deepest_level": info["deepest_level"], "exact_evals": info["exact_evals"],
                              "host_share": max(0.0, 1.0 - (info["ms_select"] + info["ms_rollout"] + info["ms_backprop"]) / (1e3 * wall))}
        if not a.no_host:
            for mode, overlap in (("host_tree_sync", 0), ("host_tree_1_in_flight", 1), ("host_tree_2_in_flight", 2)):
                for rep in range(2):
                    host.mcts_host_set_next_rollout_id(1 << 42)
                    ht = Session(host, m.GAME_QUORIDOR, root, rollouts_per_leaf=R, leaves_per_flush=K, policy=a.policy)
                    ht.set_gpu_expand(True)
                    ht.set_overlap(overlap)
                    t0 = time.perf_counter()
                    ht.grow(iters)
                    wall = time.perf_counter() - t0
                    tot, backend = ht.times()
                    ht.close()
                row[mode] = {"wall_ms_per_flush": 1e3 * wall / a.flushes, "rollouts_per_sec": iters * R / wall, "iterations_per_sec": iters / wall,
                             "host_share": (tot - backend) / tot}
            row["speedup_sync"] = row["host_tree_sync"]["wall_ms_per_flush"] / row["device_tree"]["wall_ms_per_flush"]
            row["speedup_1_in_flight"] = row["host_tree_1_in_flight"]["wall_ms_per_flush"] / row["device_tree_1_in_flight"]["wall_ms_per_flush"]
        out.append(row)
        print(json.dumps(row), flush=True)
    be.close()


if __name__ == "__main__":
    main()
