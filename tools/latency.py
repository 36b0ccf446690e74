#!/usr/bin/env python
"""Per-call latency floor and small-launch throughput of the rollout path through the C ABI
(mctsb_rollout_batch: host state in, H2D, kernel, D2H of the leaf sums, host clock around the call).
Prints one JSON line; kept under profiles/ (VERDICT r01 item 6)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import montecarlotreesearch_b200 as m

def main():
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    st = np.array([m.quoridor_opening()], dtype=m.QSTATE_DTYPE)
    G, P = m.GAME_QUORIDOR, m.POLICY_REF
    out = {"what": "mctsb_rollout_batch, opening position, REF policy, synchronous calls, host clock", "sizes": {}}
    for n in (1, 4, 32, 1024, 4096, 16384, 65536, 262144, 1048576):
        reps = 200 if n <= 1024 else 40 if n <= 65536 else 8
        for k in range(3):
            be.rollout_batch(G, P, st, n, k * n)
        ts = []
        for k in range(reps):
            t0 = time.perf_counter()
            be.rollout_batch(G, P, st, n, (10 + k) * n)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        med = ts[len(ts) // 2]
        out["sizes"][str(n)] = {"median_ms": med * 1e3, "min_ms": ts[0] * 1e3, "rollouts_per_s": n / med}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
