#!/usr/bin/env python
"""DEVELOPMENT TOOL (not product, not test): build variants of libmctsb.so with extra -D switches and time
the lockstep kernel of each on the GPU box in one gpurun call.

  python tools/kvar.py build name1:-DX=1,-DY name2: ...     (here: nvcc cross-compiles into tools/_var/<name>/)
  python tools/kvar.py run [rollouts] [workload]             (on the GPU box: every built variant, 4 timed launches;
                                                              KVAR_POLICY=1 times the UNI path instead of the REF kernel)
"""
import os, subprocess, sys, json, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "tools", "_var")
sys.path.insert(0, ROOT)

def build(specs):
    from montecarlotreesearch_b200 import _build as b
    shutil.rmtree(VAR, ignore_errors=True)
    for spec in specs:
        name, _, flags = spec.partition(":")
        d = os.path.join(VAR, name); os.makedirs(d, exist_ok=True)
        cmd = [b._nvcc()] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", os.path.join(d, "libmctsb.so")] + b.cuda_sources()
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            print(res.stdout); raise SystemExit("build failed: " + name)
        regs = [l for l in res.stdout.splitlines() if "Used" in l]
        info = res.stdout.split("quoridor_lockstep_kernelILb0")[1].split("Compile time")[0] if "quoridor_lockstep_kernelILb0" in res.stdout else ""
        print(name, " ".join(info.split()[-30:]))

def run_one(lib, rollouts, workload):
    import numpy as np
    import montecarlotreesearch_b200 as m
    m.LIB_PATH = lib
    be = m.RolloutBackend(0, m.DEFAULT_SEED)
    if workload == "corpus":
        state = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]; per_leaf = max(1, rollouts // 4096)
    else:
        state = np.array([m.quoridor_opening()], dtype=m.QSTATE_DTYPE); per_leaf = rollouts
    be.upload_states(m.GAME_QUORIDOR, state)
    ms = []
    for k in range(6):
        t = be.run_resident(m.GAME_QUORIDOR, int(os.environ.get("KVAR_POLICY", "0")), per_leaf, k * rollouts, timed=True)
        score, sums = be.wait()
        ms.append(t)
    c = be.counters()
    best = min(ms[2:]); n = per_leaf * len(state)
    return {"M_per_s": n / best / 1e3, "ms": best, "mean": float(np.sum(score)) / n,
            "steps_per_rollout": c["flood_steps"] / max(1, c["rollouts"])}

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        print(json.dumps(run_one(sys.argv[2], int(sys.argv[3]), sys.argv[4])))
    else:
        rollouts = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
        workload = sys.argv[3] if len(sys.argv) > 3 else "opening"
        for name in sorted(os.listdir(VAR)):
            lib = os.path.join(VAR, name, "libmctsb.so")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", lib, str(rollouts), workload], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            print(name, r.stdout.strip() or r.stderr[-400:], flush=True)
