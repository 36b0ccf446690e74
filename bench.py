#!/usr/bin/env python
"""bench.py — Quoridor 9x9 rollouts/sec on N B200s vs the reference's host-core JobScheduler.

A "step" is one pass of the hot path over one batch: `--rollouts` (default 2^20, BASELINE.json
configs[1]) independent REF-policy playouts (reference Quoridor_state::rollout semantics,
Quoridor.cpp:809-855) of the opening position per GPU, each keyed by its own Philox stream id.

  value : whole-job rollouts/s, state already resident in HBM, kernel timed with CUDA events on the
          backend's stream (max over ranks, summed over the K steps).
  e2e   : the same through the C-ABI call a host tree makes (mctsb_rollout_batch: host state in,
          H2D, kernel, D2H of the exact per-leaf sums) timed on the host clock around the call.
  roofline : the path is SM-integer-issue bound (no contraction, ~0 HBM traffic), so the
          denominator is the measured integer-issue peak (mctsb_int_issue_peak, run live), and
          `achieved` is algorithmic thread-level integer ops/s (DESIGN.md "Roofline").
  cpu_baseline / --impl reference : the UNMODIFIED reference compiled in oracle/_ref
          (RolloutJobs on its JobScheduler with every host core) timed on this box.

N>1 (torchrun): every rank plays its own id range (weak scaling: `--rollouts` per GPU; `--scaling strong`:
`--rollouts` split over the GPUs); one all-reduce of the exact integer leaf sums per step — NCCL behind the C
ABI (mctsb_allreduce_leaf_sums) — is inside the timed region of BOTH `value` and `e2e`.  torch.distributed
ships the communicator id and provides the barriers.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "quoridor_9x9_ref_policy_rollouts_per_sec"
UNIT = "rollouts/s"
# Algorithmic integer work of one REF-policy rollout from the opening position — the REFERENCE algorithm's
# work, counted by the oracle (oracle/quoridor_oracle.c work counters, 4 000 rollouts under the bench seed):
# 13 133 BFS levels ("flood-fill steps") x 24 ops + 4 533 cheap wall tests x 12 ops + 49.1 plies x 200 ops
# (DESIGN.md "Roofline").  The kernel executes fewer steps than this: exact pruning.
# The literals are TOTALS over the counted sample, so the recount (tests/test_alg_ops.py, `bench.py --recount`) can
# assert equality instead of trusting a rounded mean:
#   opening: 4 000 rollouts, Philox ids 0..3999 under DEFAULT_SEED
#   corpus : states 0, 64, 128, ... 4032 of tests/golden/corpus_q.npy x 200 rollouts each, ids 200*i ..
ALG_SAMPLES = {
    "opening": {"rollouts": 4000, "flood_steps": 52533142, "cheap_wall_tests": 18130320, "plies": 196382},
    "corpus": {"rollouts": 12800, "flood_steps": 34874581, "cheap_wall_tests": 12006688, "plies": 381357},
    # UNI policy from the opening, 1 000 rollouts, ids 0..999: the reference's generate_all_moves + play_move loop
    "opening_uni": {"rollouts": 1000, "flood_steps": 24789089, "cheap_wall_tests": 5238252, "plies": 375626},
}
OPS_PER_FLOOD_STEP, OPS_PER_CHEAP_TEST, OPS_PER_PLY = 24, 12, 200


def alg_ops_per_rollout(workload):
    c = ALG_SAMPLES[workload]
    return (c["flood_steps"] * OPS_PER_FLOOD_STEP + c["cheap_wall_tests"] * OPS_PER_CHEAP_TEST + c["plies"] * OPS_PER_PLY) / c["rollouts"]


def recount(workload):
    """The oracle's work counters over the sample ALG_SAMPLES names (test infrastructure: only the recount uses it)."""
    import montecarlotreesearch_b200 as m
    from oracle.portlib import PortLib
    port = PortLib()
    tot = {"flood_steps": 0, "cheap_wall_tests": 0, "plies": 0}
    policy = 0
    if workload == "opening":
        runs = [(m.quoridor_opening(), 0, 4000)]
    elif workload == "opening_uni":
        runs, policy = [(m.quoridor_opening(), 0, 1000)], 1
    else:
        corpus = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))
        runs = [(corpus[i], 200 * i, 200) for i in range(0, 4096, 64)]
    n = 0
    for state, first, count in runs:
        _, _, w = port.q_rollout_philox(state, m.DEFAULT_SEED, first, count, policy=policy, work=True)
        for k in tot:
            tot[k] += int(w[k])
        n += count
    tot["rollouts"] = n
    return tot


ALG_OPS_PER_ROLLOUT = alg_ops_per_rollout("opening")
# executed flood step of the kernel (qcore.cuh q_flood_step): 18 IMAD + 12 LOP3 thread instructions per 81-cell expansion
EXEC_OPS_PER_FLOOD_STEP = 30
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of 2^20 rollouts (ncu --set full, profiles/): the
# algorithmic HBM traffic is 32 B in + 24 B out per LEAF; what is measured is stack/L2 write-back noise
NCU_DRAM_BYTES_PER_LAUNCH = 433664 + 1706752   # profiles/r02_lockstep.md (ncu --set full, one launch of 2^20 rollouts: read + write)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms 50` process streaming CSV rows, started before and killed after the region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "150"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)          # let the first sample land before the region starts
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        self.rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.count(",") >= 8]

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = sorted(v for v in (num(r[1]) for r in self.rows) if v is not None)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        power = [v for v in (num(r[3]) for r in self.rows) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": num(self.rows[0][2]), "reasons": sorted(reasons),
                "samples": len(self.rows), "power_w_max": max(power) if power else None}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_rollouts_per_sec(target_seconds, threads, states=None):
    """RolloutJobs on the reference's own JobScheduler (mcts.cpp:60-66) with `threads` workers; the opening position
    unless `states` (packed records, cycled over) is given.  -> (rate, jobs, seconds, mean score, std of one score)"""
    from oracle.reflib import RefLib
    ref = RefLib("native")
    calib_jobs = 200 * threads
    secs, _ = ref.jobscheduler_bench(0, threads, calib_jobs, qstates=states)
    rate = calib_jobs / secs
    jobs = max(calib_jobs, int(rate * target_seconds))
    secs, mean, std = ref.jobscheduler_bench2(0, threads, jobs, qstates=states)
    return jobs / secs, jobs, secs, mean, std


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    states = None
    if args.workload == "corpus":
        states = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]
    per_step = []
    sample_jobs = 0
    for i in range(args.warmup + args.steps):
        rate, jobs, secs, mean, std = reference_rollouts_per_sec(2.0 if i >= args.warmup else 0.5, threads, states)
        if i >= args.warmup:
            per_step.append((jobs, secs, mean))
            sample_jobs = jobs
    total_jobs = sum(j for j, _, _ in per_step)
    total_secs = sum(s for _, s, _ in per_step)
    value = total_jobs / total_secs
    where = "opening position" if states is None else "mid-game corpus (4096 states, cycled)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": ("configs[1]: Quoridor 9x9 opening position" if states is None else "configs[2]: mid-game corpus, 4096 states")
                               + ", independent REF-policy rollouts (the reference's rollout() as RolloutJobs on its JobScheduler; each step is a bounded sample of the workload)",
                   "policy": "REF (reference Quoridor_state::rollout semantics, <=50 plies + eval)",
                   "sample_rollouts_per_step": sample_jobs, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "%d RolloutJobs per step on JobScheduler(%d), %s, stock RNG" % (sample_jobs, threads, where)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "mean_score": sum(j * mn for j, _, mn in per_step) / max(1, total_jobs),
    }
    print(json.dumps(line), flush=True)


def time_batch_calls(be, G, P, states, per_leaf, first, calls):
    """median / min host-clock milliseconds of synchronous mctsb_rollout_batch calls"""
    ts = []
    for k in range(calls):
        t0 = time.perf_counter()
        be.rollout_batch(G, P, states, per_leaf, first + k * per_leaf * len(states))
        ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rollouts", type=int, default=1 << 20, help="rollouts per step per GPU (weak scaling) / per step over all GPUs (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rollouts per GPU per step (driver contract); strong: --rollouts per step split over the GPUs")
    ap.add_argument("--workload", default="opening", choices=["opening", "corpus"],
                    help="opening = configs[1] (headline); corpus = configs[2]: 4096 mid-game states x rollouts/4096 each")
    ap.add_argument("--policy", default="ref", choices=["ref", "uni"], help="ref = headline (reference rollout()); uni = uniformly random legal moves (secondary)")
    ap.add_argument("--game", default="quoridor", choices=["quoridor", "tictactoe"], help="tictactoe = configs[0] rollouts from the empty board (secondary)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU sample for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements of the `extra` block")
    ap.add_argument("--recount", action="store_true", help="recount the roofline numerators with the oracle's work counters and compare with the literals (CPU only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.recount:
        out = {}
        for w in ("opening", "corpus", "opening_uni"):
            got = recount(w)
            out[w] = {"counted": got, "literal": ALG_SAMPLES[w], "equal": got == ALG_SAMPLES[w], "alg_ops_per_rollout": alg_ops_per_rollout(w)}
        print(json.dumps(out))
        sys.exit(0 if all(v["equal"] for v in out.values()) else 1)

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # stdout carries exactly one JSON line: anything native libraries print (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import montecarlotreesearch_b200 as m

    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    be = m.RolloutBackend(local_rank, m.DEFAULT_SEED)
    if dist is not None:
        # the C ABI's own communicator: rank 0 makes the id, torch.distributed ships it (plumbing only)
        uid = torch.zeros(m.COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            uid = torch.frombuffer(bytearray(m.comm_unique_id()), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        be.comm_init_rank(uid.cpu().numpy().tobytes(), world, rank)

    rollouts_per_gpu = args.rollouts if args.scaling == "weak" else max(1, args.rollouts // world)
    if args.workload == "corpus":
        state = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]
        per_leaf = max(1, rollouts_per_gpu // 4096)
    else:
        state = np.array([m.quoridor_opening()], dtype=m.QSTATE_DTYPE)
        per_leaf = rollouts_per_gpu
    if args.game == "tictactoe":
        state = np.zeros(1, dtype=m.TSTATE_DTYPE)
        per_leaf = rollouts_per_gpu
    n_leaves = len(state)
    n = per_leaf * n_leaves
    G = m.GAME_QUORIDOR if args.game == "quoridor" else m.GAME_TICTACTOE
    P = m.POLICY_UNI if (args.policy == "uni" and args.game == "quoridor") else m.POLICY_REF
    secondary = args.game != "quoridor" or args.policy != "ref"
    uni = args.game == "quoridor" and args.policy == "uni"
    rooflined = not secondary or (uni and args.workload == "opening")       # workloads with a counted numerator

    def barrier():
        if dist is not None:
            dist.barrier()

    def first_id(step):
        return (step * world + rank) * n

    def merged(sums):
        """exact integer merge over the GPUs (no-op on one GPU): every rank gets the sums of the whole job"""
        return be.allreduce_leaf_sums(sums) if dist is not None else sums

    def score_of(sums):
        a = np.ascontiguousarray(sums, dtype=m.LEAFSUM_DTYPE).reshape(-1)
        return float(sum(be.lib.mctsb_leaf_sum_to_double(a[i:i + 1].ctypes.data) for i in range(a.size)))

    # ---- measured integer-issue peak (roofline denominator) -----------------------------------------
    int_peak, _ = be.int_issue_peak()
    lop3_peak, _ = be.lop3_peak()

    # ---- device-resident throughput ("value") -------------------------------------------------------------
    be.upload_states(G, state)
    launches0 = be.counters()["kernel_launches"]
    for w in range(args.warmup):                       # a warm-up step is a whole step: kernel, D2H, cross-rank merge
        be.run_resident(G, P, per_leaf, first_id(w), timed=True)
        _, wsums = be.wait()
        merged(wsums)
    barrier()
    # one nvidia-smi process per job (rank 0) watching every GPU of the job
    sampler = ClockSampler(",".join(str(g) for g in range(world)) if world > 1 else local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms = []
    total_score = 0.0
    step_wall = 0.0                                    # steps only: the L2 flush between them is not part of a step
    c0 = be.counters()
    for k in range(args.steps):
        be.flush_l2()                                  # 256 MiB memset + sync between timed steps
        ta = time.perf_counter()
        ms = be.run_resident(G, P, per_leaf, first_id(args.warmup + k), timed=True)
        _, sums = be.wait()
        total_score += score_of(merged(sums))          # N>1: blocking all-reduce of 24 B per leaf over NVLink, every step
        step_wall += time.perf_counter() - ta
        kernel_ms.append(ms)
    ta = time.perf_counter()
    barrier()
    step_wall += time.perf_counter() - ta              # the closing barrier belongs to the job
    c1 = be.counters()
    if rank == 0:
        sampler.stop()
    launches = c1["kernel_launches"] - c0["kernel_launches"]

    def max_over_ranks(values):
        if dist is None:
            return [float(v) for v in values]
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    dev_ms = float(sum(kernel_ms))
    dev_ms, wall_ms = max_over_ranks([dev_ms, step_wall * 1e3])
    # N=1: CUDA events around the kernel launches.  N>1: a step = kernel + D2H + cross-rank merge, host clock around the steps, max over ranks
    step_ms = dev_ms if dist is None else wall_ms
    total_rollouts = n * args.steps * world
    value = total_rollouts / (step_ms * 1e-3)
    mean_score = total_score / total_rollouts

    # ---- end to end through the C ABI with host buffers ----------------------------------------------------
    states_host = np.ascontiguousarray(state)
    merged(be.rollout_batch(G, P, states_host, per_leaf, first_id(1000))[1])
    barrier()
    e2e_s, e2e_score = 0.0, 0.0
    for k in range(args.steps):
        be.flush_l2()
        t0 = time.perf_counter()
        # the SAME rollout ids as the `value` steps: the same playouts, plus the copies of a call with host buffers
        _, sums = be.rollout_batch(G, P, states_host, per_leaf, first_id(args.warmup + k))    # host buffers in, H2D, kernel, D2H, host sums out
        e2e_score += score_of(merged(sums))                                                   # ... and the job-wide result on every rank, as in a `value` step
        e2e_s += time.perf_counter() - t0
    t0 = time.perf_counter()
    barrier()
    e2e_s += time.perf_counter() - t0                  # as for value: the closing barrier belongs to the job
    (e2e_s,) = max_over_ranks([e2e_s])
    e2e_value = total_rollouts / e2e_s

    # ---- secondary measurements for the driver-visible line (rank 0's GPU, after the timed regions) --------
    extra = None
    if not args.no_extra and not secondary and args.workload == "opening":
        extra = {}
        opening = np.array([m.quoridor_opening()], dtype=m.QSTATE_DTYPE)
        if rank == 0:
            # configs[2]: the mid-game corpus, device-resident like `value`
            corpus = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]
            be.upload_states(G, corpus)
            ms_c = []
            for k in range(3 + 5):
                t_ms = be.run_resident(G, P, 256, (5000 + k) * (1 << 20), timed=True)
                be.wait()
                if k >= 3:
                    ms_c.append(t_ms)
            rate_c = 5 * 4096 * 256 / (1e-3 * sum(ms_c))
            extra["corpus"] = {"workload": "configs[2]: 4096 mid-game states x 256 REF-policy rollouts per step, device-resident, 5 steps after 3 warm-ups",
                               "rollouts_per_sec": rate_c, "ms_per_step": sum(ms_c) / 5,
                               "roofline_frac": rate_c * alg_ops_per_rollout("corpus") / int_peak,
                               "alg_ops_per_rollout": alg_ops_per_rollout("corpus")}
            # configs[3] flush shape and the per-call floor: synchronous C-ABI calls with host buffers, host clock
            lat = {}
            for size, calls in ((1, 200), (4, 200), (32, 200), (1024, 100), (65536, 40)):
                time_batch_calls(be, G, P, opening, size, 7000 * (1 << 20), 5)
                med, mn = time_batch_calls(be, G, P, opening, size, 7001 * (1 << 20), calls)
                lat[str(size)] = {"median_ms": med, "min_ms": mn, "rollouts_per_sec": size / (1e-3 * med)}
            extra["call_latency"] = {"what": "synchronous mctsb_rollout_batch from the opening position (host buffers, H2D + kernel + D2H), host clock; 65536 = one configs[3] flush",
                                     "sizes": lat}
            # the search around the path: MCTS from the opening, flushes of 4096 leaves x 16 playouts, the tree
            # on the host (MCTS_tree, one thread) and resident on the device (mctsb_tree_*); one flush in flight in both
            if not uni:
                try:
                    from montecarlotreesearch_b200.hosttree import Session, load as load_host
                    hostlib = load_host()
                    K, R, F = 4096, 16, 12
                    rows = {}
                    for name in ("host_tree", "device_tree"):
                        for rep in range(2):                                   # the first search warms up, the second is timed
                            if name == "host_tree":
                                hostlib.mcts_host_set_next_rollout_id(9000 * (1 << 20))
                                t = Session(hostlib, G, opening[0], rollouts_per_leaf=R, leaves_per_flush=K, device=local_rank, seed=m.DEFAULT_SEED)
                                t.set_gpu_expand(True); t.set_overlap(1)
                                t0 = time.perf_counter(); t.grow(K * F); w = time.perf_counter() - t0
                                tot, backend_s = t.times()
                                host_share = (tot - backend_s) / tot
                                best = t.best_move(); t.close()
                            else:
                                t = m.DeviceTree(be, opening[0], R, K)
                                t.set_overlap(True)
                                t0 = time.perf_counter(); t.grow(K * F, 9000 * (1 << 20)); w = time.perf_counter() - t0
                                host_share = max(0.0, 1.0 - 1e-3 * t.info()["ms_total"] / w)
                                best = t.best_move(); t.close()
                        rows[name] = {"ms_per_flush": 1e3 * w / F, "iterations_per_sec": K * F / w, "rollouts_per_sec": K * F * R / w,
                                      "host_share_of_wall": host_share, "best_move_id": best}
                    extra["tree_search"] = {"what": "MCTS from the opening, %d flushes of %d leaves x %d REF playouts, one flush in flight; same Philox ids, so both trees are the same tree (tests/test_device_tree.py)" % (F, K, R),
                                            "same_best_move": rows["host_tree"]["best_move_id"] == rows["device_tree"]["best_move_id"], **rows}
                except Exception as e:                                        # never lose the main line to the secondary block
                    extra["tree_search"] = {"error": repr(e)}
        barrier()

    counters = be.counters()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    be.close()
    if rank != 0:
        return

    # ---- CPU baseline on this box's host cores (bounded sample) ------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and uni and args.workload == "opening":
        try:
            from oracle.reflib import RefLib
            threads = host_threads()
            ref = RefLib("native")
            secs, _, _ = ref.uniform_bench(threads, 50 * threads)
            jobs = max(50 * threads, int(50 * threads / secs * args.cpu_seconds))
            secs, mean, plies_mean = ref.uniform_bench(threads, jobs)
            cpu = {"value": jobs / secs, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": "%d uniformly random playouts from the opening position driven through the reference's public API "
                             "(actions_to_try = generate_all_moves, play_move) on %d threads, %.1f s; the reference ships no such loop, "
                             "this is the one the UNI oracle restates" % (jobs, threads, secs),
                   "mean_score": mean, "mean_plies": plies_mean}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    if not args.no_cpu_baseline and not secondary:
        try:
            threads = host_threads()
            cstates = None if args.workload == "opening" else np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]
            where = "opening position" if cstates is None else "mid-game corpus (4096 states, cycled)"
            rate, jobs, secs, mean, std = reference_rollouts_per_sec(args.cpu_seconds, threads, cstates)
            sem = std / max(1.0, jobs) ** 0.5
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": "%d RolloutJobs on the reference JobScheduler(%d), %s, %.1f s, stock RNG (std::default_random_engine, unshimmed)" % (jobs, threads, where, secs),
                   "mean_score": mean, "score_std": std, "mean_score_sem": sem,
                   "gpu_mean_minus_cpu_mean_in_sem": (mean_score - mean) / sem if sem > 0 else None,
                   "agrees_within_3_sigma": bool(abs(mean_score - mean) <= 3 * sem) if sem > 0 else None}
            by_threads = {}
            for t in sorted({1, min(4, threads)}):
                if t == threads:
                    continue
                r_t, j_t, s_t, _, _ = reference_rollouts_per_sec(max(1.0, args.cpu_seconds / 6), t, cstates)
                by_threads[str(t)] = {"value": r_t, "sample": "%d RolloutJobs, %.1f s" % (j_t, s_t)}
            by_threads[str(threads)] = {"value": rate, "sample": "%d RolloutJobs, %.1f s" % (jobs, secs)}
            cpu["by_threads"] = by_threads         # JobScheduler(1), (4) = stock NUMBER_OF_THREADS, (nproc)
            if extra is not None and "corpus" in extra:
                cst = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]
                r_c, j_c, s_c, mean_c, _ = reference_rollouts_per_sec(max(1.0, args.cpu_seconds / 4), threads, cst)
                extra["corpus"]["cpu_baseline"] = {"value": r_c, "unit": UNIT, "cores": threads, "kind": "reference",
                                                   "sample": "%d RolloutJobs cycling over the 4096 corpus states, %.1f s" % (j_c, s_c), "mean_score": mean_c}
        except Exception as e:  # the compiled reference may be absent
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    per_gpu_rate = (n * args.steps) / (dev_ms * 1e-3)
    alg_ops = alg_ops_per_rollout("opening_uni" if uni else args.workload)
    achieved = per_gpu_rate * alg_ops
    d_roll = max(1, c1["rollouts"] - c0["rollouts"])
    exec_steps_per_rollout = (c1["flood_steps"] - c0["flood_steps"]) / d_roll
    exec_flood_ops = per_gpu_rate * exec_steps_per_rollout * EXEC_OPS_PER_FLOOD_STEP
    line = {
        "metric": METRIC if not secondary else "%s_%s_rollouts_per_sec" % (args.game, args.policy), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": ("Tic-Tac-Toe empty board, %d independent rollouts per GPU per step" % n) if args.game != "quoridor"
                   else ("configs[1]: Quoridor 9x9 opening position, %d independent %s-policy rollouts per GPU per step" % (n, args.policy.upper())) if args.workload == "opening"
                   else ("configs[2]: mid-game corpus, 4096 states x %d %s-policy rollouts per GPU per step" % (per_leaf, args.policy.upper())),
                   "policy": "REF (reference Quoridor_state::rollout semantics, <=50 plies + eval)" if not secondary
                             else "UNI (uniformly random legal move per ply, to the end of the game)" if args.game == "quoridor" else "reference TicTacToe_state::rollout",
                   "rollouts_per_step_per_gpu": n, "leaves": n_leaves, "seed": hex(m.DEFAULT_SEED),
                   "l2": "flushed between timed steps (256 MiB memset, outside the timed kernel); inputs are 32 B per leaf and rollout ids differ every step",
                   "merge": None if world == 1 else "every step: mctsb_allreduce_leaf_sums (NCCL behind the C ABI) of the exact per-leaf sums, inside the timed region of value and e2e"},
        "roofline": {"bound": "sm_int_issue", "achieved": achieved / 1e12 if rooflined else None, "peak": int_peak / 1e12, "unit": "Tint-op/s",
                     "frac": achieved / int_peak if rooflined else None,
                     "frac_algorithmic": achieved / int_peak if rooflined else None,
                     "frac_executed_flood_arithmetic": exec_flood_ops / int_peak if args.game == "quoridor" else None,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if not secondary else None,
                     "traffic_note": "REF policy, 2^20 rollouts: 2.1 MB per 28 ms launch = 76 MB/s of spill / counter write-back against 32 B in + 24 B out per leaf algorithmic: HBM is idle, the bound is integer issue"
                                     if not secondary else "UNI policy: 48 B hand-over record written by the wall phase and read by the walk kernel per rollout (profiles/r02_uni.md)" if uni else None,
                     "alu_pipe_lop3_peak": lop3_peak / 1e12, "alg_ops_per_rollout": alg_ops, "executed_flood_steps_per_rollout": exec_steps_per_rollout,
                     "note": "thread-level integer ops; peak measured live by mctsb_int_issue_peak (LOP3 + integer-add chains on all SMs, both pipes; alu_pipe_lop3_peak = LOP3 only). "
                             "frac (= frac_algorithmic) = rollouts/s x the REFERENCE algorithm's integer work per rollout (oracle work counters, recounted by tests/test_alg_ops.py) / peak: "
                             "it rewards exact pruning and is not SM utilisation.  frac_executed_flood_arithmetic = flood steps the kernel executed in this run (device counters) x 30 "
                             "thread instructions per step / peak: the share of the issue peak spent on flood-fill arithmetic proper; the rest of the instruction stream is bookkeeping (profiles/)"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 32 * n_leaves * world, "d2h_bytes_per_step": 24 * n_leaves * world,
                "note": "mctsb_rollout_batch with host buffers" + ("" if world == 1 else " + the cross-rank merge, every step") +
                        "; the same rollout ids as the value steps (the same playouts): it differs from value by the launch and copy latency of a call with host buffers (tens of microseconds per 28 ms step)",
                "mean_score": e2e_score / total_rollouts},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "mean_score": mean_score,
        "kernel_ms_per_step": dev_ms / args.steps,
        "extra": extra,
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
