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

N>1 (torchrun): weak scaling, every rank plays its own id range; one NCCL all-reduce of the exact
integer root statistics per step is inside the timed region.
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
ALG_OPS_PER_ROLLOUT = 13133.3 * 24 + 4532.6 * 12 + 49.1 * 200
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of 2^20 rollouts (ncu --set full, profiles/): the
# algorithmic HBM traffic is 32 B in + 24 B out per LEAF; what is measured is stack/L2 write-back noise
NCU_DRAM_BYTES_PER_LAUNCH = 818944 + 520704   # profiles/r01_final_lockstep.md (ncu, one launch)


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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


def reference_rollouts_per_sec(target_seconds, threads):
    """RolloutJobs on the reference's own JobScheduler (mcts.cpp:60-66), all host threads, opening state."""
    from oracle.reflib import RefLib
    ref = RefLib("native")
    calib_jobs = 200 * threads
    secs, _ = ref.jobscheduler_bench(0, threads, calib_jobs)
    rate = calib_jobs / secs
    jobs = max(calib_jobs, int(rate * target_seconds))
    secs, mean = ref.jobscheduler_bench(0, threads, jobs)
    return jobs / secs, jobs, secs, mean


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    per_step = []
    sample_jobs = 0
    for i in range(args.warmup + args.steps):
        rate, jobs, secs, mean = reference_rollouts_per_sec(2.0 if i >= args.warmup else 0.5, threads)
        if i >= args.warmup:
            per_step.append((jobs, secs))
            sample_jobs = jobs
    total_jobs = sum(j for j, _ in per_step)
    total_secs = sum(s for _, s in per_step)
    value = total_jobs / total_secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": "configs[1]: Quoridor 9x9 opening position, independent REF-policy rollouts "
                               "(the reference's rollout() as RolloutJobs on its JobScheduler; each step is a bounded sample of the workload)",
                   "policy": "REF (reference Quoridor_state::rollout semantics, <=50 plies + eval)",
                   "sample_rollouts_per_step": sample_jobs, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": "%d RolloutJobs per step on JobScheduler(%d), opening position, stock RNG" % (sample_jobs, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rollouts", type=int, default=1 << 20, help="rollouts per step per GPU")
    ap.add_argument("--workload", default="opening", choices=["opening", "corpus"],
                    help="opening = configs[1] (headline); corpus = configs[2]: 4096 mid-game states x rollouts/4096 each")
    ap.add_argument("--policy", default="ref", choices=["ref", "uni"], help="ref = headline (reference rollout()); uni = uniformly random legal moves (secondary)")
    ap.add_argument("--game", default="quoridor", choices=["quoridor", "tictactoe"], help="tictactoe = configs[0] rollouts from the empty board (secondary)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU sample for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # stdout carries exactly one JSON line: anything native libraries print (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import montecarlotreesearch_b200 as m
    from montecarlotreesearch_b200 import multi

    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    be = m.RolloutBackend(local_rank, m.DEFAULT_SEED)
    if args.workload == "corpus":
        state = np.load(os.path.join(ROOT, "tests", "golden", "corpus_q.npy"))[:4096]
        per_leaf = max(1, args.rollouts // 4096)
    else:
        state = np.array([m.quoridor_opening()], dtype=m.QSTATE_DTYPE)
        per_leaf = args.rollouts
    if args.game == "tictactoe":
        state = np.zeros(1, dtype=m.TSTATE_DTYPE)
        per_leaf = args.rollouts
    n_leaves = len(state)
    n = per_leaf * n_leaves
    G = m.GAME_QUORIDOR if args.game == "quoridor" else m.GAME_TICTACTOE
    P = m.POLICY_UNI if (args.policy == "uni" and args.game == "quoridor") else m.POLICY_REF
    secondary = args.game != "quoridor" or args.policy != "ref"

    def barrier():
        if dist is not None:
            dist.barrier()

    def first_id(step):
        return (step * world + rank) * n

    # ---- measured integer-issue peak (roofline denominator) -----------------------------------------
    int_peak, _ = be.int_issue_peak()
    lop3_peak, _ = be.lop3_peak()

    # ---- device-resident throughput ("value") -------------------------------------------------------------
    be.upload_states(G, state)
    launches0 = be.counters()["kernel_launches"]
    for w in range(args.warmup):                       # a warm-up step is a whole step: kernel, D2H, cross-rank merge
        be.run_resident(G, P, per_leaf, first_id(w), timed=True)
        _, wsums = be.wait()
        if dist is not None:
            multi.allreduce_leaf_sums(wsums, dist, device="cuda")
    barrier()
    # one nvidia-smi process per job (rank 0) watching every GPU of the job
    sampler = ClockSampler(",".join(str(g) for g in range(world)) if world > 1 else local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    kernel_ms, merges = [], []

    def merged_score(pending):
        ms_ = pending.result()
        return float(sum(multi.leaf_sum_to_double(lo, hi) for lo, hi in zip(ms_["lo"], ms_["hi"])))

    total_score = 0.0
    step_wall = 0.0                                    # steps only: the L2 flush between them is not part of a step
    for k in range(args.steps):
        be.flush_l2()                                  # 256 MiB memset + sync between timed steps
        ta = time.perf_counter()
        ms = be.run_resident(G, P, per_leaf, first_id(args.warmup + k), timed=True)
        score, sums = be.wait()
        if dist is not None:
            # exact integer merge over NVLink (NCCL), started now and collected one step later: the next
            # step's kernel runs while this step's statistics are exchanged
            merges.append(multi.allreduce_leaf_sums_async(sums, dist, device="cuda"))
            if len(merges) > 1:
                total_score += merged_score(merges.pop(0))
        else:
            total_score += float(np.sum(score))
        step_wall += time.perf_counter() - ta
        kernel_ms.append(ms)
    ta = time.perf_counter()
    while merges:
        total_score += merged_score(merges.pop(0))
    barrier()
    step_wall += time.perf_counter() - ta              # the last merge and the closing barrier belong to the job
    if rank == 0:
        sampler.stop()
    launches = be.counters()["kernel_launches"] - launches0 - args.warmup

    dev_ms = float(sum(kernel_ms))
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, step_wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
        step_ms = wall_ms            # at N>1 a step = kernel + D2H + cross-rank merge: host clock around the steps, max over ranks
    else:
        step_ms = dev_ms             # N=1: CUDA events around the kernel launches
    total_rollouts = n * args.steps * world
    value = total_rollouts / (step_ms * 1e-3)
    mean_score = total_score / (n * args.steps * world) if dist is not None else total_score / (n * args.steps)

    # ---- end to end through the C ABI with host buffers ----------------------------------------------------
    states_host = np.ascontiguousarray(state)
    be.rollout_batch(G, P, states_host, per_leaf, first_id(1000))
    barrier()
    e2e_s = 0.0
    for k in range(args.steps):
        be.flush_l2()
        t0 = time.perf_counter()
        be.rollout_batch(G, P, states_host, per_leaf, first_id(2000 + k))    # host buffers in, H2D, kernel, D2H, host sums out
        e2e_s += time.perf_counter() - t0
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = total_rollouts / e2e_s

    counters = be.counters()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- CPU baseline on this box's host cores (bounded sample) ------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and not secondary:
        try:
            threads = host_threads()
            rate, jobs, secs, mean = reference_rollouts_per_sec(args.cpu_seconds, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": "%d RolloutJobs on the reference JobScheduler(%d), opening position, %.1f s" % (jobs, threads, secs)}
        except Exception as e:  # the compiled reference may be absent
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}

    per_gpu_rate = (n * args.steps) / (dev_ms * 1e-3)
    # mid-game corpus: 2 709 BFS levels, 929 cheap wall tests, 29.8 plies per rollout (oracle, 64 states x 200 rollouts)
    alg_ops = ALG_OPS_PER_ROLLOUT if args.workload == "opening" else 2709.4 * 24 + 929.5 * 12 + 29.8 * 200
    achieved = per_gpu_rate * alg_ops
    exec_steps_per_rollout = counters["flood_steps"] / max(1, counters["rollouts"])
    line = {
        "metric": METRIC if not secondary else "%s_%s_rollouts_per_sec" % (args.game, args.policy), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": ("Tic-Tac-Toe empty board, %d independent rollouts per GPU per step" % n) if args.game != "quoridor"
                   else ("configs[1]: Quoridor 9x9 opening position, %d independent %s-policy rollouts per GPU per step" % (n, args.policy.upper())) if args.workload == "opening"
                   else ("configs[2]: mid-game corpus, 4096 states x %d %s-policy rollouts per GPU per step" % (per_leaf, args.policy.upper())),
                   "policy": "REF (reference Quoridor_state::rollout semantics, <=50 plies + eval)" if not secondary
                             else "UNI (uniformly random legal move per ply, to the end of the game)" if args.game == "quoridor" else "reference TicTacToe_state::rollout",
                   "rollouts_per_step_per_gpu": n, "leaves": n_leaves, "seed": hex(m.DEFAULT_SEED),
                   "l2": "flushed between timed steps (256 MiB memset, outside the timed kernel); inputs are 32 B per leaf and rollout ids differ every step"},
        "roofline": {"bound": "sm_int_issue", "achieved": None if secondary else achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tint-op/s",
                     "frac": None if secondary else achieved / int_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                     "alu_pipe_lop3_peak": lop3_peak / 1e12, "alg_ops_per_rollout": alg_ops, "executed_flood_steps_per_rollout": exec_steps_per_rollout,
                     "note": "thread-level integer ops; peak measured live by mctsb_int_issue_peak (LOP3 + integer-add chains on all SMs, both pipes; alu_pipe_lop3_peak = LOP3 only); "
                             "achieved = rollouts/s x the REFERENCE algorithm's integer work per rollout (DESIGN.md 5); executed "
                             "thread instructions per second from ncu are in profiles/ (18.7 T/s = 0.54 of this peak)"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 32 * n_leaves * world, "d2h_bytes_per_step": 24 * n_leaves * world},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "mean_score": mean_score,
        "kernel_ms_per_step": dev_ms / args.steps,
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
