"""Multi-GPU plumbing for the rollout path (one process per GPU, torch.distributed for the plumbing).

The path shards naturally: rollouts are independent units keyed by their Philox stream
id, so a flush is cut into contiguous id blocks, one per rank, with NO data-path collective; the only
exchange is one all-reduce per step / per move of a few hundred bytes:
  * leaf statistics  — the exact integer accumulators {lo, hi, n} of include/mctsb.h (int64 sum is exact,
    so the merged result is bit-identical to a single-GPU run over the same id range);
  * root-child statistics of root-parallel trees — 2 doubles per root action, slot k = k-th move of the
    root's actions_to_try() order (MCTS_tree::export_root_stats / import_root_stats).
Works with the `nccl` backend on GPUs and with `gloo` on CPU (tests).
"""
import numpy as np


def shard_range(total, world, rank):
    """Contiguous block [start, start+count) of `total` units owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


class PendingLeafMerge:
    """An all-reduce of leaf accumulators in flight (async_op): `result()` waits and returns the merged array.
    Lets a rank launch its next flush while the previous flush's statistics cross NVLink."""

    def __init__(self, tensor, work, dtype):
        self.tensor, self.work, self.dtype = tensor, work, dtype

    def result(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        out = self.tensor.cpu().numpy()
        merged = np.zeros(out.shape[0], dtype=self.dtype)
        merged["lo"], merged["hi"], merged["n"] = out[:, 0].astype(np.uint64), out[:, 1].astype(np.uint64), out[:, 2].astype(np.uint64)
        return merged


def allreduce_leaf_sums_async(sums, dist, device=None):
    """Start the exact merge of per-leaf accumulators across ranks; returns a PendingLeafMerge."""
    import torch
    flat = np.stack([sums["lo"].astype(np.int64), sums["hi"].astype(np.int64), sums["n"].astype(np.int64)], axis=1)
    t = torch.from_numpy(np.ascontiguousarray(flat))
    if device is not None:
        t = t.to(device, non_blocking=True)
    work = dist.all_reduce(t, async_op=True)
    return PendingLeafMerge(t, work, sums.dtype)


def allreduce_leaf_sums(sums, dist, device=None):
    """Exact merge of per-leaf accumulators across ranks.  `sums`: numpy structured array (lo, hi, n)."""
    return allreduce_leaf_sums_async(sums, dist, device).result()


def leaf_sum_to_double(lo, hi):
    """Correctly rounded score sum of one accumulator (the same value mctsb_leaf_sum_to_double returns)."""
    from fractions import Fraction
    return float(Fraction(int(hi) * (1 << 29) + int(lo), 1 << 57))


def allreduce_root_stats(vec, dist, device=None):
    """Sum of the root-child statistics vectors of all ranks (every rank receives the same bytes)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()
