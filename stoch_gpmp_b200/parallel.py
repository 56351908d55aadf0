"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

StochGPMP shards naturally by planning problem: problems share nothing, so the data path
has NO collective — each rank runs StochGPMPBatch on its contiguous slice with `problem_offset` set to
the slice start (RNG streams are keyed by global ids, results are bit-identical for any sharding).
A collective exists only in split-particle mode, where ONE problem's samples are divided over ranks and
the per-particle softmax needs one exchange of (m, Z, A) statistics per iteration:
    m = max_s(-c_s/tau),  Z = sum_s exp(-c_s/tau - m),  A = sum_s exp(-c_s/tau - m) eps_s
merged by log-sum-exp.  The message is NP*(M+2) reals (14.4 KB for the Panda shape): latency-bound, one
all_gather instead of a MAX-allreduce followed by a SUM-allreduce.
"""
import torch
import torch.distributed as dist

from .ops import merge_stats


def shard_range(num_problems, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of the problem axis owned by `rank`."""
    base, rem = divmod(num_problems, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_problem_results(local, num_problems, group=None):
    """All ranks receive the per-problem results [B, ...] assembled from the per-rank slices (dim 0)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(num_problems, r, world) for r in range(world)]
    bufs = [torch.empty((hi - lo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for lo, hi in sizes]
    dist.all_gather(bufs, local.contiguous(), group=group)
    return torch.cat(bufs, dim=0)


def allreduce_stats(stats, group=None):
    """Split-particle exchange: all_gather the local (m, Z, A) blocks and merge them by log-sum-exp.
    Every rank gets the same merged statistics (deterministic: fixed rank order)."""
    world = dist.get_world_size(group)
    bufs = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(bufs, stats.contiguous(), group=group)
    return merge_stats(bufs)
