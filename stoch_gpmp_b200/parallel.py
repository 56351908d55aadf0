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


class StreamShards:
    """Host-to-host pipelining on ONE GPU: the shards of a problem batch (one StochGPMPBatch each, built by the caller with
    `problem_offset` = start of the shard, so every result is bit-identical to the unsharded batch) run on their own CUDA
    streams.  `step()` copies every shard's observation H2D from pinned memory, runs optimize() and copies the shard's plan
    (particle means) D2H; the D2H of shard k overlaps the kernel of shard k+1.  `wait()` blocks the host until every plan is in
    its pinned buffer.

        shards = StreamShards([planner_0, ..., planner_7], obs_key='obstacle_spheres', host_inputs=[pinned_0, ...])
        shards.step(); shards.wait(); shards.host_means[k]    # [B_k, NP, T, d] pinned
    """

    def __init__(self, planners, obs_key=None, host_inputs=None):
        self.planners = list(planners)
        dev = self.planners[0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.planners]
        ready = torch.cuda.current_stream(dev).record_event()      # the planners were built (reset()) on the current stream
        for st in self.streams:
            st.wait_event(ready)
        self.host_means = [torch.empty(p.particle_means.shape, dtype=p.dtype).pin_memory() for p in self.planners]
        self.obs_key = obs_key
        self.host_inputs = None
        self.dev_inputs = None
        if host_inputs is not None:
            self.host_inputs = [h if h.is_pinned() else h.pin_memory() for h in host_inputs]
            self.dev_inputs = [torch.empty_like(h, device=dev) for h in self.host_inputs]

    def step(self, **optimize_kwargs):
        for k, (p, st) in enumerate(zip(self.planners, self.streams)):
            with torch.cuda.stream(st):
                obs = {}
                if self.host_inputs is not None:
                    self.dev_inputs[k].copy_(self.host_inputs[k], non_blocking=True)
                    if self.obs_key is not None:
                        obs[self.obs_key] = self.dev_inputs[k]
                p.optimize(return_samples=False, **obs, **optimize_kwargs)
                self.host_means[k].copy_(p.particle_means, non_blocking=True)

    def wait(self):
        for st in self.streams:
            st.synchronize()
