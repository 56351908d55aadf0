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


class NcclComm:
    """NCCL communicator owned by libsgpmp.so (csrc/sgpmp_nccl.cu) for the ranks of a torch.distributed group: rank 0 draws the
    unique id in C, torch.distributed carries its 128 bytes to the other ranks, every rank calls ncclCommInitRank in C.  The
    split-particle exchange is then issued from C on the compute stream (ops.iterate_split_particles) — torch.distributed is
    plumbing for the rendezvous only.  One communicator per (group, device) is cached."""

    _cache = {}

    def __init__(self, group=None, device=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.handle = None
        if self.world == 1:
            return
        try:       # name the NCCL that torch itself uses (bundled in the wheel), so that one copy serves both
            import nvidia.nccl
            import os
            cand = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
            _lib.check(lib.sgpmp_nccl_load(cand.encode() if os.path.exists(cand) else None), "sgpmp_nccl_load")
        except ImportError:
            _lib.check(lib.sgpmp_nccl_load(None), "sgpmp_nccl_load")
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(lib.sgpmp_comm_unique_id(C.cast(buf, C.c_void_p)), "sgpmp_comm_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        obj = [uid.tolist()]
        dist.broadcast_object_list(obj, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(obj[0])
        handle = C.c_void_p()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(dev):
            idbuf = C.create_string_buffer(raw, 128)
            _lib.check(lib.sgpmp_comm_init(C.cast(idbuf, C.c_void_p), self.rank, self.world, C.byref(handle)), "sgpmp_comm_init")
        self.handle = handle.value

    @classmethod
    def for_group(cls, group=None, device=None):
        key = (id(group) if group is not None else None, str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(group, device)
        return cls._cache[key]

    def destroy(self):
        if self.handle:
            from . import _lib
            _lib.load().sgpmp_comm_destroy(self.handle)
            self.handle = None


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
