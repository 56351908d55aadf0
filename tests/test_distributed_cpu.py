"""CPU, world_size 2, gloo: host-side logic of the two multi-GPU modes.

  * problem sharding (no data-path collective): rank r owns problems [r*B/R, (r+1)*B/R) and keys its RNG
    by global problem ids; only results are gathered.
  * split-particle mode: ranks exchange (m, Z, A) statistics with all_gather and merge them by
    log-sum-exp (stoch_gpmp_b200.ops.merge_stats / parallel.allreduce_stats).
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stoch_gpmp_b200 import parallel
    # --- problem sharding
    B = 10
    lo, hi = parallel.shard_range(B, rank, world)
    mine = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1) * torch.ones(1, 3, dtype=torch.float64)
    gathered = parallel.gather_problem_results(mine, B)
    # --- split-particle statistics
    rs = np.random.RandomState(0)
    z = torch.tensor(rs.randn(2, 3, 40) * 5)
    eps = torch.tensor(rs.randn(2, 3, 6, 40))
    sl = slice(0, 13) if rank == 0 else slice(13, 40)
    zz, ee = z[..., sl], eps[..., sl]
    m = zz.max(-1).values
    w = torch.exp(zz - m.unsqueeze(-1))
    local = torch.cat([m.unsqueeze(-1), w.sum(-1, keepdim=True), (ee * w.unsqueeze(-2)).sum(-1)], -1)
    merged = parallel.allreduce_stats(local)
    q.put((rank, lo, hi, gathered.numpy(), merged.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 5, 5, 10)
    want = np.arange(10, dtype=np.float64).reshape(-1, 1) * np.ones((1, 3))
    assert np.array_equal(out[0][3], want) and np.array_equal(out[1][3], want)
    # merged statistics are identical on both ranks and equal the single-process result
    assert np.array_equal(out[0][4], out[1][4])
    rs = np.random.RandomState(0)
    z = rs.randn(2, 3, 40) * 5
    eps = rs.randn(2, 3, 6, 40)
    m = z.max(-1)
    w = np.exp(z - m[..., None])
    full_A = (eps * w[..., None, :]).sum(-1) / w.sum(-1)[..., None]
    got = out[0][4]
    assert np.allclose(got[..., 2:] / got[..., 1:2], full_A, atol=1e-12)


def test_shard_range_covers_everything():
    from stoch_gpmp_b200 import parallel
    for B in (1, 7, 8, 4096):
        for R in (1, 2, 4, 8):
            spans = [parallel.shard_range(B, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(R - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
