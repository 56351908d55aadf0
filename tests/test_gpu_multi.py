"""GPU, 2 devices, NCCL: the two multi-GPU modes against the single-GPU result (skipped on a 1-GPU box).

  * problem sharding: each rank runs its slice with problem_offset; gathered means == one-GPU batch, bit for bit.
  * split-particle: each rank scores half of every particle's samples; one ncclAllGather of (m, Z, A) per iteration, issued
    from C on the compute stream (csrc/sgpmp_nccl.cu);
    means == the one-GPU fused loop (fp64: 1e-9).
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, dtype, B, lo, hi, S, seed=11):
    from stoch_gpmp_b200.scenarios import panda_batch
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.costs.fields import LinkDistanceField
    from stoch_gpmp_b200.planner import StochGPMPBatch
    from stoch_gpmp_b200.robots import PandaFK
    n, T, G, K = 7, 16, 2, 1
    ta = dict(device=dev, dtype=dtype)
    start, goals, spheres = panda_batch(B, G=G, O=3, seed0=5)
    s = torch.tensor(start[lo:hi], **ta)
    g = torch.tensor(goals[lo:hi], **ta)
    comp = CostComposite(n, T, [
        CostGP(n, T, s, 0.05, dict(sigma_start=0.5, sigma_gp=0.5), ta),
        CostGoalPrior(n, T, multi_goal_states=g, num_particles_per_goal=K, num_samples=S, sigma_goal_prior=20., tensor_args=ta),
        CostCollision(n, T, field=LinkDistanceField(tensor_args=ta), sigma_coll=0.3)], FK=PandaFK(), tensor_args=ta)
    pl = StochGPMPBatch(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=1, dt=0.05, n_dof=n, step_size=0.5,
                        temperature=200., start_state=s, multi_goal_states=g, initial_particle_means='const_vel', cost=comp,
                        seed=seed, tensor_args=ta, problem_offset=lo, sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                        sigma_start_sample=4.0, sigma_goal_sample=4.0, sigma_gp_sample=0.5)
    obs = {'obstacle_spheres': torch.tensor(spheres[lo:hi], **ta)}
    return pl, obs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from stoch_gpmp_b200 import parallel
    B, S = 4, 64
    # problem sharding (fp32)
    lo, hi = parallel.shard_range(B, rank, world)
    pl, obs = _build(dev, torch.float32, B, lo, hi, S)
    pl.optimize(opt_iters=3, **obs)
    sharded = parallel.gather_problem_results(pl.particle_means, B)
    # split-particle (fp64): every rank holds all problems, half of the samples
    ps, obs2 = _build(dev, torch.float64, B, 0, B, S)
    out = ps.optimize_split(opt_iters=3, return_samples=True, **obs2)
    q.put((rank, sharded.cpu().numpy(), ps.particle_means.cpu().numpy(), tuple(out[2].shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_modes_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    B, S = 4, 64
    one, obs = _build(dev, torch.float32, B, 0, B, S)
    one.optimize(opt_iters=3, **obs)
    assert np.array_equal(res[0][1], one.particle_means.cpu().numpy())          # sharding: bit-identical
    assert np.array_equal(res[0][1], res[1][1])
    one64, obs64 = _build(dev, torch.float64, B, 0, B, S)
    one64.optimize(opt_iters=3, **obs64)
    want = one64.particle_means.cpu().numpy()
    for r in res:
        assert r[3] == (B, 2, S // 2, 16, 7)                                    # each rank holds half of the samples
        assert np.abs(r[2] - want).max() / np.abs(want).max() < 1e-9            # split mode == single GPU
    assert np.array_equal(res[0][2], res[1][2])                                 # identical on every rank
