"""The reference's examples/planar_environment.py on stoch_gpmp_b200 (headless: prints instead of plotting).

Identical to the reference script (examples/planar_environment.py:12-110) except for the package name, a fixed
seed, and the removed matplotlib block.  Run: python examples/planar_environment.py [--float32]
"""
import random
import sys
import time

import numpy as np
import torch

from stoch_gpmp_b200.envs.map_generator import generate_obstacle_map
from stoch_gpmp_b200.planner import StochGPMP, print_info
from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior


if __name__ == "__main__":
    device = torch.device('cuda:0')
    tensor_args = {'device': device, 'dtype': torch.float32 if '--float32' in sys.argv else torch.float64}

    n_dof = 2
    traj_len = 64
    dt = 0.02
    num_particles_per_goal = 5
    num_samples = 128
    seed = 0
    start_q = torch.Tensor([-9, -9]).to(**tensor_args)
    start_state = torch.cat((start_q, torch.zeros(2, **tensor_args)))
    multi_goal_states = torch.tensor([
        [9, 6, 0., 0.],
        [9, -3, 0., 0.],
        [-3, 9, 0., 0.],
    ]).to(**tensor_args)

    obst_params = dict(map_dim=[20, 20], obst_list=[], cell_size=0.1, random_gen=True, num_obst=15,
                       rand_limits=[[-7.5, 7.5], [-7.5, 7.5]], rand_rect_shape=[2, 2], tensor_args=tensor_args)
    random.seed(seed)
    np.random.seed(seed)
    obst_map = generate_obstacle_map(**obst_params)[0]

    cost_sigmas = dict(sigma_start=0.001, sigma_gp=0.1)
    sigma_coll = 1e-5
    sigma_goal_prior = 0.001
    cost_prior = CostGP(n_dof, traj_len, start_state, dt, cost_sigmas, tensor_args)
    cost_goal_prior = CostGoalPrior(n_dof, traj_len, multi_goal_states=multi_goal_states,
                                    num_particles_per_goal=num_particles_per_goal, num_samples=num_samples,
                                    sigma_goal_prior=sigma_goal_prior, tensor_args=tensor_args)
    cost_obst_2D = CostCollision(n_dof, traj_len, field=obst_map, sigma_coll=sigma_coll)
    cost_composite = CostComposite(n_dof, traj_len, [cost_prior, cost_goal_prior, cost_obst_2D])

    stochgpmp_params = dict(
        num_particles_per_goal=num_particles_per_goal, num_samples=num_samples, traj_len=traj_len, dt=dt, n_dof=n_dof,
        opt_iters=1, temperature=1., start_state=start_state, multi_goal_states=multi_goal_states, cost=cost_composite,
        step_size=0.5, sigma_start_init=1e-3, sigma_goal_init=1e-3, sigma_gp_init=20., sigma_start_sample=1e-3,
        sigma_goal_sample=1e-3, sigma_gp_sample=3, seed=seed, tensor_args=tensor_args,
    )
    planner = StochGPMP(**stochgpmp_params)
    obs = {}

    opt_iters = 500
    torch.cuda.synchronize()
    start_time = time.time()
    traj_history = []
    for i in range(opt_iters + 1):
        start_time_iter = time.time()
        _, _, _, _, costs, _ = planner.optimize(**obs)
        if i == 1 or i % 50 == 0:
            print_info(i, opt_iters, start_time_iter, start_time, costs)
            trajectories, controls = planner.get_recent_samples()
            traj_history.append(trajectories)
    torch.cuda.synchronize()
    print(f'plan: {opt_iters + 1} optimize() calls in {(time.time() - start_time) * 1e3:.1f} ms (Python loop, one launch each)')

    # the same plan as ONE fused launch
    planner = StochGPMP(**stochgpmp_params)
    torch.cuda.synchronize()
    t0 = time.time()
    planner.optimize(opt_iters=opt_iters + 1, **obs)
    torch.cuda.synchronize()
    print(f'plan: one optimize(opt_iters={opt_iters + 1}) call in {(time.time() - t0) * 1e3:.1f} ms')

    # collision check of the final mean trajectories against the occupancy grid (the reference plots them)
    means = planner.particle_means[..., :2]
    occ = obst_map.compute_cost(means)
    print('final mean trajectories: cells in collision per particle =', occ.sum(-1).int().tolist())
