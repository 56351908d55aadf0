"""The reference's Gauss-Newton planner (stoch_gpmp/planner.py:352-661, `GPMP`) on stoch_gpmp_b200 — the reference ships no
example for it; this one uses the cost list of examples/panda_environment.py (GP prior, goal prior, self-collision, obstacle
spheres, end-effector SE(3) goal) with sigmas soft enough for a Gauss-Newton step to make progress.

    PYTHONPATH=. python examples/panda_gpmp.py
"""
import math
import time

import numpy as np
import torch

from stoch_gpmp_b200.planner import GPMP
from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoal, CostGoalPrior
from stoch_gpmp_b200.costs.fields import EESE3DistanceField, LinkDistanceField, LinkSelfDistanceField
from stoch_gpmp_b200.robots import PandaFK
from stoch_gpmp_b200 import ops


if __name__ == '__main__':
    device = torch.device('cuda:0')
    tensor_args = {'device': device, 'dtype': torch.float32}
    np.random.seed(0)
    num_particles_per_goal, traj_len, dt, num_obst = 8, 64, 0.05, 5

    c, s = math.cos(-math.pi), math.sin(-math.pi)                     # target frame of examples/panda_environment.py:38-43
    target_H = torch.eye(4, dtype=torch.float64)
    target_H[:3, :3] = torch.tensor([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]], dtype=torch.float64) @ \
        torch.tensor([[c, 0., s], [0., 1., 0.], [-s, 0., c]], dtype=torch.float64)
    target_H[:3, 3] = torch.tensor([.3, .3, .3], dtype=torch.float64)
    target_H = target_H.to(**tensor_args).unsqueeze(0)

    panda_fk = PandaFK()
    n_dof = panda_fk._n_dofs
    start_q = torch.tensor([0.012, -0.57, 0., -2.81, 0., 3.037, 0.741], **tensor_args)
    start_state = torch.cat((start_q, torch.zeros_like(start_q)))
    q_goal = torch.tensor([-0.0138, -0.4637, 0.7626, -2.5, 0.371, 2.1212, 2.8481], **tensor_args)     # an IK solution of the target
    multi_goal_states = torch.cat([q_goal, torch.zeros_like(q_goal)]).unsqueeze(0)

    cost = CostComposite(n_dof, traj_len, [
        CostGP(n_dof, traj_len, start_state, dt, dict(sigma_start=0.001, sigma_gp=0.1), tensor_args),
        CostGoalPrior(n_dof, traj_len, multi_goal_states=multi_goal_states, num_particles_per_goal=num_particles_per_goal,
                      num_samples=1, sigma_goal_prior=0.01, tensor_args=tensor_args),
        CostCollision(n_dof, traj_len, field=LinkSelfDistanceField(margin=0.03, tensor_args=tensor_args), sigma_coll=0.1),
        CostCollision(n_dof, traj_len, field=LinkDistanceField(tensor_args=tensor_args), sigma_coll=0.02),
        CostGoal(n_dof, traj_len, field=EESE3DistanceField(target_H, tensor_args=tensor_args), sigma_goal=0.05),
    ], FK=panda_fk)

    planner = GPMP(num_particles_per_goal=num_particles_per_goal, traj_len=traj_len, opt_iters=1, dt=dt, n_dof=n_dof, step_size=0.3,
                   start_state=start_state, multi_goal_states=multi_goal_states, cost=cost,
                   sigma_start_init=0.0001, sigma_goal_init=0.1, sigma_gp_init=0.8,
                   sigma_start_sample=0.001, sigma_goal_sample=0.07, sigma_gp_sample=0.1,
                   solver_params=dict(delta=1e-2, trust_region=True, method='inverse'), seed=0, tensor_args=tensor_args)

    obstacle_spheres = np.zeros((1, num_obst, 4))
    obstacle_spheres[0, :, :3] = np.random.uniform([0.3, -0.3, 0.4], [0.7, 0.3, 0.9], (num_obst, 3))
    obstacle_spheres[0, :, 3] = np.random.uniform(0.05, 0.12, num_obst)
    obs = {'obstacle_spheres': torch.from_numpy(obstacle_spheres).to(**tensor_args)}

    _, _, c0 = planner.optimize(opt_iters=1, **obs)
    torch.cuda.synchronize()
    t0 = time.time()
    vel, pos, costs = planner.optimize(opt_iters=200, **obs)
    torch.cuda.synchronize()
    print(f'GPMP: 200 Gauss-Newton iterations of {planner.num_particles} particles in {(time.time() - t0) * 1e3:.1f} ms; '
          f'cost b^T K b {c0.mean().item():.4e} -> {costs.mean().item():.4e}')
    q = planner.particle_means[..., :n_dof].reshape(-1, n_dof).contiguous()
    links = ops.fk_link_positions(panda_fk, q)
    sph = obs['obstacle_spheres'][0]
    dist = (links[:, :, None, :] - sph[None, None, :, :3]).norm(dim=-1) - sph[None, None, :, 3]
    print('final means: min link-origin clearance to the spheres = %.3f m' % dist.min().item())
    ee = links.reshape(-1, traj_len, links.shape[1], 3)[:, -1, -1]
    print('final means: EE position error to the target = %s m' % (ee - target_H[0, :3, 3]).norm(dim=-1).cpu().numpy().round(4))
