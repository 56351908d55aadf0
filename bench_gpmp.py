#!/usr/bin/env python
"""Measurement of the Gauss-Newton GPMP path: ms per iteration of sgpmp_gpmp_step on one B200 for
the shipped Panda cost list (CostGP + CostGoalPrior + self-collision + obstacle spheres + EE SE(3) goal, T = 64, 4 goals
x K particles) at several problem counts, next to the reference algorithm on the host cores (oracle/gpmp.py: dense
A^T K A + dense solve per particle, the reference's own formulation, one problem).

    python bench_gpmp.py [--problems 1 64 1024] [--K 4] [--iters 20] [--no-cpu]

One JSON line per configuration.  The kernels are latency-bound (a sequential block Cholesky over T pivot blocks per
particle), so the figure of merit is ms per iteration and particle-iterations/s, not a roofline fraction.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, nargs="*", default=[1, 64, 1024])
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    from stoch_gpmp_b200 import scenarios as sc
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoal, CostGoalPrior
    from stoch_gpmp_b200.costs.fields import EESE3DistanceField, LinkDistanceField, LinkSelfDistanceField
    from stoch_gpmp_b200.planner import GPMPBatch
    from stoch_gpmp_b200.robots import PandaFK

    dev = torch.device("cuda", 0)
    n, T, G, K, dt = 7, 64, 4, args.K, 0.05
    tH = np.eye(4)
    tH[:3, :3] = np.diag([1., -1., -1.])
    tH[:3, 3] = [.3, .3, .3]
    cost_sig = dict(sigma_start=0.01, sigma_gp=0.3)
    solver = dict(delta=1e-2, trust_region=True)
    for method in ("inverse", "cholesky"):
        for B in args.problems:
            ta = dict(device=dev, dtype=torch.float32)
            start, goals, spheres = sc.panda_batch(B, G=G, O=5, seed0=0)
            s, g = torch.tensor(start, **ta), torch.tensor(goals, **ta)
            cl = [CostGP(n, T, s, dt, cost_sig, ta),
                  CostGoalPrior(n, T, multi_goal_states=g, num_particles_per_goal=K, num_samples=1, sigma_goal_prior=0.5, tensor_args=ta),
                  CostCollision(n, T, field=LinkSelfDistanceField(margin=0.03, tensor_args=ta), sigma_coll=0.2),
                  CostCollision(n, T, field=LinkDistanceField(tensor_args=ta), sigma_coll=0.1),
                  CostGoal(n, T, field=EESE3DistanceField(torch.tensor(tH, **ta)), sigma_goal=0.05)]
            pl = GPMPBatch(num_particles_per_goal=K, traj_len=T, opt_iters=args.iters, dt=dt, n_dof=n, step_size=0.3, start_state=s,
                           multi_goal_states=g, initial_particle_means='const_vel', cost=CostComposite(n, T, cl, FK=PandaFK()),
                           sigma_start_init=1e-3, sigma_start_sample=1e-3, sigma_goal_init=1e-3, sigma_goal_sample=1e-3,
                           sigma_gp_init=1., sigma_gp_sample=1., solver_params=dict(solver, method=method), tensor_args=ta)
            obs = {"obstacle_spheres": torch.tensor(spheres, **ta)}
            pl.optimize(opt_iters=3, **obs)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _, _, costs = pl.optimize(opt_iters=args.iters, **obs)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / args.iters)
            ms = min(ts)
            print(json.dumps({"kernel": "sgpmp_gpmp_step (assemble + solve)", "method": method, "problems": B, "particles": B * G * K,
                              "traj_len": T, "n_dof": n, "dtype": "f32 storage, f64 arithmetic", "ms_per_iteration": ms,
                              "particle_iterations_per_s": B * G * K / (ms * 1e-3), "launches_per_iteration": 2,
                              "mean_cost": float(costs.mean())}), flush=True)
    if not args.no_cpu:
        from oracle import gpmp as OG
        from oracle import prior as P
        start, goals, spheres = sc.panda_batch(1, G=G, O=5, seed0=0)
        spec = dict(n_dof=n, T=T, dt=dt, G=G, K=K, step_size=0.3, start=start[0], goals=goals[0], cost_sigma_start=0.01, cost_sigma_gp=0.3,
                    sigma_goal_prior=0.5, sigma_coll=0.1, spheres=spheres[0], self_margin=0.03, sigma_self=0.2, ee_target=tH, sigma_ee_goal=0.05,
                    delta=1e-2, trust_region=True, method="inverse")
        means = P.const_vel_trajectories(start[0], goals[0], dt, T, n, K).reshape(-1, T, 2 * n)
        OG.step(spec, means)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            means = OG.step(spec, means)["means_post"]
        sec = (time.perf_counter() - t0) / reps
        print(json.dumps({"cpu_baseline": "oracle/gpmp.py (numpy dense normal equations + dense solve, the reference's formulation)",
                          "kind": "port", "cores": os.cpu_count(), "problems": 1, "particles": G * K, "ms_per_iteration": sec * 1e3,
                          "particle_iterations_per_s": G * K / sec}), flush=True)


if __name__ == "__main__":
    main()
