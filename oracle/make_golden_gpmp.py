"""Generate tests/golden/gpmp_*.npz by running the UNMODIFIED reference `GPMP` planner (Gauss-Newton GPMP,
stoch_gpmp/planner.py:352-661) on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.make_golden_gpmp

Each case builds the cost list like examples/panda_environment.py:83-98 (smaller shapes), constructs
GPMP(..., solver_params=dict(delta, trust_region, method)) with explicit initial particle means and calls
optimize(opt_iters=1) a few times, recording per iteration: means before/after, the step d_theta, the costs
optimize() returns (b^T K b of the LAST linear system, planner.py:545,635-637) and, for the first iteration,
the dense normal equations J^T J, g the reference solved.  The field Jacobians come from torch autograd through
oracle/fk.py's torch FK injected via the reference's own hook CostComposite(FK=...) (FK parity unpinned, as for
StochGPMP).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle import fk as ofk  # noqa: E402
from oracle import prior as P  # noqa: E402
from stoch_gpmp_b200.scenarios import PANDA_START, panda_goals, panda_spheres  # noqa: E402
from oracle.make_golden import shipped_target_H  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def run_case(name, *, n_dof, T, dt, G, K, dtype, start, goals, cost_sigmas, sigma_goal_prior, solver, step_size, iters,
             spheres=None, sigma_coll=None, self_field=None, interp=None, ee_goal=None, mean_noise=0.05, seed=0):
    ref = ref_loader.load()
    from stoch_gpmp.planner import GPMP
    ta = {'device': torch.device('cpu'), 'dtype': dtype}
    start_state = torch.tensor(start, **ta)
    goal_states = torch.tensor(goals, **ta)
    d = 2 * n_dof
    rec = dict(kind='gpmp', n_dof=n_dof, T=T, dt=dt, G=G, K=K, dtype=str(dtype).split('.')[-1], step_size=step_size,
               start=_np(start_state), goals=_np(goal_states), sigma_goal_prior=sigma_goal_prior,
               cost_sigma_start=cost_sigmas['sigma_start'], cost_sigma_gp=cost_sigmas['sigma_gp'],
               sigma_coll=-1.0 if sigma_coll is None else sigma_coll,
               delta=solver['delta'], trust_region=bool(solver['trust_region']), method=solver['method'])
    cost_list = [ref.CostGP(n_dof, T, start_state, dt, cost_sigmas, ta),
                 ref.CostGoalPrior(n_dof, T, multi_goal_states=goal_states, num_particles_per_goal=K, num_samples=1,
                                   sigma_goal_prior=sigma_goal_prior, tensor_args=ta)]
    FK = None
    obs = {}
    if self_field is not None:
        FK = ofk.fk_all_links_torch()
        cost_list.append(ref.CostCollision(n_dof, T, field=ref.LinkSelfDistanceField(margin=self_field[0], tensor_args=ta),
                                           sigma_coll=self_field[1], tensor_args=ta))
        rec['self_margin'], rec['sigma_self'] = self_field
    if spheres is not None:
        FK = ofk.fk_all_links_torch()
        fkw = {}
        if interp is not None:
            fkw = dict(num_interpolate=interp[0], link_interpolate_range=list(interp[1]))
            rec['num_interpolate'], rec['interp_range'] = interp[0], np.array(interp[1])
        cost_list.append(ref.CostCollision(n_dof, T, field=ref.LinkDistanceField(tensor_args=ta, **fkw), sigma_coll=sigma_coll,
                                           tensor_args=ta))
        sph = torch.tensor(spheres, **ta).reshape(1, -1, 4)
        obs = {'obstacle_spheres': sph}
        rec['spheres'] = _np(sph[0])
    if ee_goal is not None:     # CostGoal + EESE3DistanceField (SE3_distance = oracle/se3.py through the import stub)
        FK = ofk.fk_all_links_torch()
        tH = torch.tensor(ee_goal['target_H'], **ta).reshape(1, 4, 4)
        fld = ref.EESE3DistanceField(tH, w_pos=ee_goal.get('w_pos', 1.), w_rot=ee_goal.get('w_rot', 1.),
                                     square=ee_goal.get('square', True), tensor_args=ta)
        cost_list.append(ref.CostGoal(n_dof, T, field=fld, sigma_goal=ee_goal['sigma_goal'], tensor_args=ta))
        rec['ee_target'] = np.asarray(ee_goal['target_H'], dtype=np.float64).reshape(4, 4)
        rec['sigma_ee_goal'] = ee_goal['sigma_goal']
        rec['ee_w_pos'], rec['ee_w_rot'] = ee_goal.get('w_pos', 1.), ee_goal.get('w_rot', 1.)
        rec['ee_square'] = ee_goal.get('square', True)
    cost = ref.CostComposite(n_dof, T, cost_list, FK=FK, tensor_args=ta)

    # initial particle means [G, K, T, d]: straight lines start -> goal plus a smooth perturbation (so that the GP,
    # goal and collision residuals are all non-zero)
    rs = np.random.RandomState(seed)
    mu = P.const_vel_trajectories(np.array(start, dtype=np.float64), np.array(goals, dtype=np.float64), dt, T, n_dof, K)   # [G,K,T,d]
    tt = np.linspace(0, 1, T)[None, None, :, None]
    mu = mu + mean_noise * np.sin(np.pi * tt * rs.uniform(0.5, 2.0, (G, K, 1, d))) * rs.normal(0, 1, (G, K, 1, d))
    ipm = torch.tensor(mu, **ta)
    rec['initial_particle_means'] = _np(ipm)

    planner = GPMP(num_particles_per_goal=K, traj_len=T, opt_iters=1, dt=dt, n_dof=n_dof, step_size=step_size,
                   start_state=start_state, multi_goal_states=goal_states, initial_particle_means=ipm.clone(), cost=cost,
                   sigma_start_init=1e-3, sigma_start_sample=1e-3, sigma_goal_init=1e-3, sigma_goal_sample=1e-3,
                   sigma_gp_init=1., sigma_gp_sample=1., solver_params=solver, tensor_args=ta)
    for it in range(iters):
        pre = f'it{it}_'
        means_pre = _np(planner.particle_means)
        if it == 0:
            A, b, Kw = cost.get_linear_system(planner.particle_means.detach().clone(), **obs)
            JtJ, g = planner._get_grad_terms(A, b, Kw, delta=solver['delta'], trust_region=solver['trust_region'])
            rec['it0_JtJ'] = _np(JtJ)
            rec['it0_g'] = _np(g.squeeze(-1))
            rec['it0_AtKA'] = _np(A.transpose(1, 2) @ Kw @ A)
        vel_m, pos_m, costs = planner.optimize(**obs)
        means_post = _np(planner.particle_means)
        rec[pre + 'means_pre'] = means_pre
        rec[pre + 'means_post'] = means_post
        rec[pre + 'd_theta'] = (means_post.astype(np.float64) - means_pre.astype(np.float64)) / step_size
        rec[pre + 'costs'] = _np(costs)
        assert np.array_equal(_np(pos_m), means_post[..., :n_dof]) and np.array_equal(_np(vel_m), means_post[..., n_dof:])
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **rec)
    print(f'{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); costs it0 = {rec["it0_costs"]}, '
          f'last = {rec[f"it{iters - 1}_costs"]}, |d_theta| it0 = {np.abs(rec["it0_d_theta"]).max():.3e}')


def main(only=None):
    cases = dict(
        # CostGP + CostGoalPrior only (planar point robot; the occupancy map has no gradient, obst_map.py:164-182, so the
        # reference's GPMP cannot take it — field_factor.py:35 would differentiate a floor())
        gpmp_planar_f64=dict(n_dof=2, T=16, dt=0.1, G=2, K=2, dtype=torch.float64, start=[0.3, -0.2, 0., 0.],
                             goals=[[1.0, 0.6, 0, 0], [-0.8, 0.9, 0, 0]], cost_sigmas=dict(sigma_start=0.05, sigma_gp=0.5),
                             sigma_goal_prior=0.1, solver=dict(delta=1e-2, trust_region=True, method='cholesky'),
                             step_size=0.5, iters=3, mean_noise=0.2),
        # Panda: sphere field, trust region + Cholesky
        gpmp_panda_f64=dict(n_dof=7, T=12, dt=0.05, G=2, K=2, dtype=torch.float64, start=PANDA_START, goals=panda_goals(2, 20),
                            cost_sigmas=dict(sigma_start=0.01, sigma_gp=0.3), sigma_goal_prior=0.5,
                            solver=dict(delta=1e-2, trust_region=True, method='cholesky'), step_size=0.5, iters=3,
                            spheres=[[0.35, 0.0, 0.55, 0.18], [0.5, 0.1, 0.35, 0.15], [0.2, -0.1, 0.8, 0.12]], sigma_coll=0.1),
        # Panda: sphere field (interpolated) + self-collision, plain damping + dense solve
        gpmp_panda_self_f64=dict(n_dof=7, T=12, dt=0.05, G=2, K=1, dtype=torch.float64, start=PANDA_START, goals=panda_goals(2, 21),
                                 cost_sigmas=dict(sigma_start=0.01, sigma_gp=0.3), sigma_goal_prior=0.5,
                                 solver=dict(delta=1e-1, trust_region=False, method='inverse'), step_size=0.3, iters=2,
                                 spheres=panda_spheres(4, 21) + [[0.4, 0.05, 0.5, 0.2]], sigma_coll=0.1, self_field=(0.15, 0.2),
                                 interp=(2, (5, 7))),
        # the shipped Panda cost list [CostGP, CostGoalPrior, self, spheres, CostGoal(EE SE(3))] under GPMP
        gpmp_panda_ee_f64=dict(n_dof=7, T=12, dt=0.05, G=1, K=3, dtype=torch.float64, start=PANDA_START, goals=panda_goals(1, 22),
                               cost_sigmas=dict(sigma_start=0.01, sigma_gp=0.3), sigma_goal_prior=0.5,
                               solver=dict(delta=1e-2, trust_region=True, method='inverse'), step_size=0.3, iters=2,
                               spheres=panda_spheres(3, 22), sigma_coll=0.1, self_field=(0.15, 0.2),
                               ee_goal=dict(target_H=shipped_target_H(), sigma_goal=0.05, w_pos=1.5, w_rot=0.7)),
        gpmp_panda_ee_nosq_f64=dict(n_dof=7, T=10, dt=0.05, G=2, K=1, dtype=torch.float64, start=PANDA_START, goals=panda_goals(2, 23),
                                    cost_sigmas=dict(sigma_start=0.01, sigma_gp=0.3), sigma_goal_prior=0.5,
                                    solver=dict(delta=1e-2, trust_region=False, method='cholesky'), step_size=0.3, iters=1,
                                    ee_goal=dict(target_H=shipped_target_H(), sigma_goal=0.1, square=False)),
        # fp32 run of the first Panda case (the reference solves in fp32; compared loosely)
        gpmp_panda_f32=dict(n_dof=7, T=12, dt=0.05, G=2, K=2, dtype=torch.float32, start=PANDA_START, goals=panda_goals(2, 20),
                            cost_sigmas=dict(sigma_start=0.01, sigma_gp=0.3), sigma_goal_prior=0.5,
                            solver=dict(delta=1e-2, trust_region=True, method='cholesky'), step_size=0.5, iters=2,
                            spheres=[[0.35, 0.0, 0.55, 0.18], [0.5, 0.1, 0.35, 0.15], [0.2, -0.1, 0.8, 0.12]], sigma_coll=0.1),
    )
    for name, kw in cases.items():
        if only and name not in only:
            continue
        run_case(name, **kw)


if __name__ == '__main__':
    main(set(sys.argv[1:]) or None)
