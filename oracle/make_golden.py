"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container only:

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

Every case builds the planner exactly like examples/planar_environment.py:62-110 /
examples/panda_environment.py:83-146 (smaller shapes), calls `optimize()` with opt_iters=1 a few
times, and records inputs, the eps torch drew (re-drawn from the saved generator state — SURVEY.md
Appendix A, bit-exact), the reference's scale_tril / Sigma_inv, samples, per-term costs, total costs,
weights, grad and means before/after.  The Panda cases inject oracle/fk.py's torch FK through the
reference's own hook CostComposite(FK=...) because torch_robotics is absent (FK parity unpinned).
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle import fk as ofk  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def run_case(name, *, n_dof, T, dt, G, K, S, dtype, seed, start, goals, planner_sigmas, cost_sigmas,
             temperature, step_size, iters, map_params=None, spheres=None, initial_particle_means=None,
             sigma_coll=None, sigma_goal_prior=None, store_L=True, self_field=None, field_type='rbf', clamp_sdf=False,
             interp=None, self_interp=None, ee_goal=None, gp_trajectory=False, compact=False):
    """compact=True (full-size shapes, e.g. the C3 shape T=64, S=256): the big arrays — eps, samples, reset draws, dense Sigma_inv / L —
    are NOT stored; instead the torch CPU generator state before each draw is (5 KB), from which a test re-draws the identical eps
    (`torch.set_rng_state(...); torch.empty(S, NP, M).normal_()`: bit-exact for the same torch build), plus the first 2 samples of
    every particle for a direct sample check."""
    ref = ref_loader.load()
    ta = {'device': torch.device('cpu'), 'dtype': dtype}
    start_state = torch.tensor(start, **ta)
    goal_states = None if goals is None else torch.tensor(goals, **ta)
    rec = dict(n_dof=n_dof, T=T, dt=dt, G=G, K=K, S=S, seed=seed, temperature=temperature,
               step_size=step_size, dtype=str(dtype).split('.')[-1],
               start=_np(start_state), sigma_coll=-1.0 if sigma_coll is None else sigma_coll,
               sigma_goal_prior=-1.0 if sigma_goal_prior is None else sigma_goal_prior)
    if goals is not None:
        rec['goals'] = _np(goal_states)
    for k, v in planner_sigmas.items():
        rec[k] = v
    for k, v in cost_sigmas.items():
        rec['cost_' + k] = v

    if gp_trajectory:        # CostGPTrajectory (cost_functions.py:171-218): GP transition factors without the start factor
        rec['cost_sigma_start'] = -1.0
        cost_list = [ref.CostGPTrajectory(n_dof, T, start_state, dt, cost_sigmas, ta)]
    else:
        cost_list = [ref.CostGP(n_dof, T, start_state, dt, cost_sigmas, ta)]
    term_names = ['gp']
    if goals is not None and sigma_goal_prior is not None:
        cost_list.append(ref.CostGoalPrior(n_dof, T, multi_goal_states=goal_states, num_particles_per_goal=K,
                                           num_samples=S, sigma_goal_prior=sigma_goal_prior, tensor_args=ta))
        term_names.append('goal')
    FK = None
    obs = {}
    if map_params is not None:
        random.seed(map_params['seed'])
        np.random.seed(map_params['seed'])
        obst_map = ref.generate_obstacle_map(
            map_dim=map_params['map_dim'], obst_list=[], cell_size=map_params['cell_size'], random_gen=True,
            num_obst=map_params['num_obst'], rand_limits=map_params['rand_limits'],
            rand_rect_shape=[2, 2], tensor_args=ta)[0]
        cost_list.append(ref.CostCollision(n_dof, T, field=obst_map, sigma_coll=sigma_coll))
        term_names.append('coll')
        rec['map'] = obst_map.map.copy()
        rec['map_cell_size'] = map_params['cell_size']
        rec['map_origin'] = np.array([obst_map.origin_xi, obst_map.origin_yi])
    if self_field is not None:      # (margin, sigma_self): examples/panda_environment.py:67,88 — listed BEFORE the obstacle cost
        from stoch_gpmp.costs.fields import LinkSelfDistanceField
        FK = ofk.fk_all_links_torch()
        skw = {}
        if self_interp is not None:      # (num_interpolate, [lo, hi])  costs/fields.py:117-123
            skw = dict(num_interpolate=self_interp[0], link_interpolate_range=list(self_interp[1]))
            rec['self_num_interpolate'], rec['self_interp_range'] = self_interp[0], np.array(self_interp[1])
        cost_list.append(ref.CostCollision(n_dof, T, field=LinkSelfDistanceField(margin=self_field[0], tensor_args=ta, **skw),
                                           sigma_coll=self_field[1]))
        term_names.append('self')
        rec['self_margin'], rec['sigma_self'] = self_field
    if spheres is not None:
        FK = ofk.fk_all_links_torch()
        fkw = {}
        if interp is not None:           # (num_interpolate, [lo, hi])  costs/fields.py:68-74
            fkw = dict(num_interpolate=interp[0], link_interpolate_range=list(interp[1]))
            rec['num_interpolate'], rec['interp_range'] = interp[0], np.array(interp[1])
        field = ref.LinkDistanceField(field_type=field_type, clamp_sdf=clamp_sdf, tensor_args=ta, **fkw)
        if field_type != 'rbf':
            rec['field_type'], rec['clamp_sdf'] = field_type, clamp_sdf
        cost_list.append(ref.CostCollision(n_dof, T, field=field, sigma_coll=sigma_coll))
        term_names.append('coll')
        sph = torch.tensor(spheres, **ta).reshape(1, -1, 4)
        obs = {'obstacle_spheres': sph}
        rec['spheres'] = _np(sph[0])
    if ee_goal is not None:
        # CostGoal + EESE3DistanceField, last in the shipped list (examples/panda_environment.py:69,89-90); SE3_distance
        # itself is the restatement of oracle/se3.py bound through the import stub (parity unpinned there)
        FK = ofk.fk_all_links_torch()
        tH = torch.tensor(ee_goal['target_H'], **ta).reshape(1, 4, 4)
        fld = ref.EESE3DistanceField(tH, w_pos=ee_goal.get('w_pos', 1.), w_rot=ee_goal.get('w_rot', 1.),
                                     square=ee_goal.get('square', True), tensor_args=ta)
        cost_list.append(ref.CostGoal(n_dof, T, field=fld, sigma_goal=ee_goal['sigma_goal'], tensor_args=ta))
        term_names.append('ee')
        rec['ee_target'] = np.asarray(ee_goal['target_H'], dtype=np.float64).reshape(4, 4)
        rec['sigma_ee_goal'] = ee_goal['sigma_goal']
        rec['ee_w_pos'], rec['ee_w_rot'] = ee_goal.get('w_pos', 1.), ee_goal.get('w_rot', 1.)
        rec['ee_square'] = ee_goal.get('square', True)
    cost = ref.CostComposite(n_dof, T, cost_list, FK=FK, tensor_args=ta)

    ipm = initial_particle_means
    if ipm is not None and not isinstance(ipm, str):
        ipm = torch.tensor(ipm, **ta)
        rec['initial_particle_means'] = _np(ipm)
    elif isinstance(ipm, str):
        rec['initial_particle_means_mode'] = ipm

    def _run(tag):
        torch.manual_seed(seed)
        st0 = torch.get_rng_state()
        planner = ref.StochGPMP(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=1, dt=dt, n_dof=n_dof,
                                step_size=step_size, temperature=temperature, start_state=start_state,
                                multi_goal_states=goal_states, initial_particle_means=ipm, cost=cost,
                                seed=seed, tensor_args=ta, **planner_sigmas)
        NP = planner.num_particles
        M = T * 2 * n_dof
        st_after_reset = torch.get_rng_state()
        # re-draw what reset() drew (planner.py:213 then :227)
        torch.set_rng_state(st0)
        if initial_particle_means is None:
            rec[tag + 'init_eps'] = _np(torch.empty(K, NP // K, M, dtype=dtype).normal_())       # [K, G, M]
        discard = _np(torch.empty(S, NP, M, dtype=dtype).normal_())
        assert torch.equal(torch.get_rng_state(), st_after_reset)
        rec[tag + 'means_reset'] = _np(planner.particle_means)
        if not compact:
            rec[tag + 'reset_discard_eps'] = discard
            rec[tag + 'state_samples_reset'] = _np(planner.state_samples)
            rec[tag + 'Sigma_inv'] = _np(planner.Sigma_inv)
        if store_L and not compact:
            rec[tag + 'L'] = _np(planner._sample_dist.dist._unbroadcasted_scale_tril[0])

        for it in range(iters):
            st = torch.get_rng_state()
            means_pre = _np(planner.particle_means)
            out = planner.optimize(**obs)
            st_post = torch.get_rng_state()
            torch.set_rng_state(st)
            eps = torch.empty(S, NP, M, dtype=dtype).normal_()
            assert torch.equal(torch.get_rng_state(), st_post)
            pos_m, vel_m, pos_s, vel_s, costs, grad = out
            samples = planner.state_samples
            # bit-exact eps reproduction check
            L = planner._sample_dist.dist._unbroadcasted_scale_tril[0]
            x_chk = (torch.tensor(means_pre, **ta).reshape(1, NP, M) + (L @ eps.unsqueeze(-1)).squeeze(-1))
            x_chk = x_chk.view(S, NP, T, 2 * n_dof).transpose(1, 0)
            assert torch.equal(x_chk, samples), "eps reproduction is not bit-exact"
            pre = f'it{it}_'
            rec[tag + pre + 'means_pre'] = means_pre
            if compact:
                rec[tag + pre + 'rng_state'] = st.numpy().copy()        # eps = set_rng_state(this); empty(S, NP, M).normal_()
                rec[tag + pre + 'samples_head'] = _np(samples[:, :2])   # [NP, 2, T, d]
                rec[tag + pre + 'eps_head'] = _np(eps[:2])              # [2, NP, M] guards the re-draw itself
            else:
                rec[tag + pre + 'eps'] = _np(eps)                       # [S, NP, M]  (reference layout)
                rec[tag + pre + 'samples'] = _np(samples)               # [NP, S, T, d]
            rec[tag + pre + 'costs'] = _np(costs)                       # [NP, S]
            rec[tag + pre + 'grad'] = _np(grad)
            rec[tag + pre + 'weights'] = _np(planner._weights).reshape(NP, S)
            rec[tag + pre + 'means_post'] = _np(planner.particle_means)
            assert np.array_equal(_np(pos_m), means_pre[..., :n_dof])
            # per-term costs from the reference's own cost objects
            trajs = samples.reshape(-1, T, 2 * n_dof)
            x_trajs = None
            if FK is not None:
                x_trajs = FK(trajs.reshape(-1, 2 * n_dof)[:, :n_dof]).reshape(trajs.shape[0], T, -1, 4, 4)
            tot = 0
            for nm, c in zip(term_names, cost_list):
                val = c(trajs, x_trajs=x_trajs, **obs)
                tot = tot + val
                rec[tag + pre + 'term_' + nm] = _np(val.reshape(NP, S))
            if term_names[0] == 'gp' and not gp_trajectory:
                # split CostGP into its start and transition parts with the reference's own factors
                cgp = cost_list[0]
                err_p = cgp.start_prior.get_error(trajs[:, [0]], calc_jacobian=False)
                sc = (err_p @ cgp.start_prior.K.unsqueeze(0) @ err_p.transpose(1, 2)).squeeze()
                rec[tag + pre + 'term_start'] = _np(sc.reshape(NP, S))
            rec[tag + pre + 'term_is'] = _np(costs - tot.reshape(NP, S))

    _run('')
    if dtype == torch.float32:
        # "the same prior factor": re-run the unmodified reference with torch's precision->scale_tril
        # evaluated in fp64 (on the reference's own fp64-built precision) and cast to fp32.
        import torch.distributions.multivariate_normal as mvn
        orig = mvn._precision_to_scale_tril
        ta64 = {'device': torch.device('cpu'), 'dtype': torch.float64}
        s64 = torch.tensor(start, **ta64)
        g64 = None if goals is None else torch.tensor(goals, **ta64)
        cands = []
        for which in ('init', 'sample'):
            Ks = torch.eye(2 * n_dof, **ta64) / planner_sigmas[f'sigma_start_{which}'] ** 2
            Kg = None if goals is None else torch.eye(2 * n_dof, **ta64) / planner_sigmas[f'sigma_goal_{which}'] ** 2
            from stoch_gpmp.costs.factors.gp_factor import GPFactor
            Qi = GPFactor(n_dof, planner_sigmas[f'sigma_gp_{which}'], dt, T - 1, ta64).Q_inv[0]
            pr = ref.MultiMPPrior(T - 1, dt, 2 * n_dof, n_dof, Ks, Qi, s64, K_g_inv=Kg, goal_states=g64, tensor_args=ta64)
            cands.append((pr.Sigma_inv, orig(pr.Sigma_inv)))

        def patched(P):
            if P.dtype != torch.float32:
                return orig(P)
            P0 = P.reshape(-1, P.shape[-2], P.shape[-1])[0].double()
            errs = [float((P0 - c[0]).abs().max() / c[0].abs().max()) for c in cands]
            i = int(np.argmin(errs))
            assert errs[i] < 1e-3, errs
            return cands[i][1].to(torch.float32).expand(P.shape).contiguous()
        mvn._precision_to_scale_tril = patched
        try:
            _run('sameL_')
        finally:
            mvn._precision_to_scale_tril = orig
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **rec)
    print(f'{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); '
          f'max weight it0 = {rec["it0_weights"].max():.3f}, cost range = '
          f'[{rec["it0_costs"].min():.3e}, {rec["it0_costs"].max():.3e}]')


from stoch_gpmp_b200.scenarios import (PANDA_SIGMAS, PANDA_START, PLANAR_GOALS, PLANAR_SIGMAS, panda_goals,  # noqa: E402
                              panda_spheres)

PLANAR_MAP = dict(map_dim=[20, 20], cell_size=0.1, num_obst=15, rand_limits=[[-7.5, 7.5], [-7.5, 7.5]], seed=0)
SOFT_MAP = dict(map_dim=[8, 8], cell_size=0.1, num_obst=4, rand_limits=[[-1.5, 1.5], [-1.5, 1.5]], seed=1)


def shipped_target_H():
    """target frame of examples/panda_environment.py:39-43: rot = Rz(-pi) Ry(-pi), trans = (.3, .3, .3)."""
    cz, sz = np.cos(-np.pi), np.sin(-np.pi)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Ry = np.array([[cz, 0, sz], [0, 1, 0], [-sz, 0, cz]])
    H = np.eye(4)
    H[:3, :3] = Rz @ Ry
    H[:3, 3] = [.3, .3, .3]
    return H


def main(only=None):
    global run_case
    if only:
        _rc = run_case

        def run_case(name, **kw):      # noqa: F811  (filter: regenerate only the named cases)
            if name in only:
                _rc(name, **kw)
    # C1-like: shipped planar sigmas (examples/planar_environment.py:52-96), fp64, small shapes
    run_case('planar_f64', n_dof=2, T=16, dt=0.02, G=3, K=2, S=6, dtype=torch.float64, seed=0,
             start=[-9, -9, 0, 0], goals=PLANAR_GOALS, planner_sigmas=PLANAR_SIGMAS,
             cost_sigmas=dict(sigma_start=0.001, sigma_gp=0.1), sigma_coll=1e-5, sigma_goal_prior=0.001,
             temperature=1., step_size=0.5, iters=2, map_params=PLANAR_MAP)
    # shipped traj_len (64) and K=5; fewer samples to keep the fixture small
    run_case('planar_f64_T64', n_dof=2, T=64, dt=0.02, G=3, K=5, S=4, dtype=torch.float64, seed=1,
             start=[-9, -9, 0, 0], goals=PLANAR_GOALS, planner_sigmas=PLANAR_SIGMAS,
             cost_sigmas=dict(sigma_start=0.001, sigma_gp=0.1), sigma_coll=1e-5, sigma_goal_prior=0.001,
             temperature=1., step_size=0.5, iters=1, map_params=PLANAR_MAP)
    # soft sigmas, start/goals near the origin and a high temperature: softmax weights are NOT one-hot
    # (exercises the weighted update; with the shipped sigmas every example is one-hot)
    for nm, dt_ in (('planar_soft_f64', torch.float64), ('planar_soft_f32', torch.float32)):
        run_case(nm, n_dof=2, T=16, dt=0.1, G=2, K=2, S=8, dtype=dt_, seed=2,
                 start=[0.3, -0.2, 0, 0], goals=[[1.0, 0.6, 0, 0], [-0.8, 0.9, 0, 0]],
                 planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.1,
                                     sigma_start_sample=1.0, sigma_goal_sample=1.0, sigma_gp_sample=2.0),
                 cost_sigmas=dict(sigma_start=0.5, sigma_gp=2.), sigma_coll=0.5, sigma_goal_prior=1.0,
                 temperature=20., step_size=0.5, iters=2, map_params=SOFT_MAP)
    # not goal-directed (multi_goal_states=None -> G=1, no goal factor; planner.py:56-58)
    run_case('planar_nogoal_f64', n_dof=2, T=12, dt=0.05, G=1, K=3, S=5, dtype=torch.float64, seed=3,
             start=[-2, 1, 0.5, -0.5], goals=None, planner_sigmas=dict(PLANAR_SIGMAS, sigma_gp_sample=1., sigma_gp_init=2.),
             cost_sigmas=dict(sigma_start=0.01, sigma_gp=0.5), sigma_coll=0.1, sigma_goal_prior=None,
             temperature=2., step_size=1.0, iters=1, map_params=PLANAR_MAP)
    # 'const_vel' initial means (planner.py:142-155,197-201)
    run_case('planar_constvel_f64', n_dof=2, T=10, dt=0.02, G=3, K=2, S=4, dtype=torch.float64, seed=4,
             start=[-9, -9, 0, 0], goals=PLANAR_GOALS, planner_sigmas=PLANAR_SIGMAS,
             cost_sigmas=dict(sigma_start=0.001, sigma_gp=0.1), sigma_coll=1e-5, sigma_goal_prior=0.001,
             temperature=1., step_size=0.5, iters=1, map_params=PLANAR_MAP, initial_particle_means='const_vel')
    # Panda (examples/panda_environment.py:72-118 sigmas), FK injected, fp32 and fp64
    for nm, dt_ in (('panda_f32', torch.float32), ('panda_f64', torch.float64)):
        run_case(nm, n_dof=7, T=16, dt=0.05, G=2, K=2, S=8, dtype=dt_, seed=5,
                 start=PANDA_START, goals=panda_goals(2, 0), planner_sigmas=PANDA_SIGMAS,
                 cost_sigmas=dict(sigma_start=0.0001, sigma_gp=0.0007), sigma_coll=0.01, sigma_goal_prior=20.,
                 temperature=1., step_size=0.1, iters=2, spheres=panda_spheres(3, 0))
    # shipped Panda cost list minus the EE-goal term: + LinkSelfDistanceField(margin=0.03), sigma_self=0.01
    # (examples/panda_environment.py:67,78,88-90), fp32 and fp64; a soft fp64 variant with live weights
    for nm, dt_ in (('panda_self_f32', torch.float32), ('panda_self_f64', torch.float64)):
        run_case(nm, n_dof=7, T=16, dt=0.05, G=2, K=2, S=8, dtype=dt_, seed=7,
                 start=PANDA_START, goals=panda_goals(2, 2), planner_sigmas=PANDA_SIGMAS,
                 cost_sigmas=dict(sigma_start=0.0001, sigma_gp=0.0007), sigma_coll=0.01, sigma_goal_prior=20.,
                 temperature=1., step_size=0.1, iters=2, spheres=panda_spheres(5, 2), self_field=(0.03, 0.01))
    run_case('panda_self_soft_f64', n_dof=7, T=16, dt=0.05, G=2, K=1, S=16, dtype=torch.float64, seed=8,
             start=PANDA_START, goals=panda_goals(2, 3),
             planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                 sigma_start_sample=4.0, sigma_goal_sample=4.0, sigma_gp_sample=0.5),
             initial_particle_means='const_vel',
             cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.3, sigma_goal_prior=20.,
             temperature=200., step_size=0.5, iters=2, spheres=None, self_field=(0.15, 0.1))
    # LinkDistanceField field_type variants (costs/fields.py:80-86): sdf, clamped sdf, occupancy — soft sigmas so that
    # the collision term matters; spheres placed ON the arm's workspace so that contacts really occur
    near = [[0.35, 0.0, 0.55, 0.18], [0.5, 0.1, 0.35, 0.15], [0.2, -0.1, 0.8, 0.12]]
    for nm, ft, cl, dt_ in (('panda_sdf_f64', 'sdf', False, torch.float64), ('panda_sdfclamp_f32', 'sdf', True, torch.float32),
                            ('panda_occ_f64', 'occupancy', False, torch.float64)):
        run_case(nm, n_dof=7, T=12, dt=0.05, G=2, K=1, S=12, dtype=dt_, seed=9,
                 start=PANDA_START, goals=panda_goals(2, 4),
                 planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                     sigma_start_sample=0.5, sigma_goal_sample=0.5, sigma_gp_sample=0.5),
                 initial_particle_means='const_vel',
                 cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.05, sigma_goal_prior=20.,
                 temperature=200., step_size=0.5, iters=1, spheres=near, field_type=ft, clamp_sdf=cl)
    # Panda with soft sigmas / temperature (non-degenerate weights).  fp32 only constructs for mild
    # conditioning (torch's fp32 Cholesky of the precision fails otherwise — SURVEY §6/§7), so the
    # softer variant is fp64.
    run_case('panda_soft_f32', n_dof=7, T=16, dt=0.05, G=2, K=1, S=16, dtype=torch.float32, seed=6,
             start=PANDA_START, goals=panda_goals(2, 1),
             planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                 sigma_start_sample=0.5, sigma_goal_sample=0.5, sigma_gp_sample=0.5),
             initial_particle_means='const_vel',
             cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.3, sigma_goal_prior=20.,
             temperature=200., step_size=0.5, iters=2, spheres=panda_spheres(5, 1))
    run_case('panda_soft_f64', n_dof=7, T=16, dt=0.05, G=2, K=1, S=16, dtype=torch.float64, seed=6,
             start=PANDA_START, goals=panda_goals(2, 1),
             planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                 sigma_start_sample=4.0, sigma_goal_sample=4.0, sigma_gp_sample=0.5),
             initial_particle_means='const_vel',
             cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.3, sigma_goal_prior=20.,
             temperature=200., step_size=0.5, iters=2, spheres=panda_spheres(5, 1))


    # CostGPTrajectory in place of CostGP (no start factor), planar
    run_case('planar_gptraj_f64', n_dof=2, T=12, dt=0.1, G=2, K=2, S=6, dtype=torch.float64, seed=13,
             start=[0.3, -0.2, 0, 0], goals=[[1.0, 0.6, 0, 0], [-0.8, 0.9, 0, 0]],
             planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.1,
                                 sigma_start_sample=1.0, sigma_goal_sample=1.0, sigma_gp_sample=2.0),
             cost_sigmas=dict(sigma_gp=2.), sigma_coll=0.5, sigma_goal_prior=1.0,
             temperature=20., step_size=0.5, iters=1, map_params=SOFT_MAP, gp_trajectory=True)
    # the COMPLETE shipped Panda cost list [CostGP, CostGoalPrior, self, obstacle, CostGoal(EESE3DistanceField)]
    # (examples/panda_environment.py:66-90, sigma_goal = 0.00007), fp64 and fp32
    for nm, dt_ in (('panda_shipped_f64', torch.float64), ('panda_shipped_f32', torch.float32)):
        run_case(nm, n_dof=7, T=16, dt=0.05, G=1, K=3, S=8, dtype=dt_, seed=10,
                 start=PANDA_START, goals=panda_goals(1, 5), planner_sigmas=PANDA_SIGMAS,
                 cost_sigmas=dict(sigma_start=0.0001, sigma_gp=0.0007), sigma_coll=0.01, sigma_goal_prior=20.,
                 temperature=1., step_size=0.1, iters=2, spheres=panda_spheres(5, 5), self_field=(0.03, 0.01),
                 ee_goal=dict(target_H=shipped_target_H(), sigma_goal=0.00007))
    # soft variant with live weights, non-default w_pos / w_rot and the un-squared distance
    tH = shipped_target_H()
    tH[:3, 3] = [0.45, 0.1, 0.5]
    run_case('panda_ee_soft_f64', n_dof=7, T=12, dt=0.05, G=2, K=1, S=12, dtype=torch.float64, seed=11,
             start=PANDA_START, goals=panda_goals(2, 6),
             planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                 sigma_start_sample=4.0, sigma_goal_sample=4.0, sigma_gp_sample=0.5),
             initial_particle_means='const_vel',
             cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.3, sigma_goal_prior=20.,
             temperature=200., step_size=0.5, iters=1, spheres=panda_spheres(3, 6),
             ee_goal=dict(target_H=tH, sigma_goal=0.05, w_pos=2.0, w_rot=0.5, square=False))
    # link interpolation on both link fields (costs/fields.py:68-74, :117-123): 3 extra points between frames 5-6
    # and 6-7 for the sphere field (the reference's default range), 2 between frames 3-4 .. 5-6 for the self field
    for nm, dt_ in (('panda_interp_f64', torch.float64), ('panda_interp_f32', torch.float32)):
        run_case(nm, n_dof=7, T=12, dt=0.05, G=2, K=1, S=12, dtype=dt_, seed=12,
                 start=PANDA_START, goals=panda_goals(2, 7),
                 planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                     sigma_start_sample=4.0 if dt_ == torch.float64 else 0.5,
                                     sigma_goal_sample=4.0 if dt_ == torch.float64 else 0.5, sigma_gp_sample=0.5),
                 initial_particle_means='const_vel',
                 cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.3, sigma_goal_prior=20.,
                 temperature=200., step_size=0.5, iters=1, spheres=panda_spheres(4, 7), self_field=(0.15, 0.1),
                 interp=(3, (5, 7)), self_interp=(2, (3, 6)))
    # BASELINE shape (C3: 4 goals x K = 1, T = 64, n = 7; S = 256) with the shipped Panda sigmas and 5 spheres, fp64, compact
    # record (see run_case): pins the full-size shape by the reference itself, not only by the numpy restatement
    run_case('c3_panda_f64', n_dof=7, T=64, dt=0.05, G=4, K=1, S=256, dtype=torch.float64, seed=21,
             start=PANDA_START, goals=panda_goals(4, 0), planner_sigmas=PANDA_SIGMAS, initial_particle_means='const_vel',
             cost_sigmas=dict(sigma_start=0.0001, sigma_gp=0.0007), sigma_coll=0.01, sigma_goal_prior=20.,
             temperature=1., step_size=0.1, iters=2, spheres=panda_spheres(5, 0), compact=True)
    # the same shape with soft sigmas (live softmax weights at full size)
    run_case('c3_panda_soft_f64', n_dof=7, T=64, dt=0.05, G=4, K=1, S=256, dtype=torch.float64, seed=22,
             start=[0.05 * v for v in PANDA_START], goals=[[0.05 * v for v in g] for g in panda_goals(4, 3)],
             planner_sigmas=dict(sigma_start_init=0.5, sigma_goal_init=0.5, sigma_gp_init=0.8,
                                 sigma_start_sample=4.0, sigma_goal_sample=4.0, sigma_gp_sample=0.5),
             initial_particle_means='const_vel',
             cost_sigmas=dict(sigma_start=0.5, sigma_gp=0.5), sigma_coll=0.3, sigma_goal_prior=20.,
             temperature=200., step_size=0.5, iters=1, spheres=panda_spheres(5, 3), compact=True)


if __name__ == '__main__':
    main(set(sys.argv[1:]) or None)
