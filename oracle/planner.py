"""Oracle: one whole StochGPMP iteration (sample -> costs -> softmax -> update) per problem.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: StochGPMP.reset / sample_and_eval / _get_costs / _update_distribution / optimize
(stoch_gpmp/planner.py:181-317).  A "spec" is a plain dict describing ONE planning problem:

  n_dof, T, dt, G, K, S, temperature, step_size,
  sigma_start_init, sigma_gp_init, sigma_goal_init        INIT prior   (planner.py:87-112)
  sigma_start_sample, sigma_gp_sample, sigma_goal_sample  SAMPLING prior (planner.py:115-140)
  start [d], goals [G,d] or None
  cost_sigma_start, cost_sigma_gp                         CostGP  (cost_functions.py:90-126)
  sigma_goal_prior or None                                CostGoalPrior (cost_functions.py:342-374)
  sigma_coll or None, and ONE of
     map [H,W], map_cell_size, map_origin (xi, yi)        ObstacleMap (envs/obst_map.py:112-147)
     spheres [O,4]                                        LinkDistanceField 'rbf' (costs/fields.py:63-79)
  self_margin, sigma_self (optional)                      LinkSelfDistanceField (costs/fields.py:89-127)
  num_interpolate, interp_range / self_num_interpolate, self_interp_range (optional)   link interpolation (fields.py:68-74)
  ee_target [4,4], sigma_ee_goal, ee_w_pos, ee_w_rot, ee_square (optional)   CostGoal + EESE3DistanceField
                                                          (cost_functions.py:282-321, fields.py:130-153)
"""
import numpy as np

from . import costs as C
from . import fk as FK
from . import prior as P
from . import sampler as SMP
from . import update as U


def spec_from_golden(g):
    """Build a spec from a tests/golden/*.npz record."""
    s = dict(n_dof=int(g['n_dof']), T=int(g['T']), dt=float(g['dt']), G=int(g['G']), K=int(g['K']), S=int(g['S']),
             temperature=float(g['temperature']), step_size=float(g['step_size']), start=g['start'],
             goals=g['goals'] if 'goals' in g.files else None,
             cost_sigma_start=float(g['cost_sigma_start']) if float(g['cost_sigma_start']) > 0 else None,      # None: CostGPTrajectory
             cost_sigma_gp=float(g['cost_sigma_gp']),
             sigma_goal_prior=float(g['sigma_goal_prior']) if float(g['sigma_goal_prior']) > 0 else None,
             sigma_coll=float(g['sigma_coll']) if float(g['sigma_coll']) > 0 else None,
             dtype=str(g['dtype']))
    for k in ('sigma_start_init', 'sigma_gp_init', 'sigma_goal_init',
              'sigma_start_sample', 'sigma_gp_sample', 'sigma_goal_sample'):
        s[k] = float(g[k])
    if 'map' in g.files:
        s['map'] = g['map']
        s['map_cell_size'] = float(g['map_cell_size'])
        s['map_origin'] = (int(g['map_origin'][0]), int(g['map_origin'][1]))
    if 'spheres' in g.files:
        s['spheres'] = g['spheres']
    if 'field_type' in g.files:
        s['field_type'] = str(g['field_type'])
        s['clamp_sdf'] = bool(g['clamp_sdf'])
    if 'self_margin' in g.files:
        s['self_margin'] = float(g['self_margin'])
        s['sigma_self'] = float(g['sigma_self'])
    for pre in ('', 'self_'):
        if pre + 'num_interpolate' in g.files:
            s[pre + 'num_interpolate'] = int(g[pre + 'num_interpolate'])
            s[pre + 'interp_range'] = (int(g[pre + 'interp_range'][0]), int(g[pre + 'interp_range'][1]))
    if 'ee_target' in g.files:
        s['ee_target'] = g['ee_target']
        s['sigma_ee_goal'] = float(g['sigma_ee_goal'])
        s['ee_w_pos'], s['ee_w_rot'], s['ee_square'] = float(g['ee_w_pos']), float(g['ee_w_rot']), bool(g['ee_square'])
    return s


def sampling_prior(spec):
    goal = spec['sigma_goal_sample'] if spec.get('goals') is not None else None
    D, O = P.precision_blocks(spec['T'], spec['dt'], spec['sigma_start_sample'], spec['sigma_gp_sample'], goal)
    return D, O, P.banded_factor(D, O)


def init_prior(spec):
    goal = spec['sigma_goal_init'] if spec.get('goals') is not None else None
    D, O = P.precision_blocks(spec['T'], spec['dt'], spec['sigma_start_init'], spec['sigma_gp_init'], goal)
    return D, O, P.banded_factor(D, O)


def initial_means(spec, init_eps=None, mode=None):
    """Particle means after reset() [NP,T,d].  init_eps in the reference's draw layout [K, G, M]
    (planner.py:213: _init_dist.sample(K) -> dist.sample((K,)) with G modes)."""
    n, T, dt, K = spec['n_dof'], spec['T'], spec['dt'], spec['K']
    if mode == 'const_vel':
        return P.const_vel_trajectories(spec['start'], spec['goals'], dt, T, n, K).reshape(-1, T, 2 * n)
    _, _, fac = init_prior(spec)
    if spec.get('goals') is not None:
        mu = P.const_vel_mean(spec['start'], spec['goals'], dt, T, n)              # [G,T,d]
    else:
        mu = np.repeat(spec['start'][None, None, :], T, axis=1)                    # [1,T,d]
    Gm = mu.shape[0]
    e = np.transpose(init_eps.reshape(K, Gm, T, 2 * n), (1, 0, 2, 3))              # [G,K,T,d]
    x = SMP.sample_banded(mu, fac['G'], fac['H'], e)                               # [G,K,T,d]
    return x.reshape(Gm * K, T, 2 * n)


def eval_costs(spec, samples, means, D, O, dtype=np.float64):
    """Per-term costs [NP,S] each, and the total in the reference's summation order
    (CostGP[start + gp], CostGoalPrior, self-collision, obstacle collision, EE goal — the order of the shipped
    lists, examples/panda_environment.py:90 — then the IS term)."""
    x = samples.astype(dtype)
    terms = {}
    # CostGPTrajectory (cost_functions.py:171-218) has no start factor
    terms['start'] = (C.cost_start(x, spec['start'].astype(dtype), spec['cost_sigma_start']) if spec.get('cost_sigma_start') is not None
                      else np.zeros(x.shape[:2], dtype=dtype))
    terms['gp'] = C.cost_gp(x, spec['dt'], spec['cost_sigma_gp'])
    total = terms['start'] + terms['gp']
    if spec.get('goals') is not None and spec.get('sigma_goal_prior') is not None:
        terms['goal'] = C.cost_goal_prior(x, spec['goals'].astype(dtype), spec['K'], spec['sigma_goal_prior'])
        total = total + terms['goal']
    if spec.get('self_margin') is not None:
        terms['self'] = C.cost_self_collision(x, spec['self_margin'], spec['sigma_self'], lambda q: FK.fk_all_links(q),
                                              spec.get('self_num_interpolate', 0), spec.get('self_interp_range', (5, 7)))
        total = total + terms['self']
    if spec.get('sigma_coll') is not None and 'map' in spec:
        terms['coll'] = C.cost_collision_map(x, spec['map'], spec['map_cell_size'], spec['map_origin'][0],
                                             spec['map_origin'][1], spec['sigma_coll'])
        total = total + terms['coll']
    if spec.get('sigma_coll') is not None and 'spheres' in spec:
        terms['coll'] = C.cost_collision_spheres(x, spec['spheres'].astype(dtype), spec['sigma_coll'],
                                                 lambda q: FK.fk_all_links(q), spec.get('field_type', 'rbf'),
                                                 spec.get('clamp_sdf', False), spec.get('num_interpolate', 0),
                                                 spec.get('interp_range', (5, 7)))
        total = total + terms['coll']
    if spec.get('ee_target') is not None:
        terms['ee'] = C.cost_ee_goal(x, spec['ee_target'], spec['sigma_ee_goal'], lambda q: FK.fk_all_links(q),
                                     spec.get('ee_w_pos', 1.), spec.get('ee_w_rot', 1.), spec.get('ee_square', True))
        total = total + terms['ee']
    terms['is'] = C.cost_importance(x, means.astype(dtype), D, O, spec['temperature'])
    total = total + terms['is']
    return terms, total


def iterate(spec, means, eps, dtype=np.float64):
    """One optimize() iteration.  means [NP,T,d]; eps [NP,S,T,d] (trajectory layout).
    Returns dict(samples, terms, costs, weights, grad, means_post)."""
    D, O, fac = sampling_prior(spec)
    samples = SMP.sample_banded(means.astype(np.float64), fac['G'], fac['H'], eps.astype(np.float64))
    terms, costs = eval_costs(spec, samples, means, D, O, dtype)
    means_post, grad, w = U.update(means.astype(np.float64), samples, costs.astype(np.float64),
                                   spec['temperature'], spec['step_size'])
    return dict(samples=samples, terms=terms, costs=costs, weights=w, grad=grad, means_post=means_post)


def eps_ref_to_traj(eps_ref, T, d):
    """[S, NP, M] (torch draw layout) -> [NP, S, T, d]."""
    S, NP, M = eps_ref.shape
    return np.transpose(eps_ref.reshape(S, NP, T, d), (1, 0, 2, 3))
