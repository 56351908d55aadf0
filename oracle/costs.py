"""Oracle: per-trajectory-sample cost terms.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference call sites restated (all in /root/reference/stoch_gpmp):
  CostComposite.eval   costs/cost_functions.py:47-58   sum of children in list order, b = p*S + s
  CostGP.eval          costs/cost_functions.py:128-146 start + GP transition quadratic forms
  GPFactor.get_error   costs/factors/gp_factor.py:54-58   e_t = x_{t+1} - Phi x_t
  UnaryFactor          costs/factors/unary_factor.py:19-23
  CostGoalPrior.eval   costs/cost_functions.py:376-388 goal of particle p is p // K
  CostCollision.eval   costs/cost_functions.py:247-261 + FieldFactor.get_error
                       costs/factors/field_factor.py:18-32  (time steps 1..T-1 only)
  ObstacleMap.get_collisions  envs/obst_map.py:164-182
  LinkDistanceField.compute_cost ('rbf')  costs/fields.py:63-79
  LinkSelfDistanceField.compute_cost      costs/fields.py:114-124
  link interpolation          costs/fields.py:68-74, :117-123
  CostGoal.eval + EESE3DistanceField.compute_cost   costs/cost_functions.py:308-321, costs/fields.py:142-150
  importance-sampling term    planner.py:229-237   tau * x^T Sigma^-1 mu

Every function takes samples x [NP, S, T, d] and returns [NP, S].
"""
import numpy as np

from . import prior as _prior


def cost_start(x, start_state, sigma_start):
    e = start_state[None, None, :] - x[:, :, 0, :]
    return (e * e).sum(-1) / sigma_start ** 2


def cost_gp(x, dt, sigma_gp):
    d = x.shape[-1]
    n = d // 2
    p, v = x[..., :n], x[..., n:]
    ep = p[:, :, 1:] - p[:, :, :-1] - dt * v[:, :, :-1]
    ev = v[:, :, 1:] - v[:, :, :-1]
    qc = 1.0 / sigma_gp ** 2
    q11, q12, q22 = 12.0 * dt ** -3.0 * qc, -6.0 * dt ** -2.0 * qc, 4.0 * dt ** -1.0 * qc
    return (q11 * ep * ep + 2.0 * q12 * ep * ev + q22 * ev * ev).sum((-1, -2))


def cost_goal_prior(x, goal_states, K, sigma_goal_prior):
    NP = x.shape[0]
    g = np.arange(NP) // K
    e = goal_states[g][:, None, :] - x[:, :, -1, :]
    return (e * e).sum(-1) / sigma_goal_prior ** 2


def map_lookup(xy, occ_map, cell_size, origin_xi, origin_yi):
    """Occupancy lookup with the reference's exact arithmetic order, in xy's dtype:
    X*(1/cell) + offset  (two roundings, no FMA) -> floor -> int -> clamp -> map[iy, ix].
    Note the reference clamps ix with shape[0] and iy with shape[1] (obst_map.py:177-178)."""
    dt = xy.dtype.type
    inv = dt(1.0 / cell_size)
    xo = (xy[..., 0] * inv + dt(origin_xi))
    yo = (xy[..., 1] * inv + dt(origin_yi))
    ix = np.clip(np.floor(xo).astype(np.int64), 0, occ_map.shape[0] - 1)
    iy = np.clip(np.floor(yo).astype(np.int64), 0, occ_map.shape[1] - 1)
    return occ_map[iy, ix]


def cost_collision_map(x, occ_map, cell_size, origin_xi, origin_yi, sigma_coll):
    vals = map_lookup(x[:, :, 1:, :2], occ_map, cell_size, origin_xi, origin_yi)
    return vals.sum(-1).astype(x.dtype) * (1.0 / sigma_coll ** 2)


def sphere_rbf(link_pos, spheres):
    """link_pos [..., L, 3], spheres [O, 4] -> [...]: sum_l sum_o exp(-0.5 |p-c|^2 / r^2)."""
    diff = link_pos[..., :, None, :] - spheres[None, :, :3]
    return np.exp(-0.5 * (diff * diff).sum(-1) / (spheres[:, 3] ** 2)).sum((-1, -2))


def sphere_field(link_pos, spheres, field_type='rbf', clamp_sdf=False):
    """LinkDistanceField.compute_cost for every field_type (costs/fields.py:78-86)."""
    if field_type == 'rbf':
        return sphere_rbf(link_pos, spheres)
    dist = np.linalg.norm(link_pos[..., :, None, :] - spheres[None, :, :3], axis=-1)
    if field_type == 'sdf':
        sdf = -dist + spheres[:, 3]
        if clamp_sdf:
            sdf = np.minimum(sdf, 0.)
        return sdf.max(-1).max(-1)
    if field_type == 'occupancy':
        return (dist < spheres[:, 3]).sum((-1, -2)).astype(link_pos.dtype)
    raise ValueError(field_type)


def interp_alpha(num_interpolate, dtype):
    """alpha of costs/fields.py:69: torch.linspace(0, 1, n + 2) is evaluated in torch's DEFAULT dtype (float32)
    and only then cast to the link dtype, so fp64 runs see float32-rounded fractions.  Restated: float32
    step = 1/(n+1); first half start + i*step, second half end - (steps-1-i)*step with the product fused into the
    subtraction (ATen RangeFactories, vectorised CPU kernel); checked against torch for n = 1..8 in
    tests/test_oracle_golden.py."""
    steps = num_interpolate + 2
    step = np.float32(1.0) / np.float32(steps - 1)
    half = steps // 2
    a = np.array([np.float32(i) * step if i < half else np.float32(1.0 - np.float64(step) * (steps - 1 - i))
                  for i in range(steps)], dtype=np.float32)
    return a[1:num_interpolate + 1].astype(dtype)


def interpolate_links(link_pos, num_interpolate=0, interp_range=(5, 7)):
    """link_pos [..., L, 3] -> [..., L + n*(hi-lo), 3]: extra evaluation points X_i + (X_{i+1} - X_i) alpha between
    link frames i and i+1 for i in [lo, hi), appended after the frames (costs/fields.py:68-74)."""
    if not num_interpolate:
        return link_pos
    alpha = interp_alpha(num_interpolate, link_pos.dtype)[:, None]
    out = [link_pos]
    for i in range(interp_range[0], interp_range[1]):
        X1, X2 = link_pos[..., i:i + 1, :], link_pos[..., i + 1:i + 2, :]
        out.append(X1 + (X2 - X1) * alpha)
    return np.concatenate(out, axis=-2)


def cost_collision_spheres(x, spheres, sigma_coll, fk_fn, field_type='rbf', clamp_sdf=False, num_interpolate=0,
                           interp_range=(5, 7)):
    """fk_fn: q [N, n] -> H [N, L, 4, 4]."""
    NP, S, T, d = x.shape
    n = d // 2
    q = x[:, :, 1:, :n].reshape(-1, n)
    H = fk_fn(q)
    pos = interpolate_links(H[:, :, :3, 3].reshape(NP, S, T - 1, H.shape[1], 3), num_interpolate, interp_range)
    return sphere_field(pos, spheres, field_type, clamp_sdf).sum(-1) * (1.0 / sigma_coll ** 2)


def self_rbf(link_pos, margin):
    """link_pos [..., L, 3] -> [...]: sum over ALL ordered pairs (i == j included) of
    exp(-|p_i - p_j|^2 / (2 margin^2)).  LinkSelfDistanceField.compute_cost, costs/fields.py:114-124."""
    diff = link_pos[..., :, None, :] - link_pos[..., None, :, :]
    return np.exp((diff * diff).sum(-1) / (-margin ** 2 * 2)).sum((-1, -2))


def cost_self_collision(x, margin, sigma_self, fk_fn, num_interpolate=0, interp_range=(5, 7)):
    NP, S, T, d = x.shape
    n = d // 2
    H = fk_fn(x[:, :, 1:, :n].reshape(-1, n))
    pos = interpolate_links(H[:, :, :3, 3].reshape(NP, S, T - 1, H.shape[1], 3), num_interpolate, interp_range)
    return self_rbf(pos, margin).sum(-1) * (1.0 / sigma_self ** 2)


def cost_ee_goal(x, target_H, sigma_goal, fk_fn, w_pos=1., w_rot=1., square=True):
    """CostGoal.eval (cost_functions.py:308-321): the field is evaluated on the LAST time step only
    (FieldFactor range [T-1, T], cost_functions.py:300-304); EESE3DistanceField takes the last link frame as the
    end-effector (fields.py:142-144), squares the SE(3) distance (fields.py:147-150); weight 1/sigma_goal^2."""
    from .se3 import se3_distance
    NP, S, T, d = x.shape
    n = d // 2
    H = fk_fn(x[:, :, -1, :n].reshape(-1, n))
    dist = se3_distance(H[:, -1], np.asarray(target_H, dtype=x.dtype), w_pos, w_rot).reshape(NP, S)
    if square:
        dist = dist * dist
    return dist * (1.0 / sigma_goal ** 2)


def cost_importance(x, means, D, O, temperature):
    """tau * x^T P mu, P given by its per-DoF blocks."""
    b = _prior.precision_times(D, O, means)          # [NP, T, d]
    return temperature * (x * b[:, None]).sum((-1, -2))
