"""Cost objects with the reference's constructor signatures (stoch_gpmp/costs/cost_functions.py).

Here they are PARAMETER CARRIERS: `StochGPMP` lowers `CostComposite.cost_list` into the POD descriptor
`sgpmp_cost_desc_t` consumed by the fused CUDA kernel (csrc/sgpmp_cost.cuh).  `eval()` is kept for
drop-in use and runs the standalone CUDA cost kernel (K3); fields / FK callables that a lowered term needs
but the kernels cannot evaluate raise NotImplementedError — there is no CPU fallback.  A `cost_list` entry
that is not one of these classes but is callable is a USER term (the reference accepts any callable,
cost_functions.py:47-56): it runs as the user's own torch code on the materialised samples, next to the
kernels' terms (LoweredCost.custom_costs, planner.StochGPMPBatch._optimize_separate).

  CostGP          <- cost_functions.py:88-146     CostGoalPrior <- cost_functions.py:340-388
  CostCollision   <- cost_functions.py:221-261    CostComposite <- cost_functions.py:32-58
  CostGoal        <- cost_functions.py:282-321
"""
import ctypes as C

import torch

from .. import _lib
from .. import ops
from ..envs.occupancy import ObstacleMap
from ..robots.serial_chain import SerialChainFK
from .fields import EESE3DistanceField, LinkDistanceField, LinkSelfDistanceField


class Cost:
    def __init__(self, n_dof, traj_len):
        self.n_dof = n_dof
        self.dim = 2 * n_dof
        self.traj_len = traj_len

    def __call__(self, trajs, **observation):
        return self.eval(trajs, **observation)

    def eval(self, trajs, **observation):
        """Evaluate this term alone through the CUDA cost kernel (trajs [..., T, d] on a CUDA device)."""
        comp = CostComposite(self.n_dof, self.traj_len, [self], FK=observation.pop('FK', None))
        return comp._eval_terms(trajs, **observation)[self._term]

    def get_linear_system(self, trajs, **observation):
        raise NotImplementedError("the dense (A, b, K) of cost_functions.py:60-85 is never formed here: stoch_gpmp_b200.gpmp.GPMP "
                                  "works on the block-tridiagonal normal equations inside csrc/sgpmp_gpmp.cu")


class CostGP(Cost):
    """Start-state + GP transition factors.  start_state may be [d] or, for a problem batch, [B, d]."""
    _term = 'gp+start'

    def __init__(self, n_dof, traj_len, start_state, dt, sigma_params, tensor_args, **kwargs):
        super().__init__(n_dof, traj_len)
        self.start_state = start_state
        self.dt = dt
        self.sigma_start = sigma_params['sigma_start']
        self.sigma_gp = sigma_params['sigma_gp']
        self.tensor_args = tensor_args


class CostGoalPrior(Cost):
    """Goal-state factor per goal.  multi_goal_states [G, d] or [B, G, d]."""
    _term = 'goal'

    def __init__(self, n_dof, traj_len, multi_goal_states=None, num_particles_per_goal=None, num_samples=None,
                 sigma_goal_prior=None, tensor_args=None):
        super().__init__(n_dof, traj_len)
        self.multi_goal_states = multi_goal_states
        self.num_goals = multi_goal_states.shape[-2]
        self.num_particles_per_goal = num_particles_per_goal
        self.num_particles = num_particles_per_goal * self.num_goals
        self.num_samples = num_samples
        self.sigma_goal_prior = sigma_goal_prior
        self.tensor_args = tensor_args


class CostCollision(Cost):
    """Obstacle factor over time steps 1..T-1.  field: ObstacleMap (or a list of them, one per problem of a
    batch), LinkDistanceField or LinkSelfDistanceField; None disables the term as in the reference."""

    @property
    def _term(self):
        return 'self' if isinstance(self.field, LinkSelfDistanceField) else 'coll'

    def __init__(self, n_dof, traj_len, field=None, sigma_coll=None, tensor_args=None):
        super().__init__(n_dof, traj_len)
        self.field = field
        self.sigma_coll = sigma_coll
        self.tensor_args = tensor_args

    def eval_torch(self, trajs, x_trajs=None, **observation):
        """This term in torch over time steps 1..T-1 (see _field_term_torch): the path of an un-lowerable FK / field."""
        return _field_term_torch(self, self.field, self.sigma_coll, 1, self.traj_len, trajs, x_trajs,
                                 obstacle_spheres=observation.get('obstacle_spheres', None))


def _field_term_torch(cost, field, sigma, lo, hi, trajs, x_trajs, **observations):
    """(1 / sigma^2) * sum_t field.compute_cost(states_t) over the time range [lo, hi) — the arithmetic of the reference's
    FieldFactor.get_error + CostCollision / CostGoal.eval (factors/field_factor.py:20-34, cost_functions.py:247-261, :308-321) in torch,
    for a term the kernels cannot lower (an FK callable that is not a SerialChainFK, or a user-written field object)."""
    batch = trajs.shape[0]
    if x_trajs is not None:
        states = x_trajs[:, lo:hi]
    else:
        states = trajs[:, lo:hi, :cost.n_dof].reshape(-1, cost.n_dof)
    err = field.compute_cost(states, **observations)
    if not isinstance(err, torch.Tensor):       # e.g. LinkDistanceField without spheres returns 0 (fields.py:64-65)
        return torch.zeros(batch, dtype=trajs.dtype, device=trajs.device)
    return err.reshape(batch, hi - lo).sum(1) / float(sigma) ** 2


class CostGoal(Cost):
    """End-effector goal factor on the LAST time step (FieldFactor range [T-1, T], cost_functions.py:300-304):
    (1/sigma_goal^2) * field.compute_cost(x_{T-1}).  field: EESE3DistanceField; None disables the term as in the
    reference (cost_functions.py:309-310)."""
    _term = 'ee'

    def __init__(self, n_dof, traj_len, field=None, sigma_goal=None, tensor_args=None):
        super().__init__(n_dof, traj_len)
        self.field = field
        self.sigma_goal = sigma_goal
        self.tensor_args = tensor_args

    def eval_torch(self, trajs, x_trajs=None, **observation):
        """This term in torch on the last time step (see _field_term_torch): the path of an un-lowerable FK / field."""
        return _field_term_torch(self, self.field, self.sigma_goal, self.traj_len - 1, self.traj_len, trajs, x_trajs)


class CostGPTrajectory(Cost):
    """GP transition factors WITHOUT the start-state factor (cost_functions.py:171-218); `start_state` is stored but, as in the
    reference, unused.  Lowered as a CostGP with no start term (sgpmp_cost_desc_t.sigma_start <= 0)."""
    _term = 'gp+start'

    def __init__(self, n_dof, traj_len, start_state, dt, sigma_params, tensor_args, **kwargs):
        super().__init__(n_dof, traj_len)
        self.start_state = start_state
        self.dt = dt
        self.sigma_start = None
        self.sigma_gp = sigma_params['sigma_gp']
        self.tensor_args = tensor_args


class LoweredCost:
    """Device-resident form of a CostComposite for a batch of B problems."""

    def __init__(self, composite, B, G, device, dtype):
        self.B, self.G, self.device, self.dtype = B, G, device, dtype
        gp = goal = coll = selfc = ee = None
        self.custom = []
        self.composite_fk = composite.FK
        # an FK callable that is not a SerialChainFK descriptor (the reference takes ANY callable, cost_functions.py:39-52): the
        # kernels cannot evaluate it, so every term that needs link frames runs in torch on the materialised samples
        # (Cost*.eval_torch, fields.*.compute_cost), next to the terms the kernels do evaluate
        fk_torch_only = callable(composite.FK) and not isinstance(composite.FK, SerialChainFK)
        for c in composite.cost_list:
            if isinstance(c, (CostGP, CostGPTrajectory)):
                if gp is not None:
                    raise NotImplementedError("more than one CostGP / CostGPTrajectory in cost_list")
                gp = c
            elif isinstance(c, CostGoalPrior):
                if goal is not None:
                    raise NotImplementedError("more than one CostGoalPrior in cost_list")
                goal = c
            elif isinstance(c, CostGoal):
                if c.field is None:
                    continue
                if callable(composite.FK) and hasattr(c.field, 'compute_cost') and (fk_torch_only or not isinstance(c.field, EESE3DistanceField)):
                    self.custom.append(c.eval_torch)
                    continue
                if not isinstance(c.field, EESE3DistanceField):
                    raise NotImplementedError("CostGoal field %s cannot be lowered to the CUDA path" % type(c.field).__name__)
                if ee is not None:
                    raise NotImplementedError("more than one CostGoal in cost_list")
                ee = c
            elif isinstance(c, CostCollision):
                if c.field is None:
                    continue                      # the reference returns 0 for a field-less collision cost
                known = isinstance(c.field, (LinkDistanceField, LinkSelfDistanceField, ObstacleMap, list, tuple))
                if hasattr(c.field, 'compute_cost') and ((fk_torch_only and isinstance(c.field, (LinkDistanceField, LinkSelfDistanceField)))
                                                         or not known):
                    self.custom.append(c.eval_torch)      # link field under an arbitrary FK callable, or a user-written field object
                    continue
                if isinstance(c.field, LinkSelfDistanceField):
                    if selfc is not None:
                        raise NotImplementedError("more than one self-collision field in cost_list")
                    selfc = c
                    continue
                if coll is not None:
                    raise NotImplementedError("more than one obstacle field in cost_list")
                coll = c
            elif callable(c) or hasattr(c, 'eval'):
                # a user-defined cost (the reference accepts any object with eval(trajs, x_trajs=..., **observation),
                # cost_functions.py:47-56): it cannot be compiled into the kernels, so the planner runs the separate-kernel
                # iteration (K2 -> K3 -> K4) and adds this term, evaluated by the user's own torch code on the materialised
                # samples, to the kernel's costs before the softmax (planner.py: _custom_costs)
                self.custom.append(c)
            else:
                raise NotImplementedError("cost object %s cannot be lowered to the CUDA path and is not callable"
                                          % type(c).__name__)
        if gp is None:
            raise NotImplementedError("cost_list needs a CostGP (start + GP factors) or a CostGPTrajectory")
        n, d = composite.n_dof, 2 * composite.n_dof
        self.n_dof, self.T = n, composite.traj_len
        kw = dict(device=device, dtype=dtype)
        self.dt, self.sigma_gp = float(gp.dt), float(gp.sigma_gp)
        self.sigma_start = float(gp.sigma_start) if gp.sigma_start is not None else -1.0      # CostGPTrajectory: no start factor
        self.start = torch.as_tensor(gp.start_state).to(**kw).reshape(-1, d).expand(B, d).contiguous()
        self.goals, self.sigma_goal_prior = None, -1.0
        self.goal_K = self.goal_S = None
        if goal is not None:
            self.goals = torch.as_tensor(goal.multi_goal_states).to(**kw).reshape(-1, G, d).expand(B, G, d).contiguous()
            self.sigma_goal_prior = float(goal.sigma_goal_prior)
            self.goal_K, self.goal_S = goal.num_particles_per_goal, goal.num_samples
        self.map = self.map_index = self.map_u8 = None
        self.map_meta = None
        self.sphere_sigma = None
        self.fk = None
        self.self_margin = self.self_sigma = None
        self.self_interp = self.sphere_interp = None          # (n, lo, hi, alpha list) when num_interpolate > 0
        self.ee = None
        if selfc is not None:
            selfc.field.check_lowerable()
            self._need_fk(composite, n)
            self.self_margin, self.self_sigma = float(selfc.field.margin), float(selfc.sigma_coll)
            if selfc.field.num_interpolate:
                r = selfc.field.link_interpolate_range
                self.self_interp = (int(selfc.field.num_interpolate), int(r[0]), int(r[1]), selfc.field.interp_alpha())
        if ee is not None:
            ee.field.check_lowerable()
            self._need_fk(composite, n)
            self.ee = (ee.field, float(ee.sigma_goal))
        if coll is not None:
            fields = coll.field if isinstance(coll.field, (list, tuple)) else [coll.field]
            if all(isinstance(f, ObstacleMap) for f in fields):
                f0 = fields[0]
                for f in fields:
                    if f.map.shape != f0.map.shape or f.cell_size != f0.cell_size:
                        raise NotImplementedError("all occupancy maps of a batch must share shape and cell size")
                if len(fields) not in (1, B):
                    raise ValueError("need 1 or B=%d occupancy maps, got %d" % (B, len(fields)))
                if f0.map.shape[0] != f0.map.shape[1]:
                    raise NotImplementedError("non-square occupancy maps are not supported (obst_map.py:177-178 clamps "
                                              "x with shape[0] and y with shape[1])")
                import numpy as np
                stacked = np.stack([f.map for f in fields])
                self.map = torch.as_tensor(stacked).to(**kw).contiguous()
                # occupancy counts are small integers: a byte copy (4x smaller, L1-resident) gives identical values
                self.map_u8 = None
                if np.all(stacked == np.floor(stacked)) and stacked.min() >= 0 and stacked.max() <= 255:
                    self.map_u8 = torch.as_tensor(stacked.astype(np.uint8)).to(device).contiguous()
                if len(fields) == B and B > 1:
                    self.map_index = torch.arange(B, dtype=torch.int32, device=device)
                # 1/cell_size as the reference forms it: X * (1/self.cell_size) (obst_map.py:172)
                self.map_meta = dict(h=f0.map.shape[0], w=f0.map.shape[1], oxi=f0.origin_xi, oyi=f0.origin_yi,
                                     inv_cell=1.0 / f0.cell_size, sigma=float(coll.sigma_coll))
            elif len(fields) == 1 and isinstance(fields[0], LinkDistanceField):
                fields[0].check_lowerable()
                self._need_fk(composite, n)
                self.sphere_sigma = float(coll.sigma_coll)
                self.sphere_field_type = fields[0].field_code()
                if fields[0].num_interpolate:
                    r = fields[0].link_interpolate_range
                    self.sphere_interp = (int(fields[0].num_interpolate), int(r[0]), int(r[1]), fields[0].interp_alpha())
            else:
                raise NotImplementedError("collision field %s cannot be lowered to the CUDA path"
                                          % type(fields[0]).__name__)
        self._spheres = None
        self._spheres_src = None
        self._desc_cache = None
        self._ee_key = None

    def custom_costs(self, trajs, **observation):
        """Sum of the user-defined terms of cost_list, in list order, on trajs [N, T, d] -> [N] (the calling convention of
        cost_functions.py:47-56: cost(trajs, x_trajs=FK(q) or None, **observation))."""
        T, d, n = self.T, 2 * self.n_dof, self.n_dof
        x_trajs = None
        if self.composite_fk is not None:
            x_trajs = self.composite_fk(trajs.reshape(-1, d)[:, :n]).reshape(trajs.shape[0], T, -1, 4, 4)
        total = 0
        for c in self.custom:
            f = c if callable(c) else c.eval
            total = total + f(trajs, x_trajs=x_trajs, **observation)
        return total

    def _need_fk(self, composite, n):
        if not isinstance(composite.FK, SerialChainFK):
            raise NotImplementedError(
                "link distance fields need CostComposite(FK=<stoch_gpmp_b200.robots.SerialChainFK>), e.g. PandaFK(); "
                "arbitrary FK callables cannot be lowered to the CUDA kernel")
        self.fk = composite.FK
        if self.fk.n_dofs != n:
            raise ValueError("FK chain has %d joints but n_dof=%d" % (self.fk.n_dofs, n))
        if len(self.fk.joint) > _lib.MAX_FRAMES:
            raise NotImplementedError("FK chain longer than %d frames" % _lib.MAX_FRAMES)

    def _fill_chain(self, d):
        fk = self.fk
        d.n_frames = len(fk.joint)
        d.include_base = 1 if fk.include_base else 0
        for f in range(len(fk.joint)):
            for k in range(9):
                d.chain_R[f][k] = fk.R[f][k]
            for k in range(3):
                d.chain_p[f][k] = fk.xyz[f][k]
            d.chain_joint[f] = fk.joint[f]

    def desc(self, temperature, obstacle_spheres=None):
        """The sgpmp_cost_desc_t of this cost (keeps the tensors it points to alive on self).  The constant part
        (sigmas, pointers, map metadata, FK chain) is filled once and cached; a call only patches the temperature and
        the obstacle-sphere observation, so the per-optimize() host cost stays at a few microseconds."""
        d = self._desc_cache
        if d is None:
            d = _lib.CostDesc()
            d.dt, d.sigma_start, d.sigma_gp = self.dt, self.sigma_start, self.sigma_gp
            d.sigma_goal_prior = self.sigma_goal_prior
            d.start = self.start.data_ptr()
            d.goals = self.goals.data_ptr() if self.goals is not None else None
            if self.map is not None:
                m = self.map_meta
                d.occ_map = self.map.data_ptr()
                d.occ_map_u8 = self.map_u8.data_ptr() if self.map_u8 is not None else None
                d.map_of_problem = self.map_index.data_ptr() if self.map_index is not None else None
                d.n_maps, d.map_h, d.map_w = self.map.shape[0], m['h'], m['w']
                d.origin_xi, d.origin_yi = m['oxi'], m['oyi']
                d.map_inv_cell, d.map_sigma_coll = m['inv_cell'], m['sigma']
            if self.fk is not None:
                self._fill_chain(d)
            if self.self_margin is not None:
                d.self_margin, d.self_sigma_coll = self.self_margin, self.self_sigma
            for pre, itp in (('self', self.self_interp), ('sphere', self.sphere_interp)):
                if itp is not None:
                    setattr(d, pre + '_interp_n', itp[0])
                    setattr(d, pre + '_interp_lo', itp[1])
                    setattr(d, pre + '_interp_hi', itp[2])
                    arr = getattr(d, pre + '_interp_alpha')
                    for k, a in enumerate(itp[3]):
                        arr[k] = a
            self._desc_cache = d
        d.temperature = float(temperature)
        if self.ee is not None and (not isinstance(self.ee[0].target_H, torch.Tensor) or
                                    self._ee_key != (id(self.ee[0].target_H), self.ee[0].target_H._version)):
            # the target is re-read whenever EESE3DistanceField.update_target (fields.py:139-140) replaced the tensor or the tensor
            # was modified in place (torch version counter) — not on every call: target_matrix() is a device-to-host copy
            fld, sig = self.ee
            self._ee_key = (id(fld.target_H), fld.target_H._version) if isinstance(fld.target_H, torch.Tensor) else None
            H = fld.target_matrix()
            d.ee_sigma_goal = sig
            for r in range(3):
                for c in range(3):
                    d.ee_target_R[3 * r + c] = float(H[r, c])
                d.ee_target_p[r] = float(H[r, 3])
            d.ee_w_pos, d.ee_w_rot, d.ee_square = float(fld.w_pos), float(fld.w_rot), 1 if fld.square else 0
        if self.sphere_sigma is not None:
            if obstacle_spheres is None:
                # the reference's LinkDistanceField.compute_cost returns 0 without spheres (fields.py:64-65)
                d.spheres, d.n_spheres = None, 0
                self._spheres_src = None        # the next call WITH spheres must patch the descriptor again (the same
                #                                 tensor passed after a None call used to be taken for "unchanged" -> no obstacles)
            elif obstacle_spheres is not self._spheres_src:
                sp = torch.as_tensor(obstacle_spheres).to(device=self.device, dtype=self.dtype)
                if sp.dim() == 2:
                    sp = sp.unsqueeze(0)
                if sp.shape[0] not in (1, self.B) or sp.shape[-1] != 4:
                    raise ValueError("obstacle_spheres must be [1,O,4] or [B,O,4], got %s" % (tuple(sp.shape),))
                if sp.shape[1] > _lib.MAX_SPHERES:
                    raise NotImplementedError("more than %d obstacle spheres" % _lib.MAX_SPHERES)
                self._spheres = sp.contiguous()
                # the same tensor object passed again (the usual optimize(**obs) loop) is not re-validated, but its
                # CONTENT is read by the kernel on every call when no copy was needed
                self._spheres_src = obstacle_spheres if self._spheres.data_ptr() == getattr(obstacle_spheres, 'data_ptr', lambda: 0)() else None
                d.spheres = self._spheres.data_ptr()
                d.n_spheres = sp.shape[1]
                d.spheres_per_problem = 1 if (sp.shape[0] == self.B and self.B > 1) else 0
                d.sphere_sigma_coll = self.sphere_sigma
                d.sphere_field_type = self.sphere_field_type
        return d


class CostComposite(Cost):

    def __init__(self, n_dof, traj_len, cost_list, FK=None, tensor_args=None):
        super().__init__(n_dof, traj_len)
        self.cost_list = cost_list
        self.FK = FK
        self.tensor_args = tensor_args

    def lower(self, B, G, device, dtype):
        return LoweredCost(self, B, G, device, dtype)

    def _eval_terms(self, trajs, **observation):
        if not trajs.is_cuda:
            raise RuntimeError("cost.eval: trajs is on %s — CUDA only, no CPU fallback" % trajs.device)
        T, d = self.traj_len, self.dim
        x = trajs.reshape(-1, T, d)
        nb = x.shape[0]
        goal = next((c for c in self.cost_list if isinstance(c, CostGoalPrior)), None)
        if goal is not None:
            G, K, S = goal.num_goals, goal.num_particles_per_goal, goal.num_samples
            if G * K * S != nb:
                raise RuntimeError("CostGoalPrior was built for %d x %d x %d trajectories, got %d" % (G, K, S, nb))
        else:
            G, K, S = 1, 1, nb
        low = self.lower(1, G, x.device, x.dtype)
        shape = ops.make_shape(1, G, K, S, T, self.n_dof, x.dtype)
        xs = x.reshape(1, G * K, S, T, d).permute(0, 1, 3, 4, 2).contiguous()
        desc = low.desc(0.0, observation.get('obstacle_spheres', None))
        costs, terms = ops.cost(shape, desc, None, xs, None, want_terms=True)
        out = {nm: terms[i].reshape(-1) for i, nm in enumerate(_lib.TERM_NAMES)}
        out['gp+start'] = out['start'] + out['gp']
        out['total'] = costs.reshape(-1)
        if low.custom:       # user-defined terms of cost_list (torch code, the reference's calling convention)
            out['total'] = out['total'] + torch.as_tensor(low.custom_costs(x, **observation), device=x.device).to(x.dtype).reshape(-1)
        return out

    def eval(self, trajs, **observation):
        return self._eval_terms(trajs, **observation)['total']


def _map_lookup_cuda(obst_map, X):
    """ObstacleMap.compute_cost through the cost kernel: a 2-step, 1-DoF-pair 'trajectory' per point whose
    collision term is exactly the lookup at the point."""
    if not X.is_cuda:
        raise RuntimeError("ObstacleMap.compute_cost: X is on %s — CUDA only, no CPU fallback" % X.device)
    pts = X.reshape(-1, 2)
    n = pts.shape[0]
    traj = torch.zeros(n, 2, 4, dtype=pts.dtype, device=pts.device)
    traj[:, 1, :2] = pts
    ta = dict(device=pts.device, dtype=pts.dtype)
    comp = CostComposite(2, 2, [CostGP(2, 2, torch.zeros(4, **ta), 1.0, dict(sigma_start=1.0, sigma_gp=1.0), ta),
                                CostCollision(2, 2, field=obst_map, sigma_coll=1.0)])
    return comp._eval_terms(traj)['coll'].reshape(X.shape[:-1])
