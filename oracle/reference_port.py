"""Oracle: dense torch-CPU restatement of the reference's OWN algorithm (cost included) — the CPU baseline.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).  Unlike oracle/planner.py (which uses the banded
closed forms the CUDA kernels use), this module does the work the way the reference does it, so that its
wall time is a fair stand-in for `StochGPMP.optimize()` on CPU when the reference tree itself is not
available (the GPU box):

  dense precision P = A^T W A                         mp_priors_multi.py:170-202
  NP copies, MultivariateNormal(precision_matrix=...)  mp_priors_multi.py:97-110 -> torch
      multivariate_normal.py:79-85 (chol(flip P), triangular solve) + PositiveDefinite validation
      (constraints.py: symmetry isclose + cholesky_ex), REBUILT on every set_mean (planner.py:273)
  eps = empty(S,NP,M).normal_(); x = mu + L @ eps      multivariate_normal.py:251-254
  costs via small batched matmuls                      cost_functions.py:128-146, 247-261, 376-388
  FK hook + LinkDistanceField rbf / ObstacleMap lookup cost_functions.py:51-52, fields.py:63-79, obst_map.py:164-182
  IS term V @ Sigma_inv @ U^T (dense GEMM)             planner.py:233-236
  softmax, weighted-mean update                        planner.py:263-275

It is pinned against the real reference in tests/test_reference_port.py (bit-for-bit on the golden runs
where the arithmetic order is the same, 1e-12 otherwise) and timed beside it in DESIGN.md §8.
"""
import torch


def dense_precision(T, dt, n, sigma_start, sigma_gp, sigma_goal, dtype):
    """AᵀWA assembled the reference's way: dense A ((T+1)d x M), block-diagonal W, two GEMMs."""
    d = 2 * n
    M = T * d
    eye_n = torch.eye(n, dtype=dtype)
    Phi = torch.eye(d, dtype=dtype)
    Phi[:n, n:] = eye_n * dt
    A = torch.eye(M, dtype=dtype)
    for t in range(T - 1):
        A[(t + 1) * d:(t + 2) * d, t * d:(t + 1) * d] += -1. * Phi
    rows = [A]
    if sigma_goal is not None:
        b = torch.zeros(d, M, dtype=dtype)
        b[:, -d:] = torch.eye(d, dtype=dtype)
        rows.append(b)
    A = torch.cat(rows)
    Qc = eye_n / sigma_gp ** 2
    Qinv = torch.cat((torch.cat((12. * (dt ** -3.) * Qc, -6. * (dt ** -2.) * Qc), dim=-1),
                      torch.cat((-6. * (dt ** -2.) * Qc, 4. * (dt ** -1.) * Qc), dim=-1)), dim=-2)
    W = torch.zeros(A.shape[0], A.shape[0], dtype=dtype)
    W[:d, :d] = torch.eye(d, dtype=dtype) / sigma_start ** 2
    for t in range(T - 1):
        W[(t + 1) * d:(t + 2) * d, (t + 1) * d:(t + 2) * d] = Qinv
    if sigma_goal is not None:
        W[-d:, -d:] = torch.eye(d, dtype=dtype) / sigma_goal ** 2
    return A.t() @ W @ A, Qinv, Phi


def precision_to_scale_tril(P):
    """torch/distributions/multivariate_normal.py:79-85, restated."""
    Lf = torch.linalg.cholesky(torch.flip(P, (-2, -1)))
    L_inv = torch.transpose(torch.flip(Lf, (-2, -1)), -2, -1)
    Id = torch.eye(P.shape[-1], dtype=P.dtype)
    return torch.linalg.solve_triangular(L_inv, Id, upper=False)


def validate_precision(P):
    """The PositiveDefinite check MultivariateNormal runs on its argument (symmetry + cholesky_ex)."""
    sym = torch.isclose(P, P.mT, atol=1e-6).all(-2).all(-1)
    ok = torch.linalg.cholesky_ex(P).info.eq(0)
    if not bool((sym & ok).all()):
        raise ValueError("precision matrix is not positive definite")


class ReferencePort:
    """One planning problem, reference semantics, CPU tensors.  `spec` as in oracle/planner.py."""

    def __init__(self, spec, dtype=torch.float32, fk=None, validate=True):
        self.s = spec
        self.dtype = dtype
        n, T = spec['n_dof'], spec['T']
        self.n, self.T, self.d, self.M = n, T, 2 * n, T * 2 * n
        self.G, self.K, self.S = spec['G'], spec['K'], spec['S']
        self.NP = self.G * self.K
        self.validate = validate
        goal = spec['sigma_goal_sample'] if spec.get('goals') is not None else None
        self.Sigma_inv, _, _ = dense_precision(T, spec['dt'], n, spec['sigma_start_sample'], spec['sigma_gp_sample'], goal, dtype)
        self.Sigma_invs = self.Sigma_inv.repeat(self.NP, 1, 1)
        _, self.Qinv_cost, self.Phi = dense_precision(2, spec['dt'], n, 1.0, spec['cost_sigma_gp'], None, dtype)
        self.start = torch.as_tensor(spec['start'], dtype=dtype)
        self.goals = None if spec.get('goals') is None else torch.as_tensor(spec['goals'], dtype=dtype)
        # cost_sigma_start None: CostGPTrajectory (cost_functions.py:171-218), no start factor
        self.K_start = (torch.eye(self.d, dtype=dtype) / spec['cost_sigma_start'] ** 2 if spec.get('cost_sigma_start') is not None
                        else torch.zeros(self.d, self.d, dtype=dtype))
        self.K_goal = None if spec.get('sigma_goal_prior') is None else torch.eye(self.d, dtype=dtype) / spec['sigma_goal_prior'] ** 2
        self.fk = fk
        self.map = torch.as_tensor(spec['map']).to(dtype) if 'map' in spec else None
        self.spheres = torch.as_tensor(spec['spheres'], dtype=dtype).reshape(1, -1, 4) if 'spheres' in spec else None
        self.means = None
        self.L = None

    def set_mean(self, means):
        """MultiMPPrior.set_mean -> update_dist: the MVN (validation + factor of NP dense copies) is rebuilt."""
        self.means = means.reshape(self.NP, self.T, self.d).clone()
        if self.validate:
            validate_precision(self.Sigma_invs)
        self.L = precision_to_scale_tril(self.Sigma_invs)

    def sample(self, eps=None):
        if eps is None:
            eps = torch.empty(self.S, self.NP, self.M, dtype=self.dtype).normal_()
        x = self.means.reshape(self.NP, self.M) + torch.matmul(self.L, eps.unsqueeze(-1)).squeeze(-1)
        return x.view(self.S, self.NP, self.T, self.d).transpose(1, 0), eps

    def eval_costs(self, samples):
        s, n, T, d = self.s, self.n, self.T, self.d
        trajs = samples.reshape(-1, T, d)
        nb = trajs.shape[0]
        x_trajs = None
        if self.fk is not None:
            x_trajs = self.fk(trajs.view(-1, d)[:, :n]).reshape(nb, T, -1, 4, 4)
        # CostGP
        err_p = self.start - trajs[:, [0]]
        start_c = (err_p @ self.K_start.unsqueeze(0) @ err_p.transpose(1, 2)).squeeze()
        e = (trajs[:, 1:].unsqueeze(-1) - self.Phi @ trajs[:, :-1].unsqueeze(-1))
        gp_c = (e.transpose(2, 3) @ self.Qinv_cost.reshape(1, 1, d, d) @ e).sum(1).squeeze()
        costs = start_c + gp_c
        # CostGoalPrior
        if self.goals is not None and self.K_goal is not None:
            xg = trajs.reshape(self.G, self.K * self.S, T, d)
            gc = torch.zeros(self.G, self.K * self.S, dtype=self.dtype)
            for i in range(self.G):
                eg = self.goals[i] - xg[i, :, [-1]]
                gc[i] += (eg @ self.K_goal.unsqueeze(0) @ eg.transpose(1, 2)).squeeze()
            costs = costs + gc.flatten()
        # CostCollision(LinkSelfDistanceField)   costs/fields.py:114-124
        def interp(lt, num, rng):      # link interpolation, costs/fields.py:68-74
            if not num:
                return lt
            alpha = torch.linspace(0, 1, num + 2).type_as(lt)[1:num + 1]
            alpha = alpha.view(tuple([1] * (lt.dim() - 2) + [-1, 1]))
            for i in range(rng[0], rng[1]):
                X1, X2 = lt[..., i, :].unsqueeze(-2), lt[..., i + 1, :].unsqueeze(-2)
                lt = torch.cat([lt, X1 + (X2 - X1) * alpha], dim=-2)
            return lt
        if s.get('self_margin') is not None:
            lt = interp(x_trajs[:, 1:T][..., :3, -1], s.get('self_num_interpolate', 0), s.get('self_interp_range', (5, 7)))
            selfc = torch.exp(torch.square(lt.unsqueeze(-2) - lt.unsqueeze(-3)).sum(-1) / (-s['self_margin'] ** 2 * 2)).sum((-1, -2))
            costs = costs + (1. / s['sigma_self'] ** 2) * selfc.sum(1)
        # CostCollision
        if s.get('sigma_coll') is not None and self.map is not None:
            X = trajs[:, 1:T, :n].reshape(-1, n)
            off = torch.tensor([s['map_origin'][0], s['map_origin'][1]], dtype=self.dtype)
            occ = (X * (1 / s['map_cell_size']) + off).floor().int()
            occ[..., 0] = occ[..., 0].clamp(0, self.map.shape[0] - 1)
            occ[..., 1] = occ[..., 1].clamp(0, self.map.shape[1] - 1)
            vals = self.map[occ[..., 1], occ[..., 0]].reshape(nb, T - 1)
            costs = costs + (1. / s['sigma_coll'] ** 2) * vals.sum(1)
        if s.get('sigma_coll') is not None and self.spheres is not None:
            link = interp(x_trajs[:, 1:T][..., :3, -1], s.get('num_interpolate', 0), s.get('interp_range', (5, 7))).unsqueeze(-2)
            sp = self.spheres.unsqueeze(0)
            ft = s.get('field_type', 'rbf')
            if ft == 'rbf':
                fld = torch.exp(-0.5 * torch.square(link - sp[..., :3]).sum(-1) / torch.square(sp[..., 3])).sum((-1, -2))
            elif ft == 'sdf':
                sdf = -torch.linalg.norm(link - sp[..., :3], dim=-1) + sp[..., 3]
                if s.get('clamp_sdf', False):
                    sdf = sdf.clamp(max=0.)
                fld = sdf.max(-1)[0].max(-1)[0]
            else:
                fld = (torch.linalg.norm(link - sp[..., :3], dim=-1) < sp[..., 3]).sum((-1, -2))
            costs = costs + (1. / s['sigma_coll'] ** 2) * fld.sum(1)
        # CostGoal(EESE3DistanceField)   cost_functions.py:308-321, fields.py:142-150 (last step, last link frame)
        if s.get('ee_target') is not None:
            from .se3 import se3_distance_torch
            Ht = torch.as_tensor(s['ee_target'], dtype=self.dtype).reshape(1, 4, 4)
            dist = se3_distance_torch(x_trajs[:, T - 1:T][..., -1, :, :], Ht, w_pos=s.get('ee_w_pos', 1.), w_rot=s.get('ee_w_rot', 1.)).squeeze()
            if s.get('ee_square', True):
                dist = torch.square(dist)
            costs = costs + (1. / s['sigma_ee_goal'] ** 2) * dist.reshape(nb, 1).sum(1)
        costs = costs.reshape(self.NP, self.S)
        V = samples.reshape(-1, self.S, self.M)
        U = self.means.view(-1, 1, self.M)
        costs = costs + s['temperature'] * (V @ self.Sigma_inv @ U.transpose(1, 2)).squeeze(2)
        return costs

    def iterate(self, eps=None):
        """One optimize() iteration; returns (samples, costs, weights, grad)."""
        samples, eps = self.sample(eps)
        costs = self.eval_costs(samples)
        w = torch.softmax(-costs / self.s['temperature'], dim=1)
        grad = (w.reshape(-1, self.S, 1, 1) * (samples - self.means.unsqueeze(1))).sum(1)
        self.set_mean(self.means + self.s['step_size'] * grad)
        return samples, costs, w, grad


def time_port(spec, dtype, n_problems, iters, warmup=1, fk=None, threads=None, device="cpu"):
    """Seconds per optimize() iteration of ONE problem (mean over `n_problems` sequential problems, `iters` timed
    iterations each after `warmup`).  Used by bench.py (cpu_baseline / --impl reference; device="cuda:0" is the
    informational --impl reference-cuda arm: the same dense formulation on stock torch CUDA ops)."""
    import time
    import os
    from . import prior as P
    if threads:
        torch.set_num_threads(threads)
    else:
        torch.set_num_threads(os.cpu_count() or 1)
    cuda = str(device).startswith("cuda")
    total, count = 0.0, 0
    with torch.device(device):
        for b in range(n_problems):
            port = ReferencePort(spec, dtype=dtype, fk=fk)
            mu0 = P.const_vel_trajectories(spec['start'], spec['goals'], spec['dt'], spec['T'], spec['n_dof'], spec['K'])
            port.set_mean(torch.as_tensor(mu0, dtype=dtype))
            for it in range(warmup + iters):
                if cuda:
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                port.iterate()
                if cuda:
                    torch.cuda.synchronize()
                t1 = time.perf_counter()
                if it >= warmup:
                    total += t1 - t0
                    count += 1
    return total / count, torch.get_num_threads()
