"""GPMP — the reference's Gauss-Newton planner (stoch_gpmp/planner.py:352-661) with its Python API, on the CUDA path.

  GPMP        drop-in for stoch_gpmp/planner.py:352-661 (one planning problem)
  GPMPBatch   the same iteration over B independent problems

Per iteration the reference builds the dense linear system of every cost (`cost.get_linear_system`, autograd for the
field Jacobians), forms dense `A^T K A` and solves it; here `sgpmp_gpmp_step` (csrc/sgpmp_gpmp.cu) works on the
block-tridiagonal structure with analytic field gradients.  Construction, priors, the initial particle means and
`sample_trajectories` are shared with StochGPMP (the reference's GPMP.reset/get_dist are copies of StochGPMP's).

Kept quirks: `optimize()` returns (velocity means, position means, costs) in THAT order (planner.py:551-555), the costs
are b^T K b at the means BEFORE the last update (planner.py:545), and solver_params['method'] == 'cholesky' reproduces
the reference's arithmetic as written (diag(l)^-1 l^-1 g, see csrc/sgpmp_gpmp.cu) — use 'inverse' for the Gauss-Newton
step proper.
"""
import time

import torch

from . import ops
from .planner import StochGPMPBatch, print_info, prior_blocks


class GPMPBatch(StochGPMPBatch):

    _batched = True
    _warm = False

    def __init__(self, num_particles_per_goal, traj_len, opt_iters, dt=None, n_dof=None, step_size=1., temperature=1.,
                 start_state=None, multi_goal_states=None, initial_particle_means=None, cost=None,
                 sigma_start_init=None, sigma_start_sample=None, sigma_goal_init=None, sigma_goal_sample=None,
                 sigma_goal=None, sigma_gp_init=None, sigma_gp_sample=None, seed=None, solver_params=None,
                 tensor_args=None, problem_offset=0, **kwargs):
        if solver_params is None:
            raise ValueError("GPMP needs solver_params = dict(delta=..., trust_region=..., method='inverse'|'cholesky') "
                             "(planner.py:581-586 indexes it unconditionally)")
        self.solver_params = solver_params
        self.sigma_goal = sigma_goal
        self.costs = None
        super().__init__(num_particles_per_goal, 1, traj_len, opt_iters, dt=dt, n_dof=n_dof, step_size=step_size,
                         temperature=temperature, start_state=start_state, multi_goal_states=multi_goal_states,
                         initial_particle_means=initial_particle_means, cost=cost,
                         sigma_start_init=sigma_start_init, sigma_start_sample=sigma_start_sample,
                         sigma_goal_init=sigma_goal_init, sigma_goal_sample=sigma_goal_sample,
                         sigma_gp_init=sigma_gp_init, sigma_gp_sample=sigma_gp_sample, seed=seed, tensor_args=tensor_args,
                         problem_offset=problem_offset)
        self.N = self.d_state_opt * self.traj_len

    def reset(self, start_state=None, multi_goal_states=None, initial_particle_means=None, _init_eps=None):
        super().reset(start_state, multi_goal_states, initial_particle_means, _init_eps=_init_eps)
        # The reference's GPMP builds its sampling distribution WITHOUT goal_states (planner.py:533-539): MultiMPPrior is then not
        # goal-directed and Sigma^-1 carries no goal block, so sample_trajectories() is not pinned at the goal.
        if self.goal_directed:
            self.goal_directed = False
            try:
                self._tables, (self._D, self._O) = self._factor(self.sigma_start_sample, self.sigma_gp_sample, None, "sampling")
            finally:
                self.goal_directed = True
            self._Sigma_inv = None
            self._state_samples = None
        low = self._lowered
        if low is None:
            raise NotImplementedError("GPMP needs a CostComposite (cost=...)")
        if low.custom:
            raise NotImplementedError("GPMP needs the linear system of every cost term; user-defined cost objects (%s) only work "
                                      "with StochGPMP" % ", ".join(type(c).__name__ for c in low.custom))
        if not low.sigma_start > 0:
            # the reference's CostGPTrajectory.get_linear_system is `pass` (cost_functions.py:217-218): GPMP fails on it too
            raise NotImplementedError("GPMP cannot take CostGPTrajectory (no linear system in the reference either); use CostGP")
        # start/GP/goal part of A^T K A: the prior's closed form evaluated with the COST sigmas
        # (CostGP cost_functions.py:148-168, CostGoalPrior :390-405)
        D, O = prior_blocks(self.traj_len, low.dt, low.sigma_start, low.sigma_gp,
                            low.sigma_goal_prior if (low.goals is not None and low.sigma_goal_prior > 0) else None)
        self._gn_D = torch.tensor(D, dtype=torch.float64, device=self.device)
        self._gn_O = torch.tensor(O, dtype=torch.float64, device=self.device)

    def _step(self, n_iters=1, **observation):
        """planner.py:575-600, n_iters times.  Returns (costs [B,NP], d_theta [B,NP,T,d])."""
        sp = self.solver_params
        desc = self._lowered.desc(0.0, observation.get('obstacle_spheres', None))
        costs, d_theta, not_pd = ops.gpmp_step(self._shape(S=1), desc, self._gn_D, self._gn_O, self._means, sp['delta'],
                                               sp['trust_region'], sp['method'], self.step_size, n_iters)
        self._not_pd = not_pd
        return costs, d_theta

    def optimize(self, opt_iters=None, debug=False, **observation):
        """planner.py:524-555: (velocity means, position means, costs [NP])."""
        if opt_iters is None:
            opt_iters = self.opt_iters
        start_time = time.time()
        n = self.n_dof
        if debug:
            costs = None
            for opt_step in range(opt_iters):
                t_iter = time.time()
                costs, _ = self._step(1, **observation)
                if opt_step % 50 == 0:
                    print_info(opt_step, opt_iters, t_iter, start_time, costs)
        else:
            costs, _ = self._step(opt_iters, **observation)
        if debug or self.solver_params.get('check_pd', True):
            bad = int(self._not_pd.max().item())
            if bad:          # torch.linalg.cholesky raises in the reference (planner.py:627)
                raise torch.linalg.LinAlgError("GPMP: J^T J is not positive definite (pivot block t=%d)" % (bad - 1))
        self.costs = self._out(costs)
        pos = self._out(self._means[..., :n]).clone()
        vel = self._out(self._means[..., -n:]).clone()
        self._recent_control_particles = vel
        self._recent_state_trajectories = pos
        return vel, pos, self.costs.clone()

    def get_recent_samples(self):
        """planner.py:639-651: (position means, velocity means) — GPMP has no samples."""
        n = self.n_dof
        return self._out(self._means[..., :n]).detach().clone(), self._out(self._means[..., -n:]).detach().clone()


class GPMP(GPMPBatch):
    """Single-problem planner with exactly the reference's shapes (no leading B axis)."""

    _batched = False

    def _out(self, t):
        return None if t is None else t[0]

    def reset(self, start_state=None, multi_goal_states=None, initial_particle_means=None, _init_eps=None):
        if initial_particle_means is not None and not isinstance(initial_particle_means, str):
            initial_particle_means = torch.as_tensor(initial_particle_means).unsqueeze(0)
        if start_state is not None:
            start_state = torch.as_tensor(start_state).reshape(1, -1)
        if multi_goal_states is not None:
            multi_goal_states = torch.as_tensor(multi_goal_states).unsqueeze(0)
        return super().reset(start_state, multi_goal_states, initial_particle_means, _init_eps=_init_eps)
