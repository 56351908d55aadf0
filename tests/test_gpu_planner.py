"""GPU: the Python drop-in (StochGPMP / StochGPMPBatch) end to end against the reference's golden runs.

These tests read like the reference's example scripts: build cost objects, build the planner, call
optimize() / get_recent_samples().  eps is injected (the reference's torch-drawn normals) through the
private `_eps` / `_init_eps` hooks.
"""
import numpy as np
import pytest
import torch

from oracle import planner as OP

from helpers import (GOLDEN, FACTOR_TOL_F64, TOL_F32, TOL_F64, eps_ref_to_traj, load, n_iters, rel, to_sminor)
from test_gpu_kernels import _lowered, _obs

pytestmark = pytest.mark.gpu


def _planner(g, spec, dev, dtype, **over):
    from stoch_gpmp_b200.planner import StochGPMP
    comp, _ = _lowered(spec, dev, dtype)
    ta = dict(device=dev, dtype=dtype)
    mode = str(g['initial_particle_means_mode']) if 'initial_particle_means_mode' in g.files else None
    kw = dict(num_particles_per_goal=spec['K'], num_samples=spec['S'], traj_len=spec['T'], opt_iters=1, dt=spec['dt'],
              n_dof=spec['n_dof'], step_size=spec['step_size'], temperature=spec['temperature'],
              start_state=torch.tensor(spec['start'], **ta),
              multi_goal_states=None if spec.get('goals') is None else torch.tensor(spec['goals'], **ta),
              initial_particle_means=mode, cost=comp, seed=int(g['seed']), tensor_args=ta,
              **{k: spec[k] for k in ('sigma_start_init', 'sigma_gp_init', 'sigma_goal_init', 'sigma_start_sample',
                                      'sigma_gp_sample', 'sigma_goal_sample')})
    kw.update(over)
    return StochGPMP(**kw)


@pytest.mark.parametrize("name", GOLDEN)
def test_planner_matches_reference_run(name, cuda):
    g = load(name)
    spec = OP.spec_from_golden(g)
    f32 = spec['dtype'] == 'float32'
    dtype = torch.float32 if f32 else torch.float64
    pre0 = 'sameL_' if f32 else ''
    ftol = TOL_F32 if f32 else FACTOR_TOL_F64.get(name, TOL_F64)
    T, n = spec['T'], spec['n_dof']
    d = 2 * n
    pl = _planner(g, spec, cuda, dtype)
    assert pl.num_particles == spec['G'] * spec['K']
    # reset(): initial means from the INIT prior with the reference's init draw
    if pre0 + 'init_eps' in g.files:
        pl.reset(_init_eps=torch.tensor(g[pre0 + 'init_eps'], device=cuda))
    assert tuple(pl.particle_means.shape) == (pl.num_particles, T, d)
    assert rel(pl.particle_means.cpu().numpy(), g[pre0 + 'means_reset']) < max(ftol, 1e-10)
    assert rel(pl.Sigma_inv.cpu().numpy(), g['Sigma_inv']) < (1e-6 if f32 else 1e-15)
    obs = _obs(spec, cuda, dtype)
    for it in range(n_iters(g, pre0)):
        pre = f'{pre0}it{it}_'
        # continue from the reference's own pre-iteration means so that each iteration is checked alone
        pl.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
        eps = torch.tensor(to_sminor(eps_ref_to_traj(g[pre + 'eps'], T, d)), device=cuda)   # [1,NP,T,d,S] == [n_iters, NP, ...]
        pos_m, vel_m, pos_s, vel_s, costs, grad = pl.optimize(_eps=eps, **obs)
        assert tuple(pos_s.shape) == (pl.num_particles, spec['S'], T, n) and tuple(costs.shape) == (pl.num_particles, spec['S'])
        # optimize() returns the means from BEFORE the update (planner.py:252-253)
        assert np.array_equal(pos_m.cpu().numpy(), g[pre + 'means_pre'][..., :n])
        assert np.array_equal(vel_m.cpu().numpy(), g[pre + 'means_pre'][..., n:])
        assert rel(pos_s.cpu().numpy(), g[pre + 'samples'][..., :n]) < ftol
        assert rel(vel_s.cpu().numpy(), g[pre + 'samples'][..., n:]) < ftol
        if not f32:
            assert rel(costs.cpu().numpy(), g[pre + 'costs']) < max(ftol, 1e-10)
            assert np.abs(pl._weights.cpu().numpy().reshape(g[pre + 'weights'].shape) - g[pre + 'weights']).max() < 1e-7
            assert rel(grad.cpu().numpy(), g[pre + 'grad']) < max(10 * ftol, 1e-9)
            assert rel(pl.particle_means.cpu().numpy(), g[pre + 'means_post']) < ftol
        tr, ct = pl.get_recent_samples()
        assert torch.equal(tr, pos_s) and torch.equal(ct, vel_s) and tr.data_ptr() != pos_s.data_ptr()


def test_final_plans_agree_in_cost_and_collision_freeness(cuda):
    """north_star: 'final plans must agree in cost and collision-freeness'.  Chain the golden iterations
    WITHOUT resetting the means in between; compare the final means, their cost and their occupancy."""
    g = load('planar_f64')
    spec = OP.spec_from_golden(g)
    T, n = spec['T'], spec['n_dof']
    pl = _planner(g, spec, cuda, torch.float64)
    pl.reset(_init_eps=torch.tensor(g['init_eps'], device=cuda))
    for it in range(n_iters(g)):
        eps = torch.tensor(to_sminor(eps_ref_to_traj(g[f'it{it}_eps'], T, 2 * n)), device=cuda)
        pl.optimize(_eps=eps)
    final = pl.particle_means.cpu().numpy()
    want = g[f'it{n_iters(g) - 1}_means_post']
    assert rel(final, want) < 1e-10
    from oracle import costs as C
    occ_got = C.map_lookup(final[:, 1:, :2], spec['map'], spec['map_cell_size'], *spec['map_origin'])
    occ_want = C.map_lookup(want[:, 1:, :2], spec['map'], spec['map_cell_size'], *spec['map_origin'])
    assert np.array_equal(occ_got, occ_want)


def test_spheres_then_none_then_same_spheres(cuda):
    """optimize(obstacle_spheres=sp), optimize() (no spheres: the field contributes 0, fields.py:64-65), then the SAME
    tensor again must plan WITH the obstacles again — the descriptor cache used to take the repeated tensor for 'unchanged'."""
    g = load('panda_soft_f32')
    spec = OP.spec_from_golden(g)
    obs = _obs(spec, cuda, torch.float32)
    a = _planner(g, spec, cuda, torch.float32)
    b = _planner(g, spec, cuda, torch.float32)
    n_sph = a._desc(obs).n_spheres
    assert n_sph == obs['obstacle_spheres'].shape[-2]
    c1 = a.optimize(**obs)[4]
    a.optimize()
    assert a._desc({}).n_spheres == 0
    d3 = a._desc(obs)
    assert d3.n_spheres == n_sph and d3.spheres
    # same draws on a planner that never saw the None call: identical third iteration only if the obstacles are back
    b.optimize(**obs)
    b._means.copy_(a._means)
    b._draw = a._draw
    assert torch.equal(a.optimize(**obs)[4], b.optimize(**obs)[4])
    assert not torch.equal(c1, a.optimize()[4])


def test_lazy_samples_regenerate_identically(cuda):
    """return_samples=False writes nothing; get_recent_samples() rebuilds the same samples from the
    counter-based RNG; state_samples after reset is reproducible too."""
    g = load('panda_soft_f32')
    spec = OP.spec_from_golden(g)
    obs = _obs(spec, cuda, torch.float32)
    a = _planner(g, spec, cuda, torch.float32)
    b = _planner(g, spec, cuda, torch.float32)
    assert torch.equal(a.state_samples, b.state_samples)
    oa = a.optimize(opt_iters=3, return_samples=True, **obs)
    ob = b.optimize(opt_iters=3, return_samples=False, **obs)
    assert ob[2] is None and ob[3] is None
    assert torch.equal(oa[4], ob[4]) and torch.equal(oa[5], ob[5])
    ta, tb = a.get_recent_samples(), b.get_recent_samples()
    assert float((ta[0] - tb[0]).abs().max()) < 1e-6 and float((ta[1] - tb[1]).abs().max()) < 1e-5
    # three fused iterations == three single-iteration calls
    c = _planner(g, spec, cuda, torch.float32)
    for _ in range(3):
        oc = c.optimize(opt_iters=1, **obs)
    assert torch.equal(oc[4], oa[4]) and torch.equal(c.particle_means, a.particle_means)


def test_reference_method_path_equals_fused(cuda):
    """sample_and_eval() + _update_distribution() (K2/K3/K4) == optimize() (fused) for the same draw."""
    g = load('planar_soft_f64')
    spec = OP.spec_from_golden(g)
    a = _planner(g, spec, cuda, torch.float64)
    b = _planner(g, spec, cuda, torch.float64)
    vel_s, pos_s, vel_m, pos_m, costs = a.sample_and_eval()
    grad = a._update_distribution(costs, a.state_samples)
    out = b.optimize()
    assert rel(costs.cpu().numpy(), out[4].cpu().numpy()) < 1e-12
    assert rel(pos_s.cpu().numpy(), out[2].cpu().numpy()) < 1e-13
    assert rel(grad.cpu().numpy(), out[5].cpu().numpy()) < 1e-9
    assert rel(a.particle_means.cpu().numpy(), b.particle_means.cpu().numpy()) < 1e-12
    p, v = a.sample_trajectories(7)
    assert tuple(p.shape) == (a.num_particles, 7, spec['T'], spec['n_dof'])


def test_batch_equals_loop_of_singles(cuda):
    """StochGPMPBatch over B problems == B StochGPMP planners with problem_offset = b (new B axis)."""
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.envs.occupancy import generate_obstacle_map
    from stoch_gpmp_b200.planner import StochGPMP, StochGPMPBatch
    from stoch_gpmp_b200.scenarios import PLANAR_COST, PLANAR_SIGMAS, planar_batch
    import random
    B, G, K, S, T, n = 3, 4, 2, 32, 32, 2
    dtype = torch.float32
    ta = dict(device=cuda, dtype=dtype)
    start, goals = planar_batch(B, G, seed0=1)
    maps = []
    for b in range(B):
        random.seed(100 + b)
        np.random.seed(100 + b)
        maps.append(generate_obstacle_map(map_dim=[20, 20], cell_size=0.1, random_gen=True, num_obst=10,
                                          rand_limits=[[-7.5, 7.5], [-7.5, 7.5]], rand_rect_shape=[2, 2], tensor_args=ta)[0])

    def build(cls, sl, field, **kw):
        s = torch.tensor(start[sl], **ta)
        gl = torch.tensor(goals[sl], **ta)
        comp = CostComposite(n, T, [
            CostGP(n, T, s, 0.02, dict(sigma_start=PLANAR_COST['sigma_start'], sigma_gp=PLANAR_COST['sigma_gp']), ta),
            CostGoalPrior(n, T, multi_goal_states=gl, num_particles_per_goal=K, num_samples=S,
                          sigma_goal_prior=PLANAR_COST['sigma_goal_prior'], tensor_args=ta),
            CostCollision(n, T, field=field, sigma_coll=PLANAR_COST['sigma_coll'])])
        return cls(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=2, dt=0.02, n_dof=n, step_size=0.5,
                   temperature=1., start_state=s, multi_goal_states=gl, cost=comp, seed=3, tensor_args=ta, **PLANAR_SIGMAS, **kw)
    pb = build(StochGPMPBatch, slice(0, B), maps)
    ob = pb.optimize()
    for b in range(B):
        ps = build(StochGPMP, b, maps[b], problem_offset=b)
        os_ = ps.optimize()
        assert torch.equal(ps.particle_means, pb.particle_means[b])
        assert torch.equal(os_[4], ob[4][b]) and torch.equal(os_[5], ob[5][b])


@pytest.mark.parametrize("name", ['panda_soft_f64', 'panda_shipped_f64', 'panda_ee_soft_f64', 'panda_interp_f64', 'panda_sdf_f64',
                                  'panda_soft_f32'])
def test_arbitrary_fk_callable(name, cuda):
    """CostComposite(FK=<any callable>) as the reference allows (cost_functions.py:39-52): with an FK that is not a SerialChainFK
    descriptor the link-field terms (sphere / self-collision / EE goal, with interpolation and the sdf variant) are evaluated in torch
    on the materialised samples, the GP / goal terms and the IS term still in the CUDA cost kernel.  Same golden costs, gradient and
    means as the run of the unmodified reference (fp64 1e-10), i.e. as the fully lowered cost list."""
    from stoch_gpmp_b200.costs.cost_functions import CostComposite
    from stoch_gpmp_b200.robots import PandaFK
    g = load(name)
    spec = OP.spec_from_golden(g)
    f32 = spec['dtype'] == 'float32'
    dtype = torch.float32 if f32 else torch.float64
    pre0 = 'sameL_' if f32 else ''
    T, n = spec['T'], spec['n_dof']
    d = 2 * n
    comp, _ = _lowered(spec, cuda, dtype)
    chain = PandaFK()
    calls = []

    def fk(q):                       # a plain function: the planner cannot lower it
        calls.append(tuple(q.shape))
        return chain(q)
    comp_any = CostComposite(n, T, list(comp.cost_list), FK=fk, tensor_args=dict(device=cuda, dtype=dtype))
    pl = _planner(g, spec, cuda, dtype, cost=comp_any)
    assert pl._lowered.fk is None and pl._lowered.custom
    obs = _obs(spec, cuda, dtype)
    pre = f'{pre0}it0_'
    pl.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
    eps = torch.tensor(to_sminor(eps_ref_to_traj(g[pre + 'eps'], T, d)), device=cuda)
    out = pl.optimize(_eps=eps, **obs)
    assert calls and calls[-1] == (pl.num_particles * spec['S'] * T, n)
    ftol = FACTOR_TOL_F64.get(name, TOL_F64)
    if f32:
        ref = _planner(g, spec, cuda, dtype)
        ref.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
        want = ref.optimize(_eps=eps, **obs)[4]
        assert rel(out[4].cpu().numpy(), want.cpu().numpy()) < 2e-5
    else:
        assert rel(out[4].cpu().numpy(), g[pre + 'costs']) < max(ftol, 1e-10)
        assert rel(out[5].cpu().numpy(), g[pre + 'grad']) < max(10 * ftol, 1e-9)
        assert rel(pl.particle_means.cpu().numpy(), g[pre + 'means_post']) < ftol


def test_user_defined_term_on_a_problem_batch(cuda):
    """A user term on StochGPMPBatch sees trajs [B*NP*S, T, d] in (problem, particle, sample) order: a batch of B problems with a
    user-written velocity penalty == B single planners with the same term (problem_offset = b), over two iterations."""
    from stoch_gpmp_b200.costs.cost_functions import CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.planner import StochGPMP, StochGPMPBatch
    from stoch_gpmp_b200.scenarios import PLANAR_COST, PLANAR_SIGMAS, planar_batch
    B, G, K, S, T, n = 3, 2, 2, 16, 16, 2
    ta = dict(device=cuda, dtype=torch.float64)
    start, goals = planar_batch(B, G, seed0=5)
    seen = []

    def vel_penalty(trajs, x_trajs=None, **obs):
        seen.append(tuple(trajs.shape))
        return 0.25 * (trajs[:, :, n:] ** 2).sum((-1, -2))

    def build(cls, sl, **kw):
        s = torch.tensor(start[sl], **ta)
        gl = torch.tensor(goals[sl], **ta)
        comp = CostComposite(n, T, [
            CostGP(n, T, s, 0.02, dict(sigma_start=PLANAR_COST['sigma_start'], sigma_gp=PLANAR_COST['sigma_gp']), ta),
            CostGoalPrior(n, T, multi_goal_states=gl, num_particles_per_goal=K, num_samples=S,
                          sigma_goal_prior=PLANAR_COST['sigma_goal_prior'], tensor_args=ta),
            vel_penalty])
        return cls(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=2, dt=0.02, n_dof=n, step_size=0.5,
                   temperature=1., start_state=s, multi_goal_states=gl, cost=comp, seed=3, tensor_args=ta, **PLANAR_SIGMAS, **kw)
    pb = build(StochGPMPBatch, slice(0, B))
    ob = pb.optimize()
    assert seen and seen[-1] == (B * G * K * S, T, 2 * n)
    for b in range(B):
        ps = build(StochGPMP, b, problem_offset=b)
        os_ = ps.optimize()
        assert seen[-1] == (G * K * S, T, 2 * n)
        assert rel(ps.particle_means.cpu().numpy(), pb.particle_means[b].cpu().numpy()) < 1e-13
        assert rel(os_[4].cpu().numpy(), ob[4][b].cpu().numpy()) < 1e-13
    # the penalty is really in the costs: the same batch without it differs
    pb0 = build(StochGPMPBatch, slice(0, B))
    pb0.cost.cost_list.pop()
    pb0 = StochGPMPBatch(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=2, dt=0.02, n_dof=n, step_size=0.5, temperature=1.,
                         start_state=torch.tensor(start, **ta), multi_goal_states=torch.tensor(goals, **ta), cost=pb0.cost, seed=3,
                         tensor_args=ta, **PLANAR_SIGMAS)
    assert not torch.equal(pb0.optimize()[4], ob[4])


def test_error_behaviour(cuda, lib):
    from stoch_gpmp_b200.costs.cost_functions import CostComposite, CostGP, Cost
    from stoch_gpmp_b200.planner import StochGPMP
    g = load('planar_f64')
    spec = OP.spec_from_golden(g)
    # CPU device: no fallback
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _planner(g, spec, torch.device('cpu'), torch.float64)
    # goals must be 2-D (planner.py:60)
    with pytest.raises(AssertionError):
        _planner(g, spec, cuda, torch.float64, multi_goal_states=torch.zeros(4, device=cuda, dtype=torch.float64))
    # non-PD prior -> ValueError (the reference's failure class, mp_priors_multi.py:106)
    with pytest.raises(ValueError, match="positive definite"):
        _planner(g, spec, cuda, torch.float64, sigma_gp_sample=float('nan'))
    # a cost_list entry that is neither a known cost class nor callable -> NotImplementedError (callables are user-defined
    # terms, test_user_defined_cost_terms)
    ta = dict(device=cuda, dtype=torch.float64)
    comp = CostComposite(2, spec['T'], [CostGP(2, spec['T'], torch.zeros(4, **ta), 0.02, dict(sigma_start=1., sigma_gp=1.), ta),
                                        object()])
    with pytest.raises(NotImplementedError):
        _planner(g, spec, cuda, torch.float64, cost=comp)
    # CostGoalPrior built for another (K, S): the reference fails in its reshape (cost_functions.py:379)
    pl = _planner(g, spec, cuda, torch.float64, num_samples=spec['S'] + 1)
    with pytest.raises(RuntimeError):
        pl.optimize()
    # n_dof without an instantiation
    with pytest.raises(NotImplementedError):
        StochGPMP(1, 4, 8, 1, dt=0.1, n_dof=9, start_state=torch.zeros(18, **ta), multi_goal_states=torch.zeros(1, 18, **ta),
                  cost=None, sigma_start_init=1., sigma_start_sample=1., sigma_goal_init=1., sigma_goal_sample=1.,
                  sigma_gp_init=1., sigma_gp_sample=1., tensor_args=ta)


class _TorchGoalPrior:
    """A user-written cost term with the reference's calling convention (cost_functions.py:47-56): the arithmetic of the
    reference's CostGoalPrior (cost_functions.py:376-388) in plain torch — goal factor on x_{T-1}, one goal per G block."""

    def __init__(self, goals, K, S, sigma):
        self.goals, self.K, self.S, self.sigma = goals, K, S, sigma
        self.calls = 0

    def __call__(self, trajs, x_trajs=None, **observation):
        self.calls += 1
        G = self.goals.shape[0]
        x = trajs.reshape(G, self.K * self.S, trajs.shape[-2], trajs.shape[-1])[:, :, -1, :]
        e = self.goals[:, None, :] - x
        return ((e * e).sum(-1) / self.sigma ** 2).reshape(-1)


class _TorchEEHeight:
    """A user-written term that needs forward kinematics: squared height error of the last link frame, summed over time
    (receives x_trajs [N, T, L, 4, 4] like a reference cost, cost_functions.py:51-56)."""

    def eval(self, trajs, x_trajs=None, **observation):
        z = x_trajs[:, 1:, -1, 2, 3]
        return 3.0 * ((z - 0.4) ** 2).sum(-1)


@pytest.mark.parametrize("name", ['planar_f64', 'panda_soft_f32'])
def test_user_defined_cost_terms(name, cuda):
    """The reference accepts any callable in cost_list.  Here such terms run as the user's torch code on materialised samples
    (K2 -> K3 -> user terms -> K4).  (1) A torch restatement of CostGoalPrior in place of the lowered one gives the reference's
    golden costs / means (fp64 1e-10).  (2) A term that reads x_trajs shifts the costs by exactly its own values."""
    g = load(name)
    spec = OP.spec_from_golden(g)
    f32 = spec['dtype'] == 'float32'
    dtype = torch.float32 if f32 else torch.float64
    pre0 = 'sameL_' if f32 else ''
    T, n = spec['T'], spec['n_dof']
    d = 2 * n
    ta = dict(device=cuda, dtype=dtype)
    comp, _ = _lowered(spec, cuda, dtype)
    from stoch_gpmp_b200.costs.cost_functions import CostComposite, CostGoalPrior
    user_goal = _TorchGoalPrior(torch.tensor(spec['goals'], **ta), spec['K'], spec['S'], spec['sigma_goal_prior'])
    cl = [user_goal if isinstance(c, CostGoalPrior) else c for c in comp.cost_list]
    assert any(c is user_goal for c in cl)
    comp_user = CostComposite(n, T, cl, FK=comp.FK, tensor_args=ta)
    pl_ref = _planner(g, spec, cuda, dtype)
    pl_usr = _planner(g, spec, cuda, dtype, cost=comp_user)
    obs = _obs(spec, cuda, dtype)
    pre = f'{pre0}it0_'
    eps = torch.tensor(to_sminor(eps_ref_to_traj(g[pre + 'eps'], T, d)), device=cuda)
    for pl in (pl_ref, pl_usr):
        pl.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
    out_ref = pl_ref.optimize(_eps=eps, **obs)
    out_usr = pl_usr.optimize(_eps=eps, **obs)
    assert user_goal.calls == 1
    tol = 1e-5 if f32 else 1e-12
    assert torch.equal(out_usr[0], out_ref[0])
    assert rel(out_usr[2].cpu().numpy(), out_ref[2].cpu().numpy()) < (1e-6 if f32 else 1e-13)
    assert rel(out_usr[4].cpu().numpy(), out_ref[4].cpu().numpy()) < tol
    if not f32:
        assert rel(out_usr[4].cpu().numpy(), g[pre + 'costs']) < 1e-10
        assert rel(pl_usr.particle_means.cpu().numpy(), g[pre + 'means_post']) < 1e-10
        assert rel(out_usr[5].cpu().numpy(), g[pre + 'grad']) < 1e-9
    # in-kernel RNG: the separate-kernel path draws the same stream as the fused loop
    pl_ref.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
    pl_usr.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
    pl_usr._draw = pl_ref._draw
    k = 1 if f32 else 2
    a, b = pl_ref.optimize(opt_iters=k, **obs), pl_usr.optimize(opt_iters=k, **obs)
    assert rel(b[2].cpu().numpy(), a[2].cpu().numpy()) < (2e-5 if f32 else 1e-12)
    assert rel(b[4].cpu().numpy(), a[4].cpu().numpy()) < (1e-4 if f32 else 1e-10)
    if comp.FK is not None:
        comp_fk = CostComposite(n, T, list(comp.cost_list) + [_TorchEEHeight()], FK=comp.FK, tensor_args=ta)
        pl_fk = _planner(g, spec, cuda, dtype, cost=comp_fk)
        pl_fk.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
        pl_ref.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda)
        o_fk, o_rf = pl_fk.optimize(_eps=eps, **obs), pl_ref.optimize(_eps=eps, **obs)
        x = torch.cat([o_fk[2], o_fk[3]], -1).reshape(-1, T, d)
        extra = _TorchEEHeight().eval(x, x_trajs=comp.FK(x[..., :n])).reshape(o_rf[4].shape)
        assert float(extra.abs().max()) > 0
        # the totals differ by exactly the user term (to the rounding of the fp32 totals themselves)
        assert float(((o_fk[4] - o_rf[4]) - extra).abs().max()) < (1e-5 if f32 else 1e-12) * float(o_fk[4].abs().max())
        # CostComposite.eval() (drop-in use) carries the user term too
        ev = comp_fk.eval(x, **obs) - comp.eval(x, **obs)
        assert float((ev - extra.reshape(-1)).abs().max()) < (1e-5 if f32 else 1e-12) * float(comp.eval(x, **obs).abs().max())
        # GPMP and the split-particle mode need kernels for every term
        with pytest.raises(NotImplementedError):
            pl_fk.optimize_split(**obs)


# ------------------------------------------------------------------------------- BASELINE.json shapes
def _oracle_check_inkernel(pl, spec, obs, n_iters, tol_cost, tol_mean):
    """Run `n_iters` optimize() iterations with the in-kernel Philox stream and check each against the numpy
    oracle fed the oracle-side restatement of that stream (costs vs oracle; update on the kernel's own costs)."""
    from oracle import philox as OPH
    from oracle import update as U
    NP, S, T, n = spec['G'] * spec['K'], spec['S'], spec['T'], spec['n_dof']
    for _ in range(n_iters):
        mu = pl.particle_means.cpu().numpy().astype(np.float64)
        draw = pl._draw
        out = pl.optimize(**obs)
        eps = OPH.normals(pl.seed, draw, np.arange(NP), S, T, n)
        r = OP.iterate(spec, mu, eps)
        costs = out[4].cpu().numpy().astype(np.float64)
        assert rel(out[2].cpu().numpy(), r['samples'][..., :n]) < tol_cost
        assert rel(costs, r['costs']) < tol_cost
        mp, grad, w = U.update(mu, r['samples'], costs, spec['temperature'], spec['step_size'])
        assert np.abs(pl._weights.cpu().numpy().reshape(NP, S) - w).max() < 10 * tol_cost
        assert rel(pl.particle_means.cpu().numpy(), mp) < tol_mean


def test_c1_planar_as_shipped_shape_fp64(cuda):
    """BASELINE configs[0]: examples/planar_environment.py as shipped — G=3, K=5, S=128, T=64, n=2, fp64, the
    generator's 200x200 map (seed 0), shipped sigmas — three iterations against the oracle."""
    import random
    from stoch_gpmp_b200.scenarios import PLANAR_COST, PLANAR_GOALS, PLANAR_SIGMAS
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.envs.map_generator import generate_obstacle_map
    from stoch_gpmp_b200.planner import StochGPMP
    n, T, K, S, dt = 2, 64, 5, 128, 0.02
    ta = dict(device=cuda, dtype=torch.float64)
    start = torch.tensor([-9., -9., 0., 0.], **ta)
    goals = torch.tensor(PLANAR_GOALS, **ta)
    random.seed(0)
    np.random.seed(0)
    om = generate_obstacle_map(map_dim=[20, 20], obst_list=[], cell_size=0.1, random_gen=True, num_obst=15,
                               rand_limits=[[-7.5, 7.5], [-7.5, 7.5]], rand_rect_shape=[2, 2], tensor_args=ta)[0]
    cost = CostComposite(n, T, [
        CostGP(n, T, start, dt, dict(sigma_start=PLANAR_COST['sigma_start'], sigma_gp=PLANAR_COST['sigma_gp']), ta),
        CostGoalPrior(n, T, multi_goal_states=goals, num_particles_per_goal=K, num_samples=S,
                      sigma_goal_prior=PLANAR_COST['sigma_goal_prior'], tensor_args=ta),
        CostCollision(n, T, field=om, sigma_coll=PLANAR_COST['sigma_coll'])])
    pl = StochGPMP(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=1, dt=dt, n_dof=n, temperature=1.,
                   start_state=start, multi_goal_states=goals, cost=cost, step_size=0.5, seed=0, tensor_args=ta, **PLANAR_SIGMAS)
    spec = dict(n_dof=n, T=T, dt=dt, G=3, K=K, S=S, temperature=1., step_size=0.5, start=start.cpu().numpy(),
                goals=goals.cpu().numpy(), map=om.map, map_cell_size=0.1, map_origin=(om.origin_xi, om.origin_yi),
                cost_sigma_start=PLANAR_COST['sigma_start'], cost_sigma_gp=PLANAR_COST['sigma_gp'],
                sigma_goal_prior=PLANAR_COST['sigma_goal_prior'], sigma_coll=PLANAR_COST['sigma_coll'], **PLANAR_SIGMAS)
    # the initial means come from the INIT prior with the kernel's stream: check them against the oracle too
    from oracle import philox as OPH
    init_eps = OPH.normals(0, 0, np.arange(3), K, T, n)                                  # [G, K, T, d]
    m0 = OP.initial_means(spec, np.transpose(init_eps, (1, 0, 2, 3)).reshape(K, 3, T * 2 * n))
    assert rel(pl.particle_means.cpu().numpy(), m0) < 1e-10
    _oracle_check_inkernel(pl, spec, {}, 3, 1e-10, 1e-10)


def test_c3_panda_single_problem_shape_fp32(cuda):
    """BASELINE configs[2]: Panda single problem, 4 goals x 512 samples x T=64, fp32, shipped sigmas, O=5 spheres."""
    from stoch_gpmp_b200.scenarios import PANDA_COST, PANDA_SIGMAS, PANDA_START, panda_goals, panda_spheres
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.costs.fields import LinkDistanceField
    from stoch_gpmp_b200.planner import StochGPMP
    from stoch_gpmp_b200.robots import PandaFK
    n, T, G, K, S, dt = 7, 64, 4, 1, 512, 0.05
    ta = dict(device=cuda, dtype=torch.float32)
    start, goals, spheres = np.array(PANDA_START), np.array(panda_goals(G, 0)), np.array(panda_spheres(5, 0))
    s_t, g_t = torch.tensor(start, **ta), torch.tensor(goals, **ta)
    cost = CostComposite(n, T, [
        CostGP(n, T, s_t, dt, dict(sigma_start=PANDA_COST['sigma_start'], sigma_gp=PANDA_COST['sigma_gp']), ta),
        CostGoalPrior(n, T, multi_goal_states=g_t, num_particles_per_goal=K, num_samples=S,
                      sigma_goal_prior=PANDA_COST['sigma_goal_prior'], tensor_args=ta),
        CostCollision(n, T, field=LinkDistanceField(tensor_args=ta), sigma_coll=PANDA_COST['sigma_coll'])], FK=PandaFK(), tensor_args=ta)
    pl = StochGPMP(num_particles_per_goal=K, num_samples=S, traj_len=T, opt_iters=1, dt=dt, n_dof=n, step_size=0.1, temperature=1.,
                   start_state=s_t, multi_goal_states=g_t, initial_particle_means='const_vel', cost=cost, seed=3, tensor_args=ta,
                   **PANDA_SIGMAS)
    spec = dict(n_dof=n, T=T, dt=dt, G=G, K=K, S=S, temperature=1., step_size=0.1, start=start, goals=goals, spheres=spheres,
                cost_sigma_start=PANDA_COST['sigma_start'], cost_sigma_gp=PANDA_COST['sigma_gp'],
                sigma_goal_prior=PANDA_COST['sigma_goal_prior'], sigma_coll=PANDA_COST['sigma_coll'], **PANDA_SIGMAS)
    obs = {'obstacle_spheres': torch.tensor(spheres, **ta).unsqueeze(0)}
    _oracle_check_inkernel(pl, spec, obs, 2, 1e-5, 1e-5)


def test_user_supplied_initial_means(cuda):
    """initial_particle_means as a [G,K,T,d] tensor (planner.py:202-203) and reset() with new start/goals."""
    g = load('planar_f64')
    spec = OP.spec_from_golden(g)
    T, d, G, K = spec['T'], 2 * spec['n_dof'], spec['G'], spec['K']
    pm = torch.tensor(g['it0_means_pre'].reshape(G, K, T, d), device=cuda)
    pl = _planner(g, spec, cuda, torch.float64, initial_particle_means=pm)
    assert np.array_equal(pl.particle_means.cpu().numpy(), g['it0_means_pre'])
    eps = torch.tensor(to_sminor(eps_ref_to_traj(g['it0_eps'], T, d)), device=cuda)
    out = pl.optimize(_eps=eps)
    assert rel(out[4].cpu().numpy(), g['it0_costs']) < 1e-10
    assert rel(pl.particle_means.cpu().numpy(), g['it0_means_post']) < 1e-10
    with pytest.raises(AssertionError):
        pl.reset(initial_particle_means=torch.zeros(G, K + 1, T, d, device=cuda, dtype=torch.float64))


def test_stream_shards_equal_single_batch(cuda):
    """parallel.StreamShards (shards of a batch on their own streams, plans copied to pinned host memory) == the unsharded
    batch, bit for bit (RNG streams are keyed by global problem ids)."""
    import bench
    from stoch_gpmp_b200.parallel import StreamShards
    B, n_sh = 8, 4
    w = bench.workload('panda', B)
    w = dict(w, S=64, T=16)
    whole = bench.build_planner(w, B, cuda)
    sph = torch.tensor(w['spheres'], dtype=torch.float32, device=cuda)
    whole.optimize(opt_iters=2, return_samples=False, obstacle_spheres=sph)
    Bs = B // n_sh
    planners, host_in = [], []
    for k in range(n_sh):
        a = k * Bs
        wk = dict(w, start=w['start'][a:a + Bs], goals=w['goals'][a:a + Bs], spheres=w['spheres'][a:a + Bs])
        planners.append(bench.build_planner(wk, Bs, cuda, problem_offset=a))
        host_in.append(torch.tensor(wk['spheres'], dtype=torch.float32))
    sh = StreamShards(planners, obs_key='obstacle_spheres', host_inputs=host_in)
    sh.step(opt_iters=2)
    sh.wait()
    got = torch.cat(sh.host_means, 0)
    assert torch.equal(got, whole.particle_means.cpu())
