"""GPU parity of the Gauss-Newton GPMP planner (stoch_gpmp_b200.gpmp -> sgpmp_gpmp_step) against runs of the real
reference GPMP (tests/golden/gpmp_*.npz) and the numpy oracle (oracle/gpmp.py).  fp64: 1e-9 relative on the step
(block-tridiagonal Cholesky vs the reference's dense solve of a system with cond ~1e7); fp32 storage: 1e-4 (the kernel
solves in fp64, the reference in fp32)."""
import numpy as np
import pytest
import torch

from oracle import gpmp as G

from helpers import GOLDEN_GPMP, load, rel

pytestmark = pytest.mark.gpu


def _planner(spec, dev, dtype, means, batch=None):
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP, CostGoalPrior
    from stoch_gpmp_b200.costs.fields import LinkDistanceField, LinkSelfDistanceField
    from stoch_gpmp_b200.planner import GPMP, GPMPBatch
    from stoch_gpmp_b200.robots import PandaFK
    ta = dict(device=dev, dtype=dtype)
    n, T = spec['n_dof'], spec['T']
    start = torch.tensor(spec['start'], **ta)
    goals = torch.tensor(spec['goals'], **ta)
    if batch:
        start, goals = start.expand(batch, -1).contiguous(), goals.expand(batch, -1, -1).contiguous()
    cl = [CostGP(n, T, start, spec['dt'], dict(sigma_start=spec['cost_sigma_start'], sigma_gp=spec['cost_sigma_gp']), ta),
          CostGoalPrior(n, T, multi_goal_states=goals, num_particles_per_goal=spec['K'], num_samples=1,
                        sigma_goal_prior=spec['sigma_goal_prior'], tensor_args=ta)]
    FK = None
    if spec.get('self_margin') is not None:
        FK = PandaFK()
        cl.append(CostCollision(n, T, field=LinkSelfDistanceField(margin=spec['self_margin'], tensor_args=ta), sigma_coll=spec['sigma_self']))
    if 'spheres' in spec:
        FK = PandaFK()
        ikw = {}
        if spec.get('num_interpolate'):
            ikw = dict(num_interpolate=spec['num_interpolate'], link_interpolate_range=list(spec['interp_range']))
        cl.append(CostCollision(n, T, field=LinkDistanceField(tensor_args=ta, **ikw), sigma_coll=spec['sigma_coll']))
    if spec.get('ee_target') is not None:
        from stoch_gpmp_b200.costs.cost_functions import CostGoal
        from stoch_gpmp_b200.costs.fields import EESE3DistanceField
        FK = PandaFK()
        fld = EESE3DistanceField(torch.tensor(spec['ee_target'], **ta).reshape(1, 4, 4), w_pos=spec['ee_w_pos'], w_rot=spec['ee_w_rot'],
                                 square=spec['ee_square'], tensor_args=ta)
        cl.append(CostGoal(n, T, field=fld, sigma_goal=spec['sigma_ee_goal'], tensor_args=ta))
    comp = CostComposite(n, T, cl, FK=FK, tensor_args=ta)
    cls = GPMPBatch if batch else GPMP
    return cls(num_particles_per_goal=spec['K'], traj_len=T, opt_iters=1, dt=spec['dt'], n_dof=n, step_size=spec['step_size'],
               start_state=start, multi_goal_states=goals, initial_particle_means=means, cost=comp,
               sigma_start_init=1e-3, sigma_start_sample=1e-3, sigma_goal_init=1e-3, sigma_goal_sample=1e-3,
               sigma_gp_init=1., sigma_gp_sample=1.,
               solver_params=dict(delta=spec['delta'], trust_region=spec['trust_region'], method=spec['method']), tensor_args=ta)


def _obs(spec, dev, dtype):
    return {'obstacle_spheres': torch.tensor(spec['spheres'], device=dev, dtype=dtype).unsqueeze(0)} if 'spheres' in spec else {}


@pytest.mark.parametrize("name", GOLDEN_GPMP)
def test_gpmp_iterations_match_reference(name, cuda):
    g = load(name)
    spec = G.spec_from_golden(g)
    f64 = spec['dtype'] == 'float64'
    dt = torch.float64 if f64 else torch.float32
    pl = _planner(spec, cuda, dt, torch.tensor(g['initial_particle_means'], device=cuda, dtype=dt))
    obs = _obs(spec, cuda, dt)
    it = 0
    while f'it{it}_means_pre' in g.files:
        pre = f'it{it}_'
        assert rel(pl.particle_means.cpu().numpy(), g[pre + 'means_pre']) < (1e-9 if f64 else 1e-4)
        # one step from the REFERENCE's means, so that errors do not compound across iterations
        pl.particle_means = torch.tensor(g[pre + 'means_pre'], device=cuda, dtype=dt)
        vel, pos, costs = pl.optimize(**obs)
        post = pl.particle_means.cpu().numpy().astype(np.float64)
        dth = (post - g[pre + 'means_pre'].astype(np.float64)) / spec['step_size']
        o = G.step(spec, g[pre + 'means_pre'])
        assert rel(costs.cpu().numpy(), g[pre + 'costs']) < (1e-12 if f64 else 1e-5)
        assert rel(costs.cpu().numpy(), o['costs']) < (1e-12 if f64 else 1e-5)
        assert rel(dth, o['d_theta']) < (1e-9 if f64 else 1e-4)
        assert rel(dth, g[pre + 'd_theta']) < (1e-9 if f64 else 1e-3)
        assert rel(post, g[pre + 'means_post']) < (1e-9 if f64 else 1e-4)
        n = spec['n_dof']
        assert np.array_equal(pos.cpu().numpy(), pl.particle_means[..., :n].cpu().numpy())
        assert np.array_equal(vel.cpu().numpy(), pl.particle_means[..., n:].cpu().numpy())
        it += 1
    p2, v2 = pl.get_recent_samples()
    assert torch.equal(p2, pl.particle_means[..., :spec['n_dof']])


@pytest.mark.parametrize("method,trust", [('inverse', True), ('inverse', False), ('cholesky', False)])
def test_gpmp_solver_variants_match_oracle(method, trust, cuda):
    """Every (method, trust_region) combination against the dense numpy oracle, several iterations in ONE call."""
    g = load('gpmp_panda_self_f64')
    spec = dict(G.spec_from_golden(g), method=method, trust_region=trust, delta=0.05)
    pl = _planner(spec, cuda, torch.float64, torch.tensor(g['initial_particle_means'], device=cuda))
    obs = _obs(spec, cuda, torch.float64)
    means = g['it0_means_pre'].astype(np.float64)
    for _ in range(3):
        o = G.step(spec, means)
        means = o['means_post']
    vel, pos, costs = pl.optimize(opt_iters=3, **obs)
    assert rel(pl.particle_means.cpu().numpy(), means) < 1e-8
    assert rel(costs.cpu().numpy(), o['costs']) < 1e-8


def test_gpmp_batch_equals_singles(cuda):
    g = load('gpmp_panda_f64')
    spec = G.spec_from_golden(g)
    m0 = torch.tensor(g['initial_particle_means'], device=cuda)
    single = _planner(spec, cuda, torch.float64, m0.clone())
    obs = _obs(spec, cuda, torch.float64)
    single.optimize(opt_iters=2, **obs)
    B = 3
    batch = _planner(spec, cuda, torch.float64, m0.unsqueeze(0).expand(B, -1, -1, -1, -1).contiguous(), batch=B)
    v, p, c = batch.optimize(opt_iters=2, **obs)
    for b in range(B):
        assert torch.equal(batch.particle_means[b], single.particle_means)
    assert c.shape == (B, spec['G'] * spec['K'])


def test_gpmp_rejects_what_it_cannot_differentiate(cuda):
    from stoch_gpmp_b200.costs.cost_functions import CostCollision, CostComposite, CostGP
    from stoch_gpmp_b200.envs.occupancy import ObstacleMap
    from stoch_gpmp_b200.planner import GPMP
    ta = dict(device=cuda, dtype=torch.float64)
    start = torch.zeros(4, **ta)
    om = ObstacleMap([4, 4], 0.1, tensor_args=ta)
    comp = CostComposite(2, 8, [CostGP(2, 8, start, 0.1, dict(sigma_start=1., sigma_gp=1.), ta), CostCollision(2, 8, field=om, sigma_coll=1.)])
    pl = GPMP(num_particles_per_goal=2, traj_len=8, opt_iters=1, dt=0.1, n_dof=2, start_state=start,
              multi_goal_states=torch.ones(1, 4, **ta), cost=comp, sigma_start_init=1., sigma_start_sample=1., sigma_goal_init=1.,
              sigma_goal_sample=1., sigma_gp_init=1., sigma_gp_sample=1., solver_params=dict(delta=0.1, trust_region=False, method='inverse'),
              tensor_args=ta)
    with pytest.raises(NotImplementedError, match="no gradient"):
        pl.optimize()
    with pytest.raises(ValueError, match="solver_params"):
        GPMP(num_particles_per_goal=2, traj_len=8, opt_iters=1, dt=0.1, n_dof=2, start_state=start, cost=comp, tensor_args=ta)
