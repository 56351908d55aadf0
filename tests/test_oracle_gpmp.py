"""CPU: oracle/gpmp.py (numpy restatement of the reference's Gauss-Newton GPMP planner, analytic field Jacobians,
dense normal equations) against runs of the REAL reference GPMP (tests/golden/gpmp_*.npz, oracle/make_golden_gpmp.py)."""
import numpy as np
import pytest

from oracle import gpmp as G

from helpers import GOLDEN_GPMP, load, rel


def _iters(g):
    it = 0
    while f'it{it}_means_pre' in g.files:
        it += 1
    return it


@pytest.mark.parametrize("name", GOLDEN_GPMP)
def test_normal_equations_match_reference(name):
    """A^T K A (analytic Jacobian-transpose gradients vs the reference's autograd), the damped J^T J and g = A^T K b."""
    g = load(name)
    spec = G.spec_from_golden(g)
    r = G.step(spec, g['it0_means_pre'])
    tol = 1e-12 if spec['dtype'] == 'float64' else 2e-5
    assert rel(r['AtKA'], g['it0_AtKA']) < tol
    assert rel(r['JtJ'], g['it0_JtJ']) < tol
    assert rel(r['g'], g['it0_g']) < tol


@pytest.mark.parametrize("name", GOLDEN_GPMP)
def test_step_matches_reference(name):
    g = load(name)
    spec = G.spec_from_golden(g)
    f64 = spec['dtype'] == 'float64'
    for it in range(_iters(g)):
        pre = f'it{it}_'
        r = G.step(spec, g[pre + 'means_pre'])
        # 'inverse' goes through an LU solve of a matrix with cond ~1e7 in the reference: 1e-10 there, 1e-12 otherwise
        assert rel(r['d_theta'], g[pre + 'd_theta']) < (1e-10 if f64 else 1e-4)
        assert rel(r['means_post'], g[pre + 'means_post']) < (1e-10 if f64 else 1e-5)
        assert rel(r['costs'], g[pre + 'costs']) < (1e-13 if f64 else 1e-5)


def test_cholesky_branch_is_the_reference_quirk():
    """planner.py:626-629 returns diag(l)^-1 l^-1 g, NOT (J^T J)^-1 g: the golden run distinguishes the two."""
    g = load('gpmp_panda_f64')
    spec = G.spec_from_golden(g)
    assert spec['method'] == 'cholesky'
    r = G.step(spec, g['it0_means_pre'])
    true_solve = G.solve('inverse', r['JtJ'], r['g']).reshape(r['d_theta'].shape)
    assert rel(r['d_theta'], g['it0_d_theta']) < 1e-12
    assert rel(true_solve, g['it0_d_theta']) > 0.1


def test_field_gradients_against_finite_differences():
    rs = np.random.RandomState(0)
    q = rs.uniform(-1.5, 1.5, (5, 7))
    sph = np.array([[0.35, 0.0, 0.55, 0.18], [0.5, 0.1, 0.35, 0.15]])
    h = 1e-6
    for fn in (lambda x: G.sphere_field_and_grad(x, sph, 2, (5, 7)), lambda x: G.self_field_and_grad(x, 0.15, 3, (3, 6))):
        _, gr = fn(q)
        for j in range(7):
            dq = np.zeros(7)
            dq[j] = h
            fd = (fn(q + dq)[0] - fn(q - dq)[0]) / (2 * h)
            assert np.abs(fd - gr[:, j]).max() < 1e-6 * max(1.0, np.abs(gr).max())
