"""Synthetic problem generators shared by make_golden.py, the tests and bench.py.

Synthetic inputs for benchmarks, examples and tests (no arithmetic of the hot path lives here).  Shapes and sigmas follow the reference's two
examples (examples/planar_environment.py:14-96, examples/panda_environment.py:29-133) as laid out in
SURVEY.md §8(d): C1..C4.  Pure numpy; nothing here touches /root/reference.
"""
import numpy as np

PANDA_START = [0.012, -0.57, 0., -2.81, 0., 3.037, 0.741] + [0.] * 7          # examples/panda_environment.py:52
PANDA_LOWER = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])   # URDF joint limits
PANDA_UPPER = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
PANDA_SIGMAS = dict(sigma_start_init=0.0001, sigma_goal_init=0.1, sigma_gp_init=0.8,
                    sigma_start_sample=0.001, sigma_goal_sample=0.07, sigma_gp_sample=0.1)      # :113-118
PANDA_COST = dict(sigma_start=0.0001, sigma_gp=0.0007, sigma_coll=0.01, sigma_goal_prior=20.)  # :72-80
PLANAR_SIGMAS = dict(sigma_start_init=1e-3, sigma_goal_init=1e-3, sigma_gp_init=20.,
                     sigma_start_sample=1e-3, sigma_goal_sample=1e-3, sigma_gp_sample=3)        # planar :89-94
PLANAR_COST = dict(sigma_start=0.001, sigma_gp=0.1, sigma_coll=1e-5, sigma_goal_prior=0.001)    # planar :54-59
PLANAR_GOALS = [[9, 6, 0., 0.], [9, -3, 0., 0.], [-3, 9, 0., 0.]]                              # planar :25-29


def panda_goals(G, seed):
    """G goal configurations = start + N(0, 0.5^2), clipped to the joint limits; zero velocity."""
    rs = np.random.RandomState(seed)
    q = np.clip(np.array(PANDA_START[:7]) + rs.normal(0, 0.5, (G, 7)), PANDA_LOWER, PANDA_UPPER)
    return np.concatenate([q, np.zeros((G, 7))], axis=1).tolist()


def panda_spheres(O, seed):
    """O obstacle spheres (cx, cy, cz, r): centres U([0.6,1]x[-0.2,0.2]x[0.6,1]), r ~ U(0.1,0.2)
    (examples/panda_environment.py:125-133)."""
    rs = np.random.RandomState(seed)
    c = rs.uniform([0.6, -0.2, 0.6], [1.0, 0.2, 1.0], (O, 3))
    r = rs.uniform(0.1, 0.2, (O, 1))
    return np.concatenate([c, r], axis=1).tolist()


def panda_batch(B, G=4, O=5, seed0=0):
    """C3/C4 inputs: per-problem start (shared), goals [B,G,14], spheres [B,O,4]."""
    start = np.tile(np.array(PANDA_START)[None], (B, 1))
    goals = np.stack([np.array(panda_goals(G, seed0 + b)) for b in range(B)])
    spheres = np.stack([np.array(panda_spheres(O, seed0 + b)) for b in range(B)])
    return start, goals, spheres


def planar_batch(B, G=4, seed0=0):
    """C2 inputs: start ~ U([-9.5,-8.5]^2), goals = shipped three + [6, 9], jittered +-0.5; zero velocity."""
    rs = np.random.RandomState(seed0)
    start = np.concatenate([rs.uniform(-9.5, -8.5, (B, 2)), np.zeros((B, 2))], axis=1)
    base = np.array(PLANAR_GOALS + [[6, 9, 0., 0.]])[:G]
    goals = np.tile(base[None], (B, 1, 1))
    goals[:, :, :2] += rs.uniform(-0.5, 0.5, (B, G, 2))
    return start, goals
