"""Shared helpers of the parity tests: golden-vector loading and layout conversion."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ALL = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN_GPMP = [g for g in _ALL if g.startswith("gpmp_")]          # runs of the reference's Gauss-Newton GPMP planner
GOLDEN_C3 = [g for g in _ALL if g.startswith("c3_")]              # full-size (T = 64, S = 256) runs of StochGPMP, compact records
GOLDEN = [g for g in _ALL if not g.startswith("gpmp_") and not g.startswith("c3_")]           # runs of StochGPMP
GOLDEN_F64 = [g for g in GOLDEN if g.endswith("f64") or g.endswith("f64_T64")]
GOLDEN_F32 = [g for g in GOLDEN if g.endswith("f32")]

# Relative tolerance for quantities that depend on the prior FACTOR (L, samples, grad, means) in fp64.
# The reference factors a dense precision with cond(P) up to 1e9, so its own L carries
# ~cond * 2^-53 of error; measured against 50-digit arithmetic the reference is off by 1.2e-9 on
# panda_soft_f64 and this oracle by 2.2e-10.  Everything else holds 1e-10.
FACTOR_TOL_F64 = {"panda_soft_f64": 1e-8, "panda_self_soft_f64": 1e-8, "panda_sdf_f64": 1e-9, "panda_occ_f64": 1e-9,
                  "panda_ee_soft_f64": 1e-8, "panda_interp_f64": 1e-8}
TOL_F64 = 1e-10
TOL_F32 = 1e-5


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def n_iters(g, prefix=""):
    it = 0
    while f"{prefix}it{it}_eps" in g.files:
        it += 1
    return it


def eps_from_rng_state(g, pre, dtype=None):
    """Compact records (oracle/make_golden.py run_case(compact=True)): re-draw the eps the reference drew from the torch CPU
    generator state stored before the draw.  Returns [S, NP, M] (torch draw layout) and checks its head against the record."""
    import torch
    S, NP, M = int(g['S']), int(g['G']) * int(g['K']), int(g['T']) * 2 * int(g['n_dof'])
    keep = torch.get_rng_state()
    try:
        torch.set_rng_state(torch.from_numpy(g[pre + 'rng_state'].copy()))
        eps = torch.empty(S, NP, M, dtype=dtype or getattr(torch, str(g['dtype']))).normal_().numpy()
    finally:
        torch.set_rng_state(keep)
    if not np.array_equal(eps[:2], g[pre + 'eps_head']):
        raise AssertionError("the torch CPU generator does not reproduce the recorded draw (different torch build?)")
    return eps


def rel(a, b):
    """max |a-b| / max |b| — the norm-wise relative error used throughout."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def eps_ref_to_traj(eps_ref, T, d):
    """torch draw layout [S, NP, M] -> [NP, S, T, d]."""
    S, NP, M = eps_ref.shape
    return np.transpose(eps_ref.reshape(S, NP, T, d), (1, 0, 2, 3))


def to_sminor(x):
    """[NP, S, T, d] -> [1, NP, T, d, S] (the kernels' layout)."""
    return np.ascontiguousarray(np.transpose(x, (0, 2, 3, 1)))[None]


def from_sminor(x):
    """[B, NP, T, d, S] -> [B, NP, S, T, d]."""
    return np.transpose(x, (0, 1, 4, 2, 3))
