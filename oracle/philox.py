"""Oracle: the counter-based normal stream of the CUDA sampler.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws eps with torch's global generator
(torch/distributions/multivariate_normal.py:253, `_standard_normal` -> `normal_()`),
which cannot be reproduced bit-for-bit on a GPU.  Parity with the reference therefore
uses *injected* eps; this module pins the kernel's own stream so that the in-kernel
RNG path is checked too (integer words bit-exact, normals to float rounding).

Stream definition (stoch_gpmp_b200/csrc/sgpmp_rng.cuh follows this):
  Philox4x32-R (Salmon et al., SC'11; constants below) with R = ROUNDS = 7 rounds (the paper's
  Crush-resistant count; `philox4x32(..., rounds=10)` is pinned by the Random123 known answers),
  key = (seed_lo, seed_hi),
  counter = ((t << 8) | k,  sample s,  global particle id,  draw index)
  where k is the DoF PAIR (2k, 2k+1) and global particle id = problem_gid * NP + p.
  The 4 output words give 4 normals by two Box-Muller pairs
      u1 = (w0 + 0.5) * 2^-32, u2 = (w1 + 0.5) * 2^-32
      r = sqrt(-2 ln u1), th = 2 pi u2 - pi
      eps[t, pos, 2k] = r cos th,   eps[t, pos, 2k+1] = r sin th         from (w0, w1)
      eps[t, vel, 2k], eps[t, vel, 2k+1]  the same                       from (w2, w3)
  For an odd DoF count the last pair (2k+1 == n) uses (w0, w1) only:
      eps[t, pos, 2k] = r cos th,   eps[t, vel, 2k] = r sin th.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


ROUNDS = 7


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=ROUNDS):
    """Vectorised Philox4x32-R.  Inputs broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def box_muller(wa, wb):
    """Two N(0,1) from two uint32 words (float64 arithmetic)."""
    u1 = (wa.astype(np.float64) + 0.5) * 2.0 ** -32
    u2 = (wb.astype(np.float64) + 0.5) * 2.0 ** -32
    r = np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * np.pi * u2 - np.pi
    return r * np.cos(th), r * np.sin(th)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    return philox4x32(c0, c1, c2, c3, k0, k1, rounds=10)


def words(seed, draw, particle_gid, S, T, n_dof):
    """Raw Philox words [NPg, S, T, ceil(n_dof/2), 4] for the given global particle ids."""
    particle_gid = np.asarray(particle_gid, dtype=np.uint64).reshape(-1, 1, 1, 1)
    s = np.arange(S, dtype=np.uint64).reshape(1, -1, 1, 1)
    t = np.arange(T, dtype=np.uint64).reshape(1, 1, -1, 1)
    k = np.arange((n_dof + 1) // 2, dtype=np.uint64).reshape(1, 1, 1, -1)
    c0 = (t << np.uint64(8)) | k
    w = philox4x32(c0, s, particle_gid, np.uint64(draw), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=-1)


def normals(seed, draw, particle_gid, S, T, n_dof):
    """eps [NPg, S, T, d] (float64) in the trajectory layout [.., t, a*n + i]."""
    w = words(seed, draw, particle_gid, S, T, n_dof)
    z0, z1 = box_muller(w[..., 0], w[..., 1])
    z2, z3 = box_muller(w[..., 2], w[..., 3])
    npg = w.shape[0]
    n_full = n_dof // 2
    eps = np.zeros((npg, S, T, 2, n_dof))
    eps[:, :, :, 0, 0:2 * n_full:2] = z0[..., :n_full]
    eps[:, :, :, 0, 1:2 * n_full:2] = z1[..., :n_full]
    eps[:, :, :, 1, 0:2 * n_full:2] = z2[..., :n_full]
    eps[:, :, :, 1, 1:2 * n_full:2] = z3[..., :n_full]
    if n_dof % 2:
        eps[:, :, :, 0, n_dof - 1] = z0[..., n_full]
        eps[:, :, :, 1, n_dof - 1] = z1[..., n_full]
    return eps.reshape(npg, S, T, 2 * n_dof)
