"""Oracle: the counter-based normal stream of the CUDA sampler.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws eps with torch's global generator
(torch/distributions/multivariate_normal.py:253, `_standard_normal` -> `normal_()`),
which cannot be reproduced bit-for-bit on a GPU.  Parity with the reference therefore
uses *injected* eps; this module pins the kernel's own stream so that the in-kernel
RNG path is checked too (integer words bit-exact, normals to float rounding).

Stream definition (stoch_gpmp_b200/csrc/sgpmp_rng.cuh follows this):
  Philox4x32-10 (Salmon et al., SC'11; constants below), key = (seed_lo, seed_hi),
  counter = ((tpair << 8) | dof,  sample s,  global particle id,  draw index)
  where tpair = t // 2 and global particle id = problem_gid * NP + p.
  The 4 output words give 4 normals by two Box-Muller pairs
      u1 = (w0 + 0.5) * 2^-32, u2 = (w1 + 0.5) * 2^-32
      r = sqrt(-2 ln u1), th = 2 pi u2 - pi
      eps[t=2*tpair,   pos, dof] = r cos th,   eps[t=2*tpair,   vel, dof] = r sin th
  and the same from (w2, w3) for t = 2*tpair + 1.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
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


def words(seed, draw, particle_gid, S, T, n_dof):
    """Raw Philox words [NPg, S, ceil(T/2), n_dof, 4] for the given global particle ids."""
    particle_gid = np.asarray(particle_gid, dtype=np.uint64).reshape(-1, 1, 1, 1)
    s = np.arange(S, dtype=np.uint64).reshape(1, -1, 1, 1)
    tp = np.arange((T + 1) // 2, dtype=np.uint64).reshape(1, 1, -1, 1)
    dof = np.arange(n_dof, dtype=np.uint64).reshape(1, 1, 1, -1)
    c0 = (tp << np.uint64(8)) | dof
    w = philox4x32_10(c0, s, particle_gid, np.uint64(draw), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=-1)


def normals(seed, draw, particle_gid, S, T, n_dof):
    """eps [NPg, S, T, d] (float64) in the trajectory layout [.., t, a*n + i]."""
    w = words(seed, draw, particle_gid, S, T, n_dof)
    z0, z1 = box_muller(w[..., 0], w[..., 1])
    z2, z3 = box_muller(w[..., 2], w[..., 3])
    npg = w.shape[0]
    TP = w.shape[2]
    eps = np.zeros((npg, S, 2 * TP, 2, n_dof))
    eps[:, :, 0::2, 0, :] = z0
    eps[:, :, 0::2, 1, :] = z1
    eps[:, :, 1::2, 0, :] = z2
    eps[:, :, 1::2, 1, :] = z3
    return eps[:, :, :T].reshape(npg, S, T, 2 * n_dof)
