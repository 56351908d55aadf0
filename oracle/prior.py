"""Oracle: constant-velocity GP prior — precision blocks and banded factor.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference being restated
  * precision  P = A^T W A           stoch_gpmp/costs/factors/mp_priors_multi.py:170-202
  * Phi, Q^-1                        stoch_gpmp/costs/factors/gp_factor.py:36-52
  * K = I / sigma^2                  stoch_gpmp/costs/factors/unary_factor.py:19
  * P -> scale_tril L                torch/distributions/multivariate_normal.py:79-85
                                     (Lf = chol(flip P); L = (flip(Lf)^T)^-1, i.e. P = U U^T, L = U^-T)

Because every factor weight is (2x2) (x) I_n, P decouples per DoF into n identical
2T x 2T block-tridiagonal matrices with 2x2 blocks (state order [pos, vel]); the flat
index of (t, a, i) in the reference's trajectory vector is t*d + a*n + i, d = 2n.
All functions here work on the per-DoF 2x2 blocks in float64.
"""
import numpy as np


def gp_blocks(dt, sigma_gp):
    """Phi (2x2) and Q^-1 (2x2) of one DoF.  gp_factor.py:36-52."""
    phi = np.array([[1.0, dt], [0.0, 1.0]])
    qc = 1.0 / sigma_gp ** 2
    qinv = np.array([[12.0 * dt ** -3.0, -6.0 * dt ** -2.0],
                     [-6.0 * dt ** -2.0, 4.0 * dt ** -1.0]]) * qc
    return phi, qinv


def precision_blocks(T, dt, sigma_start, sigma_gp, sigma_goal=None):
    """Per-DoF blocks of P.  Returns (D [T,2,2], O [T-1,2,2]) with
    D[t] = P[t,t], O[t] = P[t+1,t] = -Q^-1 Phi.   sigma_goal=None <=> not goal directed
    (mp_priors_multi.py:187-196)."""
    assert T >= 2
    phi, qinv = gp_blocks(dt, sigma_gp)
    ptqp = phi.T @ qinv @ phi
    D = np.zeros((T, 2, 2))
    for t in range(T):
        if t >= 1:
            D[t] += qinv
        if t <= T - 2:
            D[t] += ptqp
    D[0] += np.eye(2) / sigma_start ** 2
    if sigma_goal is not None:
        D[T - 1] += np.eye(2) / sigma_goal ** 2
    O = np.repeat((-qinv @ phi)[None], T - 1, axis=0)
    return D, O


def dense_from_blocks(D, O, n_dof):
    """Expand per-DoF blocks to the reference's dense [M,M] precision (M = T*2n)."""
    T = D.shape[0]
    d = 2 * n_dof
    P = np.zeros((T * d, T * d))
    eye = np.eye(n_dof)
    for t in range(T):
        P[t * d:(t + 1) * d, t * d:(t + 1) * d] = np.kron(D[t], eye)
        if t + 1 < T:
            blk = np.kron(O[t], eye)
            P[(t + 1) * d:(t + 2) * d, t * d:(t + 1) * d] = blk
            P[t * d:(t + 1) * d, (t + 1) * d:(t + 2) * d] = blk.T
    return P


def banded_factor(D, O):
    """Reverse block Cholesky P = U U^T, U upper block-bidiagonal
    (U[t,t] = A_t upper-triangular, U[t,t+1] = C_t), and the sampler tables

        y_t = G_t eps_t - H_t y_{t-1},   G_t = A_t^-T (lower),  H_t = A_t^-T C_{t-1}^T

    so that y = U^-T eps = L eps with L the reference's scale_tril
    (torch multivariate_normal.py:79-85: chol(flip P) flipped back is exactly this U).

    Written with explicit scalar IEEE-double operations in a fixed order (no FMA) so that the CUDA
    kernel `sgpmp_prior_factor_kernel` can mirror it operation for operation: the factor of an
    ill-conditioned P (cond 1e7..1e9 here) moves by cond*2^-53 under any re-association.
    Returns dict(A [T,2,2], C [T-1,2,2], G [T,2,2], H [T,2,2] (H[0]=0))."""
    T = D.shape[0]
    A = np.zeros((T, 2, 2))
    C = np.zeros((max(T - 1, 0), 2, 2))
    G = np.zeros((T, 2, 2))
    H = np.zeros((T, 2, 2))
    g11 = g21 = g22 = 0.0
    for t in range(T - 1, -1, -1):
        d11, d12, d22 = float(D[t, 0, 0]), float(D[t, 0, 1]), float(D[t, 1, 1])
        if t < T - 1:
            o11, o12, o21, o22 = (float(O[t, 0, 0]), float(O[t, 0, 1]), float(O[t, 1, 0]), float(O[t, 1, 1]))
            # C_t = O_t^T G_{t+1}
            c11 = o11 * g11 + o21 * g21
            c12 = o21 * g22
            c21 = o12 * g11 + o22 * g21
            c22 = o22 * g22
            C[t] = [[c11, c12], [c21, c22]]
            s11 = d11 - (c11 * c11 + c12 * c12)
            s12 = d12 - (c11 * c21 + c12 * c22)
            s22 = d22 - (c21 * c21 + c22 * c22)
        else:
            s11, s12, s22 = d11, d12, d22
        # S = A A^T, A upper
        if not (s22 > 0):
            raise ValueError("prior precision is not positive definite (pivot <= 0 at t=%d)" % t)
        a22 = float(np.sqrt(s22))
        a12 = s12 / a22
        r = s11 - a12 * a12
        if not (r > 0):
            raise ValueError("prior precision is not positive definite (pivot <= 0 at t=%d)" % t)
        a11 = float(np.sqrt(r))
        A[t] = [[a11, a12], [0.0, a22]]
        g11 = 1.0 / a11
        g22 = 1.0 / a22
        g21 = -(a12 * g11) * g22
        G[t] = [[g11, 0.0], [g21, g22]]
    for t in range(1, T):
        c = C[t - 1]
        g11, g21, g22 = G[t, 0, 0], G[t, 1, 0], G[t, 1, 1]
        H[t] = [[g11 * c[0, 0], g11 * c[1, 0]],
                [g21 * c[0, 0] + g22 * c[0, 1], g21 * c[1, 0] + g22 * c[1, 1]]]
    return dict(A=A, C=C, G=G, H=H)


def dense_scale_tril(G, H, n_dof):
    """Dense L [M,M] (= reference `_unbroadcasted_scale_tril[0]`) from the tables:
    column j of L is the recurrence applied to e_j."""
    T = G.shape[0]
    d = 2 * n_dof
    M = T * d
    Ldof = np.zeros((2 * T, 2 * T))
    for j in range(2 * T):
        e = np.zeros((T, 2))
        e[j // 2, j % 2] = 1.0
        y = np.zeros((T, 2))
        prev = np.zeros(2)
        for t in range(T):
            prev = G[t] @ e[t] - H[t] @ prev
            y[t] = prev
        Ldof[:, j] = y.reshape(-1)
    L = np.zeros((M, M))
    for i in range(n_dof):
        idx = np.array([t * d + a * n_dof + i for t in range(T) for a in range(2)])
        L[np.ix_(idx, idx)] = Ldof
    return L


def precision_times(D, O, mu):
    """b = P mu for mu [..., T, d] using the per-DoF blocks (banded mat-vec).
    Used for the importance-sampling term x^T Sigma^-1 mu (planner.py:234-236)."""
    T = D.shape[0]
    d = mu.shape[-1]
    n = d // 2
    m = mu.reshape(mu.shape[:-1] + (2, n)).astype(np.float64)      # [..., T, a, i]
    b = np.einsum('tab,...tbi->...tai', D, m)
    b[..., 1:, :, :] += np.einsum('tab,...tbi->...tai', O, m[..., :-1, :, :])
    b[..., :-1, :, :] += np.einsum('tba,...tbi->...tai', O, m[..., 1:, :, :])
    return b.reshape(mu.shape)


def const_vel_mean(start_state, goal_states, dt, T, n_dof):
    """Straight-line means of the INIT prior, one per goal [G,T,d].
    mp_priors_multi.py:130-144 (divides by (T-1)*dt; the planner's own
    'const_vel' initialiser, planner.py:142-155, divides by T*dt — see below)."""
    num_steps = T - 1
    G = goal_states.shape[0]
    out = np.zeros((G, T, 2 * n_dof), dtype=goal_states.dtype)
    for g in range(G):
        vel = (goal_states[g, :n_dof] - start_state[:n_dof]) / (num_steps * dt)
        for i in range(T):
            out[g, i, :n_dof] = start_state[:n_dof] * (num_steps - i) * 1. / num_steps \
                + goal_states[g, :n_dof] * i * 1. / num_steps
        out[g, :, n_dof:] = vel[None]
    return out


def const_vel_trajectories(start_state, goal_states, dt, T, n_dof, K):
    """initial_particle_means='const_vel' [G,K,T,d].  planner.py:142-155."""
    G = goal_states.shape[0]
    out = np.zeros((G, K, T, 2 * n_dof), dtype=goal_states.dtype)
    vel = (goal_states[:, :n_dof] - start_state[:n_dof]) / (T * dt)
    for i in range(T):
        interp = start_state[:n_dof] * (T - i - 1) / (T - 1) + goal_states[:, :n_dof] * i / (T - 1)
        out[:, :, i, :n_dof] = interp[:, None]
    out[:, :, :, n_dof:] = vel[:, None, None]
    return out
