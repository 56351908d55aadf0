"""Oracle: x = mu + L eps.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: MultiMPPrior.sample (stoch_gpmp/costs/factors/mp_priors_multi.py:204-207) ->
MultivariateNormal.rsample (torch/distributions/multivariate_normal.py:251-254):
    eps = normal(S, NP, M);  x = loc + L @ eps;  view [S,NP,T,d] -> transpose -> [NP,S,T,d]
"""
import numpy as np


def banded_transform(G, H, eps):
    """y = L eps by the block-bidiagonal recurrence, eps [..., T, d] -> y [..., T, d].

        y_t = G_t eps_t - H_t y_{t-1}   per DoF, state order [pos, vel]
    """
    T = G.shape[0]
    d = eps.shape[-1]
    n = d // 2
    e = eps.reshape(eps.shape[:-1] + (2, n)).astype(np.float64)
    y = np.zeros_like(e)
    prev = np.zeros_like(e[..., 0, :, :])
    for t in range(T):
        cur = np.einsum('ab,...bi->...ai', G[t], e[..., t, :, :]) - np.einsum('ab,...bi->...ai', H[t], prev)
        y[..., t, :, :] = cur
        prev = cur
    return y.reshape(eps.shape)


def sample_banded(means, G, H, eps):
    """means [NP,T,d], eps [NP,S,T,d] -> samples [NP,S,T,d]."""
    return means[:, None] + banded_transform(G, H, eps)


def sample_dense(means, L, eps_ref_layout):
    """Reference-shaped evaluation: eps [S,NP,M] (sample-major, as torch draws it),
    L dense [M,M]; returns the logical [NP,S,T,d] tensor."""
    S, NP, M = eps_ref_layout.shape
    T, d = means.shape[-2:]
    x = means.reshape(1, NP, M) + eps_ref_layout @ L.T
    return np.transpose(x.reshape(S, NP, T, d), (1, 0, 2, 3))
