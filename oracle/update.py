"""Oracle: per-particle softmax weights and mean update.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: StochGPMP._update_distribution (stoch_gpmp/planner.py:263-275)
    w = softmax(-costs / tau, dim=1)            over the S samples of each particle
    grad = sum_s w_s (x_s - mu);  mu += step_size * grad
Identity used by the CUDA path: grad = L (sum_s w_s eps_s)   (x_s - mu = L eps_s).
"""
import numpy as np


def softmax_weights(costs, temperature):
    z = -costs / temperature
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def update(means, samples, costs, temperature, step_size):
    """means [NP,T,d], samples [NP,S,T,d], costs [NP,S] -> (new_means, grad, weights)."""
    w = softmax_weights(costs, temperature)
    grad = (w[:, :, None, None] * (samples - means[:, None])).sum(1)
    return means + step_size * grad, grad, w
