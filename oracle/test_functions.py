"""Ishigami and Sobol' G test functions (oracle, NumPy).

Restates /root/reference/src/SALib/test_functions/Ishigami.py:55-59 and
Sobol_G.py:41-86 (per-row product in column order)."""
import numpy as np


def ishigami(X, A=7.0, B=0.1):
    return np.sin(X[:, 0]) + A * np.power(np.sin(X[:, 1]), 2) + B * np.power(X[:, 2], 4) * np.sin(X[:, 0])


def sobol_g(X, a=None, delta=None, alpha=None):
    if a is None:
        a = np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    delta = np.zeros_like(a) if delta is None else np.asarray(delta, dtype=np.float64)
    alpha = np.ones_like(a) if alpha is None else np.asarray(alpha, dtype=np.float64)
    if (X < 0).any():
        raise ValueError("Sobol G function called with values less than zero")
    if (X > 1).any():
        raise ValueError("Sobol G function called with values greater than one")
    sx = X + delta[None, :]
    mod_x = sx - np.modf(sx)[1]
    t = np.abs(2 * mod_x - 1) ** alpha[None, :]
    el = ((1 + alpha)[None, :] * t + a[None, :]) / (1 + a)[None, :]
    # sequential product over columns, like ndarray.prod on one row (Sobol_G.py:84)
    Y = np.ones(X.shape[0])
    for j in range(X.shape[1]):
        Y = Y * el[:, j]
    return Y
