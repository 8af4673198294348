"""Saltelli cross-matrix build + scaling (oracle, NumPy).

Restates /root/reference/src/SALib/sample/sobol.py:103-188 (and the same loop in
sample/saltelli.py:147-203), /root/reference/src/SALib/util/__init__.py:25-90
(``scale_samples``: multiply, then add -- two roundings) and the group helpers
util/util_funcs.py:118-129, util/__init__.py:317-350.
"""
import numpy as np

from . import sobol_seq


def check_groups(problem):
    """util/util_funcs.py:118-129: False unless >1 distinct group and groups != names."""
    groups = problem.get("groups")
    if not groups:
        return False
    if groups == problem["names"]:
        return False
    if len(set(groups)) == 1:
        return False
    return groups


def group_index(groups):
    """First-appearance order of the labels (pd.unique, util/__init__.py:339-342).
    Returns (group_of_col[D], unique_names)."""
    names = []
    for g in groups:
        if g not in names:
            names.append(g)
    return np.array([names.index(g) for g in groups], dtype=np.int32), names


def cross_matrix(base, D, group_of_col, Dg, calc_second_order):
    """base [N, 2D] -> [step*N, D] in the order A, AB_0.., (BA_0..,) B per base row
    (sample/sobol.py:149-186)."""
    N = base.shape[0]
    step = 2 * Dg + 2 if calc_second_order else Dg + 2
    A = base[:, :D]
    B = base[:, D : 2 * D]
    out = np.empty((N, step, D), dtype=np.float64)
    out[:, 0, :] = A
    for k in range(Dg):
        m = group_of_col == k
        out[:, 1 + k, :] = np.where(m[None, :], B, A)
        if calc_second_order:
            out[:, 1 + Dg + k, :] = np.where(m[None, :], A, B)
    out[:, step - 1, :] = B
    return out.reshape(N * step, D)


def scale_uniform(X, bounds):
    """util/__init__.py:25-53: in place x*(ub-lb) then +lb; ValueError if lb >= ub."""
    b = np.array(bounds, dtype=np.float64)
    lb, ub = b[:, 0], b[:, 1]
    if np.any(lb >= ub):
        raise ValueError("Bounds are not legal")
    np.multiply(X, (ub - lb), out=X)
    np.add(X, lb, out=X)
    return X


def sample(problem, N, calc_second_order=True, skip_values=0, sv=None, shift=None, closed_form=False):
    """Unscrambled (or host-state scrambled) Saltelli sample, uniform ``bounds`` only.

    sample/sobol.py:103-188 with ``qmc.Sobol(2D).fast_forward(skip).random(N)`` replaced by
    the oracle generator (bit-identical)."""
    D = problem["num_vars"]
    groups = check_groups(problem)
    if not groups:
        gcol, Dg = np.arange(D, dtype=np.int32), D
    else:
        gcol, names = group_index(groups)
        Dg = len(names)
    gen = sobol_seq.points_closed_form if closed_form else sobol_seq.points
    base = gen(N, 2 * D, skip=skip_values, sv=sv, shift=shift)
    X = cross_matrix(base, D, gcol, Dg, calc_second_order)
    return scale_uniform(X, problem["bounds"])
