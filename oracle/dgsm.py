"""finite_diff sampling and DGSM analysis (oracle, NumPy).

Restates /root/reference/src/SALib/sample/finite_diff.py:69-93 and
/root/reference/src/SALib/analyze/dgsm.py:50-141 (``calc_vi_stats``, ``calc_dgsm``)."""
import numpy as np
from scipy.stats import norm

from . import saltelli, sobol_seq


def finite_diff_sample(problem, N, delta=0.01, skip_values=1024):
    """[N*(D+1), D]: base point, then the point with column j moved by base_j*delta and clipped to bounds_j."""
    D = problem["num_vars"]
    bounds = problem["bounds"]
    base = sobol_seq.points_closed_form(N, D, skip=skip_values)
    base = saltelli.scale_uniform(base, bounds)
    out = np.empty((N, D + 1, D))
    out[:, 0, :] = base
    for j in range(D):
        row = base.copy()
        row[:, j] = base[:, j] + base[:, j] * delta
        row[:, j] = np.maximum(np.minimum(row[:, j], bounds[j][1]), bounds[j][0])
        out[:, 1 + j, :] = row
    return out.reshape(N * (D + 1), D)


def dgsm_analyze(problem, X, Y, num_resamples=100, conf_level=0.95, r_list=None):
    """dict(vi, vi_std, dgsm, dgsm_conf); ``r_list[j]`` is the [R, N] index matrix of factor j (the reference
    draws one per factor from the global NumPy RNG, dgsm.py:133)."""
    D = problem["num_vars"]
    if Y.size % (D + 1):
        raise RuntimeError("Incorrect number of samples in model output file.")
    if not 0 < conf_level < 1:
        raise RuntimeError("Confidence level must be between 0-1.")
    N = Y.size // (D + 1)
    step = D + 1
    base = Y[0::step]
    Xb = X[0::step, :]
    S = {k: np.empty(D) for k in ("vi", "vi_std", "dgsm", "dgsm_conf")}
    Z = norm.ppf(0.5 + conf_level / 2.0)
    for j in range(D):
        pert = Y[j + 1::step]
        diff = X[j + 1::step, j] - Xb[:, j]
        dfdx = ((pert - base) / diff) ** 2
        S["vi"][j], S["vi_std"][j] = dfdx.mean(), dfdx.std()
        b = problem["bounds"][j]
        S["dgsm"][j] = dfdx.mean() * (b[1] - b[0]) ** 2 / (np.var(base) * np.pi ** 2)
        r = r_list[j] if r_list is not None else np.random.randint(N, size=(num_resamples, N))
        s = dfdx[r].mean(axis=1)
        S["dgsm_conf"][j] = Z * s.std(ddof=1)
    return S
