"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's discrepancy analyzer.

Follows src/SALib/analyze/discrepancy.py:88-116.  The discrepancy itself is third-party arithmetic:
``scipy.stats.qmc.discrepancy`` (scipy >= 1.9.3 per pyproject.toml:36; 1.18.1 installed), whose Cython source
is not under /root/reference.  Its published closed forms (Hickernell 1998 for CD/WD, Zhou et al. 2013 for MD,
Warnock 1972 for L2-star; printed in the function's docstring) are restated below; ``oracle/gen_golden.py``
pins them against scipy itself and against the reference's ``discrepancy.analyze`` outputs.
O(N^2) memory per input: small cases only."""
import numpy as np


def _pair(method, u, v):
    if method == "WD":
        t = np.abs(u - v)
        return 1.5 - t * (1.0 - t)
    if method == "CD":
        return 1.0 + 0.5 * np.abs(u - 0.5) + 0.5 * np.abs(v - 0.5) - 0.5 * np.abs(u - v)
    if method == "MD":
        return (15.0 / 8.0 - 0.25 * np.abs(u - 0.5) - 0.25 * np.abs(v - 0.5) - 0.75 * np.abs(u - v)
                + 0.5 * (u - v) ** 2)
    if method == "L2-star":
        return 1.0 - np.maximum(u, v)
    raise ValueError(f"{method!r} is not a valid method")


def _point(method, u):
    if method == "CD":
        return 1.0 + 0.5 * np.abs(u - 0.5) - 0.5 * (u - 0.5) ** 2
    if method == "MD":
        return 5.0 / 3.0 - 0.25 * np.abs(u - 0.5) - 0.25 * (u - 0.5) ** 2
    if method == "L2-star":
        return 1.0 - u * u
    return None


def discrepancy(sample, method="CD"):
    """qmc.discrepancy(sample, method=method) for a unit-cube sample [n, d]."""
    x = np.asarray(sample, dtype=np.float64)
    n, d = x.shape
    P = np.ones((n, n))
    for k in range(d):
        P *= _pair(method, x[:, k, None], x[None, :, k])
    s2 = P.sum() / n ** 2
    if method == "WD":
        return -(4.0 / 3.0) ** d + s2
    S = np.ones(n)
    for k in range(d):
        S *= _point(method, x[:, k])
    s1 = S.sum()
    if method == "CD":
        return (13.0 / 12.0) ** d - 2.0 / n * s1 + s2
    if method == "MD":
        return (19.0 / 12.0) ** d - 2.0 / n * s1 + s2
    return np.sqrt(3.0 ** -d - 2.0 ** (1 - d) / n * s1 + s2)


def analyze(problem, X, Y, method="WD"):
    """discrepancy.py:88-116: s_discrepancy (group means if the problem has groups)."""
    D = problem["num_vars"]
    b = np.asarray(problem["bounds"], dtype=np.float64).T
    Xu = (np.asarray(X, dtype=np.float64) - b[0]) / (b[1] - b[0])          # qmc.scale(reverse=True), :92
    Yc = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
    Yu = (Yc - Yc.min()) / (Yc.max() - Yc.min())                            # :94
    s = np.array([discrepancy(np.concatenate([Xu[:, i, None], Yu], axis=1), method) for i in range(D)])
    s = s / np.sum(s)                                                       # :101
    groups = problem.get("groups")
    if groups and groups != problem["names"] and len(set(groups)) > 1:      # _check_groups
        g = np.array(groups)
        s = np.array([np.mean(s[g == name]) for name in dict.fromkeys(groups)])   # :103-111
    return s
