"""Sobol' analysis: point estimates + bootstrap confidence intervals (oracle, NumPy).

Two restatements of /root/reference/src/SALib/analyze/sobol.py:113-246,267-279:

* ``analyze_reference_form`` keeps the reference's cost structure (gathers ``A[r]``,
  ``AB[r, j]``, ``B[r]`` and the variance of ``[A[r]; B[r]]`` for every j and every pair,
  :147-182).  It is the timed CPU baseline ("port") in bench.py.
* ``analyze`` is the de-duplicated form (one gather per resample column, variance once per
  resample) used as the fast checker in the tests.  Both return the same dict
  layout as ``create_Si_dict`` (:249-264).
"""
import numpy as np
from scipy.stats import norm


def split(Yn, D, N, calc_second_order):
    """separate_output_values (:267-279) as views of Yn.reshape(N, step)."""
    step = 2 * D + 2 if calc_second_order else D + 2
    M = Yn.reshape(N, step)
    A = M[:, 0]
    B = M[:, step - 1]
    AB = M[:, 1 : 1 + D]
    BA = M[:, 1 + D : 1 + 2 * D] if calc_second_order else None
    return A, B, AB, BA


def _n_from_size(size, D, calc_second_order):
    if calc_second_order and size % (2 * D + 2) == 0:
        return size // (2 * D + 2)
    if (not calc_second_order) and size % (D + 2) == 0:
        return size // (D + 2)
    raise RuntimeError("Incorrect number of samples in model output file.")


def resample_indices(N, num_resamples, seed):
    """:113-117,144 -- int64 [N, R]; falsy seed uses the legacy global RNG."""
    if seed:
        return np.random.default_rng(seed).integers(N, size=(N, num_resamples))
    return np.random.randint(N, size=(N, num_resamples))


def first_order(A, AB, B):
    y = np.r_[A, B]
    if np.ptp(y) == 0:
        return np.float64(0.0)
    return np.mean(B * (AB - A), axis=0) / np.var(y, axis=0)


def total_order(A, AB, B):
    y = np.r_[A, B]
    if np.ptp(y) == 0:
        return np.float64(0.0)
    return 0.5 * np.mean((A - AB) ** 2, axis=0) / np.var(y, axis=0)


def second_order(A, ABj, ABk, BAj, B):
    y = np.r_[A, B]
    if np.ptp(y) == 0:
        return np.float64(0.0)
    Vjk = np.mean(BAj * ABk - A * B, axis=0) / np.var(y, axis=0)
    return Vjk - first_order(A, ABj, B) - first_order(A, ABk, B)


def _empty(D, R, keep, c2):
    S = {k: np.zeros(D) for k in ("S1", "S1_conf", "ST", "ST_conf")}
    if keep:
        S["S1_conf_all"] = np.zeros((R, D))
        S["ST_conf_all"] = np.zeros((R, D))
    if c2:
        S["S2"] = np.full((D, D), np.nan)
        S["S2_conf"] = np.full((D, D), np.nan)
    return S


def analyze_reference_form(D, Y, calc_second_order=True, num_resamples=100, conf_level=0.95,
                           keep_resamples=False, seed=None, r=None):
    """Serial path of analyze/sobol.py:113-182, loop for loop."""
    Y = np.asarray(Y, dtype=np.float64)
    N = _n_from_size(Y.size, D, calc_second_order)
    if not 0 < conf_level < 1:
        raise RuntimeError("Confidence level must be between 0-1.")
    Yn = (Y - Y.mean()) / Y.std()
    A, B, AB, BA = split(Yn, D, N, calc_second_order)
    if r is None:
        r = resample_indices(N, num_resamples, seed)
    Z = norm.ppf(0.5 + conf_level / 2)
    S = _empty(D, num_resamples, keep_resamples, calc_second_order)
    with np.errstate(all="ignore"):
        for j in range(D):
            S["S1"][j] = first_order(A, AB[:, j], B)
            s1c = first_order(A[r], AB[r, j], B[r])
            if keep_resamples:
                S["S1_conf_all"][:, j] = s1c
            var_diff = np.ptp(np.r_[A[r], B[r]])
            S["S1_conf"][j] = Z * s1c.std(ddof=1) if var_diff != 0.0 else 0.0
            S["ST"][j] = total_order(A, AB[:, j], B)
            stc = total_order(A[r], AB[r, j], B[r])
            if keep_resamples:
                S["ST_conf_all"][:, j] = stc
            S["ST_conf"][j] = Z * stc.std(ddof=1) if var_diff != 0.0 else 0.0
        if calc_second_order:
            for j in range(D):
                for k in range(j + 1, D):
                    S["S2"][j, k] = second_order(A, AB[:, j], AB[:, k], BA[:, j], B)
                    S["S2_conf"][j, k] = Z * second_order(
                        A[r], AB[r, j], AB[r, k], BA[r, j], B[r]).std(ddof=1)
    return S


def analyze(D, Y, calc_second_order=True, num_resamples=100, conf_level=0.95,
            keep_resamples=False, seed=None, r=None, chunk=16):
    """De-duplicated form: identical estimators, each resample column gathered once."""
    Y = np.asarray(Y, dtype=np.float64)
    N = _n_from_size(Y.size, D, calc_second_order)
    if not 0 < conf_level < 1:
        raise RuntimeError("Confidence level must be between 0-1.")
    Yn = (Y - Y.mean()) / Y.std()
    A, B, AB, BA = split(Yn, D, N, calc_second_order)
    if r is None:
        r = resample_indices(N, num_resamples, seed)
    R = r.shape[1]
    Z = norm.ppf(0.5 + conf_level / 2)
    S = _empty(D, R, keep_resamples, calc_second_order)
    iu = np.triu_indices(D, 1)

    def estimates(a, b, ab, ba):
        # a, b: [n, C]; ab, ba: [n, C, D]
        y = np.concatenate([a, b], axis=0)
        V = np.var(y, axis=0)  # [C]
        const = np.ptp(y) == 0
        s1 = np.mean(b[:, :, None] * (ab - a[:, :, None]), axis=0) / V[:, None]
        st = 0.5 * np.mean((a[:, :, None] - ab) ** 2, axis=0) / V[:, None]
        s2 = None
        if ba is not None:
            mab = np.mean(a * b, axis=0)  # [C]
            G = np.einsum("ncj,nck->cjk", ba, ab) / a.shape[0]
            # mean(BAj*ABk - A*B) = mean(BAj*ABk) - mean(A*B) up to rounding
            s2 = (G - mab[:, None, None]) / V[:, None, None] - s1[:, :, None] - s1[:, None, :]
        return s1, st, s2, const

    with np.errstate(all="ignore"):
        s1, st, s2, const = estimates(A[:, None], B[:, None], AB[:, None, :],
                                      None if BA is None else BA[:, None, :])
        if const:
            S["S1"][:] = 0.0
            S["ST"][:] = 0.0
            if calc_second_order:
                S["S2"][iu] = 0.0
        else:
            S["S1"][:] = s1[0]
            S["ST"][:] = st[0]
            if calc_second_order:
                S["S2"][iu] = s2[0][iu]
        s1_all = np.empty((R, D))
        st_all = np.empty((R, D))
        s2_all = np.empty((R, len(iu[0]))) if calc_second_order else None
        any_var = False
        for c0 in range(0, R, chunk):
            rc = r[:, c0 : c0 + chunk]
            a, b, ab = A[rc], B[rc], AB[rc]
            ba = BA[rc] if calc_second_order else None
            e1, et, e2, cst = estimates(a, b, ab, ba)
            any_var = any_var or (not cst)
            s1_all[c0 : c0 + chunk] = e1
            st_all[c0 : c0 + chunk] = et
            if calc_second_order:
                s2_all[c0 : c0 + chunk] = e2[:, iu[0], iu[1]]
        if keep_resamples:
            S["S1_conf_all"][:] = s1_all
            S["ST_conf_all"][:] = st_all
        if any_var:
            S["S1_conf"][:] = Z * s1_all.std(ddof=1, axis=0)
            S["ST_conf"][:] = Z * st_all.std(ddof=1, axis=0)
        if calc_second_order:
            S["S2_conf"][iu] = Z * s2_all.std(ddof=1, axis=0)
    return S


# ---- process-pool variant (reference: analyze/sobol.py:183-194,282-362) --------------------------
_POOL_STATE = {}


def _pool_task(task):
    """One (kind, j, k) unit of sobol_parallel (:282-299), on arrays inherited by fork."""
    A, B, AB, BA, r, Z = (_POOL_STATE[k] for k in ("A", "B", "AB", "BA", "r", "Z"))
    kind, j, k = task
    with np.errstate(all="ignore"):
        if kind == "S1":
            return task, first_order(A, AB[:, j], B)
        if kind == "S1_conf":
            return task, Z * first_order(A[r], AB[r, j], B[r]).std(ddof=1)
        if kind == "ST":
            return task, total_order(A, AB[:, j], B)
        if kind == "ST_conf":
            return task, Z * total_order(A[r], AB[r, j], B[r]).std(ddof=1)
        if kind == "S2":
            return task, second_order(A, AB[:, j], AB[:, k], BA[:, j], B)
        return task, Z * second_order(A[r], AB[r, j], AB[r, k], BA[r, j], B[r]).std(ddof=1)


def analyze_reference_form_parallel(D, Y, calc_second_order=True, num_resamples=100, conf_level=0.95,
                                    seed=None, r=None, n_processors=None):
    """parallel=True path of the reference: a task list of (index-kind, j, k) mapped over a process pool."""
    import multiprocessing as mp

    Y = np.asarray(Y, dtype=np.float64)
    N = _n_from_size(Y.size, D, calc_second_order)
    Yn = (Y - Y.mean()) / Y.std()
    A, B, AB, BA = split(Yn, D, N, calc_second_order)
    if r is None:
        r = resample_indices(N, num_resamples, seed)
    Z = norm.ppf(0.5 + conf_level / 2)
    _POOL_STATE.update(A=np.ascontiguousarray(A), B=np.ascontiguousarray(B), AB=np.ascontiguousarray(AB),
                       BA=None if BA is None else np.ascontiguousarray(BA), r=r, Z=Z)
    tasks = [(kind, j, None) for j in range(D) for kind in ("S1", "S1_conf", "ST", "ST_conf")]
    if calc_second_order:
        tasks += [(kind, j, k) for j in range(D) for k in range(j + 1, D) for kind in ("S2", "S2_conf")]
    S = _empty(D, r.shape[1], False, calc_second_order)
    with mp.get_context("fork").Pool(n_processors) as pool:
        for (kind, j, k), val in pool.imap_unordered(_pool_task, tasks, chunksize=8):
            if k is None:
                S[kind][j] = val
            else:
                S[kind][j, k] = val
    return S
