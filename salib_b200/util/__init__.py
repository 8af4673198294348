"""Host-side helpers of the Sobol' path: group bookkeeping, bounds validation, result container.

Mirrors the behaviour (names, return values, exception types) of the reference helpers
  _check_groups          src/SALib/util/util_funcs.py:118-129
  _check_bounds          src/SALib/util/util_funcs.py:132-152
  compute_groups_matrix  src/SALib/util/__init__.py:317-350
  extract_group_names    src/SALib/util/__init__.py:292-314
  ResultDict             src/SALib/util/results.py:5-38
so that the drop-in ``sample`` / ``analyze`` functions validate and label exactly like SALib.
"""
from typing import Dict, List, Tuple

import numpy as np

from .results import ResultDict  # noqa: F401
from .util_funcs import read_param_file  # noqa: F401

__all__ = ["ResultDict", "ProblemSpec", "extract_group_names", "compute_groups_matrix", "read_param_file",
           "scale_samples"]


def _unique_in_order(labels) -> List:
    """First-appearance order, like ``pd.unique`` (util/__init__.py:339)."""
    seen = {}
    for g in labels:
        if g not in seen:
            seen[g] = len(seen)
    return list(seen)


def _check_groups(problem: Dict):
    """``False`` unless the problem defines more than one group that differ from ``names``."""
    groups = problem.get("groups")
    if not groups or groups == problem["names"] or len(set(groups)) == 1:
        return False
    return groups


def _check_bounds(bounds) -> Tuple[np.ndarray, np.ndarray]:
    """Split ``bounds`` into lower/upper arrays; ``ValueError`` if any lower >= upper."""
    b = np.array(bounds)
    lower, upper = b[:, 0], b[:, 1]
    if np.any(lower >= upper):
        raise ValueError("Bounds are not legal")
    return lower, upper


def compute_groups_matrix(groups: List):
    """(k-by-g 0/1 membership matrix, unique group names in first-appearance order)."""
    names = _unique_in_order(groups)
    pos = {g: i for i, g in enumerate(names)}
    G = np.zeros((len(groups), len(names)), dtype=int)
    for row, g in enumerate(groups):
        G[row, pos[g]] = 1
    return G, np.array(names)


def extract_group_names(p: Dict) -> Tuple[List, int]:
    """(unique group names, count); falls back to ``names`` when no groups are defined."""
    groups = p["groups"] if ("groups" in p and p["groups"]) else p["names"]
    names = [np.str_(g) if isinstance(g, str) else g for g in _unique_in_order(groups)]
    return names, len(names)


def group_columns(problem: Dict):
    """(group_of_col int32[D] or None, Dg) as the sampler uses them (sample/sobol.py:138-142,161)."""
    D = problem["num_vars"]
    groups = _check_groups(problem)
    if not groups:
        return None, D
    names = _unique_in_order(groups)
    pos = {g: i for i, g in enumerate(names)}
    return np.array([pos[g] for g in groups], dtype=np.int32), len(names)


def scale_samples(params, problem: Dict):
    """Scale unit-cube samples [n, D] to the problem's bounds / distributions on the device
    (reference util/__init__.py:56-90).  Uniform bounds are scaled in place like the reference (and bit-exactly);
    sets ``problem["sample_scaled"] = True`` and returns the scaled array."""
    import torch

    from .. import _lib, _runtime as rt
    from ..sample.sobol import _scaling

    lower, span, code, par = _scaling(problem)
    if problem.get("dists") is not None and params.shape[1] != len(problem["dists"]):
        raise ValueError("Mismatch in number of parameters and distributions.")
    is_np = isinstance(params, np.ndarray)
    X = rt.to_device(params, torch.float64)
    if X.data_ptr() == getattr(params, "data_ptr", lambda: None)():
        pass  # a CUDA tensor is scaled in place
    elif not is_np:
        X = X.clone()
    dev = lambda v, dt: None if v is None else rt.to_device(v, dt)  # noqa: E731
    # keep the device copies alive until the call has been enqueued (their pointers cross the C ABI)
    d_lb, d_span = dev(lower, torch.float64), dev(span, torch.float64)
    d_code, d_par = dev(code, torch.int32), dev(par, torch.float64)
    _lib.call("sa_scale_samples_f64", rt.stream_ptr(), rt.ptr(X), int(X.shape[0]), int(X.shape[1]),
              rt.ptr(d_lb), rt.ptr(d_span), rt.ptr(d_code), rt.ptr(d_par))
    problem["sample_scaled"] = True
    if not is_np:
        return X
    out = X.cpu().numpy()
    if code is None:
        params[...] = out  # the reference scales uniform bounds in place (np.multiply(..., out=params))
        return params
    return out


from .problem import ProblemSpec  # noqa: E402,F401
