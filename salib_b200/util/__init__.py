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

__all__ = ["ResultDict", "ProblemSpec", "extract_group_names", "compute_groups_matrix", "read_param_file"]


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


from .problem import ProblemSpec  # noqa: E402,F401
