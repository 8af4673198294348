"""Drop-in for ``SALib.analyze.discrepancy`` (src/SALib/analyze/discrepancy.py:10-121) on the device.

For every input ``i`` the reference computes ``scipy.stats.qmc.discrepancy`` of the 2-D projection
``[X_i, Y]`` (both scaled to the unit interval) and normalises the D values by their sum.  Each discrepancy is
an N^2 pair sum; here all inputs share one pass over the pairs (``sa_discrepancy_f64``,
salib_b200/csrc/discrepancy.cu), with ``k(y_a, y_b)`` computed once per pair."""
import numpy as np
import torch

from .. import _lib, _runtime as rt
from ..util import ResultDict, _check_groups, group_columns

__all__ = ["analyze", "METHODS"]

METHODS = {"WD": 0, "CD": 1, "MD": 2, "L2-star": 3}


def analyze(problem, X, Y, method="WD", print_to_console=False, seed=None):
    """Discrepancy indices ``s_discrepancy`` (+ ``names``); same contract as the reference.

    ``method``: "WD" | "CD" | "MD" | "L2-star" (see ``scipy.stats.qmc.discrepancy``).  ``seed`` is unused, as in
    the reference.  X ``[N, D]`` and Y ``[N]`` may be host arrays or CUDA tensors."""
    D = problem["num_vars"]
    groups = _check_groups(problem)
    if method not in METHODS:
        raise ValueError(f"{method!r} is not a valid method. It must be one of {set(METHODS)!r}")
    if X.ndim != 2:
        raise ValueError("Sample is not a 2D array")
    bounds = np.asarray(problem["bounds"], dtype=np.float64).T
    try:
        lower, upper = np.broadcast_to(bounds[0], X.shape[1]), np.broadcast_to(bounds[1], X.shape[1])
    except ValueError as exc:
        raise ValueError("'l_bounds' and 'u_bounds' must be broadcastable and respect the sample dimension") from exc
    if not np.all(lower < upper):
        raise ValueError("Bounds are not consistent 'l_bounds' < 'u_bounds'")

    rt.require_cuda()
    L = _lib.lib()
    Xd = rt.to_device(X, torch.float64)
    Yd = rt.to_device(Y, torch.float64).reshape(-1)
    N = int(Xd.shape[0])
    if int(Yd.numel()) != N:
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    lb = rt.to_device(np.ascontiguousarray(lower), torch.float64)
    ub = rt.to_device(np.ascontiguousarray(upper), torch.float64)
    gcol, Dg = group_columns(problem) if groups else (None, D)
    gdev = rt.to_device(gcol, torch.int32) if gcol is not None else None
    work = rt.empty((int(L.sa_discrepancy_workspace_bytes(N, D)),), torch.uint8)
    out = rt.empty((Dg,), torch.float64)
    flag = rt.zeros((1,), torch.int32)
    _lib.call("sa_discrepancy_f64", rt.stream_ptr(), rt.ptr(Xd), rt.ptr(Yd), N, D, rt.ptr(lb), rt.ptr(ub),
              METHODS[method], Dg, rt.ptr(gdev), rt.ptr(out), None, rt.ptr(work), int(work.numel()), rt.ptr(flag))
    f = int(flag.item())
    if f & 1:
        raise ValueError("Sample is out of bounds")
    if f & 2:
        raise ValueError("Bounds are not consistent 'l_bounds' < 'u_bounds'")

    Si = ResultDict((k, np.zeros(D)) for k in ("s_discrepancy",))
    Si["names"] = problem["names"]
    Si["s_discrepancy"] = out.cpu().numpy()
    if print_to_console:
        print(Si.to_df())
    return Si
