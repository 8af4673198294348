"""Sobol' G function on the device (reference: src/SALib/test_functions/Sobol_G.py:11-86).

Argument checks and exception types follow the reference; the range check on ``values`` is done by
the kernel (a flag read back after the launch)."""
import numpy as np
import torch

from .. import _lib, _runtime as rt

_DEFAULT_A = np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64)


def _params(a, delta, alpha):
    if a is None:
        a = _DEFAULT_A
    a = np.asarray(a, dtype=np.float64)
    if delta is not None:
        if not isinstance(delta, np.ndarray):
            raise TypeError("The argument `delta` must be given as a numpy ndarray")
        if ((delta < 0) | (delta > 1)).any():
            raise ValueError("Sobol G function called with delta values less than zero or greater than one")
    if alpha is not None:
        if not isinstance(alpha, np.ndarray):
            raise TypeError("The argument `alpha` must be given as a numpy ndarray")
        if (alpha <= 0.0).any():
            raise ValueError("Sobol G function called with alpha values less than or equal to zero")
    dev = lambda v: None if v is None else rt.to_device(np.asarray(v, dtype=np.float64), torch.float64)  # noqa: E731
    return dev(a), dev(delta), dev(alpha), a.shape[0]


def _raise_on_flag(flag):
    f = int(flag.item())
    if f & 1:
        raise ValueError("Sobol G function called with values less than zero")
    if f & 2:
        raise ValueError("Sobol G function called with values greater than one")


def evaluate(values, a=None, delta=None, alpha=None):
    if not isinstance(values, (np.ndarray, torch.Tensor)):
        raise TypeError("The argument `values` must be a numpy ndarray")
    is_np = isinstance(values, np.ndarray)
    da, dd, dl, na = _params(a, delta, alpha)
    X = rt.to_device(values, torch.float64)
    if X.ndim != 2 or X.shape[1] != na:
        raise ValueError(f"`values` must be [n, {na}] to match the parameter vector `a`")
    Y = rt.empty((X.shape[0],), torch.float64)
    flag = rt.zeros((1,), torch.int32)
    _lib.call("sa_eval_sobol_g_f64", rt.stream_ptr(), rt.ptr(X), int(X.shape[0]), int(X.shape[1]), rt.ptr(da),
              rt.ptr(dd), rt.ptr(dl), rt.ptr(Y), rt.ptr(flag))
    _raise_on_flag(flag)
    return Y.cpu().numpy() if is_np else Y


def evaluate_sample(plan, first_index: int, n: int, a=None, delta=None, alpha=None, out=None, check=True):
    """Fused path: Y [step*n] for base rows [first_index, first_index+n) of ``plan``, X never materialised."""
    da, dd, dl, na = _params(a, delta, alpha)
    if na != plan.D:
        raise ValueError(f"parameter vector `a` has {na} entries, problem has {plan.D} variables")
    Y = out if out is not None else rt.empty((plan.step * n,), torch.float64)
    flag = rt.zeros((1,), torch.int32)
    _lib.call("sa_saltelli_eval_sobol_g_f64", rt.stream_ptr(), rt.ptr(plan.sv), rt.ptr(plan.shift),
              int(first_index), int(n), plan.D, plan.Dg, rt.ptr(plan.gcol), int(plan.second_order),
              rt.ptr(plan.lb), rt.ptr(plan.span), rt.ptr(plan.dist_code), rt.ptr(plan.dist_par), rt.ptr(da),
              rt.ptr(dd), rt.ptr(dl), rt.ptr(Y), rt.ptr(flag))
    if check:
        _raise_on_flag(flag)
    return Y
