"""Ishigami function on the device (reference: src/SALib/test_functions/Ishigami.py:4-61).

``evaluate(X)`` takes a host ndarray (returns ndarray) or a CUDA tensor (returns tensor).
``evaluate_sample`` is the fused path: Sobol' indices -> Y with X never written to memory."""
import numpy as np
import torch

from .. import _lib, _runtime as rt


def evaluate(X, A: float = 7.0, B: float = 0.1):
    is_np = isinstance(X, np.ndarray)
    Xd = rt.to_device(X, torch.float64)
    if Xd.ndim != 2 or Xd.shape[1] < 3:
        raise ValueError("Ishigami needs an [n, 3] input")
    if Xd.shape[1] != 3:
        Xd = Xd[:, :3].contiguous()
    Y = rt.empty((Xd.shape[0],), torch.float64)
    _lib.call("sa_eval_ishigami_f64", rt.stream_ptr(), rt.ptr(Xd), int(Xd.shape[0]), float(A), float(B), rt.ptr(Y))
    return Y.cpu().numpy() if is_np else Y


def evaluate_sample(plan, first_index: int, n: int, A: float = 7.0, B: float = 0.1, out=None):
    """Y [step*n] for base rows [first_index, first_index+n) of ``plan`` (a ``sample.sobol.SamplePlan``)."""
    if plan.D != 3:
        raise ValueError("Ishigami needs num_vars == 3")
    Y = out if out is not None else rt.empty((plan.step * n,), torch.float64)
    _lib.call("sa_saltelli_eval_ishigami_f64", rt.stream_ptr(), rt.ptr(plan.sv), rt.ptr(plan.shift),
              int(first_index), int(n), plan.Dg, rt.ptr(plan.gcol), int(plan.second_order), rt.ptr(plan.lb),
              rt.ptr(plan.span), rt.ptr(plan.dist_code), rt.ptr(plan.dist_par), float(A), float(B), rt.ptr(Y))
    return Y
