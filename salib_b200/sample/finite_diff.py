"""Drop-in for ``SALib.sample.finite_diff`` (src/SALib/sample/finite_diff.py:10-93) on the device.

DGSM design: N Sobol' points (after ``skip_values``) scaled to the problem, each followed by D rows in which
one coordinate is moved by ``x_j * delta`` and clipped to ``bounds[j]``.  Three kernel launches:
plain Sobol' points -> scaling -> row expansion."""
from typing import Dict, Optional, Union

import numpy as np
import torch

from .. import _lib, _runtime as rt
from ..util import scale_samples
from ..util.util_funcs import handle_seed
from . import sobol_sequence

__all__ = ["sample"]


def sample(problem: Dict, N: int, delta: float = 0.01,
           seed: Optional[Union[int, np.random.Generator]] = None, skip_values: int = 1024, device: bool = False):
    """[N*(D+1), D] DGSM sample matrix (ndarray; CUDA tensor with ``device=True``)."""
    handle_seed(seed)  # the reference validates the seed and otherwise ignores it (finite_diff.py:64)
    D = problem["num_vars"]
    bounds = problem["bounds"]
    base = sobol_sequence.sample(N, D, skip=skip_values, device=True)
    base = scale_samples(base, problem)
    lo = rt.to_device(np.array([b[0] for b in bounds], dtype=np.float64), torch.float64)
    hi = rt.to_device(np.array([b[1] for b in bounds], dtype=np.float64), torch.float64)
    out = rt.empty((N * (D + 1), D), torch.float64)
    _lib.call("sa_finite_diff_expand_f64", rt.stream_ptr(), rt.ptr(base), int(N), int(D), float(delta),
              rt.ptr(lo), rt.ptr(hi), rt.ptr(out))
    return out if device else out.cpu().numpy()
