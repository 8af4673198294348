"""Drop-in for ``SALib.sample.sobol_sequence`` (src/SALib/sample/sobol_sequence.py:49-100).

``sample(N, D)`` returns the first N points of the Joe-Kuo Sobol' sequence in D dimensions (point 0 is
the origin), generated on the device.  The reference scales 31-bit integers by 2^-31; the kernel uses
scipy's 30-bit words and 2^-30, which gives bit-identical values for N <= 2^30."""
import math

import numpy as np
import torch

from .. import _lib, _runtime as rt
from .directions import BITS, MAXDIM, direction_vectors

__all__ = ["sample", "index_of_least_significant_zero_bit"]


def index_of_least_significant_zero_bit(value):
    """1-based position of the lowest zero bit (reference :94-100)."""
    index = 1
    while (value & 1) != 0:
        value >>= 1
        index += 1
    return index


def sample(N, D, skip=0, device=False):
    """[N, D] float64 Sobol' points ``skip .. skip+N-1``; ``device=True`` returns a CUDA tensor."""
    if D > MAXDIM:
        raise ValueError("Error in Sobol sequence: not enough dimensions")
    if N + skip > 2 ** BITS or (N > 1 and int(math.ceil(math.log(N) / math.log(2))) > 31):
        raise ValueError("Error in Sobol sequence: not enough bits")
    rt.require_cuda()
    sv = rt.to_device(direction_vectors(D).view(np.int32), torch.int32)
    out = rt.empty((N, D), torch.float64)
    _lib.call("sa_sobol_points_f64", rt.stream_ptr(), rt.ptr(sv), None, int(skip), int(N), int(D), rt.ptr(out))
    return out if device else out.cpu().numpy()
