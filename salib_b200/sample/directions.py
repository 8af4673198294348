"""Joe-Kuo direction vectors for the first ``dims`` Sobol' dimensions.

The packed table ``data/joe_kuo_6_21201.npz`` holds the numbers of the reference's
``src/SALib/sample/directions.py`` (new-joe-kuo-6.21201); the recurrence of
``src/SALib/sample/sobol_sequence.py:62-84`` runs in the C library (``sa_direction_vectors``).
"""
import ctypes as C
import functools
import os

import numpy as np

from .. import _lib

MAXDIM = 21201
BITS = 30
_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "joe_kuo_6_21201.npz")


@functools.lru_cache(maxsize=1)
def _table():
    z = np.load(_TABLE)
    return (np.ascontiguousarray(z["a"], dtype=np.uint32), np.ascontiguousarray(z["s"], dtype=np.uint8),
            np.ascontiguousarray(z["m"], dtype=np.uint32))


@functools.lru_cache(maxsize=8)
def direction_vectors(dims: int) -> np.ndarray:
    """uint32 [dims, 30]; equals ``scipy.stats.qmc.Sobol(dims, scramble=False)._sv``."""
    if dims > MAXDIM:
        raise ValueError(f"Maximum supported dimensionality is {MAXDIM}.")
    a, s, m = _table()
    sv = np.zeros((dims, BITS), dtype=np.uint32)
    _lib.call("sa_direction_vectors", a.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
              m.ctypes.data_as(C.c_void_p), int(dims), sv.ctypes.data_as(C.c_void_p))
    sv.setflags(write=False)
    return sv
