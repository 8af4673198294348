"""Drop-in for ``SALib.sample.sobol`` on B200.

``sample`` keeps the reference signature and returns the same host ndarray
(src/SALib/sample/sobol.py:11-189); the whole body -- Sobol' draw, fast-forward, the
A/AB/BA/B triple loop and the bounds scaling -- is ONE CUDA kernel launch
(``sa_saltelli_sample_f64``).  ``sample_device`` is the same computation left in HBM as a
torch tensor, optionally sharded by Sobol' index range across the ranks of a
``torch.distributed`` job (no communication: point k is computed in closed form from k).
"""
import warnings
from typing import Dict, Optional, Union

import numpy as np

from .. import _lib, _runtime as rt
from ..util import _check_bounds, group_columns
from .directions import BITS, direction_vectors

__all__ = ["sample", "sample_device"]

_MAXN = 2 ** BITS


def _is_pow2_gt1(M: int) -> bool:
    return (M & (M - 1) == 0) and (M != 0 and M - 1 != 0)


def _validate_skip(N: int, skip_values) -> int:
    """Warnings / errors of sample/sobol.py:110-133, verbatim conditions."""
    if skip_values > 0 and isinstance(skip_values, int):
        M = skip_values
        if not _is_pow2_gt1(M):
            msg = f"""
            Convergence properties of the Sobol' sequence is only valid if
            `skip_values` ({M}) is a power of 2.
            """
            warnings.warn(msg, stacklevel=3)
        n_exp = int(np.log2(N))
        m_exp = int(np.log2(M))
        if n_exp > m_exp:
            msg = (
                "Convergence may not be valid as the number of "
                "requested samples is"
                f" > `skip_values` ({N} > {M})."
            )
            warnings.warn(msg, stacklevel=3)
        return M
    elif skip_values < 0 or not isinstance(skip_values, int):
        raise ValueError("`skip_values` must be a positive integer.")
    return 0


def _scramble_state(dims: int, scramble: bool, seed):
    """(sv uint32 [dims,30], shift uint32 [dims] or None).

    Unscrambled: our own Joe-Kuo table.  Scrambled: the LMS matrices and digital shift are
    host-supplied state (BASELINE north_star); they are taken from a ``scipy.stats.qmc.Sobol``
    engine built with the caller's arguments exactly as sample/sobol.py:107 builds it, so the seed
    semantics (int / None / Generator -> spawned child) are scipy's own."""
    if not scramble:
        return direction_vectors(dims), None
    from scipy.stats import qmc

    eng = qmc.Sobol(d=dims, scramble=True, seed=seed)
    return np.ascontiguousarray(eng._sv, dtype=np.uint32), np.ascontiguousarray(eng._shift, dtype=np.uint32)


def _uniform_bounds(problem: Dict):
    """lower bounds and spans as float64, validated like util/__init__.py:37-43."""
    lower, upper = _check_bounds(problem["bounds"])
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    if np.any(lower >= upper):
        raise ValueError("Bounds are not legal (upper bound must be greater than lower bound)")
    return lower, np.ascontiguousarray(upper - lower)


_VALID_DISTS = ["unif", "triang", "norm", "truncnorm", "lognorm", "logunif", "weibull"]


def _dist_tables(bounds, dists, D):
    """(codes int32 [D], params float64 [D, 4]) for the device inverse CDFs.

    Parameter checks, messages and the deprecated two-value triangular format follow
    ``_nonuniform_scale_samples`` (util/__init__.py:126-289) case by case."""
    from scipy.stats import norm

    if D != len(dists):
        msg = "Mismatch in number of parameters and distributions.\n"
        msg += "Num parameters: {}".format(D)
        msg += "Num distributions: {}".format(len(dists))
        raise ValueError(msg)
    code = np.zeros(D, dtype=np.int32)
    par = np.zeros((D, 4), dtype=np.float64)
    for i in range(D):
        b = bounds[i]
        b1, b2 = b[0], b[1]
        if dists[i] == "triang":
            if len(b) == 3:
                loc_start, b1, b2 = b[0], b[1], b[2]
            elif len(b) == 2:
                warnings.warn(
                    "Two-value format for triangular distributions detected.\n"
                    "To remove this message, specify the distribution start, "
                    "end, and peak (three values) "
                    "instead of the current two-value format "
                    "(distribution end and peak, with start assumed to be 0)\n"
                    "The two-value format will be deprecated in SALib v1.5.1",
                    DeprecationWarning, stacklevel=4)
                loc_start, b1, b2 = 0, b[0], b[1]
            else:
                raise ValueError("Unknown triangular distribution specification. Check problem specification.")
            if b1 < 0 or b2 < 0 or b2 >= 1 or loc_start > b1:
                raise ValueError(
                    """Triangular distribution bound error: Scale must be
                    greater than zero; peak on interval [0,1], triangular
                    start value must be smaller than end value""")
            code[i], par[i, :3] = 1, (loc_start, b1 - loc_start, b2)
        elif dists[i] == "unif":
            if b1 >= b2:
                raise ValueError(
                    """Uniform distribution: lower bound
                    must be less than upper bound""")
            code[i], par[i, :2] = 0, (b1, b2 - b1)
        elif dists[i] == "logunif":
            with np.errstate(all="ignore"):
                la, lb_ = np.log(np.float64(b1)), np.log(np.float64(b2))
            ok = 0 < b1 < b2
            code[i], par[i, :2] = 2, ((la, lb_ - la) if ok else (np.nan, np.nan))  # scipy returns nan
        elif dists[i] == "norm":
            if b2 <= 0:
                raise ValueError("""Normal distribution: stdev must be > 0""")
            code[i], par[i, :2] = 3, (b1, b2)
        elif dists[i] == "truncnorm":
            b3, b4 = b[2], b[3]
            if b4 <= 0:
                raise ValueError(
                    """Truncated normal distribution: stdev must
                    be > 0""")
            if b1 >= b2:
                raise ValueError(
                    """Truncated normal distribution: lower bound
                    must be less than upper bound""")
            lo, hi = (b1 - b3) / b4, (b2 - b3) / b4
            # mass of the standard normal on [lo, hi], from the tail that keeps precision
            mass = norm.cdf(hi) - norm.cdf(lo) if lo < 0 else norm.sf(lo) - norm.sf(hi)
            if lo < 0:
                code[i], par[i] = 4, (norm.cdf(lo), mass, b3, b4)
            else:
                code[i], par[i] = 5, (norm.cdf(-hi), mass, b3, b4)
        elif dists[i] == "lognorm":
            if b2 <= 0:
                raise ValueError("""Lognormal distribution: stdev must be > 0""")
            code[i], par[i, :2] = 6, (b1, b2)
        elif dists[i] == "weibull":
            if len(b) == 3:
                shape, scale, loc_start = b[0], b[1], b[2]
            elif len(b) == 2:
                shape, scale, loc_start = b[0], b[1], 0
            else:
                raise ValueError("Unknown Weibull distribution specification. Check problem specification.")
            if shape <= 0 or scale <= 0:
                raise ValueError("""Weibull distribution: shape and scale must be > 0""")
            code[i], par[i, :3] = 7, (1.0 / shape, scale, loc_start)
        else:
            raise ValueError("Distributions: choose one of %s" % ", ".join(_VALID_DISTS))
    return code, par


def _scaling(problem: Dict):
    """(lower, span, dist_code, dist_par); the last two are None for plain uniform bounds."""
    dists = problem.get("dists")
    if dists is None:
        lower, span = _uniform_bounds(problem)
        return lower, span, None, None
    code, par = _dist_tables(problem["bounds"], dists, problem["num_vars"])
    return None, None, code, par


class SamplePlan:
    """Everything the kernels need for one problem, resident on the device."""

    def __init__(self, problem: Dict, calc_second_order: bool, sv: np.ndarray, shift: Optional[np.ndarray]):
        self.D = int(problem["num_vars"])
        gcol, self.Dg = group_columns(problem)
        self.second_order = bool(calc_second_order)
        self.step = 2 * self.Dg + 2 if self.second_order else self.Dg + 2
        lower, span, code, par = _scaling(problem)
        import torch

        self.sv = rt.to_device(sv.astype(np.uint32, copy=False).view(np.int32), torch.int32)
        self.shift = None if shift is None else rt.to_device(shift.astype(np.uint32, copy=False).view(np.int32), torch.int32)
        self.gcol = None if gcol is None else rt.to_device(gcol, torch.int32)
        self.lb = None if lower is None else rt.to_device(lower, torch.float64)
        self.span = None if span is None else rt.to_device(span, torch.float64)
        self.dist_code = None if code is None else rt.to_device(code, torch.int32)
        self.dist_par = None if par is None else rt.to_device(par, torch.float64)

    def generate(self, first_index: int, n: int, out=None):
        """Rows of base indices [first_index, first_index+n) -> [step*n, D] float64 on the device."""
        import torch

        if out is None:
            out = rt.empty((self.step * n, self.D), torch.float64)
        _lib.call("sa_saltelli_sample_f64", rt.stream_ptr(), rt.ptr(self.sv), rt.ptr(self.shift),
                  int(first_index), int(n), self.D, self.Dg, rt.ptr(self.gcol), int(self.second_order),
                  rt.ptr(self.lb), rt.ptr(self.span), rt.ptr(self.dist_code), rt.ptr(self.dist_par), rt.ptr(out))
        return out


def _plan(problem, N, calc_second_order, scramble, skip_values, seed, caller_depth=3):
    D = problem["num_vars"]
    first = _validate_skip(N, skip_values)
    _scaling(problem)  # host-side validation before any device work
    sv, shift = _scramble_state(2 * D, scramble, seed)
    if first == 0 and N > 0 and (N & (N - 1)) != 0:
        # scipy's own warning on the first draw of an engine (qmc.Sobol._random)
        warnings.warn("The balance properties of Sobol' points require n to be a power of 2.",
                      stacklevel=caller_depth)
    if first + N > _MAXN:
        raise ValueError(
            f"At most 2**{BITS}={_MAXN} distinct points can be generated. {first} points have been "
            f"previously generated, then: n={first}+{N}={first + N}. Consider increasing `bits`.")
    rt.require_cuda()
    return SamplePlan(problem, calc_second_order, sv, shift), first


def sample_device(problem: Dict, N: int, *, calc_second_order: bool = True, scramble: bool = True,
                  skip_values: int = 0, seed=None, shard: Union[bool, tuple] = False):
    """Like :func:`sample` but returns a CUDA ``torch.Tensor`` (no host copy).

    ``shard=True`` makes each rank of the initialised ``torch.distributed`` job generate only the base
    rows ``[lo, hi)`` of its contiguous share (``shard=(rank, world)`` does the same explicitly); the
    returned tensor then holds rows ``[lo*step, hi*step)`` of the full matrix."""
    plan, first = _plan(problem, N, calc_second_order, scramble, skip_values, seed)
    lo, hi = 0, N
    if shard:
        rank, world = rt.world() if shard is True else shard
        lo, hi = rt.shard_range(N, rank, world)
    X = plan.generate(first + lo, hi - lo)
    problem["sample_scaled"] = True
    return X


def sample(
    problem: Dict,
    N: int,
    *,
    calc_second_order: bool = True,
    scramble: bool = True,
    skip_values: int = 0,
    seed: Optional[Union[int, np.random.Generator]] = None,
):
    """Saltelli's extension of the Sobol' sequence -- same contract as the reference.

    Returns a float64 ndarray ``[N*(2Dg+2), D]`` (``[N*(Dg+2), D]`` without second order), bit-identical
    to ``SALib.sample.sobol.sample`` for ``scramble=False`` and for ``scramble=True`` with an int seed.
    """
    plan, first = _plan(problem, N, calc_second_order, scramble, skip_values, seed)
    X = plan.generate(first, N)
    out = X.cpu().numpy()
    problem["sample_scaled"] = True
    return out
