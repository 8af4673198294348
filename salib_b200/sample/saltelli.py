"""Drop-in for the deprecated ``SALib.sample.saltelli`` (src/SALib/sample/saltelli.py:12-203).

Same kernel as :mod:`salib_b200.sample.sobol`, unscrambled, with the legacy defaults: a mandatory
``DeprecationWarning`` and ``skip_values = max(2^ceil(log2 N), 16)`` when not given."""
import math
import warnings
from typing import Dict, Optional

from .. import _runtime as rt
from .directions import BITS, direction_vectors
from .sobol import SamplePlan

__all__ = ["sample"]


def _is_pow2_gt1(M: int) -> bool:
    return (M & (M - 1) == 0) and (M != 0 and M - 1 != 0)


def sample(problem: Dict, N: int, calc_second_order: bool = True, skip_values: Optional[int] = None):
    warnings.warn(
        "`salib.sample.saltelli` will be removed in SALib 1.5.1 Please use"
        " `salib.sample.sobol`",
        category=DeprecationWarning,
        stacklevel=2,
    )
    if not _is_pow2_gt1(N):
        msg = f"""
        Convergence properties of the Sobol' sequence is only valid if
        `N` ({N}) is equal to `2^n`.
        """
        warnings.warn(msg)

    if skip_values is None:
        skip_values = max(int(2 ** math.ceil(math.log(N) / math.log(2))), 16)
    elif skip_values > 0:
        M = skip_values
        if not _is_pow2_gt1(M):
            msg = f"""
            Convergence properties of the Sobol' sequence is only valid if
            `skip_values` ({M}) is a power of 2.
            """
            warnings.warn(msg)
        n_exp = int(math.log(N, 2))
        m_exp = int(math.log(M, 2))
        if n_exp > m_exp:
            msg = (
                "Convergence may not be valid as the number of "
                "requested samples is"
                f" > `skip_values` ({N} > {M})."
            )
            warnings.warn(msg)
    elif skip_values == 0:
        warnings.warn("Duplicate samples will be taken as no points are skipped.")
    else:
        assert (
            isinstance(skip_values, int) and skip_values >= 0
        ), "`skip_values` must be a positive integer."

    if N + skip_values > 2 ** BITS:
        raise ValueError("Error in Sobol sequence: not enough bits")
    D = problem["num_vars"]
    if 2 * D > 21201:
        raise ValueError("Error in Sobol sequence: not enough dimensions")
    rt.require_cuda()
    plan = SamplePlan(problem, calc_second_order, direction_vectors(2 * D), None)
    out = plan.generate(int(skip_values), N).cpu().numpy()
    problem["sample_scaled"] = True
    return out
