"""salib_b200 -- B200-native Sobol'/Saltelli sampling and ``sobol.analyze``.

Drop-ins for the SALib hot path (same names, arguments, return values and exceptions):

    from salib_b200.sample import sobol, saltelli      # SALib.sample.sobol / saltelli
    from salib_b200.analyze import sobol               # SALib.analyze.sobol
    from salib_b200.util import ProblemSpec            # sample_sobol / analyze_sobol
    from salib_b200.test_functions import Ishigami, Sobol_G

Python here is host glue only: every number is produced by hand-written sm_100a CUDA kernels behind
the C ABI of ``include/salib_b200.h`` (``salib_b200/lib/libsalib_b200.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def library_path():
    return _lib.LIB_PATH
