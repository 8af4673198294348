"""One pass of the non-GEMM kernels of the integer engine at the headline shape (moments, feature maxima, digit
split, Philox multiplicities) -- the target of `ncu -k regex:...` captures.

    python tools/prof_features.py [--log2n 20] [--resamples 2048] [--rows-of 8]
"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from salib_b200 import _runtime as rt  # noqa: E402
from salib_b200.analyze import sobol  # noqa: E402
from salib_b200.sample import sobol as sample_sobol  # noqa: E402
from salib_b200.test_functions import Sobol_G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=64)
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--resamples", type=int, default=2048)
ap.add_argument("--rows-of", type=int, default=0, help="also draw the multiplicities of the first of G row shards")
a = ap.parse_args()
D, N, R = a.dim, 1 << a.log2n, a.resamples
step = 2 * D + 2
problem = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
avec = np.resize(np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64), D)
plan, first = sample_sobol._plan(problem, N, True, False, N, None)
Yd = rt.empty((step * N,), torch.float64)
Sobol_G.evaluate_sample(plan, first, N, a=avec, out=Yd, check=False)
for rep in range(2):
    eng = sobol.SobolAnalyzer(D, N, True, algo="digits", resamples=R)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    eng.load(Yd)
    e[1].record()
    rs = sobol._Resampler(N, R, 100, "philox")
    w = rt.empty(((R + 1) * eng.Npad,), torch.uint8)
    rs.counts(0, R, w)
    e[2].record()
    if a.rows_of:
        n = N // a.rows_of
        w2 = rt.empty(((R + 1) * n,), torch.uint8)
        rs.counts(0, R, w2, rows=(0, n))
    e[3].record()
    torch.cuda.synchronize()
    print(f"load (moments+features) {e[0].elapsed_time(e[1]):.3f} ms, counts x{R} {e[1].elapsed_time(e[2]):.3f} ms, "
          f"counts of shard 1/{a.rows_of or 1} {e[2].elapsed_time(e[3]):.3f} ms")
