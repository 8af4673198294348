"""Host-side profile of one warm headline analyze call (where the time between device work goes)."""
import cProfile
import os
import pstats
import sys
import time
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

D, N, R = 64, 1 << 20, 10000
problem = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
avec = np.resize(np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64), D)
plan, first = sample_sobol._plan(problem, N, True, False, N, None)
Yd = rt.empty(((2 * D + 2) * N,), torch.float64)
Sobol_G.evaluate_sample(plan, first, N, a=avec, out=Yd, check=False)


def once():
    return sobol.analyze_device(problem, Yd, num_resamples=R, seed=100, resample="philox")


for _ in range(2):
    once()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    once()
    torch.cuda.synchronize()
    print("wall ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile()
pr.enable()
once()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
