"""Fused Sobol' -> G evaluation at the headline shape (D=64, N=2^20, second order): ms per pass."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200.sample import sobol as sm  # noqa: E402
from salib_b200.test_functions import Sobol_G  # noqa: E402

D = int(os.environ.get("DIM", 64))
N = 1 << int(os.environ.get("LOG2N", 20))
p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
plan, first = sm._plan(p, N, True, False, N, None)
a = np.tile([0, 1, 4.5, 9, 99, 99, 99, 99], D // 8).astype(np.float64)
Y = torch.empty((plan.step * N,), dtype=torch.float64, device="cuda")
f = lambda: Sobol_G.evaluate_sample(plan, first, N, a=a, out=Y, check=False)  # noqa: E731
f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5e3
print("D=%d fused G: %.3f ms, %.1f Grows/s" % (D, t * 1e3, plan.step * N / t / 1e9))
