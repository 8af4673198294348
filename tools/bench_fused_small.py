"""Fused sample -> model evaluation at small D (Ishigami D=3, G function D=8): rows/s. One JSON line each."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200.sample import sobol as sm  # noqa: E402
from salib_b200.test_functions import Ishigami, Sobol_G  # noqa: E402

N = 1 << 22


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5e3


p3 = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}
plan, first = sm._plan(p3, N, True, False, 0, None)
Y = torch.empty((plan.step * N,), dtype=torch.float64, device="cuda")
t = timed(lambda: Ishigami.evaluate_sample(plan, first, N, out=Y))
print(json.dumps({"model": "Ishigami fused", "D": 3, "rows": plan.step * N, "ms": t * 1e3, "Grows_per_s": plan.step * N / t / 1e9}))
X = plan.generate(first, N)
t = timed(lambda: Ishigami.evaluate(X))
print(json.dumps({"model": "Ishigami from X", "D": 3, "rows": plan.step * N, "ms": t * 1e3, "Grows_per_s": plan.step * N / t / 1e9,
                  "read_GBps": X.numel() * 8 / t / 1e9}))
for D, c2 in ((8, False), (8, True), (16, True)):
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    plan, first = sm._plan(p, N, c2, False, 0, None)
    a = np.tile([0, 1, 4.5, 9, 99, 99, 99, 99], D // 8).astype(np.float64)
    Y = torch.empty((plan.step * N,), dtype=torch.float64, device="cuda")
    t = timed(lambda: Sobol_G.evaluate_sample(plan, first, N, a=a, out=Y, check=False))
    print(json.dumps({"model": "G fused", "D": D, "second_order": c2, "rows": plan.step * N, "ms": t * 1e3,
                      "Grows_per_s": plan.step * N / t / 1e9}))
    X = plan.generate(first, N)
    t = timed(lambda: Sobol_G.evaluate(X, a=a))
    print(json.dumps({"model": "G from X", "D": D, "second_order": c2, "rows": plan.step * N, "ms": t * 1e3,
                      "Grows_per_s": plan.step * N / t / 1e9, "read_GBps": X.numel() * 8 / t / 1e9}))
