"""Sampling throughput at small D (the generic Saltelli kernel) vs the HBM write roofline. One JSON line per D."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200.sample import sobol as sm  # noqa: E402

peak = 6546.2
p_json = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p_json):
    peak = json.load(open(p_json))["hbm_gbs"]
N = 1 << 22
for D, c2 in [(2, True), (3, True), (5, True), (8, False), (8, True), (12, True), (16, True), (24, True)]:
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    plan, first = sm._plan(p, N, c2, False, 0, None)
    X = plan.generate(first, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.generate(first, N, out=X)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5e3
    nbytes = X.numel() * 8
    print(json.dumps({"D": D, "second_order": c2, "N": N, "GB": nbytes / 1e9, "ms": t * 1e3, "GBps": nbytes / t / 1e9,
                      "hbm_frac": nbytes / t / 1e9 / peak, "Grows_per_s": X.shape[0] / t / 1e9}))
