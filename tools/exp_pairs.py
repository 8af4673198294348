"""Second order at small Dg: pair-feature GEMV vs the 8x8 register-tile kernel (time of one analyze call).
    python tools/exp_pairs.py [--log2n 18] [--resamples 1184]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200.analyze import sobol  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=18)
ap.add_argument("--resamples", type=int, default=1184)
ap.add_argument("--dims", type=str, default="8,16,24,32,40,48,56,64")
a = ap.parse_args()
N, R = 1 << a.log2n, a.resamples
for D in [int(x) for x in a.dims.split(",")]:
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    g = torch.Generator(device="cuda").manual_seed(D)
    Y = torch.randn((N * (2 * D + 2),), dtype=torch.float64, device="cuda", generator=g)
    row = {"D": D, "N": N, "R": R}
    for algo in ("gemv", "tile8"):
        def run():
            return sobol.analyze_device(p, Y, num_resamples=R, seed=5, resample="philox", algo=algo)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S = run()
        e1.record()
        torch.cuda.synchronize()
        row[algo + "_ms"] = e0.elapsed_time(e1)
        row[algo + "_S1_0"] = float(S["S1"][0])
    print(json.dumps(row))
