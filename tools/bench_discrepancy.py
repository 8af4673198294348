"""Throughput of discrepancy.analyze on one GPU, with the scipy call the reference makes
timed on a bounded sample beside it.  One JSON line per method.

    python tools/bench_discrepancy.py [--log2n 17] [--dims 64]

Unit: pair-inputs/s = N^2 * D / time (all ordered pairs, although the kernel visits each unordered pair once).
FP64-pipe instructions per visited (pair, input): WD 4, CD 4, MD 5, L2-star 4 (DSETP.MAX + DADD + DFMA + select;
+ the shared k(y_a, y_b) per pair).  Pipe peak: 148 SMs * 64 lanes/clk * 1.965 GHz = 1.86e13 instructions/s."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200 import _lib, _runtime as rt  # noqa: E402
from salib_b200.analyze import discrepancy  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=17)
ap.add_argument("--dims", type=int, default=64)
ap.add_argument("--cpu-n", type=int, default=8192)
ap.add_argument("--methods", type=str, default="WD,CD,MD,L2-star")
a = ap.parse_args()
N, D = 1 << a.log2n, a.dims
p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand((N, D), dtype=torch.float64, device="cuda", generator=g)
Y = (X * torch.linspace(0.0, 1.0, D, dtype=torch.float64, device="cuda")).sum(dim=1)
L = _lib.lib()
peak = None
for m in a.methods.split(","):
    discrepancy.analyze(p, X, Y, method=m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.sa_launch_count(1)
    e0.record()
    S = discrepancy.analyze(p, X, Y, method=m)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3
    launches = int(L.sa_launch_count(1))
    instr = {"WD": 4, "CD": 4, "MD": 5, "L2-star": 4}[m]
    # visited pairs ~ N^2/2; FP64 pipe peak = 148 SMs * 64 lanes/clk * clock (FMA = 1 instruction)
    from scipy.stats import qmc
    n = min(a.cpu_n, N)
    xs = np.concatenate([X[:n, :1].cpu().numpy(), np.linspace(0, 1, n)[:, None]], axis=1)
    t0 = time.perf_counter()
    qmc.discrepancy(xs, method=m)
    tc = time.perf_counter() - t0
    print(json.dumps({"method": m, "N": N, "D": D, "seconds": t, "pair_inputs_per_s": N * N * D / t,
                      "fp64_instr_per_s_visited": (N * (N + 256) / 2) * (D * instr + 4) / t,
                      "gpu_launches": launches, "s_sum": float(S["s_discrepancy"].sum()),
                      "scipy_1thread_pair_inputs_per_s": n * n / tc, "scipy_sample": f"one input, n={n}"}))
