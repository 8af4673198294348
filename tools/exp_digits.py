"""Headline shape (G function D=64, N=2^20, second order, R=10^4) through the FP64 tile engine and the opt-in
exact-integer tensor-core engine (algo="digits"): whole-call time, GEMM time, result agreement.

    python tools/exp_digits.py [--log2n 20] [--resamples 10000] [--dim 64] [--digits 6] [--skip-fp64]
"""
import argparse
import json
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
ap.add_argument("--resamples", type=int, default=10000)
ap.add_argument("--digits", type=int, default=0, help="0 = the library default (7)")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--skip-fp64", action="store_true")
ap.add_argument("--first-order-only", action="store_true")
ap.add_argument("--trace", action="store_true", help="torch.profiler timeline of one digits call (device ops, gaps)")
a = ap.parse_args()

D, N, R = a.dim, 1 << a.log2n, a.resamples
c2 = not a.first_order_only
step = 2 * D + 2 if c2 else D + 2
problem = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
avec = np.resize(np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64), D)
plan, first = sample_sobol._plan(problem, N, c2, False, N, None)
Yd = rt.empty((step * N,), torch.float64)
Sobol_G.evaluate_sample(plan, first, N, a=avec, out=Yd, check=False)
sobol.DIGITS = a.digits or None
a.digits = sobol.digits_for(N)


def run(algo):
    events = []

    def once():
        return sobol.analyze_device(problem, Yd, calc_second_order=c2, num_resamples=R, seed=100, resample="philox",
                                    algo=algo, kernel_events=events, keep_resamples=True)

    S = once()
    torch.cuda.synchronize()
    events.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        S = once()
    e1.record()
    torch.cuda.synchronize()
    k_ms = sum(x.elapsed_time(y) for _, x, y in events) / a.reps
    return S, e0.elapsed_time(e1) / a.reps, k_ms


out = {"D": D, "N": N, "R": R, "digits": a.digits, "second_order": c2}
if a.trace:
    from torch.profiler import ProfilerActivity, profile

    run("digits")
    import time as _time
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        stamps = []
        for _ in range(3):
            t0 = _time.perf_counter()
            sobol.analyze_device(problem, Yd, calc_second_order=c2, num_resamples=R, seed=100, resample="philox",
                                 algo="digits")
            stamps.append((_time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
    print("host wall per call (ms):", [round(x, 2) for x in stamps], file=sys.stderr)
    dev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    # idle gaps of the device longer than 0.2 ms (between calls, or inside one)
    pe = dev[0].time_range.end
    for e in dev[1:]:
        if (e.time_range.start - pe) / 1e3 > 0.2:
            print(f"IDLE {(e.time_range.start - pe) / 1e3:8.3f} ms before {e.name[:60]} at "
                  f"{(e.time_range.start - dev[0].time_range.start) / 1e3:9.3f} ms", file=sys.stderr)
        pe = max(pe, e.time_range.end)
    n3 = len(dev) // 3
    dev = dev[2 * n3:]          # the timeline of the last call
    t0 = dev[0].time_range.start
    prev_end = t0
    for e in dev:
        gap = (e.time_range.start - prev_end) / 1e3
        print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  dur {(e.time_range.end - e.time_range.start) / 1e3:9.3f} ms  "
              f"gap {gap:7.3f} ms  {e.name[:70]}", file=sys.stderr)
        prev_end = max(prev_end, e.time_range.end)
    print(f"device span {(prev_end - t0) / 1e3:.3f} ms", file=sys.stderr)
Sd, ms_d, k_d = run("digits")
nacc = sobol._nacc(D, c2)
ops = 2.0 * (R + 1) * N * nacc * a.digits          # +1: the point estimate runs the same GEMM over one row of ones
out["digits_ms"] = ms_d
out["digits_accumulate_ms"] = k_d
out["digits_int8_TOPS"] = ops / (k_d / 1e3) / 1e12
if not a.skip_fp64:
    Sf, ms_f, k_f = run("fp64")
    out["fp64_ms"] = ms_f
    out["fp64_accumulate_ms"] = k_f
    out["speedup"] = ms_f / ms_d
    iu = np.triu_indices(D, 1)

    def rel(x, y):
        x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        return float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-300)))

    def absd(x, y):
        return float(np.max(np.abs(np.asarray(x) - np.asarray(y))))

    for k in ("S1", "ST", "S1_conf", "ST_conf", "S1_conf_all", "ST_conf_all"):
        out[f"maxrel_{k}"] = rel(Sd[k], Sf[k])
    if c2:
        for k in ("S2", "S2_conf"):
            out[f"maxrel_{k}"] = rel(Sd[k][iu], Sf[k][iu])
            out[f"maxabs_{k}"] = absd(Sd[k][iu], Sf[k][iu])
print(json.dumps(out))
