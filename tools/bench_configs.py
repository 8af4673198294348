#!/usr/bin/env python
"""Measurements of the BASELINE.json configs other than the headline (which bench.py covers), 1 GPU:

  cfg1  Ishigami D=3, N=1024, second order, R=100 (README quick start) -- drop-in calls, host arrays
  cfg2  G function D=8, N=2^20, first+total order, R=1000             -- first-order kernel vs L2 roofline
  cfg4  sample-generation sweep D=16..1024, N=2^16..2^24, scrambled / unscrambled vs HBM roofline
        (each point capped at --max-gb of output; larger points are generated in index chunks into
        a reused buffer, which is how a caller streams them)

One JSON object per line on stdout.  `python tools/bench_configs.py [--only cfg2] [--max-gb 64]`
"""
import argparse
import ctypes
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from salib_b200 import _lib, _runtime as rt  # noqa: E402
from salib_b200.analyze import sobol as an  # noqa: E402
from salib_b200.sample import saltelli, sobol as sm  # noqa: E402
from salib_b200.test_functions import Ishigami, Sobol_G  # noqa: E402

warnings.simplefilter("ignore")


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def cuda_time(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / steps


def gprob(D):
    return {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}


def cfg1():
    p = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}
    for _ in range(3):   # warm-up: CUDA context, library load, allocator
        an.analyze(p, Ishigami.evaluate(saltelli.sample(dict(p), 1024)), num_resamples=100, seed=100)
    t0 = time.perf_counter()
    for _ in range(20):
        X = saltelli.sample(dict(p), 1024)
    ts = (time.perf_counter() - t0) / 20
    Y = Ishigami.evaluate(X)
    t0 = time.perf_counter()
    for _ in range(20):
        S = an.analyze(p, Y, num_resamples=100, seed=100)
    ta = (time.perf_counter() - t0) / 20
    print(json.dumps({"config": "cfg1 Ishigami D=3 N=1024 c2 R=100 (host in/out, wall clock)",
                      "saltelli_sample_s": ts, "sample_rows_per_s": X.shape[0] / ts, "analyze_s": ta,
                      "S1": S["S1"].tolist(), "ST": S["ST"].tolist()}), flush=True)


def l2_peak():
    buf = rt.zeros((32 << 20) // 8, torch.float64)  # 32 MB, L2 resident
    out = rt.zeros((1,), torch.float64)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("sa_microbench_l2_read", rt.stream_ptr(), rt.ptr(buf), buf.numel() * 8, 64, rt.ptr(out))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, buf.numel() * 8 * 64 / (e0.elapsed_time(e1) / 1e3) / 1e9)
    return best


def cfg2():
    D, N, R = 8, 1 << 20, 1000
    p = gprob(D)
    plan, first = sm._plan(p, N, False, False, N, None)
    Y = Sobol_G.evaluate_sample(plan, first, N)
    events = []
    t = cuda_time(lambda: an.analyze_device(p, Y, calc_second_order=False, num_resamples=R, seed=100,
                                            resample="philox", kernel_events=events), steps=5, warmup=3)
    k_ms = sum(a.elapsed_time(b) for _, a, b in events[-5 * (len(events) // 8):]) / 5 if events else None
    evs = events[len(events) * 3 // 8:]
    k_s = sum(a.elapsed_time(b) for _, a, b in evs) / 1e3 / 5
    S = an.analyze_device(p, Y, calc_second_order=False, num_resamples=R, seed=100, resample="philox")
    l2 = l2_peak()
    gather_bytes = N * R * (D + 2) * 8  # one logical row per (resample, base row)
    Vi = 1.0 / (3.0 * (1.0 + np.array([0, 1, 4.5, 9, 99, 99, 99, 99.0])) ** 2)
    var = np.prod(1 + Vi) - 1
    print(json.dumps({"config": "cfg2 G-function D=8 N=2^20 first+total order R=1000 (Y resident, philox)",
                      "engine": getattr(S, "engine", None),
                      "analyze_s": t, "resample_rows_per_s": N * R / t, "accumulate_kernel_s": k_s,
                      "roofline": {"bound": "l2", "achieved": gather_bytes / k_s / 1e9, "peak": l2, "unit": "GB/s",
                                   "frac": gather_bytes / k_s / 1e9 / l2,
                                   "note": "algorithmic gather bytes N*R*(D+2)*8 over the first-order kernel time; "
                                           "peak = sa_microbench_l2_read over a 32 MB resident buffer"},
                      "S1": S["S1"].tolist(), "analytic_S1": (Vi / var).tolist(), "S1_conf": S["S1_conf"].tolist()}),
          flush=True)


def cfg4(max_gb):
    peak, src = hbm_peak()
    for scramble in (False, True):
        for D in (16, 32, 64, 128, 256, 512, 1024):
            for log2n in (16, 18, 20, 22, 24):
                N = 1 << log2n
                step = 2 * D + 2
                total_bytes = step * N * D * 8
                chunk_n = N
                while step * chunk_n * D * 8 > max_gb * 1e9:
                    chunk_n //= 2
                if chunk_n < 1024 or total_bytes > 40 * max_gb * 1e9:
                    continue
                p = gprob(D)
                sv, shift = sm._scramble_state(2 * D, scramble, 1234)
                plan = sm.SamplePlan(p, True, sv, shift)
                buf = rt.empty((step * chunk_n, D), torch.float64)

                def run():
                    for lo in range(0, N, chunk_n):
                        plan.generate(lo, chunk_n, out=buf)

                t = cuda_time(run, steps=3, warmup=1)
                print(json.dumps({"config": "cfg4", "D": D, "log2N": log2n, "scramble": scramble,
                                  "rows": step * N, "bytes": total_bytes, "chunks": N // chunk_n, "seconds": t,
                                  "Grows_per_s": step * N / t / 1e9, "GBps": total_bytes / t / 1e9,
                                  "hbm_frac": total_bytes / t / 1e9 / peak, "peak_source": src}), flush=True)
                del buf
                torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--max-gb", type=float, default=64.0)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    if a.only in ("", "cfg1"):
        cfg1()
    if a.only in ("", "cfg2"):
        cfg2()
    if a.only in ("", "cfg4"):
        cfg4(a.max_gb)
