"""Does the multiplicity draw of batch b+1 hide behind the GEMM of batch b?  Times, at the headline shape and batch
size: the pair GEMM alone, the slim Philox launch alone (per thread count), and both at once on two streams
(wall time from before both to after both, and the GEMM's own duration).

    python tools/exp_overlap.py [--log2n 20] [--batch 2560] [--threads 128,256,512,1024]
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
ap.add_argument("--batch", type=int, default=2560)
ap.add_argument("--threads", default="128,256,512,1024")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
D, N, B = a.dim, 1 << a.log2n, a.batch
step = 2 * D + 2
problem = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
avec = np.resize(np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64), D)
plan, first = sample_sobol._plan(problem, N, True, False, N, None)
Yd = rt.empty((step * N,), torch.float64)
Sobol_G.evaluate_sample(plan, first, N, a=avec, out=Yd, check=False)
eng = sobol.SobolAnalyzer(D, N, True, algo="digits", resamples=B)
eng.load(Yd)
del Yd
rs = sobol._Resampler(N, 4 * B, 100, "philox")
w0 = rt.empty(((B + 1) * eng.Npad,), torch.uint8)
w1 = rt.empty(((B + 1) * eng.Npad,), torch.uint8)
rs.counts(0, B, w0)
torch.cuda.synchronize()
main, side = torch.cuda.current_stream(), torch.cuda.Stream()


def ev():
    return torch.cuda.Event(enable_timing=True)


def gemm_alone():
    e0, e1 = ev(), ev()
    e0.record()
    eng.digit_sums(w0, B)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def counts_alone(slim):
    e0, e1 = ev(), ev()
    e0.record()
    rs.counts(B, B, w1, slim=slim)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def both(order):
    """order: which launch is enqueued first."""
    start, g0, g1, c0, c1, end = ev(), ev(), ev(), ev(), ev(), ev()
    start.record(main)
    side.wait_event(start)

    def do_gemm():
        g0.record(main)
        eng.digit_sums(w0, B)
        g1.record(main)

    def do_counts():
        with torch.cuda.stream(side):
            c0.record(side)
            rs.counts(B, B, w1, slim=True)
            c1.record(side)

    (do_gemm, do_counts)[order == "counts"]()
    (do_counts, do_gemm)[order == "counts"]()
    main.wait_event(c1)
    end.record(main)
    torch.cuda.synchronize()
    return start.elapsed_time(end), g0.elapsed_time(g1), c0.elapsed_time(c1)


gemm_alone()
out = {"batch": B, "gemm_alone_ms": [round(gemm_alone(), 3) for _ in range(a.reps)],
       "counts_wide_ms": [round(counts_alone(False), 3) for _ in range(a.reps)]}
print(json.dumps(out), flush=True)
for nt in a.threads.split(","):
    os.environ["SALIB_B200_SLIM_THREADS"] = nt
    counts_alone(True)
    rec = {"slim_threads": int(nt), "counts_slim_alone_ms": [round(counts_alone(True), 3) for _ in range(a.reps)]}
    for order in ("gemm", "counts"):
        both(order)
        rec[f"both_{order}_first(wall,gemm,counts)"] = [[round(x, 3) for x in both(order)] for _ in range(a.reps)]
    print(json.dumps(rec), flush=True)
