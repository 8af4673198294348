#!/usr/bin/env python
"""Experiment: tile-kernel cycles per processed row with random multiplicities vs all-ones weights
(no skipping, no per-warp work variance).  Separates the cost of warps drifting apart on the shared
tile ring from the per-row cost."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from salib_b200 import _lib, _runtime as rt
from salib_b200.analyze import sobol as an
from salib_b200.sample import sobol as sm
from salib_b200.test_functions import Sobol_G
warnings.simplefilter("ignore")
D, N, R = 64, 1 << int(os.environ.get("LOG2N", 18)), 1184
p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
plan, first = sm._plan(p, N, True, False, N, None)
Y = Sobol_G.evaluate_sample(plan, first, N, a=np.resize([0, 1, 4.5, 9, 99, 99, 99, 99.0], D))
eng = an.SobolAnalyzer(D, N, True)
eng.load(Y)
w = rt.empty((R * eng.Npad,), torch.uint8)
psum = rt.empty((R * eng.nacc,), torch.float64)
sms = eng.sms
def run(label, rows):
    def f():
        _lib.call("sa_sobol_accumulate_ex_f64", rt.stream_ptr(), rt.ptr(eng.Yn), rt.ptr(eng.Yt), N, D, 1, rt.ptr(w), R, 1, rt.ptr(psum), int(os.environ.get("ALGO", 4)))
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 2e3
    print(f"{label}: {t*1e3:.2f} ms, {t * sms * 1.965e9 / rows:.1f} clk per processed row per SM")
_lib.call("sa_counts_philox", rt.stream_ptr(), 7, N, 0, R, rt.ptr(w), None)
wc = w.reshape(R, eng.Npad)[:, :N]
nz = int((wc != 0).sum().item())
run("philox multiplicities", nz)
w.fill_(1)
run("all ones            ", R * N)
# same multiplicity pattern for all 8 warps of a block (no drift, still 63 % of rows)
_lib.call("sa_counts_philox", rt.stream_ptr(), 7, N, 0, R, rt.ptr(w), None)
wv = w.reshape(R, eng.Npad)
wv[:] = wv[(torch.arange(R, device=wv.device) // 8) * 8]
nz = int((wv[:, :N] != 0).sum().item())
run("block-uniform pattern", nz)
