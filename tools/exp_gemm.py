"""A/B timing of the integer engine's pair GEMM at the headline shape (G function D=64, N=2^20, second order,
R=10^4): the real digit matrix and the real Philox multiplicities are built once, then ``sa_digit_sums_f64``
(GEMM + combine) is timed under several kernel configurations chosen through the SALIB_B200_GEMM_* variables
(read by the library at call time).  Every configuration must return bit-identical sums.

    python tools/exp_gemm.py [--log2n 20] [--resamples 10000] [--reps 3] [--configs a,b,...] [--one NAME]
"""
import argparse
import json
import os
import sys
import threading
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from salib_b200 import _lib, _runtime as rt  # noqa: E402
from salib_b200.analyze import sobol  # noqa: E402
from salib_b200.sample import sobol as sample_sobol  # noqa: E402
from salib_b200.test_functions import Sobol_G  # noqa: E402

CONFIGS = {
    "oneshot": {"SALIB_B200_GEMM_PERSIST": "0"},
    "persist_nosync": {"SALIB_B200_GEMM_SLACK": "0"},
    "persist_e16_s2": {},
    "persist_e16_s1": {"SALIB_B200_GEMM_SLACK": "1"},
    "persist_e8_s2": {"SALIB_B200_GEMM_EPOCH": "8"},
    "persist_e32_s2": {"SALIB_B200_GEMM_EPOCH": "32"},
    "persist_e16_s4": {"SALIB_B200_GEMM_SLACK": "4"},
    "persist_band10": {"SALIB_B200_GEMM_BAND": "10"},
    "persist_band13": {"SALIB_B200_GEMM_BAND": "13"},
    "persist_band5": {"SALIB_B200_GEMM_BAND": "5"},
    "persist_band20": {"SALIB_B200_GEMM_BAND": "20"},
    "persist_norsplit": {"SALIB_B200_GEMM_RSPLIT_MAX": "1"},
}
KEYS = ["SALIB_B200_GEMM_PERSIST", "SALIB_B200_GEMM_SLACK", "SALIB_B200_GEMM_EPOCH", "SALIB_B200_GEMM_BAND",
        "SALIB_B200_GEMM_RSPLIT_MAX", "SALIB_B200_GEMM_PAIRS"]

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=64)
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--resamples", type=int, default=10000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--configs", default=",".join(CONFIGS))
ap.add_argument("--env", default="", help="extra NAME=VALUE,... applied to every configuration")
a = ap.parse_args()

D, N, R = a.dim, 1 << a.log2n, a.resamples
step = 2 * D + 2
problem = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
avec = np.resize(np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64), D)
plan, first = sample_sobol._plan(problem, N, True, False, N, None)
Yd = rt.empty((step * N,), torch.float64)
Sobol_G.evaluate_sample(plan, first, N, a=avec, out=Yd, check=False)

eng = sobol.SobolAnalyzer(D, N, True, algo="digits", resamples=R)
eng.load(Yd)
del Yd
rs = sobol._Resampler(N, R, 100, "philox")
w = rt.empty(((R + 1) * eng.Npad,), torch.uint8)
rs.counts(0, R, w)
w[R * eng.Npad:].fill_(1)
torch.cuda.synchronize()


class Clocks:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        self.mhz, self.watts, self.stop = [], [], threading.Event()

    def __enter__(self):
        def loop():
            while not self.stop.is_set():
                self.mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.watts.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                self.stop.wait(0.02)
        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.t.join()


nacc, S = eng.nacc, eng.ndigits
ops = 2.0 * nacc * S * (R + 1) * N
ref = None
for name in a.configs.split(","):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(CONFIGS[name])
    for kv in filter(None, a.env.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    psum, _ = eng._partial_sums(w, R + 1)      # warm-up (also the correctness sample)
    torch.cuda.synchronize()
    time.sleep(0.5)
    eng.kernel_events = []
    with Clocks() as clk:
        for _ in range(a.reps):
            psum, _ = eng._partial_sums(w, R + 1)
        torch.cuda.synchronize()
    ms = [e0.elapsed_time(e1) for _, e0, e1 in eng.kernel_events]
    eng.kernel_events = None
    if ref is None:
        ref = psum.clone()
    same = bool(torch.equal(ref, psum))
    busy = [m for m in clk.mhz[len(clk.mhz) // 4:]] or [0]
    print(json.dumps({"config": name, "env": CONFIGS[name], "ms": [round(m, 3) for m in ms],
                      "ms_best": round(min(ms), 3), "tops": round(ops / (min(ms) / 1e3) / 1e12, 1),
                      "sm_mhz_median": float(np.median(busy)), "watts_max": max(clk.watts or [0]),
                      "identical_sums": same}), flush=True)
    assert same, f"{name}: sums differ from the first configuration"
