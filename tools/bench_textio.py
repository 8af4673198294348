"""Throughput of the device text formatter / parser with np.savetxt / np.loadtxt timed
beside them on a bounded sample.  One JSON line.

    python tools/bench_textio.py [--rows 262144] [--cols 64]"""
import argparse
import io
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200 import _lib  # noqa: E402
from salib_b200.util import textio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=262144)
ap.add_argument("--cols", type=int, default=64)
ap.add_argument("--cpu-rows", type=int, default=8192)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(3)
X = (torch.rand((a.rows, a.cols), dtype=torch.float64, device="cuda", generator=g) - 0.5) * 6.283185307179586


def cuda_time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps, out


L = _lib.lib()
L.sa_launch_count(1)
t_fmt, text = cuda_time(lambda: textio.format_rows(X))
t_parse, back = cuda_time(lambda: textio.parse_text(text))
launches = int(L.sa_launch_count(1))
nbytes = int(text.numel())
Xh = X[: a.cpu_rows].cpu().numpy()
t0 = time.perf_counter()
buf = io.BytesIO()
np.savetxt(buf, Xh, fmt="%.8e")
t_np_fmt = time.perf_counter() - t0
t0 = time.perf_counter()
Xb = np.loadtxt(io.BytesIO(buf.getvalue()))
t_np_parse = time.perf_counter() - t0
ok = bool(np.array_equal(back[: a.cpu_rows].cpu().numpy(), Xb)) and text[: len(buf.getvalue())].cpu().numpy().tobytes() == buf.getvalue()
n = a.rows * a.cols
print(json.dumps({"values": n, "text_bytes": nbytes, "format_s": t_fmt, "format_values_per_s": n / t_fmt,
                  "format_text_GBps": nbytes / t_fmt / 1e9, "parse_s": t_parse, "parse_values_per_s": n / t_parse,
                  "parse_text_GBps": nbytes / t_parse / 1e9, "gpu_launches": launches,
                  "numpy_savetxt_values_per_s": Xh.size / t_np_fmt, "numpy_loadtxt_values_per_s": Xh.size / t_np_parse,
                  "numpy_sample": f"{a.cpu_rows} x {a.cols}, 1 core", "identical_to_numpy": ok}))
