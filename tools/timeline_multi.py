"""Per-kernel device timeline of one row-sharded analyze call on rank 0 of a torchrun job (headline shape by
default): where a multi-GPU step goes -- features, draws, GEMM batches, NCCL kernels, estimator algebra.

    torchrun --nproc-per-node 8 tools/timeline_multi.py [--log2n 20] [--resamples 10000] > timeline.txt
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=64)
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--resamples", type=int, default=10000)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

from salib_b200 import _runtime as rt  # noqa: E402
from salib_b200.analyze import sobol  # noqa: E402
from salib_b200.sample import sobol as sample_sobol  # noqa: E402
from salib_b200.test_functions import Sobol_G  # noqa: E402

D, N, R = a.dim, 1 << a.log2n, a.resamples
step = 2 * D + 2
problem = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
avec = np.resize(np.array([0, 1, 4.5, 9, 99, 99, 99, 99], dtype=np.float64), D)
plan, first = sample_sobol._plan(problem, N, True, False, N, None)
lo, hi = rt.shard_range(N, rank, world)
Y = rt.empty((step * (hi - lo),), torch.float64)
Sobol_G.evaluate_sample(plan, first + lo, hi - lo, a=avec, out=Y, check=False)


def once():
    return sobol.analyze_row_sharded(problem, [(lo, Y)], N, num_resamples=R, seed=100, resample="philox")


for _ in range(3):
    once()
torch.cuda.synchronize()
dist.barrier()
walls = []
for _ in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    once()
    torch.cuda.synchronize()
    walls.append((time.perf_counter() - t0) * 1e3)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    once()
    torch.cuda.synchronize()
if rank == 0:
    print(f"# row-sharded analyze, {world} ranks, D={D} N=2^{a.log2n} R={R}; wall per call on rank 0 (ms): "
          f"{[round(w, 2) for w in walls]}")
    dev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    t0 = dev[0].time_range.start
    prev_end = t0
    for e in dev:
        gap = (e.time_range.start - prev_end) / 1e3
        print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  dur {(e.time_range.end - e.time_range.start) / 1e3:9.3f} ms  "
              f"gap {gap:7.3f} ms  {e.name[:80]}")
        prev_end = max(prev_end, e.time_range.end)
    print(f"device span {(prev_end - t0) / 1e3:.3f} ms")
dist.barrier()
dist.destroy_process_group()
