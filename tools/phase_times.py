"""Where a headline analyze call goes, WITHOUT a profiler attached: CUDA events around the call and around every GEMM
batch (the library's own `kernel_events` hook) give prologue (call start -> first GEMM), GEMM phase (first GEMM start
-> last GEMM end, i.e. GEMMs + what sits between them) and tail (last GEMM end -> results on the host), next to the
host wall clock.  One process, or any torchrun world size (row shards; max over ranks).

    python tools/phase_times.py [--first-batch 0,512,1024] [--reps 5]
    torchrun --nproc-per-node 2 tools/phase_times.py
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=64)
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--resamples", type=int, default=10000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--first-batch", default="0", help="comma list of DIGITS_FIRST_BATCH values to compare")
ap.add_argument("--drop-in", action="store_true", help="go through analyze() (auto routing) instead of row shards")
ap.add_argument("--no-gc", action="store_true", help="gc.disable() around the timed calls (as timeit does)")
ap.add_argument("--numa", action="store_true", help="bind the process to the GPU's NUMA node before pinning Y")
ap.add_argument("--batch", default="0", help="comma list of forced GEMM batch sizes (0 = the library's rule)")
ap.add_argument("--sampler", action="store_true", help="run bench.py's NVML clock sampler thread beside the calls")
ap.add_argument("--host-y", action="store_true", help="the e2e leg: Y in pinned host memory, through the drop-in")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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
Yfull = None
if a.drop_in or a.host_y:
    Yfull = rt.empty((step * N,), torch.float64)
    Sobol_G.evaluate_sample(plan, first, N, a=avec, out=Yfull, check=False)
if a.host_y:
    if a.numa:
        print("numa cpus", rank, (rt.bind_host_to_gpu() or [])[:4], flush=True)
    Yh = torch.empty(Yfull.shape, dtype=torch.float64, pin_memory=True)
    Yh.copy_(Yfull)
    torch.cuda.synchronize()
    Yfull = Yh.numpy()


def once(events=None):
    if a.drop_in or a.host_y:
        return sobol.analyze_device(problem, Yfull, num_resamples=R, seed=100, resample="philox",
                                    kernel_events=events)
    return sobol.analyze_row_sharded(problem, [(lo, Y)], N, num_resamples=R, seed=100, resample="philox",
                                     kernel_events=events)


import gc  # noqa: E402

_gc_log, _gc_t = [], [0.0]


def _gc_cb(phase, info):
    if phase == "start":
        _gc_t[0] = time.perf_counter()
    else:
        _gc_log.append((info["generation"], round((time.perf_counter() - _gc_t[0]) * 1e3, 2)))


gc.callbacks.append(_gc_cb)
_host_marks = []
_orig_digit_sums = sobol.SobolAnalyzer.digit_sums


def _marked_digit_sums(self, *args, **kw):
    _host_marks.append(time.perf_counter())      # when the HOST reaches each GEMM launch
    return _orig_digit_sums(self, *args, **kw)


sobol.SobolAnalyzer.digit_sums = _marked_digit_sums


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


_sampler = None
if a.sampler:
    import bench  # noqa: E402

    _sampler = bench.ClockSampler(local)
    _sampler.__enter__()
_rule = sobol.digits_batch
for fb, forced in [(int(x), int(y)) for x in a.first_batch.split(",") for y in a.batch.split(",")]:
    sobol.DIGITS_FIRST_BATCH = fb
    sobol.digits_batch = (lambda R_, npad_, ncols_=15043, _f=forced: min(_f, max(R_, 1))) if forced else _rule
    for _ in range(3):
        S = once()
    rows = []
    del _gc_log[:]
    if a.no_gc:
        gc.collect()
        gc.disable()
    for _ in range(a.reps):
        barrier()
        ev = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        del _host_marks[:]
        t0 = time.perf_counter()
        e0.record()
        S = once(ev)
        t_ret = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        gemm = sum(x[1].elapsed_time(x[2]) for x in ev)
        rows.append([wall, e0.elapsed_time(e1), e0.elapsed_time(ev[0][1]), ev[0][1].elapsed_time(ev[-1][2]),
                     ev[-1][2].elapsed_time(e1), gemm, len(ev), (_host_marks[0] - t0) * 1e3,
                     (_host_marks[-1] - t0) * 1e3, (t_ret - t0) * 1e3])
    gc.enable()
    if rank == 0 and os.environ.get("PHASE_VERBOSE"):
        print("gc collections during the timed calls (generation, ms):", _gc_log)
    t = torch.tensor(rows, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0 and os.environ.get("PHASE_VERBOSE"):
        st = torch.cuda.memory_stats()
        print("per-rep [wall, device, prologue, gemm_phase, tail, gemm_sum, batches, host_first, host_last, host_ret]:")
        for r_ in t.tolist():
            print("   ", [round(x, 2) for x in r_])
        print("allocator:", {k: st[k] for k in ("num_alloc_retries", "num_device_alloc", "num_device_free",
                                                "reserved_bytes.all.current", "allocated_bytes.all.peak")})
    if rank == 0:
        med = t.median(dim=0).values.tolist()
        print(json.dumps({"world": world, "first_batch": fb, "batch": forced, "drop_in": a.drop_in, "host_y": a.host_y, "reps": a.reps,
                          "median_ms": {"wall": med[0], "device": med[1], "prologue": med[2], "gemm_phase": med[3],
                                        "tail": med[4], "gemm_calls_sum": med[5], "batches": int(med[6]),
                                        "host_reaches_first_gemm": med[7], "host_reaches_last_gemm": med[8],
                                        "host_returns": med[9]},
                          "S1_0": float(S["S1"][0]), "ST_conf_0": float(S["ST_conf"][0])}), flush=True)
if _sampler is not None:
    _sampler.__exit__()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
