#!/usr/bin/env python
"""Headline benchmark of the Sobol' hot path (BASELINE.json): bootstrapped ``sobol.analyze`` and Saltelli
sample generation on the Sobol' G function, D=64, N=2^20, second order, R=10^4 (configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference ...                     # the unmodified reference on the host cores

One JSON line on stdout (rank 0).  A *step* is one full ``sobol.analyze`` of the resident model output Y
(moments, digit features, multiplicities, exact integer GEMM, estimator algebra, CIs; with several ranks each holds
a base-row range and the partial sums are reduce-scattered).  ``value`` = N*R / step time (resample-rows/s, the unit
BASELINE.md section 3 defines for both arms); ``analyze_s`` is the same number as seconds; ``e2e`` is the drop-in
``analyze(problem, Y_host, ...)`` call with the host-to-device copy of Y and the read-back inside the timed region.
Sample generation is timed in its own loop and reported under ``roofline.sample`` (Grows/s, HBM roofline).
The run exits non-zero if the result is wrong (see ``result_check``).  See DESIGN.md "Measurement".
"""
import argparse
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

G_A8 = [0, 1, 4.5, 9, 99, 99, 99, 99]

# stdout must carry exactly ONE JSON line.  Libraries print there too (NCCL's "NCCL version ..." banner comes
# from C code), so file descriptor 1 is pointed at stderr for the whole run and the result line is written to
# the saved original descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="salib_b200", choices=["salib_b200", "reference"])
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--resamples", type=int, default=10000)
    ap.add_argument("--first-order-only", action="store_true")
    ap.add_argument("--algo", default="auto", help="analyze engine of the measured arm: auto | digits | fp64")
    ap.add_argument("--no-fp64-leg", action="store_true", help="skip the extra FP64-engine measurement")
    ap.add_argument("--shard", default="auto", choices=["auto", "rows", "resamples"],
                    help="multi-GPU decomposition of analyze: base-row ranges (all-reduce of partial sums; what "
                         "the integer engine scales with) or resample ids (all-gather of estimates)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sample", action="store_true", help="skip the 70 GB sample-generation leg")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short legs for BASELINE configs[1] (cfg2), [3] (cfg4 points) and [4] (cfg5, 8 GPUs)")
    ap.add_argument("--ref-log2n", type=int, default=11, help="bounded sample of the reference arm (SURVEY 8(d): "
                                                              "D=64, N=2^11..2^13, R=100)")
    ap.add_argument("--ref-resamples", type=int, default=100)
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="wall-clock budget of the reference arm")
    ap.add_argument("--cpu-log2n", type=int, default=10, help="bounded sample of the in-line cpu_baseline leg")
    return ap.parse_args()


def workload_name(a):
    order = "first+total order" if a.first_order_only else "second order"
    return f"Sobol' G-function D={a.dim}, N=2^{a.log2n}, {order}, num_resamples={a.resamples}"


def g_problem(D):
    return {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}


def g_a(D):
    return np.resize(np.array(G_A8, dtype=np.float64), D)


def measured_peaks():
    """(HBM GB/s, bf16 sustained TFLOP/s, bf16 burst TFLOP/s, source) from the driver-written MEASURED_PEAKS.json,
    else the fallback B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            return (float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]),
                    "measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.sh) on a bounded sample
# ----------------------------------------------------------------------------------------------
def load_reference():
    """-> (analyze module, sample module, Sobol_G module, kind).  kind "reference": SALib itself from oracle/_ref
    (a copy of /root/reference/src/SALib made by oracle/make_ref.sh; it travels to the GPU box with the snapshot);
    kind "port": the oracle's restatement, only if the staged reference is missing."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(ref, "SALib")):
        if ref not in sys.path:
            sys.path.insert(0, ref)
        from SALib.analyze import sobol as ref_analyze
        from SALib.sample import sobol as ref_sample
        from SALib.test_functions import Sobol_G as ref_g
        return ref_analyze, ref_sample, ref_g, "reference"
    return None, None, None, "port"


def reference_inputs(a, log2n):
    """(problem, Y, sample seconds, rows) of the bounded sample, made by the reference's own sampler and model."""
    import warnings

    ref_analyze, ref_sample, ref_g, kind = load_reference()
    D, N = a.dim, 1 << log2n
    c2 = not a.first_order_only
    problem = g_problem(D)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        if kind == "reference":
            X = ref_sample.sample(dict(problem), N, calc_second_order=c2, scramble=False, skip_values=N)
        else:
            from oracle import saltelli as o_saltelli
            X = o_saltelli.sample(problem, N, c2, skip_values=N, closed_form=True)
        t_sample = time.perf_counter() - t0
        if kind == "reference":
            Y = ref_g.evaluate(X, a=g_a(D))
        else:
            from oracle import test_functions as o_tf
            Y = o_tf.sobol_g(X, a=g_a(D))
    return problem, Y, t_sample, X.shape[0]


def reference_analyze_once(a, problem, Y, R, processes):
    """One ``SALib.analyze.sobol.analyze`` call (serial, or its own process pool) -> seconds."""
    import warnings

    ref_analyze, _, _, kind = load_reference()
    c2 = not a.first_order_only
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        if kind == "reference":
            ref_analyze.analyze(dict(problem), Y, calc_second_order=c2, num_resamples=R, seed=100,
                                parallel=processes > 1, n_processors=processes if processes > 1 else None)
        else:
            from oracle import analyze as o_analyze
            if processes > 1:
                o_analyze.analyze_reference_form_parallel(a.dim, Y, c2, R, seed=100, n_processors=processes)
            else:
                o_analyze.analyze_reference_form(a.dim, Y, c2, R, seed=100)
        return time.perf_counter() - t0


def sample_desc(a, log2n, R, processes, kind):
    what = "SALib.analyze.sobol.analyze (unmodified reference, oracle/_ref)" if kind == "reference" \
        else "oracle port of analyze/sobol.py"
    how = f"parallel=True, n_processors={processes}" if processes > 1 else "serial"
    return (f"{what}, {how}, on G-function D={a.dim}, N=2^{log2n}, R={R}, "
            f"{'first+total' if a.first_order_only else 'second'} order; normalised to resample-rows/s (N*R/s); "
            f"the headline itself needs 84 GB for r alone on the CPU (BASELINE.md)")


def run_reference_arm(a):
    """``--impl reference``: the reference's own CPU implementation (process pool on all host cores) on a bounded
    sample of the workload, sized so that warmup + steps fit ``--ref-budget-s``."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = load_reference()[3]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    R = a.ref_resamples
    total = max(1, a.steps) + max(0, a.warmup)
    # size the sample: probe at N = 2^8, time grows ~ N (each of the ~2100 tasks gathers [N, R] arrays)
    problem, Y, _, _ = reference_inputs(a, 8)
    reference_analyze_once(a, problem, Y, R, procs)                      # pool start-up, imports
    t_probe = reference_analyze_once(a, problem, Y, R, procs)
    log2n = a.ref_log2n
    while log2n > 8 and t_probe * (1 << (log2n - 8)) * 1.3 * total > a.ref_budget_s:
        log2n -= 1
    problem, Y, t_sample, rows = reference_inputs(a, log2n)
    N = 1 << log2n
    times = []
    for it in range(total):
        dt = reference_analyze_once(a, problem, Y, R, procs)
        if it >= a.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    val = N * R / sec
    desc = sample_desc(a, log2n, R, procs, kind)
    Nh, Rh = 1 << a.log2n, a.resamples
    line = {
        "metric": "sobol.analyze resample-rows/s (N*R/s)", "value": val, "unit": "resample-rows/s",
        "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": bench_config(a),
        "cpu_baseline": {"value": val, "unit": "resample-rows/s", "cores": procs, "kind": kind, "sample": desc,
                         "seconds_per_step": sec,
                         "headline_seconds_extrapolated": Nh * Rh / val,
                         "extrapolation": "linear in N*R from the bounded sample (labelled estimate; SURVEY section 6)"},
        "sample": {"value": rows / t_sample / 1e9, "unit": "Grows/s", "rows": rows, "seconds": t_sample,
                   "what": f"SALib.sample.sobol.sample(scramble=False) D={a.dim}, N=2^{log2n}, 1 core"
                   if kind == "reference" else "oracle port"},
        "e2e": {"value": val, "unit": "resample-rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cpu_baseline_leg(a):
    """The in-line ``cpu_baseline`` of this repo's arm (rank 0, N=1): the reference serial on one core, ~10-30 s."""
    kind = load_reference()[3]
    R = a.ref_resamples
    problem, Y, t_sample, rows = reference_inputs(a, a.cpu_log2n)
    sec = reference_analyze_once(a, problem, Y, R, 1)
    N = 1 << a.cpu_log2n
    return {"value": N * R / sec, "unit": "resample-rows/s", "cores": 1, "kind": kind,
            "sample": sample_desc(a, a.cpu_log2n, R, 1, kind), "seconds": sec,
            "sample_generation": {"value": rows / t_sample / 1e9, "unit": "Grows/s", "seconds": t_sample,
                                  "what": f"reference sobol.sample(scramble=False), D={a.dim}, N=2^{a.cpu_log2n}"}}


def bench_config(a):
    """The workload both arms are quoted on (identical dict in the two JSON lines)."""
    return {"workload": workload_name(a), "a": "tile([0,1,4.5,9,99,99,99,99])", "scramble": False,
            "skip_values": 1 << a.log2n, "seed": 100,
            "l2": "Y (1.09 GB) and the operands of the GEMM (26 GB) exceed the 126 MB L2; no explicit flush",
            "gc": "Python's cyclic GC is disabled inside timed regions (collected before), as timeit does"}


def profile_traffic(key):
    """DRAM bytes per launch of the dominant kernel for this configuration, from the committed ncu capture
    (profiles/r02_traffic.json; null when that configuration has not been profiled)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        e = t.get(key)
        return (e["bytes"], e["source"]) if e else (None, None)
    except Exception:
        return None, None


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from salib_b200 import _lib, _runtime as rt
    from salib_b200.analyze import sobol as analyze_sobol
    from salib_b200.sample import sobol as sample_sobol
    from salib_b200.test_functions import Sobol_G

    D, N, R = a.dim, 1 << a.log2n, a.resamples
    c2 = not a.first_order_only
    step_rows = 2 * D + 2 if c2 else D + 2
    problem = g_problem(D)
    avec = g_a(D)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        """W untimed calls, then K calls between CUDA events with barrier+sync on both sides; max over ranks."""
        for _ in range(warmup):
            fn()
        gc.collect()
        gc.disable()          # as timeit does: a generation-2 collection is a ~60 ms host stall on any one rank
        try:
            barrier()
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            marks[0].record()
            for i in range(steps):
                fn()
                marks[i + 1].record()
            barrier()
        finally:
            gc.enable()
        # the contract's number is the whole bracket / K; the per-step spread (this rank) shows host stalls
        per_step[:] = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        return max_over_ranks(marks[0].elapsed_time(marks[steps]) / 1e3)

    per_step = []

    import warnings

    warnings.simplefilter("ignore")
    # skip_values = N: the default of saltelli.sample (sample/saltelli.py:112-118); with skip 0 the first
    # base point is the all-zeros corner, whose 2D+2 identical rows dominate var(Y) of this function
    skip = N
    plan, first = sample_sobol._plan(problem, N, c2, False, skip, None)
    lo, hi = rt.shard_range(N, rank, world)
    hbm_peak, bf16_sustained, bf16_burst, peak_src = measured_peaks()

    # ---- leg 1: Saltelli sample generation into HBM (sharded by Sobol' index range, no communication)
    sample = None
    if not a.no_sample:
        X = rt.empty((step_rows * (hi - lo), D), torch.float64)
        L.sa_launch_count(1)
        t = timed(lambda: plan.generate(first + lo, hi - lo, out=X), a.steps, a.warmup)
        launches = L.sa_launch_count(1)
        rows = step_rows * N
        sample_bytes_rank = step_rows * (hi - lo) * D * 8
        gbs = sample_bytes_rank * a.steps / t / 1e9  # per GPU
        tr, tr_src = profile_traffic(f"sample_d{D}_n{a.log2n}_c{int(c2)}_g{world}")
        sample = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                  "traffic": tr, "traffic_source": tr_src, "peak_source": peak_src,
                  "kernel": "saltelli_reg_kernel" if D % 2 == 0 else "saltelli_generic_kernel",
                  "value": rows * a.steps / t / 1e9, "value_unit": "Grows/s (whole job)",
                  "ms_per_pass": t / a.steps * 1e3, "rows": rows, "bytes_per_gpu": sample_bytes_rank,
                  "launches": int(launches)}
        del X
        torch.cuda.empty_cache()

    # ---- model output: fused Sobol' -> G evaluation of this rank's index range
    Ypart = rt.empty((step_rows * (hi - lo),), torch.float64)
    t_eval = timed(lambda: Sobol_G.evaluate_sample(plan, first + lo, hi - lo, a=avec, out=Ypart, check=False),
                   a.steps, a.warmup)
    fused = {"value": step_rows * N * a.steps / t_eval / 1e9, "unit": "Grows/s", "ms_per_pass": t_eval / a.steps * 1e3}
    # multi-GPU analyze: by base-row range for the integer engine (features, multiplicities and GEMM all shrink
    # with the shard; reduce-scatter of [R, SA_NACC] partial sums), by resample id for the FP64 engines.  This is
    # the decision analyze() makes by itself under torchrun (analyze_device(shard="auto")); the device-resident
    # leg calls the route directly because its Y shard is already on the GPU.
    row_mode = world > 1 and (a.shard == "rows" or (a.shard == "auto" and a.algo in ("auto", "digits")))

    _full = []

    def full_y():
        if world == 1:
            return Ypart
        if not _full:
            sizes = [step_rows * (rt.shard_range(N, g, world)[1] - rt.shard_range(N, g, world)[0])
                     for g in range(world)]
            parts = [torch.empty((s,), dtype=torch.float64, device="cuda") for s in sizes]
            dist.all_gather(parts, Ypart)
            _full.append(torch.cat(parts))
        return _full[0]

    # ---- leg 2 (the metric): sobol.analyze on resident Y
    import ctypes

    def measure(algo, steps, warmup, sample_clocks):
        """Times `steps` analyze calls; returns (seconds per step, accumulate-kernel ms per step, resample-rows the
        accumulate launches covered per step, launches in the timed region, last result, clocks)."""
        events = []
        rows_mode = row_mode and algo != "fp64"
        Yres = Ypart if (rows_mode or world == 1) else full_y()

        def analyze_step():
            if rows_mode:
                return analyze_sobol.analyze_row_sharded(problem, [(lo, Yres)], N, calc_second_order=c2,
                                                         num_resamples=R, seed=100, resample="philox",
                                                         kernel_events=events, algo=algo)
            return analyze_sobol.analyze_device(problem, Yres, calc_second_order=c2, num_resamples=R, seed=100,
                                                resample="philox", kernel_events=events, algo=algo,
                                                shard="resamples")

        for _ in range(warmup):
            analyze_step()
        events.clear()
        L.sa_launch_count(1)
        if sample_clocks:
            with ClockSampler(local_rank) as clk:
                t = timed(analyze_step, steps, 0)
        else:
            clk = None
            t = timed(analyze_step, steps, 0)
        n_launch = int(L.sa_launch_count(1))
        timed_events = list(events)
        step_ms[algo] = [round(x, 3) for x in per_step]
        S_ = analyze_step()
        ms = max_over_ranks(sum(e0.elapsed_time(e1) for _, e0, e1 in timed_events)) / steps
        rows = sum(c for c, _, _ in timed_events) * ((hi - lo) if rows_mode else N) / steps
        return t / steps, ms, rows, n_launch, S_, clk

    engine_fallback = None
    step_ms = {}
    try:
        t_step, k_ms, k_rows, launches, S, clocks = measure(a.algo, a.steps, a.warmup, True)
    except _lib.SalibB200Error as exc:
        # an argument-level refusal of the integer engine (not a CUDA fault): measure the FP64 engines and say so
        if a.algo == "fp64" or row_mode:
            raise
        engine_fallback = str(exc)
        a.algo = "fp64"
        t_step, k_ms, k_rows, launches, S, clocks = measure(a.algo, a.steps, a.warmup, True)
    t_an = t_step * a.steps
    engine = getattr(S, "engine", "?")
    nacc = analyze_sobol._nacc(D, c2)
    clk_sum = clocks.summary()

    def fp64_peak_tflops():
        out = rt.zeros((1,), torch.float64)
        nthreads = ctypes.c_uint64(0)
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("sa_microbench_fp64", rt.stream_ptr(), 40000, rt.ptr(out), ctypes.addressof(nthreads))
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 32.0 * 40000 * nthreads.value / (e0.elapsed_time(e1) / 1e3) / 1e12)
        return best

    def i8_peak_tops():
        """The digits GEMM's own MMA mix with the operand ring filled once: burst (best short run) and sustained
        (one ~0.1 s run, power-limited clocks like the real kernel)."""
        pairs = torch.cuda.get_device_properties(local_rank).multi_processor_count // 2
        w = torch.ones((pairs * 256, 1024), dtype=torch.uint8, device="cuda")
        q = torch.ones((512, 1024), dtype=torch.int8, device="cuda")
        c = torch.empty((pairs * 256, 512), dtype=torch.int32, device="cuda")
        ops = ctypes.c_double(0.0)

        def run(ksteps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("sa_microbench_i8_tensor", rt.stream_ptr(), rt.ptr(w), rt.ptr(q), rt.ptr(c), ksteps,
                      ctypes.addressof(ops))
            e1.record()
            torch.cuda.synchronize()
            return ops.value / (e0.elapsed_time(e1) / 1e3) / 1e12

        run(2000)
        burst = max(run(4000) for _ in range(3))
        sustained = run(130000)
        return burst, sustained

    def fp64_roofline(ms, rows, t_step_):
        if c2:
            flop_per_row = 2 * (D * (D - 1) // 2) + 6 * D + 6
            kernel = "s2_accum_smem_kernel (+ scalar_sums_kernel)" if D > 48 else "fo_gemv_kernel (pair features)"
        else:
            flop_per_row = 6 * D + 6
            kernel = "fo_gemv_kernel"
        peak = fp64_peak_tflops()
        achieved = flop_per_row * rows / (ms / 1e3) / 1e12 if ms > 0 else None
        tr, tr_src = profile_traffic(f"fp64_d{D}_n{a.log2n}_c{int(c2)}")
        return {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": tr, "traffic_source": tr_src,
                "kernel": kernel, "kernel_ms_per_step": ms, "kernel_share_of_step": (ms / 1e3) / t_step_,
                "algorithmic_flop_per_resample_row": flop_per_row,
                "peak_source": "measured here: sa_microbench_fp64 (16 independent DFMA chains/thread); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "note": "FP64-FMA-pipe bound; rows with multiplicity 0 are skipped, so achieved counts "
                        "algorithmic flop of all N*R resample-rows"}

    if engine == "digits":
        S_dig = analyze_sobol.digits_for((hi - lo) if row_mode else N)
        ops_per_row = 2 * nacc * S_dig        # int8 multiply-adds x 2 per (resample, row)
        burst, sustained = i8_peak_tops()
        rows_here = (hi - lo) if row_mode else N
        n_batches = len(analyze_sobol.digits_batches(R, int(L.sa_pad_rows(rows_here)), nacc * S_dig))
        achieved = ops_per_row * k_rows / (k_ms / 1e3) / 1e12 if k_ms > 0 else None
        tr, tr_src = profile_traffic(f"digits_d{D}_n{a.log2n}_r{R}_c{int(c2)}_g{world}")
        # MEASURED_PEAKS.json has no int8 entry: the int8 tensor rate is twice the bf16 rate on this part (8192
        # vs 4096 MAC/clk/SM), so the denominator is 2 x the driver's SUSTAINED bf16 figure (the kernel runs
        # inside a long, power-capped step)
        peak = 2.0 * bf16_sustained
        sm_mhz = clk_sum.get("sm_mhz") or 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "ops": "int8 tensor operations (u8 x s8 -> s32 multiply-add = 2), counted like flop",
                    "frac": (achieved / peak) if achieved else None,
                    "peak_source": f"2 x bf16_tflops_sustained ({bf16_sustained}) of {peak_src}: int8 runs at twice "
                                   "the bf16 MAC rate and the file has no int8 entry",
                    "peak_2x_bf16_burst": 2.0 * bf16_burst,
                    "frac_of_2x_bf16_burst": (achieved / (2.0 * bf16_burst)) if achieved else None,
                    "peak_no_traffic_microbench": sustained, "peak_no_traffic_microbench_burst": burst,
                    "frac_of_no_traffic_microbench": (achieved / sustained) if achieved else None,
                    "peak_at_measured_clock": 2 * 8192 * 148 * sm_mhz * 1e6 / 1e12 if sm_mhz else None,
                    "frac_of_peak_at_measured_clock": (achieved / (2 * 8192 * 148 * sm_mhz * 1e6 / 1e12))
                    if (achieved and sm_mhz) else None,
                    "traffic": tr, "traffic_source": tr_src,
                    "algorithmic_bytes": ((R + 1) + nacc * S_dig) * ((hi - lo) if row_mode else N)
                    + 4 * (R + 1) * nacc * S_dig,
                    "operand_bytes_as_batched": ((R + 1) + n_batches * nacc * S_dig) * ((hi - lo) if row_mode else N)
                    + 4 * (R + 1) * nacc * S_dig,
                    "gemm_launches_per_step": n_batches,
                    "kernel": "digit_gemm_pair_persistent_kernel<2,4> (tcgen05.mma.cta_group::2.kind::i8)",
                    "kernel_ms_per_step": k_ms, "kernel_share_of_step": (k_ms / 1e3) / t_step,
                    "algorithmic_int8_ops_per_resample_row": ops_per_row, "digits": S_dig,
                    "note": "exact integer GEMM of the multiplicities with the int8 digits of the features; the work "
                            "moved from the FP64 pipe (4422 flop per resample-row, SURVEY 8(d)) to the int8 tensor "
                            "pipe at 6.8x the raw operation count; ncu: tensor pipe 98 % active at the clock the "
                            "1000 W cap allows (profiles/r02_step_ncu_full_summary.csv); achieved, kernel_ms and "
                            "traffic are sums over the gemm_launches_per_step launches of one step"}
    else:
        roofline = fp64_roofline(k_ms, k_rows, t_step)
    roofline["sample"] = sample

    checks = {}
    fp64_leg = None
    if engine == "digits" and not a.no_fp64_leg:
        t2, k2, rows2, launches2, S2, _ = measure("fp64", 1, 1, False)
        diffs = {k: float(np.nanmax(np.abs(np.asarray(S[k]) - np.asarray(S2[k])))) for k in S if k in S2}
        fp64_leg = {"analyze_s": t2, "value": N * R / t2, "unit": "resample-rows/s", "engine": S2.engine,
                    "gpu_launches": launches2, "roofline": fp64_roofline(k2, rows2, t2),
                    "max_abs_diff_vs_default": diffs,
                    "note": "the north-star design (no tensor cores) measured in the same run: 1 step after 1 "
                            "warm-up, same Philox multiplicities"}
        checks["max_abs_diff_digits_vs_fp64"] = max(diffs.values())
        checks["digits_vs_fp64_ok"] = bool(max(diffs.values()) <= 1e-12)

    # analytic indices of the G function (alpha = 1, delta = 0): V_i = 1 / (3 (1 + a_i)^2)
    Vi = 1.0 / (3.0 * (1.0 + avec) ** 2)
    var_total = float(np.prod(1.0 + Vi) - 1.0)
    analytic = (float(Vi[0] / var_total), float(Vi[0] * np.prod(1.0 + Vi[1:]) / var_total))
    # (i) S1/ST of x1 within 4 sigma of the analytic values (conf = 1.96 sigma)
    sig1, sigT = float(S["S1_conf"][0]) / 1.959964, float(S["ST_conf"][0]) / 1.959964
    checks.update({"S1_0": float(S["S1"][0]), "ST_0": float(S["ST"][0]), "S1_conf_0": float(S["S1_conf"][0]),
                   "ST_conf_0": float(S["ST_conf"][0]), "analytic_S1_0": analytic[0], "analytic_ST_0": analytic[1],
                   "analytic_ok": bool(abs(S["S1"][0] - analytic[0]) <= 4 * sig1
                                       and abs(S["ST"][0] - analytic[1]) <= 4 * sigT
                                       and np.all(np.isfinite(S["S1"])) and np.all(np.isfinite(S["ST_conf"])))})
    # (iii) several ranks: the point estimates must equal a single-rank recomputation on rank 0
    if world > 1:
        Yall = full_y()
        if rank == 0:
            S0 = analyze_sobol.analyze_device(problem, Yall, calc_second_order=c2, num_resamples=0, algo=a.algo,
                                              distributed=False)
            keys = ["S1", "ST"] + (["S2"] if c2 else [])
            d = max(float(np.nanmax(np.abs(np.asarray(S[k]) - np.asarray(S0[k])))) for k in keys)
            checks["max_abs_diff_point_vs_single_rank"] = d
            checks["multi_rank_ok"] = bool(d <= 1e-12)
        barrier()

    # ---- the other BASELINE configs, briefly, so that they are in the driver-parsed line (under roofline.configs)
    other = None
    if not a.no_other_configs and a.dim == 64 and a.log2n == 20 and c2:
        other = {}
        if world == 1:
            # configs[1]: G function D=8, N=2^20, first+total order, R=1000
            p8 = g_problem(8)
            plan8, first8 = sample_sobol._plan(p8, N, False, False, N, None)
            Y8 = Sobol_G.evaluate_sample(plan8, first8, N)
            t8 = timed(lambda: analyze_sobol.analyze_device(p8, Y8, calc_second_order=False, num_resamples=1000,
                                                            seed=100, resample="philox"), 9, 3)
            t8 = 3 * float(np.median(per_step)) / 1e3      # a 2 ms call: the median of 9 (a host hiccup is 5-50 ms)
            S8 = analyze_sobol.analyze_device(p8, Y8, calc_second_order=False, num_resamples=1000, seed=100,
                                              resample="philox")
            V8 = 1.0 / (3.0 * (1.0 + np.array(G_A8, dtype=np.float64)) ** 2)
            an8 = V8 / (np.prod(1.0 + V8) - 1.0)
            other["cfg2"] = {"workload": "Sobol' G-function D=8, N=2^20, first+total order, num_resamples=1000",
                             "analyze_s": t8 / 3, "value": N * 1000 * 3 / t8, "unit": "resample-rows/s",
                             "engine": S8.engine, "timing": "median of 9 calls after 3 warm-up calls (CUDA events)",
                             "S1_within_CI_of_analytic": bool(np.all(np.abs(S8["S1"] - an8) <= 2.0 * S8["S1_conf"] + 1e-3))}
            del Y8
            # configs[3]: points of the sample-generation sweep (chunked into a reused buffer beyond 64 GB)
            pts = []
            for Dq, lq, scr in ((16, 22, False), (128, 20, False), (128, 20, True), (1024, 16, False)):
                stepq, Nq = 2 * Dq + 2, 1 << lq
                chunk = Nq
                while stepq * chunk * Dq * 8 > 6.4e10:
                    chunk //= 2
                svq, shq = sample_sobol._scramble_state(2 * Dq, scr, 1234)
                planq = sample_sobol.SamplePlan(g_problem(Dq), True, svq, shq)
                buf = rt.empty((stepq * chunk, Dq), torch.float64)

                def run_q():
                    for c0 in range(0, Nq, chunk):
                        planq.generate(c0, chunk, out=buf)

                tq = timed(run_q, 2, 1) / 2
                nbytes = stepq * Nq * Dq * 8
                pts.append({"D": Dq, "log2N": lq, "scramble": scr, "Grows_per_s": stepq * Nq / tq / 1e9,
                            "GBps": nbytes / tq / 1e9, "hbm_frac": nbytes / tq / 1e9 / hbm_peak, "seconds": tq})
                del buf
                torch.cuda.empty_cache()
            other["cfg4_points"] = pts
        if world == 8:
            # configs[4]: D=64, N=2^22, second order, fused G evaluation per rank, row shards (Y never gathered)
            N5 = 1 << 22
            plan5, first5 = sample_sobol._plan(problem, N5, c2, False, N5, None)
            lo5, hi5 = rt.shard_range(N5, rank, world)
            Y5 = rt.empty((step_rows * (hi5 - lo5),), torch.float64)
            Sobol_G.evaluate_sample(plan5, first5 + lo5, hi5 - lo5, a=avec, out=Y5, check=False)

            def step5():
                return analyze_sobol.analyze_row_sharded(problem, [(lo5, Y5)], N5, calc_second_order=c2,
                                                         num_resamples=R, seed=100, resample="philox", algo=a.algo)

            t5 = timed(step5, 2, 1) / 2
            S5 = step5()
            other["cfg5"] = {"workload": f"Sobol' G-function D=64, N=2^22, second order, num_resamples={R}, 8 GPUs, "
                                         "fused evaluation, base-row shards",
                             "analyze_s": t5, "value": N5 * R / t5, "unit": "resample-rows/s", "engine": S5.engine,
                             "S1_0": float(S5["S1"][0]), "ST_0": float(S5["ST"][0])}
            del Y5
        roofline["configs"] = other

    # ---- e2e: the public drop-in call with a HOST (pinned) Y, H2D + D2H inside the timed region.  Every rank
    # holds the host array, as every process of a torchrun job would; analyze() routes by itself.
    Yfull = full_y()
    numa_cpus = rt.bind_host_to_gpu() if world > 1 else None   # staging buffer on the GPU's own NUMA node
    Yh = torch.empty(Yfull.shape, dtype=torch.float64, pin_memory=True)
    Yh.copy_(Yfull)
    torch.cuda.synchronize()
    Yh_np = Yh.numpy()
    _full.clear()
    del Yfull

    def e2e_step():
        if a.algo == "auto":
            return analyze_sobol.analyze(problem, Yh_np, calc_second_order=c2, num_resamples=R, seed=100)
        return analyze_sobol.analyze_device(problem, Yh_np, calc_second_order=c2, num_resamples=R, seed=100,
                                            algo=a.algo, shard="rows" if row_mode else "resamples")

    t_e2e = timed(e2e_step, a.steps, 1)
    S_e2e = e2e_step()
    checks["e2e_route"] = {"engine": getattr(S_e2e, "engine", "?"),
                           "resample_mode": getattr(S_e2e, "resample_mode", "?"),
                           "max_abs_diff_vs_device_leg": max(
                               float(np.nanmax(np.abs(np.asarray(S[k]) - np.asarray(S_e2e[k])))) for k in S)}
    checks["e2e_ok"] = bool(checks["e2e_route"]["max_abs_diff_vs_device_leg"] <= 1e-12)
    nest = analyze_sobol._nest(D, c2)
    h2d = int(Yh.numel() * 8) if not (world > 1 and getattr(S_e2e, "engine", "") == "digits") \
        else int(step_rows * N * 8)      # row shards: the ranks together upload Y once
    e2e = {"value": N * R * a.steps / t_e2e, "unit": "resample-rows/s", "analyze_s": t_e2e / a.steps,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int((2 * nest + 8) * 8 + 4),
           "call": "salib_b200.analyze.sobol.analyze(problem, Y_host, calc_second_order, num_resamples, seed=100)",
           "host_numa_bound": bool(numa_cpus)}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline_leg(a)

    ok = all(v for k, v in checks.items() if k.endswith("_ok"))
    if rank == 0:
        line = {
            "metric": "sobol.analyze resample-rows/s (N*R/s)", "value": N * R * a.steps / t_an,
            "unit": "resample-rows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t_an / a.steps * 1e3, "analyze_s": t_an / a.steps, "higher_is_better": True,
            "ms_of_each_step_rank0": step_ms.get(a.algo),
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (sums: exact u8 x s8 -> s32 digit GEMM)" if engine == "digits" else "f64",
            "data": "synthetic", "engine": engine, "engine_fallback": engine_fallback,
            "config": bench_config(a),
            "parallelism": (f"base-row range x{world} (reduce-scatter of partial sums)" if row_mode
                            else f"resample-id x{world}"),
            "resample": "philox seed 100",
            "fused_sample_eval": fused, "roofline": roofline, "fp64_engine": fp64_leg,
            "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": launches, "clocks": clk_sum,
            "result_check": checks, "result_ok": ok,
        }
        line["roofline"]["result_check"] = {k: v for k, v in checks.items() if not isinstance(v, dict)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.stderr.write(f"bench.py: RESULT CHECK FAILED: {checks}\n")
        sys.exit(1)


if __name__ == "__main__":
    main()
