#!/usr/bin/env python
"""Headline benchmark of the Sobol' hot path (BASELINE.json): bootstrapped ``sobol.analyze`` and Saltelli
sample generation on the Sobol' G function, D=64, N=2^20, second order, R=10^4 (configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference ...                     # reference CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  A *step* is one full ``sobol.analyze`` of the resident model output Y
(moments, normalise, point estimates, R bootstrap resamples sharded over the ranks, all-gather, CIs).
``value`` = N*R / step time (resample-rows/s, the unit BASELINE.md section 3 defines for both arms);
``analyze_s`` is the same number as seconds.  Sample generation is timed in its own loop and reported under
``sample`` (Grows/s) with its HBM roofline.  See DESIGN.md "Measurement".
"""
import argparse
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
    ap.add_argument("--ref-log2n", type=int, default=10, help="bounded sample of the reference arm")
    ap.add_argument("--ref-resamples", type=int, default=48)
    return ap.parse_args()


def workload_name(a):
    order = "first+total order" if a.first_order_only else "second order"
    return f"Sobol' G-function D={a.dim}, N=2^{a.log2n}, {order}, num_resamples={a.resamples}"


def g_problem(D):
    return {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}


def g_a(D):
    return np.resize(np.array(G_A8, dtype=np.float64), D)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's reference-form analyze on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_reference_run(a, steps, warmup, processes):
    """Times oracle.analyze.analyze_reference_form(_parallel) (a NumPy restatement of
    analyze/sobol.py:147-194, same gathers and reductions) on D=a.dim, N=2^ref_log2n, R=ref_resamples.
    Returns (resample_rows_per_s, seconds_per_step, description)."""
    from oracle import analyze as o_analyze
    from oracle import saltelli as o_saltelli
    from oracle import test_functions as o_tf

    D, N, R = a.dim, 1 << a.ref_log2n, a.ref_resamples
    c2 = not a.first_order_only
    X = o_saltelli.sample(g_problem(D), N, c2, skip_values=N, closed_form=True)
    Y = o_tf.sobol_g(X, a=g_a(D))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if processes > 1:
            o_analyze.analyze_reference_form_parallel(D, Y, c2, R, seed=100, n_processors=processes)
        else:
            o_analyze.analyze_reference_form(D, Y, c2, R, seed=100)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    desc = (f"oracle port of analyze/sobol.py on G-function D={D}, N=2^{a.ref_log2n}, R={R}, "
            f"{'second' if c2 else 'first+total'} order, {processes} process(es); "
            f"the headline itself needs 84 GB for r alone on the CPU (BASELINE.md)")
    return N * R / sec, sec, desc


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    # keep the whole run within a few minutes: warm-up is capped at 1 for the CPU arm
    val, sec, desc = cpu_reference_run(a, max(1, a.steps), min(a.warmup, 1), procs)
    line = {
        "metric": "sobol.analyze resample-rows/s (N*R/s)", "value": val, "unit": "resample-rows/s",
        "impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": desc},
        "cpu_baseline": {"value": val, "unit": "resample-rows/s", "cores": procs, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "resample-rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


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
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / 1e3)

    import warnings

    warnings.simplefilter("ignore")
    # skip_values = N: the default of saltelli.sample (sample/saltelli.py:112-118); with skip 0 the first
    # base point is the all-zeros corner, whose 2D+2 identical rows dominate var(Y) of this function
    skip = N
    plan, first = sample_sobol._plan(problem, N, c2, False, skip, None)
    lo, hi = rt.shard_range(N, rank, world)
    hbm_peak, hbm_src = measured_peaks()

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
        sample = {"value": rows * a.steps / t / 1e9, "unit": "Grows/s", "ms_per_pass": t / a.steps * 1e3,
                  "rows": rows, "bytes_per_gpu": sample_bytes_rank, "launches": int(launches),
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                               "frac": gbs / hbm_peak,
                               # ncu (profiles/r01_final_ncu_full_summary.csv): dram write 69.736 GB, read 1 MB
                               "traffic": 69.737e9 if (D == 64 and a.log2n == 20 and c2 and world == 1) else None,
                               "peak_source": hbm_src,
                               "kernel": "saltelli_reg_kernel" if D % 2 == 0 else "saltelli_generic_kernel"}}
        del X
        torch.cuda.empty_cache()

    # ---- model output: fused Sobol' -> G evaluation of this rank's index range, all-gathered
    Ypart = rt.empty((step_rows * (hi - lo),), torch.float64)
    t_eval = timed(lambda: Sobol_G.evaluate_sample(plan, first + lo, hi - lo, a=avec, out=Ypart, check=False),
                   a.steps, a.warmup)
    # multi-GPU analyze: by base-row range for the integer engine (features, multiplicities and GEMM all shrink
    # with the shard; one all-reduce of [R, SA_NACC] partial sums), by resample id for the FP64 engines
    row_mode = world > 1 and (a.shard == "rows" or (a.shard == "auto" and a.algo in ("auto", "digits")))
    if world > 1 and not row_mode:
        sizes = [step_rows * (rt.shard_range(N, g, world)[1] - rt.shard_range(N, g, world)[0]) for g in range(world)]
        parts = [torch.empty((s,), dtype=torch.float64, device="cuda") for s in sizes]
        dist.all_gather(parts, Ypart)
        Yd = torch.cat(parts)
    else:
        Yd = Ypart
    fused = {"value": step_rows * N * a.steps / t_eval / 1e9, "unit": "Grows/s", "ms_per_pass": t_eval / a.steps * 1e3}

    # ---- leg 2 (the metric): sobol.analyze on resident Y
    import ctypes

    def measure(algo, steps, warmup, sample_clocks):
        """Times `steps` analyze calls; returns (seconds per step, accumulate-kernel ms per step, resample-rows the
        accumulate launches covered per step, launches in the timed region, last result, clocks)."""
        events = []

        rows_mode = row_mode and algo != "fp64"

        def analyze_step():
            if rows_mode:
                return analyze_sobol.analyze_row_sharded(problem, [(lo, Yd)], N, calc_second_order=c2,
                                                         num_resamples=R, seed=100, resample="philox",
                                                         kernel_events=events, algo=algo)
            if row_mode:    # FP64 leg of a row-sharded run: every rank needs all of Y
                return analyze_sobol.analyze_device(problem, full_y(), calc_second_order=c2, num_resamples=R,
                                                    seed=100, resample="philox", kernel_events=events, algo=algo)
            return analyze_sobol.analyze_device(problem, Yd, calc_second_order=c2, num_resamples=R, seed=100,
                                                resample="philox", kernel_events=events, algo=algo)

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
        S_ = analyze_step()
        ms = max_over_ranks(sum(e0.elapsed_time(e1) for _, e0, e1 in timed_events)) / steps
        rows = sum(c for c, _, _ in timed_events) * ((hi - lo) if rows_mode else N) / steps
        return t / steps, ms, rows, n_launch, S_, clk

    _full = []

    def full_y():
        if not _full:
            sizes = [step_rows * (rt.shard_range(N, g, world)[1] - rt.shard_range(N, g, world)[0])
                     for g in range(world)]
            parts = [torch.empty((s,), dtype=torch.float64, device="cuda") for s in sizes]
            dist.all_gather(parts, Ypart)
            _full.append(torch.cat(parts))
        return _full[0]

    engine_fallback = None
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
        # DRAM bytes of one tile-kernel launch from the committed ncu capture (1184 resamples = one wave at
        # N = 2^20: Y in analysis layout 1.36 GB + multiplicities 1.24 GB, each read once)
        traffic = 2.635e9 if (c2 and D == 64 and a.log2n == 20) else None
        return {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "traffic_source": "profiles/r01_final_ncu_full_summary.csv (dram read+write per 1184-resample "
                                  "launch)" if traffic else None,
                "kernel": kernel, "kernel_ms_per_step": ms, "kernel_share_of_step": (ms / 1e3) / t_step_,
                "algorithmic_flop_per_resample_row": flop_per_row,
                "peak_source": "measured here: sa_microbench_fp64 (16 independent DFMA chains/thread)",
                "note": "FP64-FMA-pipe bound; rows with multiplicity 0 are skipped, so achieved counts "
                        "algorithmic flop of all N*R resample-rows"}

    if engine == "digits":
        S_dig = analyze_sobol.digits_for((hi - lo) if row_mode else N)
        ops_per_row = 2 * nacc * S_dig        # int8 multiply-adds x 2 per (resample, row)
        burst, sustained = i8_peak_tops()
        achieved = ops_per_row * k_rows / (k_ms / 1e3) / 1e12 if k_ms > 0 else None
        traffic = 436.6e9 if (c2 and D == 64 and a.log2n == 20 and R == 10000 and world == 1) else None
        roofline = {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                    "ops": "int8 tensor operations (u8 x s8 -> s32 multiply-add = 2), counted like flop",
                    "frac": (achieved / sustained) if achieved else None, "peak_burst": burst,
                    "traffic": traffic,
                    "traffic_source": "profiles/r01_digits_pair_gemm_ncu_summary.csv (dram read+write of the "
                                      "10^4-resample launch)" if traffic else None,
                    "kernel": "digit_gemm_pair_kernel<2,4> (tcgen05.mma.cta_group::2.kind::i8)",
                    "kernel_ms_per_step": k_ms, "kernel_share_of_step": (k_ms / 1e3) / t_step,
                    "algorithmic_int8_ops_per_resample_row": ops_per_row, "digits": S_dig,
                    "equivalent_fp64_tflops": (2 * (D * (D - 1) // 2) + 6 * D + 6 if c2 else 6 * D + 6) * k_rows
                    / (k_ms / 1e3) / 1e12 if k_ms > 0 else None,
                    "peak_source": "measured here: sa_microbench_i8_tensor (the kernel's own MMA mix, operand ring "
                                   "filled once, ~0.1 s run at power-limited clocks); MEASURED_PEAKS.json has no "
                                   "int8 figure (2 x its bf16 numbers: 3282 burst / 2771 sustained)",
                    "note": "exact integer GEMM of the multiplicities with the int8 digits of the features; "
                            "tensor pipe 94 % active at the 1.36 GHz the 1000 W cap allows (ncu)"}
    else:
        roofline = fp64_roofline(k_ms, k_rows, t_step)

    fp64_leg = None
    if engine == "digits" and not a.no_fp64_leg:
        t2, k2, rows2, launches2, S2, _ = measure("fp64", 1, 1, False)
        fp64_leg = {"analyze_s": t2, "value": N * R / t2, "unit": "resample-rows/s", "engine": S2.engine,
                    "gpu_launches": launches2, "roofline": fp64_roofline(k2, rows2, t2),
                    "max_abs_diff_vs_default": {k: float(np.nanmax(np.abs(np.asarray(S[k]) - np.asarray(S2[k]))))
                                                for k in S if k in S2},
                    "note": "the north-star design (no tensor cores) measured in the same run: 1 step after 1 "
                            "warm-up, same Philox multiplicities"}

    # ---- e2e: the public drop-in call with a HOST (pinned) Y, H2D + D2H inside the timed region
    Yh = torch.empty(Yd.shape, dtype=torch.float64, pin_memory=True)
    Yh.copy_(Yd)
    torch.cuda.synchronize()

    def e2e_step():
        if row_mode:   # every rank uploads only its own base-row range
            return analyze_sobol.analyze_row_sharded(problem, [(lo, Yh)], N, calc_second_order=c2, num_resamples=R,
                                                     seed=100, resample="philox", algo=a.algo)
        return analyze_sobol.analyze_device(problem, Yh, calc_second_order=c2, num_resamples=R, seed=100,
                                            resample="philox", algo=a.algo)

    t_e2e = timed(e2e_step, a.steps, 1)
    nest = analyze_sobol._nest(D, c2)
    e2e = {"value": N * R * a.steps / t_e2e, "unit": "resample-rows/s", "analyze_s": t_e2e / a.steps,
           "h2d_bytes_per_step": int(Yh.numel() * 8) * (world if row_mode else 1), "d2h_bytes_per_step": int((2 * nest + 8) * 8 + 4)}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        val, sec, desc = cpu_reference_run(a, 1, 0, 1)
        cpu = {"value": val, "unit": "resample-rows/s", "cores": 1, "kind": "port", "sample": desc,
               "seconds": sec}

    # analytic indices of the G function (alpha = 1, delta = 0): V_i = 1 / (3 (1 + a_i)^2)
    Vi = 1.0 / (3.0 * (1.0 + avec) ** 2)
    var_total = float(np.prod(1.0 + Vi) - 1.0)
    analytic = (float(Vi[0] / var_total), float(Vi[0] * np.prod(1.0 + Vi[1:]) / var_total))

    if rank == 0:
        line = {
            "metric": "sobol.analyze resample-rows/s (N*R/s)", "value": N * R * a.steps / t_an,
            "unit": "resample-rows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t_an / a.steps * 1e3, "analyze_s": t_an / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (sums: exact u8 x s8 -> s32 digit GEMM)" if engine == "digits" else "f64",
            "data": "synthetic", "engine": engine, "engine_fallback": engine_fallback,
            "config": {"workload": workload_name(a), "a": "tile([0,1,4.5,9,99,99,99,99])", "scramble": False,
                       "skip_values": skip, "resample": "philox seed 100", "parallelism": (f"base-row range x{world} (all-reduce of partial sums)" if row_mode
                                       else f"resample-id x{world}"),
                       "l2": f"Y is {Yd.numel() * 8 / 1e6:.0f} MB (> 126 MB L2), no explicit flush"
                       if Yd.numel() * 8 > 126e6 else "Y fits L2 (resident by design)"},
            "sample": sample, "fused_sample_eval": fused, "roofline": roofline, "fp64_engine": fp64_leg,
            "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
            "result_check": {"S1_0": float(S["S1"][0]), "ST_0": float(S["ST"][0]), "S1_conf_0": float(S["S1_conf"][0]),
                             "ST_conf_0": float(S["ST_conf"][0]), "analytic_S1_0": analytic[0], "analytic_ST_0": analytic[1]},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
