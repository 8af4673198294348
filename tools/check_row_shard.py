"""2+ GPU check of the row-sharded analyze path (torchrun --nproc-per-node G tools/check_row_shard.py).

Every rank evaluates its own base-row range of a G-function design on the device (fused sample + model,
X never materialised), keeps only that shard of Y, and calls analyze_row_sharded; rank 0 also runs the
resample-sharded path on the all-gathered Y and prints the largest difference between the two results."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from salib_b200 import _runtime as rt  # noqa: E402
from salib_b200.analyze import sobol  # noqa: E402
from salib_b200.sample import sobol as sampler  # noqa: E402
from salib_b200.test_functions import Sobol_G  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
D, N, R = 64, int(os.environ.get("LOG2N", 16)), int(os.environ.get("R", 2368))
N = 1 << N
p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
a = np.tile([0, 1, 4.5, 9, 99, 99, 99, 99], D // 8).astype(np.float64)
lo, hi = rt.shard_range(N, rank, world)
plan, first = sampler._plan(p, N, True, False, 0, None)
Yl = Sobol_G.evaluate_sample(plan, first + lo, hi - lo, a=a)
sobol.analyze_row_sharded(p, [(lo, Yl)], N, num_resamples=64, seed=100, resample="philox")   # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
S = sobol.analyze_row_sharded(p, [(lo, Yl)], N, num_resamples=R, seed=100, resample="philox")
torch.cuda.synchronize()
t_rows = time.perf_counter() - t0
# resample-sharded path on the full Y for comparison
if world > 1:
    per = -(-N // world) * (2 * D + 2)
    pad = torch.zeros((per,), dtype=torch.float64, device=Yl.device)
    pad[: Yl.numel()] = Yl
    full = torch.empty((world * per,), dtype=torch.float64, device=Yl.device)
    dist.all_gather_into_tensor(full, pad)
    Y = torch.cat([full[g * per: g * per + (rt.shard_range(N, g, world)[1] - rt.shard_range(N, g, world)[0]) * (2 * D + 2)]
                   for g in range(world)])
else:
    Y = Yl
sobol.analyze_device(p, Y, num_resamples=64, seed=100, resample="philox")   # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
T = sobol.analyze_device(p, Y, num_resamples=R, seed=100, resample="philox")
torch.cuda.synchronize()
t_res = time.perf_counter() - t0
if rank == 0:
    iu = np.triu_indices(D, 1)
    diff = max(float(np.max(np.abs(S[k] - T[k]))) for k in ("S1", "ST", "S1_conf", "ST_conf"))
    diff = max(diff, float(np.max(np.abs(S["S2"][iu] - T["S2"][iu]))), float(np.max(np.abs(S["S2_conf"][iu] - T["S2_conf"][iu]))))
    print(json.dumps({"world": world, "N": N, "R": R, "max_abs_diff_rows_vs_resamples": diff,
                      "seconds_row_sharded": t_rows, "seconds_resample_sharded": t_res}))
if world > 1:
    dist.destroy_process_group()
