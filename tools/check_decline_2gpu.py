"""Mixed engine decisions under torchrun: rank 1 is made to believe the integer engine does not pay, so it declines
(`_RankRecords.decline`) while rank 0 enters the row route optimistically; both must fall back to resample shards
without a mismatched collective and still reproduce the frozen reference.

    torchrun --nproc-per-node 2 tools/check_decline_2gpu.py

Written at the end of round 2 after the GPU budget was spent: NOT yet run on a GPU box (the uniform-decision paths
are covered by tests/test_gpu_r02.py::test_two_rank_drop_in_matches_the_frozen_reference).  Run it before relying on
mixed decisions; it is deliberately not part of the pytest suite.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

from salib_b200.analyze import sobol  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "analyze_g16_n256_r32.npz"))
p = {"num_vars": 16, "names": [f"x{i + 1}" for i in range(16)], "bounds": [[0.0, 1.0]] * 16}
sobol.DIGITS_MIN_WORK = 0
pays = sobol.SobolAnalyzer._digits_pays
sobol.SobolAnalyzer._digits_pays = lambda self, resamples: rank == 0 and pays(self, resamples)
for seed in (int(z["seed"]), None):                      # seeded (reference indices) and seedless (broadcast key)
    S = sobol.analyze(p, z["Y"], calc_second_order=True, num_resamples=int(z["R"]), conf_level=float(z["conf"]),
                      keep_resamples=True, seed=seed)
    assert S.engine != "digits", S.engine                # nobody may take the integer engine
    if seed:
        iu = np.triu_indices(16, 1)
        for k in ("S1", "S1_conf", "ST", "ST_conf"):
            np.testing.assert_allclose(S[k], z[k], rtol=1e-10, atol=1e-13, err_msg=k)
        np.testing.assert_allclose(S["S2"][iu], z["S2"][iu], rtol=1e-10, atol=1e-13)
dist.barrier()
dist.destroy_process_group()
print("DECLINE_OK", rank)
