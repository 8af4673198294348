import sys, json, torch, numpy as np
sys.path.insert(0, "/root/repo")
from salib_b200.analyze import sobol
N, R = 1 << 16, 1184
for D in (64, 96, 128, 256):
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    g = torch.Generator(device="cuda").manual_seed(D)
    Y = torch.randn((N * (2 * D + 2),), dtype=torch.float64, device="cuda", generator=g)
    run = lambda: sobol.analyze_device(p, Y, num_resamples=R, seed=5, resample="philox")
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flop = (2 * D * (D - 1) / 2 + 6 * D + 6) * N * R
    print(json.dumps({"D": D, "N": N, "R": R, "ms": round(ms, 2), "algorithmic_TFLOPs": round(flop / ms / 1e9, 2)}))
