import sys, json, torch, numpy as np
sys.path.insert(0, '/root/repo')
from salib_b200 import _runtime as rt
from salib_b200.sample import sobol as ss
def gp(D): return {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
for D, rows in ((1024, 1024), (1024, 2048), (1024, 4096), (512, 4096), (256, 8192), (192, 8192), (64, 1<<16)):
    step = 2*D+2
    sv, sh = ss._scramble_state(2*D, False, 1234)
    plan = ss.SamplePlan(gp(D), True, sv, sh)
    buf = rt.empty((step*rows, D), torch.float64)
    for _ in range(3): plan.generate(0, rows, out=buf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): plan.generate(i*rows, rows, out=buf)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)/5/1e3
    print(json.dumps({"D": D, "rows": rows, "GB": step*rows*D*8/1e9, "GBps": step*rows*D*8/t/1e9}), flush=True)
    del buf
