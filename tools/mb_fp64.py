import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
from salib_b200 import _lib, _runtime as rt
out = rt.zeros((1,), torch.float64); n = ctypes.c_uint64(0)
for name, fl in (("sa_microbench_fp64", 32.0), ("sa_microbench_fp64_outer", 128.0)):
    best = 0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.call(name, rt.stream_ptr(), 20000, rt.ptr(out), ctypes.addressof(n)); e1.record(); torch.cuda.synchronize()
        best = max(best, fl * 20000 * n.value / (e0.elapsed_time(e1) / 1e3) / 1e12)
    print(name, round(best, 2), "TFLOP/s")
