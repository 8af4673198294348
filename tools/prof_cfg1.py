import cProfile, pstats, sys, time, warnings
sys.path.insert(0, "/root/repo")
import numpy as np
from salib_b200.sample import saltelli
from salib_b200.analyze import sobol
from salib_b200.test_functions import Ishigami
warnings.simplefilter("ignore")
p = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}
X = saltelli.sample(p, 1024)
Y = Ishigami.evaluate(X)
for _ in range(5):
    sobol.analyze(p, Y, num_resamples=100, seed=100)
t0 = time.perf_counter()
for _ in range(50):
    sobol.analyze(p, Y, num_resamples=100, seed=100)
print("analyze ms", (time.perf_counter() - t0) / 50 * 1e3)
t0 = time.perf_counter()
for _ in range(50):
    saltelli.sample(p, 1024)
print("sample ms", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    sobol.analyze(p, Y, num_resamples=100, seed=100)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
