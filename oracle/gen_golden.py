#!/usr/bin/env python
"""Freeze golden vectors from the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py

Imports SALib from /root/reference/src, runs its own ``sobol.sample`` / ``saltelli.sample`` /
``sobol.analyze`` / test functions on small seeded inputs and writes the inputs and outputs to
``tests/golden/*.npz``.  /root/reference does not exist on the GPU box, so the tests only read
the committed fixtures.  Also copies the numbers of the reference's golden Saltelli matrix
``tests/data/model_input.txt`` into ``tests/golden/model_input_8e.npy``.
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, "/root/reference/src")
warnings.simplefilter("ignore")

from scipy.stats import qmc  # noqa: E402
from SALib.analyze import sobol as ref_analyze  # noqa: E402
from SALib.sample import saltelli as ref_saltelli  # noqa: E402
from SALib.sample import sobol as ref_sobol  # noqa: E402
from SALib.test_functions import Ishigami, Sobol_G  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)

ISHI = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}


def gprob(D, groups=None, bounds=None):
    p = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)],
         "bounds": bounds if bounds is not None else [[0.0, 1.0]] * D}
    if groups is not None:
        p["groups"] = groups
    return p


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path)} bytes")


# ---- 1. the reference's own golden Saltelli matrix (tests/data/model_input.txt, F6)
mi = np.loadtxt("/root/reference/tests/data/model_input.txt")
np.save(os.path.join(OUT, "model_input_8e.npy"), mi)
X = ref_saltelli.sample(dict(ISHI), 1024, skip_values=2048)
assert np.array_equal(np.array([float("%.8e" % v) for v in X.ravel()]).reshape(X.shape), mi)

# ---- 2. sample matrices
cases = {}
cases["ishigami_n64_c2_skip0"] = (dict(ISHI), 64, True, 0)
cases["ishigami_n64_c1_skip16"] = (dict(ISHI), 64, False, 16)
cases["g8_n128_c2_skip1000"] = (gprob(8), 128, True, 1000)
cases["g12_n32_c1_skip0"] = (gprob(12, bounds=[[-1.0 - i, 2.5 + 0.37 * i] for i in range(12)]), 32, False, 0)
cases["grp5_n64_c2_skip64"] = (gprob(5, groups=["a", "b", "a", "c", "b"],
                                      bounds=[[0.1 * i, 1.0 + i] for i in range(5)]), 64, True, 64)
cases["grp3_n32_c1_skip0"] = (gprob(3, groups=["A", "B", "A"]), 32, False, 0)
cases["g33_n16_c2_skip4096"] = (gprob(33), 16, True, 4096)
for name, (p, N, c2, skip) in cases.items():
    X = ref_sobol.sample(p, N, calc_second_order=c2, scramble=False, skip_values=skip)
    if skip > 0:
        X2 = ref_saltelli.sample(dict(p), N, calc_second_order=c2, skip_values=skip)
        assert np.array_equal(X, X2)
    save("sample_" + name, X=X, N=N, c2=c2, skip=skip, D=p["num_vars"],
         bounds=np.array(p["bounds"], dtype=np.float64),
         groups=np.array(p.get("groups", []), dtype="U8"))

# scrambled: state is host-supplied (north_star), so freeze scipy's _sv/_shift next to the output
for name, D, N, seed, skip in [("scr_g6_n64_seed7", 6, 64, 7, 0), ("scr_g40_n32_seed1234", 40, 32, 1234, 128)]:
    p = gprob(D)
    X = ref_sobol.sample(p, N, calc_second_order=True, scramble=True, skip_values=skip, seed=seed)
    eng = qmc.Sobol(d=2 * D, scramble=True, seed=seed)
    save("sample_" + name, X=X, N=N, c2=True, skip=skip, D=D, seed=seed,
         bounds=np.array(p["bounds"], dtype=np.float64),
         sv=eng._sv.astype(np.uint32), shift=eng._shift.astype(np.uint32))

# non-uniform `dists` (util/__init__.py:126-289): every distribution the reference supports, in one problem
DIST_PROBLEM = {
    "num_vars": 9, "names": [f"p{i}" for i in range(9)],
    "dists": ["unif", "triang", "norm", "truncnorm", "lognorm", "logunif", "weibull", "triang", "truncnorm"],
    "bounds": [[-2.0, 3.5], [1.0, 3.0, 0.25], [0.5, 2.0], [-1.0, 4.0, 1.0, 2.0], [0.1, 0.6], [1e-3, 10.0],
               [1.7, 2.5, 0.3], [3.0, 0.5], [2.0, 6.0, 1.0, 1.5]],
}
for name, N, c2, skip, groups in [("dists_n64_c2_skip64", 64, True, 64, None), ("dists_n32_c1_skip0", 32, False, 0, None),
                                  ("dists_grp_n32_c2_skip32", 32, True, 32, ["a", "b", "a", "c", "b", "c", "d", "d", "a"])]:
    p = {k: (list(v) if isinstance(v, list) else v) for k, v in DIST_PROBLEM.items()}
    if groups:
        p["groups"] = groups
    X = ref_sobol.sample(p, N, calc_second_order=c2, scramble=False, skip_values=skip)
    save("sample_" + name, X=X, N=N, c2=c2, skip=skip, groups=np.array(groups or [], dtype="U8"))

# first ten Joe-Kuo points (tests/test_sobol.py:24-35 checks these with atol 5e-2)
save("joe_kuo_first10", pts=qmc.Sobol(d=3, scramble=False).random(16)[:10])

# ---- 3. test functions
rng = np.random.default_rng(5)
Xi = rng.uniform(-np.pi, np.pi, size=(257, 3))
Xg = rng.uniform(0, 1, size=(129, 8))
a16 = np.tile([0, 1, 4.5, 9, 99, 99, 99, 99], 2).astype(float)
Xg16 = rng.uniform(0, 1, size=(65, 16))
delta = rng.uniform(0, 1, size=8)
alpha = rng.uniform(0.5, 2.0, size=8)
save("test_functions", Xi=Xi, Yi=Ishigami.evaluate(Xi), Xg=Xg, Yg=Sobol_G.evaluate(Xg),
     Xg16=Xg16, a16=a16, Yg16=Sobol_G.evaluate(Xg16, a=a16),
     delta=delta, alpha=alpha, Ygda=Sobol_G.evaluate(Xg, delta=delta, alpha=alpha),
     Yzero=Sobol_G.evaluate(np.zeros((1, 8))))

# ---- 4. analyze: reference outputs with keep_resamples=True, seeded
def run_analyze(name, p, N, c2, R, seed, model, conf=0.95):
    X = ref_sobol.sample(dict(p), N, calc_second_order=c2, scramble=False, skip_values=0)
    Y = model(X)
    S = ref_analyze.analyze(dict(p), Y, calc_second_order=c2, num_resamples=R, conf_level=conf,
                            keep_resamples=True, seed=seed)
    r = np.random.default_rng(seed).integers(N, size=(N, R))
    kw = {k: np.asarray(v) for k, v in S.items()}
    save("analyze_" + name, Y=Y, N=N, c2=c2, R=R, seed=seed, conf=conf, D=p["num_vars"],
         groups=np.array(p.get("groups", []), dtype="U8"), r=r.astype(np.int32), **kw)


run_analyze("ishigami_n1024_r100", dict(ISHI), 1024, True, 100, 100, Ishigami.evaluate)
run_analyze("ishigami_n256_c1_r50", dict(ISHI), 256, False, 50, 7, Ishigami.evaluate, conf=0.9)
run_analyze("g8_n512_r64", gprob(8), 512, True, 64, 11, Sobol_G.evaluate)
run_analyze("g8_n512_c1_r64", gprob(8), 512, False, 64, 12, Sobol_G.evaluate)
run_analyze("g16_n256_r32", gprob(16), 256, True, 32, 3, lambda X: Sobol_G.evaluate(X, a=a16))
run_analyze("grp3_n512_r40", {**ISHI, "groups": ["A", "B", "A"]}, 512, True, 40, 21, Ishigami.evaluate)
# D=64 through the second-order register-tile kernel.  A smooth `a` keeps the tiny-N problem well
# conditioned (with the tiled default `a`, 64..256 unscrambled points give var(Y) ~ 1e7 and S2 is pure
# cancellation, so even two CPU summation orders disagree at 1e-10).
a64 = np.linspace(1.0, 60.0, 64)
run_analyze("g64_n128_r16", gprob(64), 128, True, 16, 5, lambda X: Sobol_G.evaluate(X, a=a64))

# ---- 4b. finite_diff sample + dgsm analyze (consumers of the same Sobol' generator)
from SALib.analyze import dgsm as ref_dgsm  # noqa: E402
from SALib.sample import finite_diff as ref_fd  # noqa: E402

for name, p, N, delta, skip, R, seed in [("ishigami_n256", dict(ISHI), 256, 0.001, 1024, 50, 9),
                                         ("g8_n128", gprob(8, bounds=[[0.0, 1.0 + 0.25 * i] for i in range(8)]), 128, 0.01, 64, 30, 4)]:
    X = ref_fd.sample(dict(p), N, delta=delta, skip_values=skip)
    Y = Ishigami.evaluate(X) if p["num_vars"] == 3 else Sobol_G.evaluate(X / np.array(p["bounds"])[:, 1])
    S = ref_dgsm.analyze(dict(p), X, Y, num_resamples=R, conf_level=0.95, seed=seed)
    np.random.seed(seed)
    rs = np.stack([np.random.randint(N, size=(R, N)) for _ in range(p["num_vars"])])
    save("dgsm_" + name, X=X, Y=Y, N=N, delta=delta, skip=skip, R=R, seed=seed, D=p["num_vars"],
         bounds=np.array(p["bounds"], dtype=np.float64), r=rs.astype(np.int32),
         **{k: np.asarray(v) for k, v in S.items() if k != "names"})

# ---- 5. the CLI golden (tests/test_cli_analyze.py:250-275, F7): Y from the 8-digit X, R=1000, seed=100
Y = Ishigami.evaluate(mi)
S = ref_analyze.analyze(dict(ISHI), Y, calc_second_order=True, num_resamples=1000, conf_level=0.95, seed=100)
save("analyze_cli_golden", Y=Y, **{k: np.asarray(v) for k, v in S.items()})
print("S1", S["S1"], "ST", S["ST"])
