#!/usr/bin/env python
"""Freeze golden vectors of the UNMODIFIED reference's ``discrepancy.analyze`` (build container only).

    python oracle/gen_golden_discrepancy.py

Same rules as gen_golden.py: imports SALib from /root/reference/src, writes inputs and outputs to
``tests/golden/discrepancy_*.npz``; the tests read only the committed fixtures.  Also checks the oracle's
restated closed forms against scipy.stats.qmc.discrepancy itself."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, ".."))
warnings.simplefilter("ignore")

from scipy.stats import qmc  # noqa: E402
from SALib.analyze import discrepancy as ref_disc  # noqa: E402
from SALib.sample import latin as ref_latin  # noqa: E402
from SALib.sample import sobol as ref_sobol  # noqa: E402
from SALib.test_functions import Ishigami, Sobol_G  # noqa: E402

from oracle import discrepancy as o_disc  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
ISHI = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}
METHODS = ["WD", "CD", "MD", "L2-star"]

rng = np.random.default_rng(0)
for n, d in [(50, 2), (333, 2), (200, 5)]:
    x = rng.random((n, d))
    for m in METHODS:
        a, b = o_disc.discrepancy(x, m), qmc.discrepancy(x, method=m)
        assert abs(a - b) <= 1e-13 + 1e-9 * abs(b), (n, d, m, a, b)
print("oracle closed forms == scipy.stats.qmc.discrepancy")


def run(name, p, X, Y):
    out = {}
    for m in METHODS:
        S = ref_disc.analyze(dict(p), X, Y, method=m)
        out["s_" + m.replace("-", "_")] = np.asarray(S["s_discrepancy"], dtype=np.float64)
        mine = o_disc.analyze(p, X, Y, m)
        assert np.allclose(mine, out["s_" + m.replace("-", "_")], rtol=1e-8, atol=1e-12), (name, m)
    path = os.path.join(OUT, "discrepancy_" + name + ".npz")
    np.savez_compressed(path, X=X, Y=Y, D=p["num_vars"], bounds=np.array(p["bounds"], dtype=np.float64),
                        groups=np.array(p.get("groups", []), dtype="U8"), **out)
    print(name, os.path.getsize(path), "bytes", out["s_WD"])


# the reference's docstring example (discrepancy.py:70-84): Latin hypercube + Ishigami
X = ref_latin.sample(dict(ISHI), 1000, seed=3)
run("ishigami_latin_n1000", ISHI, X, Ishigami.evaluate(X))
# a Saltelli design of the hot path feeding the same analyzer; ragged N (not a multiple of any tile)
g8 = {"num_vars": 8, "names": [f"x{i + 1}" for i in range(8)], "bounds": [[0.0, 1.0 + 0.25 * i] for i in range(8)]}
X = ref_sobol.sample(dict(g8), 64, calc_second_order=False, scramble=False)[:601]
run("g8_sobol_n601", g8, X, Sobol_G.evaluate(X / np.array(g8["bounds"])[:, 1]))
# D >= 13 takes the 16-inputs-per-thread kernel; grouped problem
g20 = {"num_vars": 20, "names": [f"x{i + 1}" for i in range(20)], "bounds": [[-1.0, 2.0]] * 20,
       "groups": ["g%d" % (i % 3) for i in range(20)]}
X = np.random.default_rng(5).uniform(-1.0, 2.0, size=(300, 20))
run("grp20_uniform_n300", g20, X, (X ** 2).sum(axis=1) + X[:, 0] * X[:, 7])
