#!/usr/bin/env python
"""Round-2 golden vectors from the UNMODIFIED reference (build container only; adds to gen_golden.py's set).

    python oracle/gen_golden_r02.py

* analyze_g8_n65536_c1_r1000_seed1234: ``SALib.analyze.sobol.analyze`` at a size where round 1 silently left the
  reference's resample stream (N*R = 6.5e7 draws > 2^24): D=8, N=2^16, first+total order, R=1000, seed=1234.
  Only the results and a fingerprint of Y are stored (Y itself is 5 MB and is rebuilt by the test from the
  bit-exact sampler + model).
* analyze_g8_outlier_n512_r32: a heavy-tailed output -- one base row scaled by 1e6 -- second order; bounds what
  the per-column power-of-two scaling of the integer engine loses on such data against the reference itself.
* sample_sha_*: SHA-256 of the reference's sample matrices at D=512 and D=1024 (unscrambled and scrambled; the
  matrices are 17-67 MB), with scipy's scramble state frozen beside them.
* analyze_g207_n32_r4 / analyze_g208_n32_r4: the reference at the integer engine's selection boundary (Dg=207/208).
"""
import hashlib
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, "/root/reference/src")
warnings.simplefilter("ignore")

from scipy.stats import qmc  # noqa: E402
from SALib.analyze import sobol as ref_analyze  # noqa: E402
from SALib.sample import sobol as ref_sobol  # noqa: E402
from SALib.test_functions import Sobol_G  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def gprob(D):
    return {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path)} bytes")


# ---- 1. seeded numpy parity beyond 2^24 draws
D, N, R, seed = 8, 1 << 16, 1000, 1234
X = ref_sobol.sample(gprob(D), N, calc_second_order=False, scramble=False, skip_values=N)
Y = Sobol_G.evaluate(X)
S = ref_analyze.analyze(gprob(D), Y, calc_second_order=False, num_resamples=R, conf_level=0.95, seed=seed)
save("analyze_g8_n65536_c1_r1000_seed1234", N=N, R=R, seed=seed, D=D, skip=N, c2=False, conf=0.95,
     Y_sum=Y.sum(), Y_head=Y[:64], Y_tail=Y[-64:], Y_sha256=hashlib.sha256(Y.tobytes()).hexdigest(),
     **{k: np.asarray(v) for k, v in S.items()})

# ---- 2. heavy-tailed output
D, N, R, seed = 8, 512, 32, 77
X = ref_sobol.sample(gprob(D), N, calc_second_order=True, scramble=False, skip_values=N)
Y = Sobol_G.evaluate(X)
step = 2 * D + 2
Y = Y.copy()
Y[137 * step:138 * step] *= 1e6          # one base row (all of its A/AB/BA/B evaluations) is a 1e6x outlier
S = ref_analyze.analyze(gprob(D), Y, calc_second_order=True, num_resamples=R, conf_level=0.95,
                        keep_resamples=True, seed=seed)
r = np.random.default_rng(seed).integers(N, size=(N, R))
save("analyze_g8_outlier_n512_r32", Y=Y, N=N, c2=True, R=R, seed=seed, conf=0.95, D=D,
     groups=np.array([], dtype="U8"), r=r.astype(np.int32), **{k: np.asarray(v) for k, v in S.items()})

# ---- 3. wide sample matrices: hashes
for D, N, c2 in [(512, 8, False), (1024, 4, False), (512, 2, True)]:
    for scramble in (False, True):
        skip, sd = 64, 4321
        Xs = ref_sobol.sample(gprob(D), N, calc_second_order=c2, scramble=scramble, skip_values=skip,
                              seed=sd if scramble else None)
        kw = {}
        if scramble:
            eng = qmc.Sobol(d=2 * D, scramble=True, seed=sd)
            kw = {"sv": eng._sv.astype(np.uint32), "shift": eng._shift.astype(np.uint32)}
        save(f"sample_sha_d{D}_n{N}_c{int(c2)}_{'scr' if scramble else 'plain'}", D=D, N=N, c2=c2, skip=skip,
             seed=sd, scramble=scramble, shape=np.array(Xs.shape), sha256=hashlib.sha256(Xs.tobytes()).hexdigest(),
             first_row=Xs[0], last_row=Xs[-1], **kw)

# ---- 4. the integer engine's selection boundary (second order: Dg <= 207 fits the feature kernels)
for D in (207, 208):
    N, R, seed = 32, 4, 9
    a = np.linspace(1.0, 80.0, D)
    X = ref_sobol.sample(gprob(D), N, calc_second_order=True, scramble=False, skip_values=N)
    Y = Sobol_G.evaluate(X, a=a)
    S = ref_analyze.analyze(gprob(D), Y, calc_second_order=True, num_resamples=R, conf_level=0.95,
                            keep_resamples=True, seed=seed)
    r = np.random.default_rng(seed).integers(N, size=(N, R))
    iu = np.triu_indices(D, 1)
    save(f"analyze_g{D}_n32_r4", Y=Y, N=N, c2=True, R=R, seed=seed, conf=0.95, D=D, r=r.astype(np.int32),
         S1=S["S1"], ST=S["ST"], S1_conf=S["S1_conf"], ST_conf=S["ST_conf"], S2_upper=S["S2"][iu],
         S2_conf_upper=S["S2_conf"][iu], S1_conf_all=S["S1_conf_all"], ST_conf_all=S["ST_conf_all"])
