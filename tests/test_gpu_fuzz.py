"""GPU parity, randomized: analyze over random (Dg, N, order, R, kernel, row partition) against the oracle.

Seeded, so the set is fixed; it exists to cross every kernel choice with shapes nobody hand-picked (the
hand-picked shape tests missed a shared-memory limit in finalize at Dg >= 225 once)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20240)
    out = []
    for i in range(48):
        c2 = bool(rng.integers(2))
        Dg = int(rng.choice([1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 24, 31, 33, 40, 47, 48, 49, 56, 57, 63, 64, 65, 72, 90, 130,
                             230 if c2 else 400]))
        N = int(rng.choice([1, 2, 7, 31, 32, 33, 64, 100, 129, 257, 400]))
        if Dg >= 130:
            N = min(N, 64)
        R = int(rng.integers(1, 9))
        algos = ["auto", "generic", "gemv"] + (["tile8"] if c2 and Dg <= 512 else [])
        algo = str(rng.choice(algos))
        rows = bool(rng.integers(3) == 0) and N >= 4
        out.append((i, Dg, N, c2, R, algo, rows))
    return out


@pytest.mark.parametrize("i,Dg,N,c2,R,algo,rows", _cases())
def test_random_configuration_matches_oracle(i, Dg, N, c2, R, algo, rows):
    from oracle import analyze as o_analyze
    from salib_b200.analyze import sobol

    rng = np.random.default_rng(1000 + i)
    step = 2 * Dg + 2 if c2 else Dg + 2
    base = rng.normal(size=(N, 1))
    Y = (base + 0.7 * rng.normal(size=(N, step)) * rng.uniform(0.3, 1.5, size=(1, step))).ravel() * 2.0 - 3.0
    r = rng.integers(N, size=(N, R))
    p = {"num_vars": Dg, "names": [f"x{j}" for j in range(Dg)], "bounds": [[0, 1]] * Dg}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        So = o_analyze.analyze(Dg, Y, c2, R, keep_resamples=True, r=r)
        if rows:
            cuts = sorted({0, N, *rng.integers(1, N, size=2).tolist()})
            shards = [(lo, Y[lo * step:hi * step]) for lo, hi in zip(cuts, cuts[1:])]
            S = sobol.analyze_row_sharded(p, shards, N, calc_second_order=c2, keep_resamples=True, r=r, algo=algo)
        else:
            S = sobol.analyze_device(p, Y, calc_second_order=c2, keep_resamples=True, r=r, algo=algo)
    keys = ["S1", "ST", "S1_conf", "ST_conf", "S1_conf_all", "ST_conf_all"]
    for k in keys:
        np.testing.assert_allclose(S[k], So[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
    if c2 and Dg > 1:
        iu = np.triu_indices(Dg, 1)
        for k in ("S2", "S2_conf"):
            np.testing.assert_allclose(S[k][iu], np.asarray(So[k])[iu], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
