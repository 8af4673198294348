"""GPU checks at BASELINE.json's full sizes through size-independent properties (no oracle can run there):
the Saltelli matrix structure at D=64, N=2^20 (69.8 GB, verified chunk by chunk against the independent
plain-points kernel), and invariances of sobol.analyze at D=64, N=2^18..2^20 under the device Philox RNG."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gprob(D):
    return {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def test_headline_sample_structure_full_size():
    """D=64, N=2^20, second order: every base row's A / B rows are Sobol' point skip+i (dims 0..63 / 64..127,
    from sa_sobol_points_f64 -- a different kernel), AB_k differs from A exactly in column k (= B's), BA_k
    from B exactly in column k.  Generated and checked in 16 index-range shards of 4.4 GB."""
    import torch

    from salib_b200.sample import sobol, sobol_sequence

    D, N, shards, skip = 64, 1 << 20, 16, 1 << 20
    step = 2 * D + 2
    eye = torch.eye(D, dtype=torch.bool, device="cuda")
    free = torch.cuda.mem_get_info()[0]
    if free < 12e9:
        pytest.skip("needs 12 GB of free device memory")
    for r in range(shards):
        X = sobol.sample_device(gprob(D), N, scramble=False, skip_values=skip, shard=(r, shards))
        n = N // shards
        X = X.reshape(n, step, D)
        P = sobol_sequence.sample(n, 2 * D, skip=skip + r * n, device=True)
        A, B = X[:, 0, :], X[:, -1, :]
        assert torch.equal(A, P[:, :D]) and torch.equal(B, P[:, D:])
        AB, BA = X[:, 1:1 + D, :], X[:, 1 + D:1 + 2 * D, :]
        assert torch.equal(AB, torch.where(eye[None], B[:, None, :], A[:, None, :]))
        assert torch.equal(BA, torch.where(eye[None], A[:, None, :], B[:, None, :]))
        del X, P, AB, BA
    torch.cuda.empty_cache()


def _headline_y(log2n):
    from salib_b200.sample import sobol
    from salib_b200.test_functions import Sobol_G

    D, N = 64, 1 << log2n
    p = gprob(D)
    plan, first = sobol._plan(p, N, True, False, N, None)
    return p, Sobol_G.evaluate_sample(plan, first, N, a=np.resize([0, 1, 4.5, 9, 99, 99, 99, 99.0], D))


def _close(S, T, rtol, atol):
    iu = np.triu_indices(64, 1)
    for k in ("S1", "ST", "S1_conf", "ST_conf"):
        np.testing.assert_allclose(S[k], T[k], rtol=rtol, atol=atol, err_msg=k)
    for k in ("S2", "S2_conf"):
        np.testing.assert_allclose(S[k][iu], T[k][iu], rtol=rtol, atol=atol, err_msg=k)


def test_analyze_invariances_d64():
    """Properties that hold for any N: (1) affine invariance, analyze(3Y+7) == analyze(Y) with the same
    resample stream; (2) resample sharding: [0,R) in one call == two halves run separately (what
    multi-GPU does), the Philox stream being keyed by the global resample id; (3) a resample whose multiplicities are all 1 reproduces the point estimate;
    (4) the generic kernel agrees with the tile kernel."""
    import torch

    from salib_b200 import _runtime as rt
    from salib_b200.analyze import sobol

    p, Y = _headline_y(18)
    N, R = 1 << 18, 296
    S = sobol.analyze_device(p, Y, num_resamples=R, seed=100, resample="philox", keep_resamples=True, algo="fp64")
    T = sobol.analyze_device(p, 3.0 * Y + 7.0, num_resamples=R, seed=100, resample="philox", keep_resamples=True,
                             algo="fp64")
    _close(S, T, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(S["S1_conf_all"], T["S1_conf_all"], rtol=1e-9, atol=1e-12)
    assert S.engine == "fp64-tile8"

    eng = sobol.SobolAnalyzer(64, N, True, algo="fp64")
    eng.load(Y)
    rs = sobol._Resampler(N, R, 100, "philox")
    whole = eng.bootstrap(rs, 0, R)
    halves = torch.cat([eng.bootstrap(rs, 0, R // 2), eng.bootstrap(rs, R // 2, R)])
    # same multiplicities, different row splits (summation order): equal to rounding, not bitwise
    np.testing.assert_allclose(halves.cpu().numpy(), whole.cpu().numpy(), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(whole[:, :64].cpu().numpy(), S["S1_conf_all"], rtol=0, atol=0)

    ones = torch.ones((eng.Npad,), dtype=torch.uint8, device="cuda")
    est1 = rt.empty((1, eng.nest), torch.float64)
    eng._accumulate(ones, 1, est1)
    np.testing.assert_allclose(est1.cpu().numpy(), eng.point().cpu().numpy(), rtol=1e-11, atol=1e-14)

    gen = sobol.SobolAnalyzer(64, N, True, algo="generic")
    gen.load(Y)
    np.testing.assert_allclose(gen.bootstrap(rs, 0, 4).cpu().numpy(), whole[:4].cpu().numpy(), rtol=1e-9, atol=1e-12)


def test_headline_analyze_full_size_matches_analytic_indices():
    """D=64, N=2^20, second order, R=1184 (one wave): the point estimates of the important factors agree
    with the analytic G-function indices within 4 bootstrap CI half-widths, and S2 of an (a=0, a=0) pair
    with V_i V_j prod_{k != i,j}(1 + V_k)-free closed form V_i V_j / Var."""
    from salib_b200.analyze import sobol

    p, Y = _headline_y(20)
    S = sobol.analyze_device(p, Y, num_resamples=1184, seed=100, resample="philox")
    a = np.resize([0, 1, 4.5, 9, 99, 99, 99, 99.0], 64)
    Vi = 1.0 / (3.0 * (1.0 + a) ** 2)
    var = np.prod(1.0 + Vi) - 1.0
    S1 = Vi / var
    ST = Vi * np.prod(1.0 + Vi) / (1.0 + Vi) / var
    big = a <= 1.0
    assert np.all(np.abs(S["S1"][big] - S1[big]) < 4 * S["S1_conf"][big] / 1.96 + 1e-4)
    assert np.all(np.abs(S["ST"][big] - ST[big]) < 4 * S["ST_conf"][big] / 1.96 + 1e-3)
    assert abs(S["S2"][0, 8] - Vi[0] * Vi[8] / var) < 4 * S["S2_conf"][0, 8] / 1.96 + 1e-3
    assert np.isfinite(S["S2_conf"][np.triu_indices(64, 1)]).all()
