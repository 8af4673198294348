"""GPU parity: Saltelli sample generation vs the oracle / frozen reference outputs (bit-exact)."""
import warnings

import numpy as np
import pytest

from conftest import golden, problem_from

pytestmark = pytest.mark.gpu

SAMPLE_CASES = ["ishigami_n64_c2_skip0", "ishigami_n64_c1_skip16", "g8_n128_c2_skip1000", "g12_n32_c1_skip0",
                "grp5_n64_c2_skip64", "grp3_n32_c1_skip0", "g33_n16_c2_skip4096"]
ISHI = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}


def gprob(D, groups=None, bounds=None):
    p = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)],
         "bounds": bounds if bounds is not None else [[0.0, 1.0]] * D}
    if groups is not None:
        p["groups"] = groups
    return p


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


@pytest.mark.parametrize("name", SAMPLE_CASES)
def test_matches_frozen_reference(name):
    from salib_b200.sample import saltelli, sobol

    z = golden("sample_" + name)
    p = problem_from(z)
    X = sobol.sample(p, int(z["N"]), calc_second_order=bool(z["c2"]), scramble=False, skip_values=int(z["skip"]))
    assert X.dtype == np.float64 and X.shape == z["X"].shape
    assert np.array_equal(X, z["X"])
    assert p["sample_scaled"] is True
    if int(z["skip"]) > 0:
        X2 = saltelli.sample(problem_from(z), int(z["N"]), calc_second_order=bool(z["c2"]), skip_values=int(z["skip"]))
        assert np.array_equal(X2, z["X"])


@pytest.mark.parametrize("name", ["scr_g6_n64_seed7", "scr_g40_n32_seed1234"])
def test_scrambled_matches_frozen_reference(name):
    """scramble=True with an int seed: the LMS/shift state is host-supplied, the draw is ours."""
    from salib_b200.sample import sobol

    z = golden("sample_" + name)
    X = sobol.sample(problem_from(z), int(z["N"]), calc_second_order=True, scramble=True,
                     skip_values=int(z["skip"]), seed=int(z["seed"]))
    assert np.array_equal(X, z["X"])


DIST_PROBLEM = {
    "num_vars": 9, "names": [f"p{i}" for i in range(9)],
    "dists": ["unif", "triang", "norm", "truncnorm", "lognorm", "logunif", "weibull", "triang", "truncnorm"],
    "bounds": [[-2.0, 3.5], [1.0, 3.0, 0.25], [0.5, 2.0], [-1.0, 4.0, 1.0, 2.0], [0.1, 0.6], [1e-3, 10.0],
               [1.7, 2.5, 0.3], [3.0, 0.5], [2.0, 6.0, 1.0, 1.5]],
}


@pytest.mark.parametrize("name", ["dists_n64_c2_skip64", "dists_n32_c1_skip0", "dists_grp_n32_c2_skip32"])
def test_nonuniform_dists_match_frozen_reference(name):
    """Device inverse CDFs (triang, norm, truncnorm, lognorm, logunif, weibull) vs scipy's ppf in the
    reference (util/__init__.py:126-289).  Transcendental: tolerance, not bit-exact; the `unif` column is
    bit-exact; u = 0 (first unscrambled point) gives -inf / the lower bound exactly like scipy."""
    from salib_b200.sample import sobol

    z = golden("sample_" + name)
    p = {k: (list(v) if isinstance(v, list) else v) for k, v in DIST_PROBLEM.items()}
    if len(z["groups"]):
        p["groups"] = [str(g) for g in z["groups"]]
    X = sobol.sample(p, int(z["N"]), calc_second_order=bool(z["c2"]), scramble=False, skip_values=int(z["skip"]))
    assert X.shape == z["X"].shape
    assert np.array_equal(X[:, 0], z["X"][:, 0])
    assert np.array_equal(np.isinf(X), np.isinf(z["X"]))
    np.testing.assert_allclose(X, z["X"], rtol=1e-11, atol=1e-13)


def test_reference_golden_saltelli_file():
    """reference tests/data/model_input.txt (8192 x 3, %.8e) == saltelli.sample(Ishigami, 1024, skip 2048)."""
    from salib_b200.sample import saltelli

    X = saltelli.sample(dict(ISHI), 1024, skip_values=2048)
    mi = golden("model_input_8e.npy")
    assert X.shape == mi.shape
    assert np.array_equal(np.array([float("%.8e" % v) for v in X.ravel()]).reshape(X.shape), mi)


def test_saltelli_default_skip():
    from oracle import saltelli as o
    from salib_b200.sample import saltelli

    for N, skip in ((8, 16), (1024, 1024), (100, 128)):
        X = saltelli.sample(dict(ISHI), N)
        assert np.array_equal(X, o.sample(dict(ISHI), N, True, skip_values=skip, closed_form=True))


@pytest.mark.parametrize("D,N,c2,skip,groups", [
    (2, 64, True, 0, None), (4, 33, False, 7, None), (8, 1000, True, 1024, None), (16, 257, True, 0, None),
    (32, 100, False, 64, None), (64, 130, True, 4096, None), (128, 40, True, 0, None), (256, 9, False, 3, None),
    (3, 4096, True, 0, None), (5, 777, False, 1 << 20, None), (6, 100, True, 0, None), (10, 64, True, 16, None),
    (40, 50, True, 0, None), (100, 20, False, 0, None), (200, 11, True, 5, None), (1, 16, True, 0, None),
    (8, 64, True, 0, ["a", "a", "b", "c", "b", "d", "d", "a"]),
    (64, 32, True, 8, [f"g{i % 5}" for i in range(64)]),
    (7, 90, False, 0, ["u", "v", "u", "w", "v", "u", "w"]),
])
def test_matches_oracle_over_shapes(D, N, c2, skip, groups):
    """Register path (D in {2..32} u {64k}), generic path (odd / other D), grouped and not, both orders."""
    from oracle import saltelli as o
    from salib_b200.sample import sobol

    rng = np.random.default_rng(D * 1000 + N)
    bounds = [[float(a), float(a + b)] for a, b in zip(rng.uniform(-5, 5, D), rng.uniform(0.1, 7, D))]
    p = gprob(D, groups, bounds)
    X = sobol.sample(p, N, calc_second_order=c2, scramble=False, skip_values=skip)
    Xo = o.sample(gprob(D, groups, bounds), N, c2, skip_values=skip, closed_form=True)
    assert X.shape == Xo.shape
    assert np.array_equal(X, Xo)


def test_index_range_sharding_concatenates_to_full_matrix():
    """Multi-GPU sampling shards by Sobol' index range with no communication: the shards of a
    4-rank job, computed here one after the other, concatenate to the single-GPU matrix."""
    from salib_b200.sample import sobol

    p = gprob(16)
    full = sobol.sample_device(p, 1000, scramble=False, skip_values=64).cpu().numpy()
    parts = [sobol.sample_device(gprob(16), 1000, scramble=False, skip_values=64, shard=(r, 4)).cpu().numpy()
             for r in range(4)]
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_large_structure_properties():
    """BASELINE-size style checks that need no oracle: D=64, N=2^14 second order (1.1 GB would be
    N=2^17; here 136 MB).  A-rows and B-rows are the plain Sobol' points, AB_k differs from A only in
    column k, BA_k from B only in column k, and every column of A is a permutation of k/2^m."""
    import torch

    from salib_b200.sample import sobol

    D, N = 64, 1 << 14
    X = sobol.sample_device(gprob(D), N, scramble=False).reshape(N, 2 * D + 2, D)
    A, B = X[:, 0, :], X[:, -1, :]
    for k in (0, 17, 63):
        ab, ba = X[:, 1 + k, :], X[:, 1 + D + k, :]
        mask = torch.ones(D, dtype=torch.bool, device=X.device)
        mask[k] = False
        assert torch.equal(ab[:, mask], A[:, mask]) and torch.equal(ab[:, k], B[:, k])
        assert torch.equal(ba[:, mask], B[:, mask]) and torch.equal(ba[:, k], A[:, k])
    grid = torch.arange(N, dtype=torch.float64, device=X.device) / N
    for M in (A, B):
        assert torch.equal(torch.sort(M, dim=0).values, grid[:, None].expand(N, D))


def test_scale_samples_standalone():
    """util.scale_samples (reference util/__init__.py:56-90, tests/test_util.py:123-166): uniform bounds bit-exact
    and in place, distributions against the frozen reference matrix."""
    from oracle import saltelli as o
    from salib_b200.sample import sobol_sequence
    from salib_b200.util import scale_samples

    rng = np.random.default_rng(3)
    U = rng.random((500, 4))
    p = gprob(4, bounds=[[10.0, 100.0], [-3.0, 7.5], [0.0, 1e-3], [5.0, 6.0]])
    expect = o.scale_uniform(U.copy(), p["bounds"])
    got = U.copy()
    ret = scale_samples(got, p)
    assert ret is got and np.array_equal(got, expect) and p["sample_scaled"] is True
    with pytest.raises(ValueError):
        scale_samples(U.copy(), gprob(4, bounds=[[0, 1], [2, 1], [0, 1], [0, 1]]))
    # distributions: rebuild the frozen dists matrix from plain Sobol' points + cross matrix + scale_samples
    z = golden("sample_dists_n32_c1_skip0")
    base = sobol_sequence.sample(32, 18)
    X01 = o.cross_matrix(base, 9, np.arange(9, dtype=np.int32), 9, False)
    pd_ = {k: (list(v) if isinstance(v, list) else v) for k, v in DIST_PROBLEM.items()}
    X = scale_samples(X01, pd_)
    np.testing.assert_allclose(X, z["X"], rtol=1e-11, atol=1e-13)


def test_sobol_sequence_dropin():
    """reference tests/test_sobol.py:21-36 (first ten Joe-Kuo points) + oracle over many dims / a skip."""
    from oracle import sobol_seq
    from salib_b200.sample import sobol_sequence

    expected = np.array([[0, 0, 0], [0.5, 0.5, 0.5], [0.75, 0.25, 0.25], [0.25, 0.75, 0.75],
                         [0.375, 0.375, 0.625], [0.875, 0.875, 0.125], [0.625, 0.125, 0.875],
                         [0.125, 0.625, 0.375], [0.1875, 0.3125, 0.9375], [0.6875, 0.8125, 0.4375]])
    assert np.array_equal(sobol_sequence.sample(10, 3), expected)
    assert np.array_equal(sobol_sequence.sample(1000, 300), sobol_seq.points_closed_form(1000, 300))
    assert np.array_equal(sobol_sequence.sample(77, 5, skip=12345), sobol_seq.points_closed_form(77, 5, skip=12345))
    with pytest.raises(ValueError, match="not enough dimensions"):
        sobol_sequence.sample(4, 21202)
    assert sobol_sequence.index_of_least_significant_zero_bit(0b1011) == 3


def test_errors():
    from salib_b200 import _lib
    from salib_b200.sample import sobol

    with pytest.raises(ValueError):
        sobol.sample(gprob(2, bounds=[[0, 1], [3, 3]]), 8, scramble=False)
    L = _lib.lib()
    assert L.sa_saltelli_sample_f64(None, None, None, 0, 8, 2, 2, None, 1, None, None, None, None, None) == -1
