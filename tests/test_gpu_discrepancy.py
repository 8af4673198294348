"""GPU parity: discrepancy.analyze vs the frozen reference outputs and the oracle.

Tolerance: the discrepancy is ``c0 - 2/n*s1 + s2/n^2`` with O(1) terms and a result of 1e-2..1e-5, so the
last ~3-5 digits of a float64 evaluation depend on the summation order (scipy's own value moves by 1e-11
relative at n = 333 against a NumPy evaluation of the same formula).  rtol 1e-9 on the normalised indices."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
METHODS = ["WD", "CD", "MD", "L2-star"]
CASES = ["ishigami_latin_n1000", "g8_sobol_n601", "grp20_uniform_n300"]


def _problem(z):
    D = int(z["D"])
    p = {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": z["bounds"].tolist()}
    if z["groups"].size:
        p["groups"] = [str(g) for g in z["groups"]]
    return p


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("method", METHODS)
def test_matches_frozen_reference(name, method):
    from salib_b200.analyze import discrepancy

    z = golden("discrepancy_" + name)
    p = _problem(z)
    S = discrepancy.analyze(p, z["X"], z["Y"], method=method)
    np.testing.assert_allclose(S["s_discrepancy"], z["s_" + method.replace("-", "_")], rtol=1e-9, atol=1e-13)
    assert S["names"] == p["names"]
    if not z["groups"].size:
        assert list(S.to_df().index) == p["names"]


@pytest.mark.parametrize("N,D", [(1, 3), (2, 1), (63, 4), (64, 5), (257, 13), (1500, 17), (5000, 3)])
@pytest.mark.parametrize("method", METHODS)
def test_matches_oracle_over_shapes(N, D, method):
    """Tile edges: N around the 64 / 256 point tiles, D around the 4 / 16 input chunks, row splits."""
    import torch

    from oracle import discrepancy as o_disc
    from salib_b200.analyze import discrepancy

    rng = np.random.default_rng(N * 31 + D)
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[-1.0 - i, 2.0 + 0.5 * i] for i in range(D)]}
    b = np.array(p["bounds"])
    X = rng.uniform(b[:, 0], b[:, 1], size=(N, D))
    Y = np.sin(X).sum(axis=1) + X[:, 0] ** 2 if N > 1 else np.array([1.0])
    if N == 1:
        with pytest.raises(ValueError):   # a single output value: min == max (scipy: inconsistent bounds)
            discrepancy.analyze(p, X, Y, method=method)
        return
    ref = o_disc.analyze(p, X, Y, method)
    S = discrepancy.analyze(p, torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(), method=method)
    np.testing.assert_allclose(S["s_discrepancy"], ref, rtol=1e-9, atol=1e-13)


def test_errors_match_reference():
    from salib_b200.analyze import discrepancy

    p = {"num_vars": 2, "names": ["a", "b"], "bounds": [[0.0, 1.0], [0.0, 2.0]]}
    X = np.array([[0.1, 0.2], [0.5, 2.5], [0.3, 0.3]])
    Y = np.array([1.0, 2.0, 3.0])
    with pytest.raises(ValueError, match="Sample is out of bounds"):
        discrepancy.analyze(p, X, Y)
    X[1, 1] = 1.5
    with pytest.raises(ValueError, match="not a valid method"):
        discrepancy.analyze(p, X, Y, method="XX")
    with pytest.raises(ValueError, match="Bounds are not consistent"):
        discrepancy.analyze(p, X, np.ones(3))
    with pytest.raises(ValueError, match="Bounds are not consistent"):
        discrepancy.analyze({**p, "bounds": [[0.0, 1.0], [2.0, 2.0]]}, X, Y)
    S = discrepancy.analyze(p, X, Y)
    assert abs(S["s_discrepancy"].sum() - 1.0) < 1e-14


def test_problem_spec_and_cli(tmp_path, capsys):
    from salib_b200.util import ProblemSpec
    from salib_b200.sample import sobol as sampler
    from salib_b200.scripts import salib as cli
    from salib_b200.test_functions import Ishigami

    sp = ProblemSpec({"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3})
    sp.sample(sampler.sample, 128, calc_second_order=False, scramble=False).evaluate(Ishigami.evaluate)
    sp.analyze_discrepancy(method="CD")
    s = sp.analysis["s_discrepancy"]
    assert s.shape == (3,) and abs(s.sum() - 1.0) < 1e-14
    pf = tmp_path / "p.txt"
    pf.write_text("x1 -3.14159265359 3.14159265359\nx2 -3.14159265359 3.14159265359\nx3 -3.14159265359 3.14159265359\n")
    np.savetxt(tmp_path / "X.txt", sp.samples, fmt="%.8e")
    np.savetxt(tmp_path / "Y.txt", sp.results, fmt="%.8e")
    cli.main(["analyze", "discrepancy", "-p", str(pf), "-X", str(tmp_path / "X.txt"), "-Y", str(tmp_path / "Y.txt")])
    out = capsys.readouterr().out
    assert "s_discrepancy" in out and "x3" in out
