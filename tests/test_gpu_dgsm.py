"""GPU parity of the DGSM path (consumers of the same Sobol' generator and bootstrap):
finite_diff.sample bit-exact, dgsm.analyze to 1e-10 with the reference's own NumPy resample draws."""
import numpy as np
import pytest

from conftest import golden, problem_from

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["dgsm_ishigami_n256", "dgsm_g8_n128"])
def test_finite_diff_sample_matches_frozen_reference(name):
    from salib_b200.sample import finite_diff

    z = golden(name)
    X = finite_diff.sample(problem_from(z), int(z["N"]), delta=float(z["delta"]), skip_values=int(z["skip"]))
    assert X.shape == z["X"].shape and np.array_equal(X, z["X"])


@pytest.mark.parametrize("name", ["dgsm_ishigami_n256", "dgsm_g8_n128"])
def test_dgsm_analyze_matches_frozen_reference(name):
    from salib_b200.analyze import dgsm

    z = golden(name)
    p = problem_from(z)
    S = dgsm.analyze(p, z["X"], z["Y"], num_resamples=int(z["R"]), conf_level=0.95, seed=int(z["seed"]))
    for k in ("vi", "dgsm", "dgsm_conf"):
        np.testing.assert_allclose(S[k], z[k], rtol=1e-10, err_msg=k)
    np.testing.assert_allclose(S["vi_std"], z["vi_std"], rtol=1e-11, err_msg="vi_std")  # two-pass, like np.std
    assert S["names"] == p["names"] and S.to_df().shape == (p["num_vars"], 4)
    with pytest.raises(RuntimeError):
        dgsm.analyze(p, z["X"], z["Y"][:-1])
    with pytest.raises(RuntimeError):
        dgsm.analyze(p, z["X"], z["Y"], conf_level=0.0)


def test_dgsm_philox_and_problem_spec():
    """Device resampling shares one multiplicity set between the factors: CI widths agree with the NumPy
    mode within the bootstrap's own error; ProblemSpec chaining (sample_finite_diff -> analyze_dgsm)."""
    from salib_b200.analyze import dgsm
    from salib_b200.test_functions import Ishigami
    from salib_b200.util import ProblemSpec

    sp = ProblemSpec({"names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3})
    sp.sample_finite_diff(2048, delta=0.001).evaluate(Ishigami.evaluate)
    old = dgsm.RESAMPLE_MODE
    try:
        dgsm.RESAMPLE_MODE = "numpy"
        Sn = dgsm.analyze(sp, sp.samples, sp.results, num_resamples=400, seed=5)
        dgsm.RESAMPLE_MODE = "philox"
        sp.analyze_dgsm(num_resamples=400, seed=5)
    finally:
        dgsm.RESAMPLE_MODE = old
    Sp = sp.analysis
    np.testing.assert_allclose(Sp["dgsm"], Sn["dgsm"], rtol=1e-12)
    np.testing.assert_allclose(Sp["dgsm_conf"], Sn["dgsm_conf"], rtol=5.0 / np.sqrt(2 * 399))
    # reference tests/test_regression.py:465-474 (N = 10000): dgsm ~ [2.229, 7.066, 3.180]
    np.testing.assert_allclose(Sp["dgsm"], [2.229, 7.066, 3.180], atol=5e-2, rtol=1e-1)


def test_dgsm_cli(tmp_path, capsys):
    """reference tests/test_cli_analyze.py:77-105: finite_diff -> Ishigami -> `salib analyze dgsm`; the expected
    values lie within the printed confidence intervals."""
    import pandas as pd
    from io import StringIO

    from salib_b200.scripts import salib as cli
    from salib_b200.test_functions import Ishigami

    params = tmp_path / "Ishigami.txt"
    params.write_text("x1 -3.14159265359 3.14159265359\nx2 -3.14159265359 3.14159265359\nx3 -3.14159265359 3.14159265359\n")
    xin, yout = tmp_path / "x.txt", tmp_path / "y.txt"
    cli.main(f"sample finite_diff -p {params} -o {xin} -d 0.001 --precision=8 -n 1000 -s 100".split())
    np.savetxt(yout, Ishigami.evaluate(np.loadtxt(xin)))
    capsys.readouterr()
    cli.main(f"analyze dgsm -p {params} -X {xin} -Y {yout} -c 0 -r 1000 --seed=100".split())
    df = pd.read_csv(StringIO(capsys.readouterr().out), sep=r"\s+")
    expected = np.array([2.207554, 7.092019, 3.238259])
    assert ((df["dgsm"] - df["dgsm_conf"] <= expected) & (expected <= df["dgsm"] + df["dgsm_conf"])).all()


def test_dgsm_vi_std_of_a_nearly_linear_model():
    """np.std(dfdx) (analyze/dgsm.py:109) for derivatives that are constant up to rounding: a one-pass
    E[g^4] - E[g^2]^2 returns cancellation noise there; the two-pass form follows the oracle."""
    from oracle import dgsm as o_dgsm
    from salib_b200.analyze import dgsm
    from salib_b200.sample import finite_diff

    p = {"num_vars": 3, "names": ["a", "b", "c"], "bounds": [[0.5, 2.0], [1.0, 3.0], [0.1, 0.9]]}
    X = finite_diff.sample(p, 512, delta=0.01, skip_values=64)
    Y = 1000.0 + 3.0 * X[:, 0] - 2.0 * X[:, 1] + 1e-7 * X[:, 2] ** 2
    S = dgsm.analyze(p, X, Y, num_resamples=10, seed=1)
    ref = o_dgsm.dgsm_analyze(p, X, Y, num_resamples=10)
    np.testing.assert_allclose(S["vi"], ref["vi"], rtol=1e-12)
    # the spread is pure rounding noise of the finite differences (~1e-10 of vi): same noise, to a few digits
    np.testing.assert_allclose(S["vi_std"], ref["vi_std"], rtol=1e-6, atol=1e-30)
