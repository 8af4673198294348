"""GPU parity: Ishigami / Sobol' G evaluation, unfused (X -> Y) and fused (indices -> Y)."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
# float64 transcendental / product rounding differs from NumPy by a few ulp: tolerance, not bit-exact
RTOL, ATOL = 1e-12, 1e-13


def test_ishigami_matches_frozen_reference():
    from salib_b200.test_functions import Ishigami

    z = golden("test_functions")
    np.testing.assert_allclose(Ishigami.evaluate(z["Xi"]), z["Yi"], rtol=RTOL, atol=ATOL)


def test_sobol_g_matches_frozen_reference():
    from salib_b200.test_functions import Sobol_G

    z = golden("test_functions")
    np.testing.assert_allclose(Sobol_G.evaluate(z["Xg"]), z["Yg"], rtol=RTOL)
    np.testing.assert_allclose(Sobol_G.evaluate(z["Xg16"], a=z["a16"]), z["Yg16"], rtol=RTOL)
    np.testing.assert_allclose(Sobol_G.evaluate(z["Xg"], delta=z["delta"], alpha=z["alpha"]), z["Ygda"], rtol=1e-11)
    np.testing.assert_allclose(Sobol_G.evaluate(np.zeros((1, 8))), [4.0583], atol=1e-4)  # reference test_test_functions.py:17-23


def test_sobol_g_error_behaviour():
    from salib_b200.test_functions import Sobol_G

    with pytest.raises(ValueError, match="less than zero"):
        Sobol_G.evaluate(np.full((3, 8), -0.1))
    with pytest.raises(ValueError, match="greater than one"):
        Sobol_G.evaluate(np.full((3, 8), 1.1))
    with pytest.raises(TypeError):
        Sobol_G.evaluate([[0.5] * 8])
    with pytest.raises(ValueError):
        Sobol_G.evaluate(np.zeros((2, 8)), alpha=np.zeros(8))
    with pytest.raises(TypeError):
        Sobol_G.evaluate(np.zeros((2, 8)), delta=[0.1] * 8)


@pytest.mark.parametrize("D,N,c2,groups", [(8, 300, True, None), (8, 300, False, None), (64, 50, True, None),
                                           (5, 128, True, ["a", "b", "a", "c", "b"]), (33, 20, False, None)])
def test_fused_sobol_g_equals_sample_then_evaluate(D, N, c2, groups):
    from oracle import test_functions as o_tf
    from salib_b200.sample import sobol
    from salib_b200.test_functions import Sobol_G

    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    if groups:
        p["groups"] = groups
    a = np.linspace(0.0, 30.0, D)
    plan, first = sobol._plan(p, N, c2, False, 16, None)
    Yf = Sobol_G.evaluate_sample(plan, first, N, a=a).cpu().numpy()
    X = plan.generate(first, N).cpu().numpy()
    assert Yf.shape == (X.shape[0],)
    # the fused kernel multiplies the factors in column order like ndarray.prod on one row
    np.testing.assert_allclose(Yf, o_tf.sobol_g(X, a=a), rtol=1e-14)
    np.testing.assert_allclose(Sobol_G.evaluate(X, a=a), Yf, rtol=RTOL)


def test_fused_ishigami_equals_sample_then_evaluate():
    from oracle import test_functions as o_tf
    from salib_b200.sample import sobol
    from salib_b200.test_functions import Ishigami

    p = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}
    for c2 in (True, False):
        plan, first = sobol._plan(p, 512, c2, False, 512, None)
        Yf = Ishigami.evaluate_sample(plan, first, 512).cpu().numpy()
        X = plan.generate(first, 512).cpu().numpy()
        np.testing.assert_allclose(Yf, o_tf.ishigami(X), rtol=RTOL, atol=ATOL)
