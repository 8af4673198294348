"""GPU parity: device "%.Pe" text output vs np.savetxt -- byte-identical."""
import io

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _ref(X, delimiter=" ", precision=8):
    buf = io.BytesIO()
    np.savetxt(buf, X, delimiter=delimiter, fmt="%." + str(precision) + "e")
    return buf.getvalue()


def _mine(X, delimiter=" ", precision=8):
    from salib_b200.util.textio import savetxt

    buf = io.BytesIO()
    savetxt(buf, X, delimiter=delimiter, precision=precision)
    return buf.getvalue()


def test_reference_golden_matrix_round_trips():
    """The reference's golden Saltelli matrix (tests/data/model_input.txt, %.8e): same text as NumPy."""
    X = golden("model_input_8e.npy")
    assert _mine(X) == _ref(X)
    from salib_b200.sample import saltelli

    p = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-3.14159265359, 3.14159265359]] * 3}
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Xs = saltelli.sample(p, 1024, skip_values=2048)
    txt = _mine(Xs)
    assert txt == _ref(Xs)
    np.testing.assert_array_equal(np.loadtxt(io.BytesIO(txt)), X)   # the file the reference's tests pin


@pytest.mark.parametrize("precision", [0, 1, 8, 15, 17])
def test_random_and_extreme_values(precision):
    rng = np.random.default_rng(precision)
    bits = rng.integers(0, 2 ** 64, size=20000, dtype=np.uint64)
    wild = bits.view(np.float64)                                   # every exponent, nan/inf included
    vals = np.concatenate([
        wild, rng.normal(size=5000), rng.uniform(-4, 4, 5000), 10.0 ** rng.uniform(-30, 30, 5000),
        np.array([0.0, -0.0, 1.0, -1.0, 0.5, 1.5, 2.5, 0.125, 0.375, 9.5, 99.5, 999999999.5, 1e22, 1e23, 5e-324,
                  2.2250738585072014e-308, 1.7976931348623157e308, np.inf, -np.inf, np.nan, 9.999999995,
                  9.9999999949999, 0.99999999995, 1e100, 9.99999999999e99, 1e-100, 123456789.0, 1234567890.0]),
        np.arange(1, 2049) / 1024.0, np.arange(1, 1025) * 0.5, 2.0 ** np.arange(-60, 60)])
    n = (vals.size // 7) * 7
    X = vals[:n].reshape(-1, 7)
    assert _mine(X, precision=precision) == _ref(X, precision=precision)


def test_shapes_and_delimiters(tmp_path):
    from salib_b200.util.textio import savetxt

    rng = np.random.default_rng(1)
    for shape in [(1, 1), (5,), (3, 70), (1000, 3), (2, 33)]:
        X = rng.normal(size=shape)
        for d in (" ", ",", "\t"):
            assert _mine(X, delimiter=d) == _ref(X, delimiter=d)
    import torch

    Xd = torch.from_numpy(rng.normal(size=(100, 8))).cuda()
    path = tmp_path / "x.txt"
    savetxt(str(path), Xd, delimiter=",", precision=6)
    assert path.read_bytes() == _ref(Xd.cpu().numpy(), delimiter=",", precision=6)
    with pytest.raises(ValueError):
        savetxt(io.BytesIO(), np.zeros((2, 2)), delimiter="::")


def test_cli_sample_writes_the_same_file(tmp_path):
    from salib_b200.scripts import salib as cli

    pf = tmp_path / "p.txt"
    pf.write_text("x1 -3.14159265359 3.14159265359\nx2 -3.14159265359 3.14159265359\nx3 -3.14159265359 3.14159265359\n")
    out = tmp_path / "X.txt"
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cli.main(["sample", "saltelli", "-n", "1024", "-p", str(pf), "-o", str(out), "--skip-values", "2048"])
    X = golden("model_input_8e.npy")
    np.testing.assert_array_equal(np.loadtxt(out), X)
    from salib_b200.sample import saltelli
    from salib_b200.util import read_param_file

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Xs = saltelli.sample(read_param_file(str(pf)), 1024, skip_values=2048)
    assert out.read_bytes() == _ref(Xs)


# ---- text input ---------------------------------------------------------------------------------------------
def _parse_tokens(tokens):
    from salib_b200.util.textio import parse_text

    return parse_text(("\n".join(tokens) + "\n").encode()).cpu().numpy().reshape(-1)


def _same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | nan))


def test_token_conversion_equals_python_float():
    """Every token converts to exactly the double Python's float() gives (what np.loadtxt uses)."""
    rng = np.random.default_rng(7)
    wild = rng.integers(0, 2 ** 64, size=20000, dtype=np.uint64).view(np.float64)
    wild = wild[np.isfinite(wild)]
    toks = [repr(float(v)) for v in wild]                                        # 17 significant digits
    toks += ["%.8e" % v for v in rng.normal(size=5000)]                           # what the CLI writes
    toks += ["%.8e" % v for v in wild[:5000]]
    toks += ["%.20e" % v for v in wild[:3000]]                                    # 21 digits: big-integer path
    toks += ["%.3f" % v for v in rng.uniform(-1e6, 1e6, 3000)]
    toks += [str(int(v)) for v in rng.integers(-10 ** 18, 10 ** 18, 2000)]
    toks += ["%d" % (2 ** 53 + k) for k in range(0, 40)]                          # ties above 2^53
    toks += ["9007199254740993", "9007199254740992.5", "9007199254740993.00000000000000000000001",
             "1e23", "8.5e22", "1e22", "123456789012345678901234567890e-10", "0.1", ".5", "5.", "+1.5", "-0",
             "0.0", "-0.0", "1e400", "-1e400", "1e-400", "4.9406564584124654e-324", "2.4703282292062327e-324",
             "2.4703282292062328e-324", "2.2250738585072011e-308", "2.2250738585072014e-308",
             "1.7976931348623157e308", "1.7976931348623159e308", "1.797693134862315807e308",
             "nan", "NaN", "inf", "-inf", "Infinity", "-INFINITY", "1E5", "1e+5", "1e-05",
             "0.000000000000000000000000000000000000001234", "1234567890123456789012345678901234567890e-1",
             "179769313486231570000000000000000000000e270", "0.00000000e+00", "1.00000000e+00"]
    got = _parse_tokens(toks)
    ref = np.array([float(t) for t in toks])
    bad = [(t, g, r) for t, g, r in zip(toks, got, ref) if not _same_bits([g], [r])]
    assert not bad, bad[:10]
    # documented limit: more than 40 significant digits are reported, not guessed
    with pytest.raises(ValueError, match="40 significant digits"):
        _parse_tokens(["9007199254740993.0000000000000000000000001"])
    assert _parse_tokens(["1." + "0" * 60])[0] == 1.0            # trailing zeros are not significant


def test_loadtxt_structure_matches_numpy(tmp_path):
    from salib_b200.util.textio import loadtxt, savetxt

    rng = np.random.default_rng(2)
    X = rng.normal(size=(257, 5))
    path = tmp_path / "x.txt"
    savetxt(str(path), X)
    np.testing.assert_array_equal(loadtxt(str(path)), np.loadtxt(path))
    np.testing.assert_array_equal(loadtxt(str(path), usecols=(2,)), np.loadtxt(path, usecols=(2,)))
    np.testing.assert_array_equal(loadtxt(str(path), delimiter=" ", usecols=(0,)), np.loadtxt(path, delimiter=" ", usecols=(0,)))
    txt = "# header\n1.0, 2.0 ,3.5\r\n\n  4,5,6  # trailing comment\n7e0,8,9"
    np.testing.assert_array_equal(loadtxt(io.BytesIO(txt.encode()), delimiter=","),
                                  np.loadtxt(io.StringIO(txt), delimiter=","))
    one = "1.5\n2.5\n3.5\n"
    assert loadtxt(io.BytesIO(one.encode())).shape == (3,)
    assert loadtxt(io.BytesIO(one.encode()), ndmin=2).shape == np.loadtxt(io.StringIO(one), ndmin=2).shape
    assert loadtxt(io.BytesIO(b"1 2 3\n"), ndmin=2).shape == (1, 3)
    assert loadtxt(io.BytesIO(b"  1\t 2   3  \n")).tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError, match="columns"):
        loadtxt(io.BytesIO(b"1 2 3\n4 5\n"))
    with pytest.raises(ValueError, match="convert"):
        loadtxt(io.BytesIO(b"1 2 x\n"))
    with pytest.raises(ValueError, match="convert"):
        loadtxt(io.BytesIO(b"1,,3\n"), delimiter=",")
    big = rng.normal(size=(50000, 3))                      # many parse chunks, lines crossing chunk borders
    buf = io.BytesIO()
    np.savetxt(buf, big, fmt="%.17e")
    np.testing.assert_array_equal(loadtxt(io.BytesIO(buf.getvalue())), big)


def test_cli_analyze_reads_with_the_device_parser(tmp_path, capsys):
    """The CLI golden table again (tests/test_cli_analyze.py:250-275), now through loadtxt on the device."""
    from salib_b200.scripts import salib as cli
    from salib_b200.test_functions import Ishigami

    X = golden("model_input_8e.npy")
    pf = tmp_path / "p.txt"
    pf.write_text("x1 -3.14159265359 3.14159265359\nx2 -3.14159265359 3.14159265359\nx3 -3.14159265359 3.14159265359\n")
    np.savetxt(tmp_path / "Y.txt", Ishigami.evaluate(X), fmt="%.8e")
    cli.main(["analyze", "sobol", "-p", str(pf), "-Y", str(tmp_path / "Y.txt"), "-c", "0", "-r", "1000", "-s", "100"])
    out = capsys.readouterr().out
    assert "x1" in out and "S1_conf" in out
