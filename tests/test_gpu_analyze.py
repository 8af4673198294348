"""GPU parity: sobol.analyze vs the frozen reference outputs and the oracle.

Tolerances (float64): rtol 1e-10 as BASELINE north_star states, plus atol 1e-13 for indices that are
differences of O(1) quantities (S2 ~ 5e-4, S1[x3] ~ -0.013)."""
import warnings

import numpy as np
import pytest

from conftest import golden, problem_from

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-10, 1e-13
ANALYZE_CASES = ["ishigami_n1024_r100", "ishigami_n256_c1_r50", "g8_n512_r64", "g8_n512_c1_r64",
                 "g16_n256_r32", "grp3_n512_r40", "g64_n128_r16"]
ISHI = {"num_vars": 3, "names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3}


def _cmp(S, ref, D, c2, keep=True, rtol=RTOL, atol=ATOL):
    keys = ["S1", "S1_conf", "ST", "ST_conf"] + (["S1_conf_all", "ST_conf_all"] if keep else [])
    for k in keys:
        np.testing.assert_allclose(S[k], ref[k], rtol=rtol, atol=atol, err_msg=k)
    if c2:
        iu, il = np.triu_indices(D, 1), np.tril_indices(D)
        for k in ("S2", "S2_conf"):
            np.testing.assert_allclose(S[k][iu], np.asarray(ref[k])[iu], rtol=rtol, atol=atol, err_msg=k)
            assert np.isnan(S[k][il]).all()


@pytest.mark.parametrize("name", ANALYZE_CASES)
def test_matches_frozen_reference(name):
    """Drop-in call, numpy resample indices from the same seed -> same dict as the reference."""
    from salib_b200.analyze import sobol

    z = golden("analyze_" + name)
    p = problem_from(z)
    D = z["S1"].shape[0]
    S = sobol.analyze(p, z["Y"], calc_second_order=bool(z["c2"]), num_resamples=int(z["R"]),
                      conf_level=float(z["conf"]), keep_resamples=True, seed=int(z["seed"]))
    _cmp(S, z, D, bool(z["c2"]))
    assert S.problem is p
    dfs = S.to_df()
    assert len(dfs) == (3 if bool(z["c2"]) else 2)


@pytest.mark.parametrize("name,algo", [("g8_n512_r64", "generic"), ("g16_n256_r32", "generic"),
                                       ("g8_n512_c1_r64", "generic"), ("g64_n128_r16", "generic"),
                                       ("grp3_n512_r40", "generic"), ("g64_n128_r16", "dmma"),
                                       ("g16_n256_r32", "dmma"), ("g8_n512_r64", "dmma"),
                                       ("ishigami_n1024_r100", "dmma"), ("g8_n512_r64", "tile8"),
                                       ("g16_n256_r32", "tile8"), ("ishigami_n1024_r100", "tile8"),
                                       ("grp3_n512_r40", "tile8"), ("g64_n128_r16", "gemv")])
def test_generic_kernel_matches_frozen_reference(name, algo):
    """The fallback kernel is an independent implementation of the same sums; "dmma" is the experimental
    opt-in FP64 tensor-core variant of the tile kernel."""
    from salib_b200.analyze import sobol

    z = golden("analyze_" + name)
    D = z["S1"].shape[0]
    S = sobol.analyze_device(problem_from(z), z["Y"], calc_second_order=bool(z["c2"]), conf_level=float(z["conf"]),
                             keep_resamples=True, r=z["r"].astype(np.int64), algo=algo)
    _cmp(S, z, D, bool(z["c2"]))


def test_cli_golden_table():
    """reference tests/test_cli_analyze.py:250-275 -- the printed table, 6 decimals, R=1000, seed=100."""
    from salib_b200.analyze import sobol

    z = golden("analyze_cli_golden")
    S = sobol.analyze(dict(ISHI), z["Y"], calc_second_order=True, num_resamples=1000, conf_level=0.95, seed=100)
    fmt = lambda v: ["%.6f" % x for x in np.asarray(v).ravel()]  # noqa: E731
    assert fmt(S["ST"]) == ["0.557947", "0.442189", "0.241402"]
    assert fmt(S["ST_conf"]) == ["0.085851", "0.041396", "0.028607"]
    assert fmt(S["S1"]) == ["0.310576", "0.443653", "-0.012962"]
    assert fmt(S["S1_conf"]) == ["0.059615", "0.053436", "0.053891"]
    iu = np.triu_indices(3, 1)
    assert fmt(S["S2"][iu]) == ["-0.014397", "0.246231", "0.000539"]
    assert fmt(S["S2_conf"][iu]) == ["0.083679", "0.103117", "0.064169"]
    _cmp(S, z, 3, True, keep=False)


@pytest.mark.parametrize("D,N,c2,R", [(3, 100, True, 10), (5, 1000, False, 37), (9, 333, True, 20),
                                      (24, 200, True, 12), (40, 96, True, 9), (64, 300, True, 40),
                                      (70, 64, True, 6), (70, 64, False, 6), (20, 5000, False, 50),
                                      (8, 20000, True, 8), (100, 300, True, 5), (130, 96, True, 4),
                                      (200, 70, True, 3), (95, 128, False, 7), (300, 40, False, 3),
                                      (260, 40, True, 2)])
def test_matches_oracle_over_shapes(D, N, c2, R):
    """Ragged N (not a multiple of 32), Dg not a multiple of 8, Dg > 64 (tile groups, 16/8-row ring tiles;
    chunked feature matrix of the GEMV kernel), row splits.  Second order with Dg <= 40 defaults to the
    pair-feature GEMV; the register-tile kernel is run on the same inputs as well."""
    from oracle import analyze as o_analyze
    from salib_b200.analyze import sobol

    rng = np.random.default_rng(D * 7 + N)
    step = 2 * D + 2 if c2 else D + 2
    base = rng.normal(size=(N, 1))
    Y = (base + 0.5 * rng.normal(size=(N, step)) * rng.uniform(0.2, 2.0, size=(1, step))).ravel() * 3.0 + 11.0
    r = rng.integers(N, size=(N, R))
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0, 1]] * D}
    S = sobol.analyze_device(p, Y, calc_second_order=c2, keep_resamples=True, r=r)
    So = o_analyze.analyze(D, Y, c2, R, keep_resamples=True, r=r)
    _cmp(S, So, D, c2, rtol=1e-9, atol=1e-12)
    if c2 and D <= 100:
        other = "tile8" if D <= sobol.PAIR_GEMV_MAX_DG else "gemv"
        S2 = sobol.analyze_device(p, Y, calc_second_order=c2, keep_resamples=True, r=r, algo=other)
        _cmp(S2, So, D, c2, rtol=1e-9, atol=1e-12)


def test_estimator_functions():
    """first_order / total_order / second_order / separate_output_values as standalone calls
    (reference :209-246,267-279; used by tests/test_sobol.py:96-145 and other analyzers)."""
    from oracle import analyze as o
    from salib_b200.analyze import sobol

    z = golden("analyze_g8_n512_r64")
    Y, D, N = z["Y"], 8, 512
    A, B, AB, BA = sobol.separate_output_values(Y, D, N, True)
    Ao, Bo, ABo, BAo = o.split(Y, D, N, True)
    assert np.array_equal(A, Ao) and np.array_equal(AB, ABo) and np.array_equal(BA, BAo) and np.array_equal(B, Bo)
    np.testing.assert_allclose(sobol.first_order(A, AB[:, 2], B), o.first_order(A, AB[:, 2], B), rtol=1e-11)
    np.testing.assert_allclose(sobol.total_order(A, AB[:, 0], B), o.total_order(A, AB[:, 0], B), rtol=1e-11)
    np.testing.assert_allclose(sobol.second_order(A, AB[:, 0], AB[:, 1], BA[:, 0], B),
                               o.second_order(A, AB[:, 0], AB[:, 1], BA[:, 0], B), rtol=1e-9, atol=1e-13)
    r = z["r"].astype(np.int64)[:, :3]
    np.testing.assert_allclose(sobol.first_order(A[r], AB[r, 1], B[r]), o.first_order(A[r], AB[r, 1], B[r]), rtol=1e-11)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert sobol.first_order(np.ones(64), np.arange(64.0), np.ones(64)) == 0.0
    assert any("Constant values" in str(x.message) for x in w)


def test_device_resident_input_and_seed_zero():
    """Y may already live on the GPU; seed=0 is falsy -> legacy global RNG like the reference (:113)."""
    import torch

    from salib_b200.analyze import sobol

    z = golden("analyze_g8_n512_r64")
    p = problem_from(z)
    Yd = torch.from_numpy(z["Y"]).cuda()
    S = sobol.analyze(p, Yd, num_resamples=64, keep_resamples=True, seed=int(z["seed"]))
    _cmp(S, z, 8, True)
    np.random.seed(123)
    S0 = sobol.analyze(p, z["Y"], num_resamples=16, seed=0)
    np.random.seed(123)
    r = np.random.randint(512, size=(512, 16))
    S1 = sobol.analyze_device(p, z["Y"], r=r)
    assert np.array_equal(S0["S1_conf"], S1["S1_conf"])


def test_constant_output():
    """reference tests/test_sobol.py:170-224: constant Y -> all indices and CIs exactly 0, with a warning."""
    from salib_b200.analyze import sobol

    for p in (dict(ISHI), {**ISHI, "groups": ["A", "B", "A"]}):
        Dg = 2 if "groups" in p else 3
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            S = sobol.analyze(p, np.full(1024 * (2 * Dg + 2), 0.4), num_resamples=10, seed=1)
        assert any("Constant values" in str(x.message) for x in w)
        for k in ("S1", "S1_conf", "ST", "ST_conf"):
            assert S[k].shape == (Dg,) and not S[k].any()
        assert S["S2"][0, 1] == 0.0 and np.isnan(S["S2_conf"][0, 1])


def test_philox_resampling_is_statistically_equivalent():
    """Device Philox indices: not the same stream as NumPy, so compare the CI widths within the
    bootstrap's own sampling error (~ conf / sqrt(2(R-1))) and the point estimates exactly."""
    from salib_b200.analyze import sobol

    z = golden("analyze_cli_golden")
    R = 2000
    Sn = sobol.analyze_device(dict(ISHI), z["Y"], num_resamples=R, seed=100, resample="numpy")
    Sp = sobol.analyze_device(dict(ISHI), z["Y"], num_resamples=R, seed=100, resample="philox", keep_resamples=True)
    Sp2 = sobol.analyze_device(dict(ISHI), z["Y"], num_resamples=R, seed=100, resample="philox")
    Sq = sobol.analyze_device(dict(ISHI), z["Y"], num_resamples=R, seed=101, resample="philox")
    for k in ("S1", "ST"):
        assert np.array_equal(Sn[k], Sp[k])
    tol = 5.0 / np.sqrt(2 * (R - 1))
    for k in ("S1_conf", "ST_conf"):
        np.testing.assert_allclose(Sp[k], Sn[k], rtol=tol)
        assert np.array_equal(Sp[k], Sp2[k])          # same seed -> same stream
        assert not np.array_equal(Sp[k], Sq[k])       # different seed -> different stream
    iu = np.triu_indices(3, 1)
    np.testing.assert_allclose(Sp["S2_conf"][iu], Sn["S2_conf"][iu], rtol=tol)
    # the resampled estimates are centred on the point estimate
    se = Sp["S1_conf_all"].std(axis=0, ddof=1) / np.sqrt(R)
    assert np.all(np.abs(Sp["S1_conf_all"].mean(axis=0) - Sp["S1"]) < 6 * se + 5e-3)


def test_multiplicities_sum_to_n():
    """Every Philox resample draws exactly N rows; index-built multiplicities equal np.bincount."""
    import torch

    from salib_b200 import _lib, _runtime as rt

    N, R = 1000, 7
    Npad = int(_lib.lib().sa_pad_rows(N))
    w = rt.empty((R * Npad,), torch.uint8)
    _lib.call("sa_counts_philox", rt.stream_ptr(), 42, N, 3, R, rt.ptr(w), None)
    wc = w.reshape(R, Npad).cpu().numpy()
    assert (wc.sum(axis=1) == N).all() and not wc[:, N:].any()
    assert len({tuple(row) for row in wc}) == R
    # global resample id keys the stream: resample 5 is the same whether it is drawn first or third
    w1 = rt.empty((Npad,), torch.uint8)
    _lib.call("sa_counts_philox", rt.stream_ptr(), 42, N, 5, 1, rt.ptr(w1), None)
    assert np.array_equal(w1.cpu().numpy(), wc[2])
    r = np.random.default_rng(0).integers(N, size=(N, R))
    rd = rt.to_device(r, torch.int64)
    flag = rt.zeros((1,), torch.int32)
    _lib.call("sa_counts_from_indices", rt.stream_ptr(), rt.ptr(rd), N, R, 2, 4, rt.ptr(w), rt.ptr(flag))
    wc = w.reshape(R, Npad)[:4].cpu().numpy()
    for c in range(4):
        assert np.array_equal(wc[c, :N], np.bincount(r[:, 2 + c], minlength=N))
    assert int(flag.item()) == 0


def test_problem_spec_chain_on_device():
    """ProblemSpec.sample_sobol -> evaluate -> analyze_sobol (reference examples/Problem/problem_spec.py)."""
    from oracle import analyze as o_analyze
    from oracle import saltelli as o_saltelli
    from oracle import test_functions as o_tf
    from salib_b200.test_functions import Ishigami
    from salib_b200.util import ProblemSpec

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp = ProblemSpec({"names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3, "outputs": ["Y"]})
        sp.sample_sobol(512, scramble=False, skip_values=512).evaluate(Ishigami.evaluate).analyze_sobol(
            num_resamples=50, seed=101)
        Xo = o_saltelli.sample(dict(ISHI), 512, True, skip_values=512, closed_form=True)
        assert np.array_equal(sp.samples, Xo)
        So = o_analyze.analyze(3, o_tf.ishigami(Xo), True, 50, seed=101)
        _cmp(sp.analysis, So, 3, True, keep=False, rtol=1e-9, atol=1e-12)
        total, first, second = sp.to_df()
        assert list(first.index) == ["x1", "x2", "x3"] and "Analysis" in str(sp)
        # device-resident chain and two outputs
        sp2 = ProblemSpec({"names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3})
        sp2.sample_sobol(256, scramble=False, device=True)
        assert sp2.samples.is_cuda
        sp2.evaluate(lambda X: __import__("torch").stack([Ishigami.evaluate(X), 2 * Ishigami.evaluate(X) + 1], dim=1))
        sp2.analyze_sobol(num_resamples=20, seed=3)
        assert set(sp2.analysis) == {"Y1", "Y2"}
        np.testing.assert_allclose(sp2.analysis["Y1"]["S1"], sp2.analysis["Y2"]["S1"], rtol=1e-9, atol=1e-12)
        # the shared-multiplicity pass equals one reference-style call per output column
        from salib_b200.analyze import sobol as an
        for i, name in enumerate(("Y1", "Y2")):
            Si = an.analyze(sp2, sp2.results[:, i], num_resamples=20, seed=3)
            for k in ("S1", "S1_conf", "ST", "ST_conf"):
                np.testing.assert_allclose(sp2.analysis[name][k], Si[k], rtol=1e-12, atol=1e-14)


def test_cli_golden_files(tmp_path, capsys):
    """reference tests/test_cli_analyze.py:250-275 end to end through the command line: the Saltelli file is
    byte-identical to the reference's tests/data/model_input.txt and the printed table equals its golden string."""
    import re

    from salib_b200.scripts import salib as cli
    from salib_b200.test_functions import Ishigami

    params = tmp_path / "Ishigami.txt"
    params.write_text("x1 -3.14159265359 3.14159265359\nx2 -3.14159265359 3.14159265359\nx3 -3.14159265359 3.14159265359\n")
    xin, yout = tmp_path / "model_input.txt", tmp_path / "model_output.txt"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cli.main(f"sample saltelli -p {params} -o {xin} -n 1024 --precision 8 --max-order 2 --skip-values 2048".split())
    ref = tmp_path / "ref.txt"
    np.savetxt(ref, golden("model_input_8e.npy"), delimiter=" ", fmt="%.8e")
    assert xin.read_bytes() == ref.read_bytes()
    np.savetxt(yout, Ishigami.evaluate(np.loadtxt(xin)))
    capsys.readouterr()
    cli.main(f"analyze sobol -p {params} -Y {yout} -c 0 --max-order 2 -r 1000 --seed=100".split())
    out = re.sub(r"[\n\t\s]*", "", capsys.readouterr().out)
    assert out == ("STST_confx10.5579470.085851x20.4421890.041396x30.2414020.028607S1S1_confx10.3105760.059615"
                   "x20.4436530.053436x3-0.0129620.053891S2S2_conf(x1,x2)-0.0143970.083679(x1,x3)0.2462310.103117"
                   "(x2,x3)0.0005390.064169")


@pytest.mark.parametrize("D,N,c2,R", [(2, 2, True, 3), (2, 1, False, 2), (3, 31, True, 1), (2, 33, False, 1),
                                      (1, 40, True, 5), (9, 5, True, 4)])
def test_edge_shapes_match_oracle(D, N, c2, R):
    """Smallest problems: N below one ring tile, a single resample (std(ddof=1) of one value is NaN like
    NumPy's), one factor (no pairs), N = 1."""
    from oracle import analyze as o_analyze
    from salib_b200.analyze import sobol

    rng = np.random.default_rng(100 * D + N)
    step = 2 * D + 2 if c2 else D + 2
    Y = rng.normal(size=N * step) * 2.0 + 1.0
    r = rng.integers(N, size=(N, R))
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0, 1]] * D}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        So = o_analyze.analyze(D, Y, c2, R, keep_resamples=True, r=r)
        runs = [sobol.analyze_device(p, Y, calc_second_order=c2, keep_resamples=True, r=r, algo=a)
                for a in (("auto", "tile8") if c2 else ("auto",))]
    for S in runs:
        for k in ("S1", "ST", "S1_conf_all", "ST_conf_all", "S1_conf", "ST_conf"):
            np.testing.assert_allclose(S[k], So[k], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg=k)
        if c2 and D > 1:
            iu = np.triu_indices(D, 1)
            np.testing.assert_allclose(S["S2"][iu], So["S2"][iu], rtol=1e-9, atol=1e-12, equal_nan=True)
            np.testing.assert_allclose(S["S2_conf"][iu], So["S2_conf"][iu], rtol=1e-9, atol=1e-12, equal_nan=True)


def test_row_window_multiplicities_tile_the_full_ones():
    """sa_counts_*_rows over windows that tile [0, N) = the unsharded multiplicities, for both sources."""
    import torch

    from salib_b200 import _lib, _runtime as rt
    from salib_b200.analyze.sobol import _Resampler

    N, R = 1000, 9
    L = _lib.lib()
    Npad = int(L.sa_pad_rows(N))
    for rs in (_Resampler(N, R, 11, "philox"),
               _Resampler.from_indices(np.random.default_rng(3).integers(N, size=(N, R)))):
        full = rt.empty((R * Npad,), torch.uint8)
        rs.counts(0, R, full)
        full = full.reshape(R, Npad)[:, :N].cpu().numpy()
        assert (full.sum(axis=1) == N).all()
        got = np.zeros_like(full)
        for lo, hi in ((0, 300), (300, 301), (301, 777), (777, 1000)):
            npad = int(L.sa_pad_rows(hi - lo))
            w = rt.empty((R * npad,), torch.uint8)
            rs.counts(0, R, w, rows=(lo, hi - lo))
            w = w.reshape(R, npad).cpu().numpy()
            assert not w[:, hi - lo:].any()
            got[:, lo:hi] = w[:, :hi - lo]
        np.testing.assert_array_equal(got, full)
    with pytest.raises(_lib.SalibB200Error):
        rs.counts(0, R, rt.empty((R * Npad,), torch.uint8), rows=(900, 200))


@pytest.mark.parametrize("name,cuts", [("g64_n128_r16", (0, 50, 128)), ("g8_n512_r64", (0, 100, 101, 512)),
                                       ("g8_n512_c1_r64", (0, 333, 512)), ("ishigami_n1024_r100", (0, 512, 1024)),
                                       ("grp3_n512_r40", (0, 64, 512))])
def test_row_sharded_matches_frozen_reference(name, cuts):
    """Y partitioned by base-row range: per-shard partial sums added up -> the reference's dict."""
    from salib_b200.analyze import sobol

    z = golden("analyze_" + name)
    p = problem_from(z)
    D, c2 = z["S1"].shape[0], bool(z["c2"])
    Y = np.asarray(z["Y"], dtype=np.float64)
    step = Y.size // (cuts[-1])
    shards = [(lo, Y[lo * step:hi * step]) for lo, hi in zip(cuts, cuts[1:])]
    S = sobol.analyze_row_sharded(p, shards, cuts[-1], calc_second_order=c2, conf_level=float(z["conf"]),
                                  keep_resamples=True, r=z["r"].astype(np.int64))
    _cmp(S, z, D, c2)
    # the one-call form (one shard per rank; a single process holds everything)
    S1 = sobol.analyze_device(p, Y, calc_second_order=c2, conf_level=float(z["conf"]), keep_resamples=True,
                              r=z["r"].astype(np.int64), shard="rows")
    _cmp(S1, z, D, c2)


def test_row_sharded_philox_equals_unsharded():
    """Device Philox draws do not depend on the row partition: same key -> same CIs to rounding."""
    from salib_b200.analyze import sobol
    from salib_b200.sample import sobol as sampler
    from salib_b200.test_functions import Sobol_G

    D, N = 16, 4096
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    X = sampler.sample(p, N, scramble=False)
    Y = Sobol_G.evaluate(X, a=np.linspace(0.0, 20.0, D))
    step = 2 * D + 2
    ref = sobol.analyze_device(p, Y, num_resamples=300, seed=7, resample="philox", keep_resamples=True)
    cuts = (0, 1000, 2600, 4096)
    S = sobol.analyze_row_sharded(p, [(lo, Y[lo * step:hi * step]) for lo, hi in zip(cuts, cuts[1:])], N,
                                  num_resamples=300, seed=7, resample="philox", keep_resamples=True)
    _cmp(S, ref, D, True)
    with pytest.raises(RuntimeError):
        sobol.analyze_row_sharded(p, [(0, Y[:1000 * step])], N, num_resamples=10, seed=7)


@pytest.mark.parametrize("N", [1 << 15, 1 << 16, 1 << 18, 100000, 99999, 40000 * 4])
def test_philox_multiplicities_are_multinomial(N):
    """Shared-memory window scheme (binomial split tree + in-window draws; N = 99999 takes the global-atomic
    fallback): every resample has exactly N draws, and the multiplicities behave like Multinomial(N, 1/N):
    mean 1, variance 1 - 1/N, P(w = k) = Poisson(1) to sampling error, no structure across windows."""
    import torch

    from salib_b200 import _lib, _runtime as rt

    R = 24
    Npad = int(_lib.lib().sa_pad_rows(N))
    w = rt.empty((R * Npad,), torch.uint8)
    _lib.call("sa_counts_philox", rt.stream_ptr(), 1234, N, 0, R, rt.ptr(w), None)
    wc = w.reshape(R, Npad).cpu().numpy()
    assert (wc[:, :N].sum(axis=1, dtype=np.int64) == N).all() and not wc[:, N:].any()
    x = wc[:, :N].astype(np.float64)
    M = R * N
    assert abs(x.var() - (1.0 - 1.0 / N)) < 6.0 * np.sqrt(2.0 / M) + 6.0 / np.sqrt(M)   # var of Poisson-like counts
    from math import exp, factorial
    for k in range(5):
        pk = exp(-1.0) / factorial(k)
        assert abs((x == k).mean() - pk) < 6.0 * np.sqrt(pk * (1 - pk) / M) + 2.0 / N
    # draws per block of 4096 rows ~ Binomial(N, 4096/N): chi-square over blocks and resamples
    nb = N // 4096
    blocks = x[:, :nb * 4096].reshape(R, nb, 4096).sum(axis=2)
    chi = ((blocks - 4096.0) ** 2 / 4096.0).sum()
    dof = R * nb
    assert abs(chi - dof) < 6.0 * np.sqrt(2.0 * dof)
    # adjacent rows are uncorrelated
    corr = np.mean((x[:, :-1] - 1.0) * (x[:, 1:] - 1.0))
    assert abs(corr) < 6.0 / np.sqrt(M)
    # determinism and independence of the batch: resample 7 drawn alone equals row 7 of the batch
    w1 = rt.empty((Npad,), torch.uint8)
    _lib.call("sa_counts_philox", rt.stream_ptr(), 1234, N, 7, 1, rt.ptr(w1), None)
    assert np.array_equal(w1.cpu().numpy(), wc[7])
