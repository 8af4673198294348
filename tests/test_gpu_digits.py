"""GPU parity of the opt-in exact-integer tensor-core engine (csrc/digits.cu, ``algo="digits"``).

The int8 GEMM must be bit-exact against an integer matmul; the engine must meet the same 1e-10 bar against the
frozen reference outputs as the FP64 engines (tests/test_gpu_analyze.py)."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden, problem_from

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-10, 1e-13


def _gemm(W, Q, ksplit):
    """C[ks][count][ldc] through sa_digit_gemm_i32 for host arrays W uint8 [count, K], Q int8 [ncols, K]."""
    import torch
    from salib_b200 import _lib, _runtime as rt

    count, K = W.shape
    ncols = Q.shape[0]
    pitch = (K + 127) // 128 * 128
    ldc = (ncols + 31) // 32 * 32
    Wd = torch.full((count, pitch), 7, dtype=torch.uint8, device="cuda")      # junk in the pad columns:
    Qd = torch.full((ncols, pitch), -3, dtype=torch.int8, device="cuda")      # TMA must zero-fill beyond K
    Wd[:, :K] = torch.from_numpy(W).cuda()
    Qd[:, :K] = torch.from_numpy(Q).cuda()
    Cd = torch.full((ksplit, count, ldc), -1, dtype=torch.int32, device="cuda")
    _lib.call("sa_digit_gemm_i32", rt.stream_ptr(), rt.ptr(Wd), pitch, count, rt.ptr(Qd), pitch, ncols, K, ksplit,
              rt.ptr(Cd), ldc)
    torch.cuda.synchronize()
    return Cd.cpu().numpy()[:, :, :ncols]


@pytest.mark.parametrize("count,ncols,K,ksplit", [(1, 7, 100, 1), (5, 40, 128, 1), (130, 129, 1000, 1),
                                                  (130, 129, 1000, 3), (300, 600, 5000, 1), (257, 300, 4097, 5),
                                                  (128, 256, 16384, 2), (64, 128, 777, 1), (200, 100, 3000, 2)])
def test_digit_gemm_is_exact(count, ncols, K, ksplit):
    rng = np.random.default_rng(count * 1000 + ncols + K)
    W = rng.integers(0, 256, size=(count, K), dtype=np.uint8)            # full uint8 range: A is unsigned
    Q = rng.integers(-128, 128, size=(ncols, K), dtype=np.int8)          # full int8 range: B is signed
    got = _gemm(W, Q, ksplit)
    ksteps = -(-K // 128)
    per = -(-ksteps // ksplit)
    for ks in range(ksplit):
        lo, hi = ks * per * 128, min(K, (ks + 1) * per * 128)
        want = W[:, lo:hi].astype(np.int64) @ Q[:, lo:hi].astype(np.int64).T
        assert np.array_equal(got[ks].astype(np.int64), want), f"K split {ks}"


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("count,ncols,K,ksplit", [(3, 7, 300, 1), (130, 129, 1000, 2), (300, 600, 5000, 1),
                                                  (513, 1030, 2500, 3)])
def test_digit_gemm_variants_are_exact(variant, count, ncols, K, ksplit, monkeypatch):
    """Every tile shape (one CTA 128 x 128 / 128 x 256, CTA pair 256 x 256 / 256 x 512) on ragged problems."""
    monkeypatch.setenv("SALIB_B200_DIGIT_GEMM", str(variant))
    rng = np.random.default_rng(variant + count + ncols + K)
    W = rng.integers(0, 256, size=(count, K), dtype=np.uint8)
    Q = rng.integers(-128, 128, size=(ncols, K), dtype=np.int8)
    got = _gemm(W, Q, ksplit)
    ksteps = -(-K // 128)
    per = -(-ksteps // ksplit)
    for ks in range(ksplit):
        lo, hi = ks * per * 128, min(K, (ks + 1) * per * 128)
        want = W[:, lo:hi].astype(np.int64) @ Q[:, lo:hi].astype(np.int64).T
        assert np.array_equal(got[ks].astype(np.int64), want), f"variant {variant}, K split {ks}"


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("count,ncols,K,pairs,epoch,slack", [
    (1100, 1100, 5120, 4, 4, 1),      # full waves + K-chunked left-over tiles, K-sync on
    (1100, 1100, 5120, 0, 16, 2),     # all pairs of the GPU: everything is left-over (atomics)
    (700, 1500, 2100, 3, 2, 1),       # ragged K, ragged rows and columns
    (2000, 3000, 4096, 7, 8, 0),      # several bands, sync off
    (257, 513, 300, 1, 1, 1),         # a single pair walks every tile
])
def test_persistent_pair_gemm_is_exact(variant, count, ncols, K, pairs, epoch, slack, monkeypatch):
    """The persistent, K-synchronised CTA-pair kernel: waves of whole tiles, the left-over tiles cut along K and
    accumulated with int32 atomics, pairs paced by the global epoch counters -- all bit-exact."""
    monkeypatch.setenv("SALIB_B200_DIGIT_GEMM", str(variant))
    monkeypatch.setenv("SALIB_B200_GEMM_EPOCH", str(epoch))
    monkeypatch.setenv("SALIB_B200_GEMM_SLACK", str(slack))
    if pairs:
        monkeypatch.setenv("SALIB_B200_GEMM_PAIRS", str(pairs))
    rng = np.random.default_rng(variant * 7 + count + ncols + K)
    W = rng.integers(0, 256, size=(count, K), dtype=np.uint8)
    Q = rng.integers(-128, 128, size=(ncols, K), dtype=np.int8)
    got = _gemm(W, Q, 1)[0]
    want = W.astype(np.int64) @ Q.astype(np.int64).T
    assert np.array_equal(got.astype(np.int64), want)
    # the one-tile-per-pair kernel of round 1 stays available and agrees
    monkeypatch.setenv("SALIB_B200_GEMM_PERSIST", "0")
    assert np.array_equal(_gemm(W, Q, 1)[0], got)


def test_digit_gemm_one_hot_layout():
    """Each output picks exactly one (row of W, column of Q, k) product: any operand-layout error shows up
    as a misplaced value rather than as noise."""
    count, ncols, K = 128, 256, 512
    W = np.zeros((count, K), dtype=np.uint8)
    Q = np.zeros((ncols, K), dtype=np.int8)
    for m in range(count):
        W[m, (m * 5) % K] = 1 + m % 100
    for n in range(ncols):
        Q[n, (n * 5) % K] = (n % 120) - 60
    got = _gemm(W, Q, 1)[0]
    want = W.astype(np.int64) @ Q.astype(np.int64).T
    assert np.array_equal(got, want)


CASES = ["ishigami_n1024_r100", "ishigami_n256_c1_r50", "g8_n512_r64", "g8_n512_c1_r64", "g16_n256_r32",
         "grp3_n512_r40", "g64_n128_r16"]


def _cmp(S, ref, D, c2, rtol=RTOL, atol=ATOL):
    for k in ["S1", "S1_conf", "ST", "ST_conf", "S1_conf_all", "ST_conf_all"]:
        np.testing.assert_allclose(S[k], ref[k], rtol=rtol, atol=atol, err_msg=k)
    if c2:
        iu = np.triu_indices(D, 1)
        for k in ("S2", "S2_conf"):
            np.testing.assert_allclose(S[k][iu], np.asarray(ref[k])[iu], rtol=rtol, atol=atol, err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_digits_engine_matches_frozen_reference(name):
    from salib_b200.analyze import sobol

    z = golden("analyze_" + name)
    D = z["S1"].shape[0]
    S = sobol.analyze_device(problem_from(z), z["Y"], calc_second_order=bool(z["c2"]), conf_level=float(z["conf"]),
                             keep_resamples=True, r=z["r"].astype(np.int64), algo="digits")
    _cmp(S, z, D, bool(z["c2"]))


@pytest.mark.parametrize("digits,rtol", [(5, 1e-7), (6, 1e-9), (8, 1e-10)])
def test_digit_count_sets_the_precision(digits, rtol, monkeypatch):
    from salib_b200.analyze import sobol

    monkeypatch.setattr(sobol, "DIGITS", digits)
    z = golden("analyze_g16_n256_r32")
    S = sobol.analyze_device(problem_from(z), z["Y"], calc_second_order=True, conf_level=float(z["conf"]),
                             keep_resamples=True, r=z["r"].astype(np.int64), algo="digits")
    _cmp(S, z, 16, True, rtol=rtol, atol=rtol * 1e-2)


def test_digits_engine_matches_fp64_engine_philox():
    """Same device Philox multiplicities through both engines, a shape with several resample tiles and K
    splits: D = 24, N = 8192, R = 300."""
    from salib_b200.analyze import sobol
    from salib_b200.sample import sobol as sample
    from salib_b200.test_functions import Sobol_G

    D = 24
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    X = sample.sample(p, 8192, scramble=False)
    Y = Sobol_G.evaluate(X, a=np.linspace(0.0, 30.0, D))
    a = sobol.analyze_device(p, Y, num_resamples=300, seed=7, resample="philox", keep_resamples=True)
    b = sobol.analyze_device(p, Y, num_resamples=300, seed=7, resample="philox", keep_resamples=True,
                             algo="digits")
    _cmp(b, a, D, True, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("Dg,N,c2,S", [(3, 1000, True, 7), (8, 517, False, 6), (16, 300, True, 7), (5, 4096, True, 8),
                                         (64, 130, True, 7)])
def test_digit_matrix_and_sums_match_the_oracle_bit_for_bit(Dg, N, c2, S):
    """Device digit columns == oracle digits (int8, every byte); device sums == the oracle's exact integer sums
    (Python integers, one rounding).  Y has a wide dynamic range so that columns get very different scales."""
    import torch
    from oracle import digits as o_digits
    from salib_b200 import _lib, _runtime as rt
    from salib_b200.analyze import sobol

    rng = np.random.default_rng(Dg * 1000 + N)
    step = 2 * Dg + 2 if c2 else Dg + 2
    Y = rng.normal(size=(N, step)) * np.exp(rng.normal(size=(1, step)) * 2.0) + rng.normal(size=(1, step))
    Y = Y.reshape(-1)
    eng = sobol.SobolAnalyzer(Dg, N, c2, algo="digits")
    eng.ndigits = S
    Yd = rt.to_device(Y, torch.float64)
    eng.load(Yd)
    stats = eng.stats.cpu().numpy()
    F = o_digits.features(Y, Dg, c2, stats[0], stats[1])
    dig, inv, _ = o_digits.split(F, S)
    pitch = int(_lib.lib().sa_digit_row_pitch(N))
    Q = eng.Q.cpu().numpy().reshape(pitch // 128, eng.nacc * S, 128)          # K-blocked [row block][column][128]
    Q = Q.transpose(1, 0, 2).reshape(eng.nacc * S, pitch)
    assert np.array_equal(Q[:, :N], dig)
    assert not Q[:, N:].any()
    assert np.array_equal(eng.inv.cpu().numpy()[:eng.nacc], inv)
    # the heavy-tail guard count rides behind the scales
    assert int(eng.inv.cpu().numpy()[eng.nacc]) == o_digits.guard_count(Y, Dg, c2, stats[0], stats[1], stats[2], stats[3])

    count = 37
    w = rng.multinomial(N, np.full(N, 1.0 / N), size=count).astype(np.uint8)
    w[0, :] = 0
    w[0, 0] = 200                                                   # a heavy row: multiplicities beyond 127
    wd = torch.zeros((count, eng.Npad), dtype=torch.uint8, device="cuda")
    wd[:, :N] = torch.from_numpy(w).cuda()
    psum, nsplit = eng._partial_sums(wd.reshape(-1), count)
    assert nsplit == 1
    got = psum.cpu().numpy().reshape(count, eng.nacc)
    assert np.array_equal(got, o_digits.sums(w, dig, inv, S))


def test_digits_invariances_d64():
    """The properties tests/test_gpu_properties.py checks for the FP64 engines, for the integer engine at
    D = 64, N = 2^16: "auto" picks it at this size and agrees with the FP64 engines; affine invariance; resample
    sharding (two halves == one call, bit for bit: integer sums do not depend on the batch); an all-ones
    multiplicity row reproduces the point estimate bit for bit."""
    import torch

    from salib_b200 import _runtime as rt
    from salib_b200.analyze import sobol
    from salib_b200.sample import sobol as sample
    from salib_b200.test_functions import Sobol_G

    D, N, R = 64, 1 << 16, 300
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    plan, first = sample._plan(p, N, True, False, N, None)
    Y = rt.empty(((2 * D + 2) * N,), torch.float64)
    Sobol_G.evaluate_sample(plan, first, N, a=np.resize([0, 1, 4.5, 9, 99, 99, 99, 99.0], D), out=Y, check=False)
    S = sobol.analyze_device(p, Y, num_resamples=R, seed=100, resample="philox", keep_resamples=True)
    assert S.engine == "digits"
    F = sobol.analyze_device(p, Y, num_resamples=R, seed=100, resample="philox", keep_resamples=True, algo="fp64")
    assert F.engine == "fp64-tile8"
    _cmp(S, F, D, True, rtol=1e-9, atol=1e-12)
    T = sobol.analyze_device(p, 3.0 * Y + 7.0, num_resamples=R, seed=100, resample="philox", keep_resamples=True)
    _cmp(S, T, D, True, rtol=1e-9, atol=1e-12)

    eng = sobol.SobolAnalyzer(D, N, True, algo="digits")
    eng.load(Y)
    rs = sobol._Resampler(N, R, 100, "philox")
    whole = eng.bootstrap(rs, 0, R)
    halves = torch.cat([eng.bootstrap(rs, 0, 100), eng.bootstrap(rs, 100, R)])
    assert torch.equal(whole, halves)
    np.testing.assert_array_equal(whole[:, :D].cpu().numpy(), S["S1_conf_all"])
    ones = torch.ones((eng.Npad,), dtype=torch.uint8, device="cuda")
    est1 = rt.empty((1, eng.nest), torch.float64)
    eng._accumulate(ones, 1, est1)
    assert torch.equal(est1, eng.point())


def test_auto_falls_back_to_fp64_for_degenerate_outputs():
    """NaN / constant outputs keep the reference's NaN propagation through the FP64 engines even at
    sizes where "auto" would take the integer engine."""
    import warnings

    from salib_b200.analyze import sobol

    D, N = 8, 1 << 16
    p = {"num_vars": D, "names": [f"x{i}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}
    Y = np.random.default_rng(0).normal(size=(2 * D + 2) * N)
    S = sobol.analyze_device(p, Y, num_resamples=1000, seed=3, resample="philox")
    assert S.engine == "digits" and np.isfinite(S["S1"]).all()
    Ynan = Y.copy()
    Ynan[12345] = np.nan
    for bad in (Ynan, np.full_like(Y, 0.5)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = sobol.analyze_device(p, bad, num_resamples=1000, seed=3, resample="philox")
            F = sobol.analyze_device(p, bad, num_resamples=1000, seed=3, resample="philox", algo="fp64")
        assert S.engine.startswith("fp64")
        for k in F:
            np.testing.assert_array_equal(S[k], F[k])


def test_digits_row_sharded_single_process():
    """Two row shards in one process (partial sums added, then the estimator algebra) through the digits engine."""
    from salib_b200.analyze import sobol

    z = golden("analyze_g8_n512_r64")
    p = problem_from(z)
    Y = np.asarray(z["Y"])
    step = 2 * 8 + 2
    cut = 200
    S = sobol.analyze_row_sharded(p, [(0, Y[:cut * step]), (cut, Y[cut * step:])], 512, calc_second_order=True,
                                  conf_level=float(z["conf"]), keep_resamples=True, r=z["r"].astype(np.int64),
                                  algo="digits", distributed=False)
    _cmp(S, z, 8, True)


def test_exports():
    from salib_b200 import _lib

    L = _lib.lib()
    assert L.sa_digit_row_pitch(1000) == 1024
    assert L.sa_digit_ldc(10, 6) == 64
    assert L.sa_digit_ksplit(1 << 20, 2149, 6, 10000) == 1
    assert isinstance(L.sa_digit_sums_workspace_bytes(1 << 14, 100, 6, 50), int)
    assert C.sizeof(C.c_size_t) == 8
