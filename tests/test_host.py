"""Host-side logic that runs without a GPU: the C-ABI library loads and exports every declared symbol,
direction vectors, argument validation / warnings of the drop-ins, ProblemSpec bookkeeping, sharding."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

from conftest import ROOT
from oracle import sobol_seq


def test_library_exports_every_declared_symbol():
    from salib_b200 import _lib

    header = open(os.path.join(ROOT, "include", "salib_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == declared
    assert _lib.lib().sa_version() >= 100
    assert _lib.lib().sa_row_stride(64, 1) == 162 and _lib.lib().sa_row_stride(3, 0) == 12
    assert _lib.lib().sa_pad_rows(1000) == 1024 and _lib.lib().sa_pad_rows(4096) == 4224


def test_direction_vectors_match_oracle_and_scipy():
    from salib_b200.sample.directions import direction_vectors
    from scipy.stats import qmc

    sv = direction_vectors(300)
    assert sv.dtype == np.uint32 and sv.shape == (300, 30)
    assert np.array_equal(sv, sobol_seq.direction_vectors(300).astype(np.uint32))
    assert np.array_equal(sv, qmc.Sobol(300, scramble=False)._sv)
    with pytest.raises(ValueError):
        direction_vectors(21202)


def test_direction_vectors_argument_errors():
    from salib_b200 import _lib

    L = _lib.lib()
    assert L.sa_direction_vectors(None, None, None, 4, None) == -1
    assert b"sa_direction_vectors" in L.sa_last_error()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from salib_b200.analyze import sobol as an
    from salib_b200.sample import sobol as sm

    p = {"num_vars": 3, "names": ["a", "b", "c"], "bounds": [[0, 1]] * 3}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sm.sample(p, 8, scramble=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        an.analyze(p, np.zeros(64))


def test_sample_validation_matches_reference():
    # reference sample/sobol.py:110-133 and tests/test_cli_sample.py:57-90 (warnings), util bounds errors
    from salib_b200.sample import sobol as sm

    p = {"num_vars": 2, "names": ["a", "b"], "bounds": [[0, 1], [0, 1]]}
    with pytest.raises(ValueError, match="positive integer"):
        sm.sample(p, 8, skip_values=-1)
    with pytest.raises(ValueError, match="positive integer"):
        sm.sample(p, 8, skip_values=2.0)
    bad = {"num_vars": 2, "names": ["a", "b"], "bounds": [[0, 1], [1, 0]]}
    with pytest.raises(ValueError, match="Bounds are not legal"):
        sm.sample(bad, 8, scramble=False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for kw in (dict(skip_values=1025), dict(skip_values=4)):
            try:
                sm.sample(p, 16, scramble=False, **kw)
            except RuntimeError:
                pass  # no GPU here: validation already ran
        msgs = " | ".join(str(x.message) for x in w)
    assert "is a power of 2" in msgs and "Convergence may not be valid" in msgs
    with pytest.raises(ValueError, match="At most 2\\*\\*30"):
        sm.sample(p, 2 ** 30, scramble=False, skip_values=2)


def test_dists_validation_matches_reference():
    # reference util/__init__.py:160-289 and tests/test_util.py (illegal distribution parameters)
    from salib_b200.sample.sobol import _dist_tables

    def tab(dist, bound):
        return _dist_tables([bound], [dist], 1)

    for dist, bound in [("unif", [1, 1]), ("triang", [1, 3, 1.0]), ("triang", [3, 1, 0.5]), ("triang", [1, 2, 3, 4]),
                        ("norm", [0, 0]), ("truncnorm", [0, 1, 0, 0]), ("truncnorm", [1, 0, 0, 1]),
                        ("lognorm", [0, -1]), ("weibull", [0, 1]), ("weibull", [1, 1, 1, 1]), ("beta", [0, 1])]:
        with pytest.raises(ValueError):
            tab(dist, bound)
    with pytest.raises(ValueError, match="Mismatch in number"):
        _dist_tables([[0, 1]], ["unif", "unif"], 1)
    with pytest.warns(DeprecationWarning, match="Two-value format"):
        code, par = tab("triang", [3.0, 0.5])
    assert code[0] == 1 and par[0, :3].tolist() == [0.0, 3.0, 0.5]
    code, par = tab("truncnorm", [2.0, 6.0, 1.0, 1.5])   # lower bound right of the mean: right-tail form
    assert code[0] == 5
    code, par = tab("weibull", [2.0, 3.0])
    assert code[0] == 7 and par[0, :3].tolist() == [0.5, 3.0, 0.0]


def test_saltelli_validation_matches_reference():
    from salib_b200.sample import saltelli

    p = {"num_vars": 2, "names": ["a", "b"], "bounds": [[0, 1], [0, 1]]}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        for kw in (dict(), dict(skip_values=0), dict(skip_values=3)):
            try:
                saltelli.sample(p, 511, **kw)
            except RuntimeError:
                pass
    cats = [x.category for x in w]
    msgs = " | ".join(str(x.message) for x in w)
    assert DeprecationWarning in cats
    assert "is equal to `2^n`" in msgs and "Duplicate samples" in msgs and "is a power of 2" in msgs


def test_analyze_validation_matches_reference():
    # reference tests/test_sobol.py:53-85
    from salib_b200.analyze import sobol as an

    p = {"num_vars": 3, "names": ["a", "b", "c"], "bounds": [[0, 1]] * 3}
    with pytest.raises(RuntimeError, match="Incorrect number of samples"):
        an.analyze(p, np.ones(7))
    with pytest.raises(RuntimeError, match="Incorrect number of samples"):
        an.analyze(p, np.ones(5 * 4 + 1), calc_second_order=False)
    with pytest.raises(RuntimeError, match="Confidence level"):
        an.analyze(p, np.ones(16), conf_level=1.0)
    # must stay a plain function without a parameter called X (util/problem.py:354)
    assert "X" not in an.analyze.__code__.co_varnames


def test_group_helpers():
    from salib_b200.util import _check_groups, compute_groups_matrix, extract_group_names, group_columns

    p = {"num_vars": 5, "names": list("vwxyz"), "groups": ["a", "b", "a", "c", "b"], "bounds": [[0, 1]] * 5}
    assert _check_groups(p) == p["groups"]
    assert _check_groups({"names": ["a"], "groups": None}) is False
    assert _check_groups({"names": ["a", "b"], "groups": ["g", "g"]}) is False
    assert _check_groups({"names": ["a", "b"], "groups": ["a", "b"]}) is False
    G, names = compute_groups_matrix(p["groups"])
    assert list(names) == ["a", "b", "c"]
    assert G.tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0]]
    assert extract_group_names(p) == (["a", "b", "c"], 3)
    assert extract_group_names({"names": ["q", "r"]}) == (["q", "r"], 2)
    gcol, Dg = group_columns(p)
    assert gcol.tolist() == [0, 1, 0, 2, 1] and Dg == 3
    assert group_columns({"num_vars": 2, "names": ["a", "b"]}) == (None, 2)


def test_problem_spec_bookkeeping():
    # reference tests/test_problem_spec.py:30-96
    from salib_b200.util import ProblemSpec, ResultDict

    sp = ProblemSpec({"names": ["x1", "x2", "x3"], "bounds": [[-np.pi, np.pi]] * 3, "outputs": ["Y"]})
    assert sp["num_vars"] == 3 and sp["groups"] == sp["names"]
    assert "does not currently contain" in str(sp)
    sp.sample(lambda prob, n: np.zeros((n, prob["num_vars"])), 10)
    assert sp.samples.shape == (10, 3)
    with pytest.raises(ValueError, match="Mismatched sample size"):
        sp.samples = np.zeros((4, 2))
    sp.evaluate(lambda X, k: X.sum(axis=1) + k, 2.0)
    assert np.all(sp.results == 2.0)
    sp.set_samples(np.ones((10, 3)))
    assert sp.results is None  # setter clears results
    with pytest.raises(RuntimeError, match="Model not yet evaluated"):
        sp.analyze(lambda prob, Y: None)
    with pytest.raises(AssertionError):
        sp.set_results(np.zeros(11))
    sp.set_results(np.arange(10.0))

    def fake_analyze(problem, Y, scale=1.0):
        S = ResultDict(S1=np.full(problem["num_vars"], Y.mean() * scale), names=problem["names"])
        return S

    sp.analyze(fake_analyze, scale=2.0)
    assert np.allclose(sp.analysis["S1"], 9.0)
    assert sp.to_df().shape == (3, 1)
    sp2 = ProblemSpec({"names": ["a", "b"], "bounds": [[0, 1]] * 2})
    sp2.set_samples(np.zeros((4, 2))).set_results(np.arange(8.0).reshape(4, 2))
    assert sp2["outputs"] == ["Y1", "Y2"]
    sp2.analyze(fake_analyze)
    assert set(sp2.analysis) == {"Y1", "Y2"} and len(sp2.to_df()) == 2
    with pytest.raises(ValueError, match="single parameter"):
        ProblemSpec({"names": ["a"], "bounds": [[0, 1]]}).set_samples(np.zeros((2, 1))).set_results(
            np.zeros(2)).analyze(fake_analyze)
    with pytest.raises(AssertionError):
        ProblemSpec({"names": ["a", "b"], "bounds": [[0, 1]]})


def test_result_assembly_and_to_df():
    # reference tests/test_to_df.py:59-83, analyze/sobol.py:249-264
    from types import MethodType

    from salib_b200.analyze import sobol as an

    D = 3
    point = np.arange(1.0, 1.0 + 2 * D + 3)
    conf = point / 10
    est = np.arange(4.0 * (2 * D)).reshape(4, 2 * D)
    S = an._assemble(D, True, True, point, conf, est, False)
    assert list(S) == ["S1", "S1_conf", "ST", "ST_conf", "S1_conf_all", "ST_conf_all", "S2", "S2_conf"]
    assert S["S2"].shape == (3, 3) and np.isnan(S["S2"][1, 0]) and np.isnan(S["S2"][0, 0])
    assert S["S2"][0, 1] == 7.0 and S["S2"][0, 2] == 8.0 and S["S2"][1, 2] == 9.0
    assert S["S1_conf_all"].shape == (4, 3) and S["ST_conf_all"][2, 1] == est[2, 4]
    S.problem = {"names": ["x1", "x2", "x3"], "num_vars": 3}
    S.to_df = MethodType(an.to_df, S)
    total, first, second = S.to_df()
    assert list(total.columns) == ["ST", "ST_conf"] and list(first.index) == ["x1", "x2", "x3"]
    assert list(second.index) == [("x1", "x2"), ("x1", "x3"), ("x2", "x3")]
    # constant output: zeros, S2 = 0, S2_conf = NaN (reference tests/test_sobol.py:170-196)
    C = an._assemble(D, True, False, point, conf, est, True)
    assert not C["S1"].any() and not C["ST_conf"].any() and C["S2"][0, 1] == 0.0 and np.isnan(C["S2_conf"][0, 1])
    F = an._assemble(D, False, False, point[:2 * D], conf[:2 * D], None, False)
    assert "S2" not in F and F["ST"].tolist() == [4.0, 5.0, 6.0]


def test_work_decomposition_plans():
    """Batching of resamples and row splits (host logic of analyze): full coverage, whole waves for the tile
    kernel, buffer cap respected, small counts split over rows so they fill the machine."""
    from salib_b200.analyze.sobol import plan_batches, plan_row_splits, tile_groups

    assert [tile_groups(d) for d in (3, 64, 65, 128, 200)] == [1, 1, 2, 5, 11]
    assert plan_row_splits(4, 100, 1 << 16, 148, 8389, groups=5) == 2

    sms, Npad = 148, (1 << 20) + 128
    b = plan_batches(4, 0, 10000, Npad, sms)
    assert b[0][0] == 0 and b[-1][1] == 10000 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert all((c1 - c0) % (8 * sms) == 0 for c0, c1 in b[:-1]) and b[-1][1] - b[-1][0] == 10000 % (8 * sms)
    small_cap = plan_batches(4, 5, 5000, Npad, sms, max_bytes=2000 * Npad)
    assert max(c1 - c0 for c0, c1 in small_cap) <= 2000 and small_cap[0][0] == 5 and small_cap[-1][1] == 5000
    assert plan_batches(3, 0, 1000, Npad, sms) == [(0, 1000)]
    assert plan_batches(4, 7, 7, Npad, sms) == []
    assert plan_row_splits(4, 1184, 1 << 20, sms, 2149) == 1          # a full wave needs no split
    assert plan_row_splits(4, 1, 1 << 20, sms, 2149) == 148           # point estimate: one block per SM
    assert plan_row_splits(4, 528, 1 << 20, sms, 2149) == 2
    assert plan_row_splits(4, 16, 100, sms, 11) == 3                  # never fewer than 32 rows per split
    assert plan_row_splits(3, 1000, 1 << 20, sms, 21) == 333             # 4 blocks of 256 resamples -> three waves of 3 per SM
    assert plan_row_splits(1, 1000, 4096, sms, 50) == 1
    # integer engine (kind 6): one batch as large as the multiplicity buffer allows, in whole 128-resample tiles
    assert plan_batches(6, 0, 10000, Npad, sms) == [(0, 10000)]
    assert plan_batches(6, 3, 10000, Npad, sms, max_bytes=1000 * Npad) == \
        [(3, 899), (899, 1795), (1795, 2691), (2691, 3587), (3587, 4483), (4483, 5379), (5379, 6275), (6275, 7171),
         (7171, 8067), (8067, 8963), (8963, 9859), (9859, 10000)]
    assert plan_batches(6, 0, 50, Npad, sms, max_bytes=10 * Npad) == [(0, 10), (10, 20), (20, 30), (30, 40), (40, 50)]
    assert plan_batches(6, 7, 7, Npad, sms) == []


def test_engine_selection_constants():
    """Names and thresholds of the analyze engines (host logic; the selection itself needs a device)."""
    from salib_b200.analyze import sobol as an

    assert an._ALGO["auto"] == 0 and an._ALGO["fp64"] < 0 and an._ALGO["digits"] == 6
    assert an.ENGINE_NAMES[6] == "digits" and an.ENGINE_NAMES[4] == "fp64-tile8"
    assert an.digits_for(1 << 20) == 7 and an.digits_for(100) == 7
    assert an._nacc(64, True) == 2149 and an._nest(64, True) == 2144
    # the headline is far above the hand-over point, the README quick start far below
    assert (1 << 20) * 10000 * an._nacc(64, True) >= an.DIGITS_MIN_WORK > 1024 * 100 * an._nacc(3, True)


def test_read_param_file_and_cli_parsers(tmp_path):
    # reference tests/conftest.py fixtures + util_funcs.read_param_file; CLI options of sample/analyze
    from salib_b200.scripts import salib as cli
    from salib_b200.util import read_param_file

    f = tmp_path / "p.txt"
    f.write_text("#x0 0 1\nx1 -3.14159265359 3.14159265359\nx2 0 1\nx3 2.5 7\n")
    p = read_param_file(str(f))
    assert p == {"names": ["x1", "x2", "x3"], "bounds": [[-3.14159265359, 3.14159265359], [0.0, 1.0], [2.5, 7.0]],
                 "num_vars": 3, "groups": None, "dists": None}
    g = tmp_path / "g.csv"
    g.write_text("a,0,1,G1,unif\nb,0,2,G2,triang\nc,1,3,NA,norm\n")
    p = read_param_file(str(g))
    assert p["groups"] == ["G1", "G2", "c"] and p["dists"] == ["unif", "triang", "norm"] and p["num_vars"] == 3
    one = tmp_path / "one.csv"
    one.write_text("a,0,1,G\nb,0,2,G\n")
    with pytest.raises(ValueError, match="Only one group"):
        read_param_file(str(one))
    a = cli._sample_parser("saltelli").parse_args("-n 8 -p p -o o --max-order 1 --skip-values 16".split())
    assert (a.samples, a.max_order, a.skip_values, a.precision, a.delimiter) == (8, 1, 16, 8, " ")
    with pytest.raises(SystemExit):
        cli._sample_parser("saltelli").parse_args("-n 8 -p p -o o --seed 3".split())   # no seed for saltelli
    b = cli._analyze_parser("sobol").parse_args("-p p -Y y -c 2 -r 1000 --seed=100".split())
    assert (b.column, b.resamples, b.seed, b.max_order, b.parallel) == (2, 1000, 100, 2, False)
    assert cli.main([]) == 0


def test_shard_range_covers_everything():
    from salib_b200._runtime import shard_range

    for total in (0, 1, 7, 100, 10000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, R, nest, q):
    import torch
    import torch.distributed as dist

    from salib_b200._runtime import shard_range
    from salib_b200.analyze.sobol import allgather_estimates

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(R * nest, dtype=torch.float64).reshape(R, nest)
    lo, hi = shard_range(R, rank, world)
    out = allgather_estimates(full[lo:hi].clone(), R, rank, world)
    ok = bool(torch.equal(out, full))
    # sharded upload of a host array: every rank ends with the full array (device() is patched to CPU here)
    import numpy as np

    from salib_b200 import _runtime as rt

    rt.device = lambda: torch.device("cpu")
    y = np.arange(1000 * R + 3, dtype=np.float64)
    got = rt.to_device_sharded(y, torch.float64, min_bytes=0)
    ok = ok and bool(torch.equal(got, torch.from_numpy(y)))
    # row-sharded analyze: per-shard moments of every rank (rank 1 holds two shards) -> statistics of the union
    from salib_b200.analyze.sobol import allgather_moments, combine_moments

    rng = np.random.default_rng(5)
    data = rng.normal(3.0, 2.0, size=(90, 4))           # 90 base rows, step 4: columns A, AB_0, AB_1, B
    cuts = [(0, 40)] if rank == 0 else [(40, 47), (47, 90)]

    def moments(block):
        ab = block[:, [0, 3]]
        return (float(block.size), np.array([block.mean(), block.std(), block.min(), block.max(), ab.min(), ab.max(), 0, 0]))

    parts = allgather_moments([moments(data[a:b]) for a, b in cuts], world)
    st = combine_moments(parts)
    ref = moments(data)[1]
    ok = ok and len(parts) == 3 and bool(np.allclose(st[:6], ref[:6], rtol=1e-14, atol=0))
    # decisions every rank must share before it issues collectives (engine family, batch sizes): MAX over ranks
    from salib_b200.analyze.sobol import _agree

    ok = ok and _agree([rank, 10 - rank, 7]) == [world - 1, 10, 7]
    # the same agreement started early and read late (what analyze_row_sharded does around the upload of Y)
    from salib_b200.analyze.sobol import _Agreement

    pending = _Agreement([3 * rank, 5, 1 - rank], world)
    ok = ok and pending.result() == [3 * (world - 1), 5, 1] and _Agreement([4, 2], 1).result() == [4, 2]
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("R", [7, 16])
def test_allgather_estimates_gloo_world2(R):
    """N>1 host logic of analyze (resample sharding + all-gather) on the CPU with gloo, world_size 2."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, R, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


@pytest.mark.parametrize("tm,tn,ksteps,pairs,band", [(40, 30, 8192, 74, 8), (5, 3, 40, 4, 8), (1, 1, 1024, 74, 8),
                                                     (3, 7, 17, 5, 2), (40, 30, 1024, 74, 10), (10, 30, 64, 74, 8),
                                                     (2, 2, 8, 3, 8)])
def test_persistent_pair_plan_covers_every_tile_once(tm, tn, ksteps, pairs, band):
    """Work list of the persistent CTA-pair GEMM (digits.cu: unit() in digit_gemm_pair_persistent_kernel,
    restated here): full waves of whole tiles plus K-chunked left-over tiles must cover every (tile, K-step)
    exactly once, and a left-over tile is only ever touched by atomic units."""
    from salib_b200 import _lib

    out = (ctypes.c_uint32 * 8)()
    _lib.call("sa_digit_pair_plan", tm, tn, ksteps, pairs, band, out)
    P, full_waves, rem_tiles, rsplit, rper, band_, epoch, slack = list(out)
    tiles = tm * tn
    assert 1 <= P <= pairs and band_ == band
    assert full_waves * P + rem_tiles == tiles or (full_waves == 0 and rem_tiles == tiles)
    assert rper * (rsplit - 1) < ksteps <= rper * rsplit
    cover = np.zeros((tiles, ksteps), dtype=np.int32)
    seen_mn = set()
    for pair in range(P):
        rem_units = rem_tiles * rsplit
        my_rem = (rem_units - pair + P - 1) // P if pair < rem_units else 0
        for idx in range(full_waves + my_rem):
            if idx < full_waves:
                t, k0, nk = idx * P + pair, 0, ksteps
            else:
                r = pair + (idx - full_waves) * P
                t = full_waves * P + r % rem_tiles
                k0 = (r // rem_tiles) * rper
                nk = min(rper, ksteps - k0)
                assert nk >= 1
            cover[t, k0:k0 + nk] += 1
            b = t // (band * tn)
            in_band = t - b * band * tn
            bm = min(band, tm - b * band)
            seen_mn.add((t, b * band + in_band % bm, in_band // bm))
    assert (cover == 1).all()
    assert sorted({(m, n) for _, m, n in seen_mn}) == [(m, n) for m in range(tm) for n in range(tn)]
    if full_waves == 0:
        assert slack == 0           # nothing to keep in step
    # the left-over wave costs about rem/pairs of a tile, not a whole one
    if rem_tiles and ksteps >= 256:
        rounds = -(-rem_tiles * rsplit // P)
        assert rounds * rper / ksteps <= rem_tiles / P + 0.15


def test_numpy_resample_stream_can_be_drawn_in_row_blocks():
    """analyze/sobol.py:113-117,144 draws r = rng(N, size=(N, R)) in one call; the drop-in streams the same draw in
    row blocks (bounded host memory).  One generator continues one stream: the blocks concatenate to the single
    call, for the seeded Generator and for the legacy global RNG (seed falsy)."""
    N, R = 5000, 37
    whole = np.random.default_rng(1234).integers(N, size=(N, R))
    g = np.random.default_rng(1234).integers
    parts = [g(N, size=(b - a, R)) for a, b in [(0, 1), (1, 700), (700, 4099), (4099, N)]]
    assert np.array_equal(np.concatenate(parts), whole)
    np.random.seed(99)
    whole = np.random.randint(N, size=(N, R))
    np.random.seed(99)
    parts = [np.random.randint(N, size=(b - a, R)) for a, b in [(0, 333), (333, 334), (334, N)]]
    assert np.array_equal(np.concatenate(parts), whole)
    # ranges that need 64-bit draws too
    big = 2 ** 33 + 5
    whole = np.random.default_rng(7).integers(big, size=(2000, 3))
    g = np.random.default_rng(7).integers
    assert np.array_equal(np.concatenate([g(big, size=(1234, 3)), g(big, size=(766, 3))]), whole)


def test_digits_batches_and_reduce_scatter_bookkeeping():
    """Host logic of the pipelined row route (analyze/sobol.py: digits_batch, batch_share): batches cover [0, R) in
    whole tiles, and the rows each rank keeps after the reduce-scatter tile every batch exactly once, the row of
    ones (the point estimate) landing on exactly one rank."""
    from salib_b200.analyze.sobol import batch_share, digits_batch, digits_batches

    assert digits_batches(10000, 1 << 20) == [(0, 2560), (2560, 5120), (5120, 7680), (7680, 10000)]
    assert digits_batches(0, 1 << 20) == [(0, 0)] and digits_batches(100, 1 << 20) == [(0, 100)]
    short = digits_batches(10000, 1 << 20, first=512)      # a short first batch (experiment knob), the rest equal
    assert short[0] == (0, 512) and short[-1][1] == 10000 and len(short) == 5
    assert all(a1 == b0 for (_, a1), (b0, _) in zip(short, short[1:]))
    assert all((c1 - c0) % 256 == 0 for c0, c1 in short[:-1])
    assert digits_batches(10000, 1 << 20, first=4096) == digits_batches(10000, 1 << 20)   # not below the batch size

    assert digits_batch(10000, 1 << 20) == 2560      # headline: 4 batches, 3 draws hidden behind a GEMM
    assert digits_batch(10000, 1 << 17) == 2560      # one of 8 row shards: the same batches
    assert digits_batch(10000, 1 << 22) == 2048      # 8.5 GB of multiplicities per buffer at most
    assert digits_batch(100, 1 << 20) == 100
    assert digits_batch(1000, 1 << 20, 147) == 1000  # cfg2 (D = 8, first order): the GEMM is tiny, one batch
    assert digits_batch(0, 1 << 20) == 1
    for R, npad in [(10000, 1 << 20), (10000, 1 << 17), (777, 1 << 22), (5000, 1 << 24), (1, 128)]:
        B = digits_batch(R, npad)
        assert B == R or B % 256 == 0
        batches = [(c, min(c + B, R)) for c in range(0, R, B)]
        assert batches[0][0] == 0 and batches[-1][1] == R
        for world in (1, 2, 3, 8):
            seen = np.zeros(R, dtype=int)
            owners = 0
            for c0, c1 in batches:
                n = c1 - c0
                count = n + (1 if c1 == R else 0)
                for g in range(world):
                    chunk, lo, valid = batch_share(n, count, world, g)
                    assert chunk * world >= count
                    seen[c0 + lo:c0 + lo + valid] += 1
                    owners += int(count > n and lo <= n < lo + chunk)
            assert (seen == 1).all() and owners == 1


def test_heavy_tail_guard_rule():
    from salib_b200.analyze.sobol import guard_trips

    assert guard_trips(18, 18, 512)              # one outlier row: a resample misses it with probability 0.37
    assert guard_trips(10 * 130, 130, 1 << 20)   # ten rows: e^-10
    assert not guard_trips(40 * 130, 130, 1 << 20)
    assert not guard_trips(32 * 410, 410, 32)    # every row is "large": no resample can miss them all
    assert guard_trips(0, 130, 1 << 20)
