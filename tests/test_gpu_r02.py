"""Round-2 GPU parity: the pipelined row route of the integer engine, seeded NumPy resamples beyond 2^24 draws,
heavy-tailed outputs, wide sample matrices, the engine-selection boundary, slim vs wide Philox launches, and a
2-rank run of the drop-in ``analyze`` (skipped with fewer than 2 GPUs)."""
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, problem_from

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-10, 1e-13


def gprob(D):
    return {"num_vars": D, "names": [f"x{i + 1}" for i in range(D)], "bounds": [[0.0, 1.0]] * D}


def _cmp(S, ref, D, c2, rtol=RTOL, atol=ATOL, keys=("S1", "S1_conf", "ST", "ST_conf", "S1_conf_all", "ST_conf_all")):
    for k in keys:
        np.testing.assert_allclose(S[k], ref[k], rtol=rtol, atol=atol, err_msg=k)
    if c2:
        iu = np.triu_indices(D, 1)
        for k in ("S2", "S2_conf"):
            np.testing.assert_allclose(S[k][iu], np.asarray(ref[k])[iu], rtol=rtol, atol=atol, err_msg=k)


def test_seeded_resamples_follow_the_reference_beyond_2p24_draws():
    """analyze(..., seed=s) at D=8, N=2^16, R=1000 (6.5e7 draws; round 1 silently switched to Philox above 2^24):
    the drop-in keeps the reference's own default_rng(seed).integers stream (analyze/sobol.py:113-117,144),
    streamed in row blocks, and reproduces the frozen reference result."""
    from salib_b200.analyze import sobol
    from salib_b200.sample import sobol as sample
    from salib_b200.test_functions import Sobol_G

    z = golden("analyze_g8_n65536_c1_r1000_seed1234")
    D, N, R, seed = int(z["D"]), int(z["N"]), int(z["R"]), int(z["seed"])
    X = sample.sample(gprob(D), N, calc_second_order=False, scramble=False, skip_values=int(z["skip"]))
    Y = Sobol_G.evaluate(X)
    np.testing.assert_allclose(Y[:64], z["Y_head"], rtol=1e-13)
    np.testing.assert_allclose(Y[-64:], z["Y_tail"], rtol=1e-13)
    np.testing.assert_allclose(Y.sum(), float(z["Y_sum"]), rtol=1e-13)
    S = sobol.analyze(gprob(D), Y, calc_second_order=False, num_resamples=R, seed=seed)
    assert S.resample_mode == "numpy"
    _cmp(S, z, D, False, keys=("S1", "S1_conf", "ST", "ST_conf"))
    # the integer engine on the same index stream
    S2 = sobol.analyze_device(gprob(D), Y, calc_second_order=False, num_resamples=R, seed=seed, algo="digits")
    assert S2.engine == "digits" and S2.resample_mode == "numpy"
    _cmp(S2, z, D, False, keys=("S1", "S1_conf", "ST", "ST_conf"))


def test_unseeded_large_problem_uses_philox_and_seeded_too_large_warns(monkeypatch):
    from salib_b200.analyze import sobol

    monkeypatch.setattr(sobol, "NUMPY_SEEDED_MAX", 1 << 12)
    monkeypatch.setattr(sobol, "NUMPY_RESAMPLE_MAX", 1 << 10)
    z = golden("analyze_g8_n512_c1_r64")
    with pytest.warns(UserWarning, match="Philox"):
        S = sobol.analyze(problem_from(z), z["Y"], calc_second_order=False, num_resamples=64, seed=5)
    assert S.resample_mode == "philox"
    S = sobol.analyze(problem_from(z), z["Y"], calc_second_order=False, num_resamples=64)
    assert S.resample_mode == "philox"


def test_outlier_row_auto_falls_back_and_matches_the_reference():
    """One base row scaled by 1e6 (frozen reference output).  The integer engine scales every feature column by its
    maximum, so the other rows lose ~40 bits; "auto" detects that only a handful of rows carry the large values
    (heavy-tail guard) and repeats the call with the FP64 engines -> the reference's 1e-10.  The explicit
    algo="digits" result is kept as the measured bound of what the per-column scaling costs on such data."""
    from salib_b200.analyze import sobol

    z = golden("analyze_g8_outlier_n512_r32")
    p = problem_from(z)
    r = z["r"].astype(np.int64)
    old = sobol.DIGITS_MIN_WORK
    sobol.DIGITS_MIN_WORK = 0           # make "auto" consider the integer engine at this small size
    try:
        S = sobol.analyze_device(p, z["Y"], calc_second_order=True, keep_resamples=True, r=r)
    finally:
        sobol.DIGITS_MIN_WORK = old
    assert S.engine != "digits"
    # This fixture is ill-conditioned in ANY arithmetic: the per-resample indices jump by factors of 1e5 depending
    # on whether the outlier row was drawn (S1_conf ~ 226), and two CPU summation orders -- the reference and the
    # oracle -- already differ by 2.4e-9 on S2_conf.  The FP64 engines land at the same level (1.5e-9); the point
    # estimates, where the outlier is always present, keep 1e-10.
    for k in ("S1", "ST"):
        np.testing.assert_allclose(S[k], z[k], rtol=1e-10, atol=1e-13, err_msg=k)
    for k in ("S1_conf", "ST_conf", "S1_conf_all", "ST_conf_all"):
        np.testing.assert_allclose(S[k], z[k], rtol=2e-8, atol=1e-13, err_msg=k)
    iu = np.triu_indices(8, 1)
    np.testing.assert_allclose(S["S2"][iu], z["S2"][iu], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(S["S2_conf"][iu], z["S2_conf"][iu], rtol=1e-7)
    F = sobol.analyze_device(p, z["Y"], calc_second_order=True, keep_resamples=True, r=r, algo="digits")
    assert F.engine == "digits" and sobol.guard_trips(F.digits_guard, 18, 512)
    # point estimates are sums over ALL rows (the outlier is always in): unaffected.  Resamples that miss the outlier
    # row are left with the cut entries: that is where the loss shows.
    np.testing.assert_allclose(F["S1"], z["S1"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(F["ST"], z["ST"], rtol=1e-9, atol=1e-12)
    worst = max(float(np.max(np.abs(F[k] - z[k]))) for k in ("S1_conf_all", "ST_conf_all"))
    assert worst < 1e-2, worst
    print("outlier row, forced digits: worst per-resample deviation", worst)


def test_benign_output_is_not_flagged_by_the_guard():
    from salib_b200.analyze import sobol
    from salib_b200.sample import sobol as sample
    from salib_b200.test_functions import Sobol_G

    D = 16
    X = sample.sample(gprob(D), 4096, scramble=False, skip_values=4096)
    Y = Sobol_G.evaluate(X, a=np.resize([0, 1, 4.5, 9, 99, 99, 99, 99], D).astype(float))
    S = sobol.analyze_device(gprob(D), Y, num_resamples=64, seed=3, resample="philox", algo="digits")
    assert not sobol.guard_trips(S.digits_guard, 2 * D + 2, 4096)


@pytest.mark.parametrize("name", ["sample_sha_d512_n8_c0_plain", "sample_sha_d512_n8_c0_scr",
                                  "sample_sha_d1024_n4_c0_plain", "sample_sha_d1024_n4_c0_scr",
                                  "sample_sha_d512_n2_c1_plain", "sample_sha_d512_n2_c1_scr"])
def test_wide_sample_matrices_bit_exact(name):
    """cfg4 goes to D = 1024: bit-exact against the reference's matrices (by SHA-256; scrambled with scipy's
    frozen LMS / shift state as host-supplied state) and against the oracle."""
    from oracle import saltelli as o_saltelli
    from salib_b200.sample import sobol as sample

    z = golden(name)
    D, N, c2, skip = int(z["D"]), int(z["N"]), bool(z["c2"]), int(z["skip"])
    if bool(z["scramble"]):
        X = sample.sample(gprob(D), N, calc_second_order=c2, scramble=True, skip_values=skip, seed=int(z["seed"]))
    else:
        X = sample.sample(gprob(D), N, calc_second_order=c2, scramble=False, skip_values=skip)
        assert np.array_equal(X, o_saltelli.sample(gprob(D), N, c2, skip_values=skip, closed_form=True))
    assert tuple(X.shape) == tuple(z["shape"])
    assert np.array_equal(X[0], z["first_row"]) and np.array_equal(X[-1], z["last_row"])
    assert hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest() == str(z["sha256"])


@pytest.mark.parametrize("D", [207, 208])
def test_engine_selection_boundary(D):
    """Second order, Dg = 207 is the widest problem the integer engine's feature kernels hold in shared memory;
    208 must be refused cleanly and served by the FP64 engines.  Both against the frozen reference."""
    from salib_b200.analyze import sobol

    z = golden(f"analyze_g{D}_n32_r4")
    p = gprob(D)
    r = z["r"].astype(np.int64)
    iu = np.triu_indices(D, 1)

    def check(S, rtol, atol):
        for k in ("S1", "ST", "S1_conf", "ST_conf", "S1_conf_all", "ST_conf_all"):
            np.testing.assert_allclose(S[k], z[k], rtol=rtol, atol=atol, err_msg=k)
        np.testing.assert_allclose(S["S2"][iu], z["S2_upper"], rtol=rtol, atol=atol)
        np.testing.assert_allclose(S["S2_conf"][iu], z["S2_conf_upper"], rtol=rtol, atol=atol)

    old = sobol.DIGITS_MIN_WORK
    sobol.DIGITS_MIN_WORK = 0
    try:
        S = sobol.analyze_device(p, z["Y"], calc_second_order=True, keep_resamples=True, r=r)
    finally:
        sobol.DIGITS_MIN_WORK = old
    assert (S.engine == "digits") == (D == 207), S.engine
    # 32 base rows of a 200-dimensional function: S2 is cancellation of O(1) terms
    check(S, 1e-9, 1e-11)
    if D == 208:
        with pytest.raises(Exception):
            sobol.analyze_device(p, z["Y"], calc_second_order=True, r=r, algo="digits")


def test_slim_and_wide_philox_launches_draw_the_same_multiplicities():
    import torch
    from salib_b200 import _lib, _runtime as rt

    N, count = 1 << 17, 300
    for row0, nrows in [(0, N), (N // 8 * 3, N // 8), (1000, 50001)]:
        npad = int(_lib.lib().sa_pad_rows(nrows))
        a = torch.full((count, npad), 9, dtype=torch.uint8, device="cuda")
        b = torch.full((count, npad), 7, dtype=torch.uint8, device="cuda")
        for buf, slim in ((a, 0), (b, 1)):
            _lib.call("sa_counts_philox_rows_ex", rt.stream_ptr(), 1234567, N, 5, count, row0, nrows, rt.ptr(buf), 0,
                      slim)
        torch.cuda.synchronize()
        assert torch.equal(a, b)
        if nrows == N:
            assert bool((a[:, :N].sum(dim=1, dtype=torch.int64) == N).all())
            full = a
        else:                        # a row window is the corresponding slice of the full draw
            assert torch.equal(a[:, :nrows], full[:, row0:row0 + nrows])


def test_pipelined_batches_equal_one_batch(monkeypatch):
    """Several GEMM batches (second buffer drawn by the slim kernel on the side stream while the first is in the
    tensor cores) give bit-identical results to a single batch, and agree with the FP64 engine."""
    from salib_b200.analyze import sobol
    from salib_b200.sample import sobol as sample
    from salib_b200.test_functions import Sobol_G

    D = 12
    p = gprob(D)
    X = sample.sample(p, 8192, scramble=False, skip_values=8192)
    Y = Sobol_G.evaluate(X, a=np.linspace(0.0, 20.0, D))
    kw = dict(num_resamples=700, seed=11, resample="philox", keep_resamples=True)
    one = sobol.analyze_device(p, Y, algo="digits", **kw)
    monkeypatch.setattr(sobol, "DIGITS_BATCH_BYTES", 256 * 8192)       # 256 resamples per batch -> 3 batches
    many = sobol.analyze_device(p, Y, algo="digits", **kw)
    for k in one:
        assert np.array_equal(np.asarray(one[k]), np.asarray(many[k]), equal_nan=True), k
    # a host Y: every batch is drawn by the stand-alone kernel behind the upload; a device Y: batches 1.. are drawn
    # by spare warps of the previous batch's GEMM.  Same multiplicities, same results.
    import torch
    dev = sobol.analyze_device(p, torch.from_numpy(Y).cuda(), algo="digits", **kw)
    assert (many.draws, dev.draws) == ("predrawn", "hosted")
    for k in one:
        assert np.array_equal(np.asarray(one[k]), np.asarray(dev[k]), equal_nan=True), k
    # two outputs share the multiplicities of every batch
    both = sobol.analyze_outputs(p, [Y, 2.0 * Y + 1.0], algo="digits", **kw)
    for k in one:
        assert np.array_equal(np.asarray(one[k]), np.asarray(both[0][k]), equal_nan=True), k
        np.testing.assert_allclose(both[1][k], one[k], rtol=1e-9, atol=1e-12)
    ref = sobol.analyze_device(p, Y, algo="fp64", **kw)
    _cmp(many, ref, D, True, rtol=1e-9, atol=1e-12)
    # row shards on one GPU walk the same pipeline (reduce over local shards instead of ranks)
    step = 2 * D + 2
    cuts = [0, 3000, 3001, 8192]
    sh = sobol.analyze_row_sharded(p, [(lo, Y[lo * step:hi * step]) for lo, hi in zip(cuts, cuts[1:])], 8192,
                                   algo="digits", **kw)
    _cmp(sh, one, D, True, rtol=1e-11, atol=1e-13)


def test_device_combined_moments_match_the_host_formula():
    import torch
    from salib_b200 import _lib, _runtime as rt
    from salib_b200.analyze.sobol import combine_moments

    rng = np.random.default_rng(3)
    parts = []
    for n in (1000.0, 37.0, 0.0, 512.0):
        st = np.array([rng.normal(), abs(rng.normal()) + 0.1, -3.0 - rng.random(), 4.0 + rng.random(),
                       -1.0 - rng.random(), 2.0 + rng.random(), 0.0, 0.0])
        parts.append((n, st))
    rec = np.array([[n, *st] for n, st in parts])
    out = torch.zeros(9, dtype=torch.float64, device="cuda")
    _lib.call("sa_combine_moments_f64", rt.stream_ptr(), rt.ptr(torch.from_numpy(rec).cuda()), len(parts), rt.ptr(out))
    want = combine_moments([(n, st) for n, st in parts if n > 0])
    got = out.cpu().numpy()
    np.testing.assert_allclose(got[:6], want[:6], rtol=1e-14)
    assert got[8] == 1549.0


def test_one_collective_ci_matches_numpy_std():
    """sa_conf_local_f64 + sa_conf_combine_f64 (the row-sharded CI with one all-gather) against np.std(ddof=1) of the
    whole estimate matrix, for uneven shares including an empty one; the trailing doubles are summed."""
    import torch
    from scipy.stats import norm
    from salib_b200 import _lib, _runtime as rt

    rng = np.random.default_rng(11)
    ncols, nextra = 77, 5
    est = rng.normal(size=(1000, ncols)) * rng.uniform(1e-6, 10.0, size=ncols) + rng.normal(size=ncols) * 100.0
    shares = [(0, 400), (400, 400), (400, 993), (993, 1000)]
    pitch = 1 + 2 * ncols + nextra + 3          # a pitch wider than the record
    recs = torch.zeros((len(shares), pitch), dtype=torch.float64, device="cuda")
    extras = rng.normal(size=(len(shares), nextra))
    for g, (lo, hi) in enumerate(shares):
        e = torch.from_numpy(np.ascontiguousarray(est[lo:hi])).cuda()
        _lib.call("sa_conf_local_f64", rt.stream_ptr(), rt.ptr(e), hi - lo, ncols, 1, rt.ptr(recs[g]))
        recs[g, 1 + 2 * ncols:1 + 2 * ncols + nextra] = torch.from_numpy(extras[g]).cuda()
    Z = norm.ppf(0.975)
    conf = torch.empty(ncols, dtype=torch.float64, device="cuda")
    extra = torch.empty(nextra, dtype=torch.float64, device="cuda")
    _lib.call("sa_conf_combine_f64", rt.stream_ptr(), rt.ptr(recs), len(shares), pitch, ncols, nextra, float(Z),
              rt.ptr(conf), rt.ptr(extra))
    np.testing.assert_allclose(conf.cpu().numpy(), Z * est.std(axis=0, ddof=1), rtol=1e-13)
    np.testing.assert_allclose(extra.cpu().numpy(), extras.sum(axis=0), rtol=1e-14, atol=1e-15)
    assert recs[2, 0].item() == 593.0 and recs[1, 0].item() == 0.0
    # row slices of one matrix (how a single rank fills the GPU): 7 slices of 1000 rows, the last one short
    sl = torch.zeros((7, 1 + 2 * ncols), dtype=torch.float64, device="cuda")
    e = torch.from_numpy(est).cuda()
    _lib.call("sa_conf_local_f64", rt.stream_ptr(), rt.ptr(e), 1000, ncols, 7, rt.ptr(sl))
    _lib.call("sa_conf_combine_f64", rt.stream_ptr(), rt.ptr(sl), 7, 1 + 2 * ncols, ncols, 0, float(Z), rt.ptr(conf), 0)
    np.testing.assert_allclose(conf.cpu().numpy(), Z * est.std(axis=0, ddof=1), rtol=1e-13)
    assert sl[:, 0].sum().item() == 1000.0 and sl[6, 0].item() == 1000 - 6 * 143


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from salib_b200.analyze import sobol
z = np.load({fixture!r})
p = {{"num_vars": 16, "names": [f"x{{i+1}}" for i in range(16)], "bounds": [[0.0, 1.0]] * 16}}
sobol.DIGITS_MIN_WORK = 0
# the drop-in, routed by itself under the process group: row shards + integer engine
S = sobol.analyze(p, z["Y"], calc_second_order=True, num_resamples=int(z["R"]), conf_level=float(z["conf"]),
                  keep_resamples=True, seed=int(z["seed"]))
assert S.engine == "digits" and S.resample_mode == "numpy", (S.engine, S.resample_mode)
iu = np.triu_indices(16, 1)
for k in ("S1", "S1_conf", "ST", "ST_conf", "S1_conf_all", "ST_conf_all"):
    np.testing.assert_allclose(S[k], z[k], rtol=1e-10, atol=1e-13, err_msg=k)
for k in ("S2", "S2_conf"):
    np.testing.assert_allclose(S[k][iu], z[k][iu], rtol=1e-10, atol=1e-13, err_msg=k)
# seedless calls: every rank must use rank 0's draw (same result on both ranks)
T = sobol.analyze(p, z["Y"], calc_second_order=True, num_resamples=50)
t = torch.tensor(np.concatenate([T["S1_conf"], T["ST_conf"]]), device="cuda")
ts = [torch.empty_like(t) for _ in range(2)]
dist.all_gather(ts, t)
assert torch.equal(ts[0], ts[1])
# resample-id shards (FP64 engines) still work
U = sobol.analyze_device(p, z["Y"], calc_second_order=True, keep_resamples=True, r=z["r"].astype(np.int64), algo="fp64")
np.testing.assert_allclose(U["S1_conf"], z["S1_conf"], rtol=1e-10, atol=1e-13)
dist.destroy_process_group()
print("RANK_OK", rank)
'''


def test_two_rank_drop_in_matches_the_frozen_reference(tmp_path):
    import subprocess
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, fixture=os.path.join(ROOT, "tests", "golden", "analyze_g16_n256_r32.npz")))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "RANK_OK 0" in out.stdout and "RANK_OK 1" in out.stdout
