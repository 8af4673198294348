"""Drop-in for ``SALib.analyze.dgsm`` (src/SALib/analyze/dgsm.py:8-141) on the device.

``vi = mean(g_j^2)``, ``g_j = (Y_j - Y_base) / (x_j - x_base)``, is linear in the bootstrap multiplicities, so
the confidence intervals reuse the machinery of ``sobol.analyze``: per-row features ``g_j^2`` (one kernel),
resample multiplicities, the skinny-GEMM kernel, ``Z * std(ddof=1)``.

Resampling: the reference draws one ``[R, N]`` index matrix PER FACTOR from the global NumPy RNG
(``np.random.seed(seed)`` if seed, :50-51,133).  ``RESAMPLE_MODE`` "numpy" reproduces exactly that (results match
to ~1e-13); "philox" draws one set of multiplicities on the device and shares it between the factors
(statistically equivalent); "auto" switches at NUMPY_RESAMPLE_MAX index draws."""
import numpy as np
import torch
from scipy.stats import norm

from .. import _lib, _runtime as rt
from ..util import ResultDict
from . import sobol as _sobol

__all__ = ["analyze"]

RESAMPLE_MODE = "auto"
NUMPY_RESAMPLE_MAX = 2 ** 24


def _feature_sums(F, N, nfeat, w, count, sms):
    """[count, nfeat] weighted column sums of the chunked feature matrix F."""
    nsplit = _sobol.plan_row_splits(3, count, N, sms, nfeat)
    psum = rt.empty((nsplit * count * nfeat,), torch.float64)
    _lib.call("sa_feature_sums_f64", rt.stream_ptr(), rt.ptr(F), N, nfeat, rt.ptr(w), int(count), int(nsplit),
              rt.ptr(psum))
    out = rt.empty((count, nfeat), torch.float64)
    _lib.call("sa_reduce_splits_f64", rt.stream_ptr(), rt.ptr(psum), int(count), nfeat, int(nsplit), 1.0 / N,
              rt.ptr(out))
    return out


def analyze(problem, X, Y, num_resamples=100, conf_level=0.95, print_to_console=False, seed=None):
    """Derivative-based global sensitivity measure; returns ``vi, vi_std, dgsm, dgsm_conf`` (+ ``names``)."""
    if seed:
        np.random.seed(seed)
    D = problem["num_vars"]
    Y_size = int(Y.numel()) if isinstance(Y, torch.Tensor) else int(np.size(Y))
    if Y_size % (D + 1) == 0:
        N = Y_size // (D + 1)
    else:
        raise RuntimeError("Incorrect number of samples in model output file.")
    if not 0 < conf_level < 1:
        raise RuntimeError("Confidence level must be between 0-1.")

    rt.require_cuda()
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(rt.device()).multi_processor_count
    Xd = rt.to_device(X, torch.float64).reshape(-1)
    Yd = rt.to_device(Y, torch.float64).reshape(-1)
    nfeat = 2 * D
    F = rt.empty((N * int(L.sa_chunked_feature_doubles(nfeat)),), torch.float64)
    _lib.call("sa_dgsm_features_f64", rt.stream_ptr(), rt.ptr(Xd), rt.ptr(Yd), N, D, rt.ptr(F))

    # point values: mean(g^2), std(g^2), var(base)
    m_dev = _feature_sums(F, N, nfeat, None, 1, sms)
    # np.std(dfdx) (:109) is two-pass: squared deviations from the mean, not E[g^4] - E[g^2]^2, which cancels
    # catastrophically for nearly constant derivatives (a linear model must give exactly 0)
    ss = rt.empty((D,), torch.float64)
    _lib.call("sa_feature_centered_ss_f64", rt.stream_ptr(), rt.ptr(F), N, nfeat, D, rt.ptr(m_dev), rt.ptr(ss))
    vi = m_dev[0].cpu().numpy()[:D]
    vi_std = np.sqrt(ss.cpu().numpy() / N)
    base = Yd[0::D + 1].contiguous()
    stats = rt.zeros((8,), torch.float64)
    work = rt.empty((int(L.sa_moments_workspace_bytes()),), torch.uint8)
    _lib.call("sa_moments_f64", rt.stream_ptr(), rt.ptr(base), N, 1, rt.ptr(stats), rt.ptr(work), int(work.numel()))
    var_base = float(stats[1].item()) ** 2
    b = np.array([[bb[0], bb[1]] for bb in problem["bounds"]], dtype=np.float64)
    dgsm = vi * (b[:, 1] - b[:, 0]) ** 2 / (var_base * np.pi ** 2)

    # bootstrap of the means
    R = int(num_resamples)
    Z = norm.ppf(0.5 + conf_level / 2.0)
    Npad = int(L.sa_pad_rows(N))
    mode = RESAMPLE_MODE
    if mode == "auto":
        mode = "numpy" if N * R * D <= NUMPY_RESAMPLE_MAX else "philox"
    conf = np.full(D, np.nan)
    if R > 0:
        w = rt.empty((R * Npad,), torch.uint8)
        flag = rt.zeros((1,), torch.int32)
        if mode == "numpy":
            for j in range(D):  # one index matrix per factor, in the reference's draw order
                r = np.random.randint(N, size=(R, N))
                rd = rt.to_device(np.ascontiguousarray(r.T, dtype=np.int64), torch.int64)  # [N, R] as sobol draws it
                _lib.call("sa_counts_from_indices", rt.stream_ptr(), rt.ptr(rd), N, R, 0, R, rt.ptr(w), rt.ptr(flag))
                s = _feature_sums(F, N, nfeat, w, R, sms)
                cj = rt.empty((nfeat,), torch.float64)
                _lib.call("sa_sobol_conf_f64", rt.stream_ptr(), rt.ptr(s), R, nfeat, float(Z), rt.ptr(cj))
                conf[j] = float(cj[j].item())
        else:
            key = int(np.random.randint(0, 2 ** 62))
            _lib.call("sa_counts_philox", rt.stream_ptr(), key, N, 0, R, rt.ptr(w), rt.ptr(flag))
            s = _feature_sums(F, N, nfeat, w, R, sms)
            cj = rt.empty((nfeat,), torch.float64)
            _lib.call("sa_sobol_conf_f64", rt.stream_ptr(), rt.ptr(s), R, nfeat, float(Z), rt.ptr(cj))
            conf[:] = cj[:D].cpu().numpy()

    S = ResultDict((("vi", vi.copy()), ("vi_std", vi_std), ("dgsm", dgsm), ("dgsm_conf", conf)))
    S["names"] = problem["names"]
    if print_to_console:
        print(S.to_df())
    return S
