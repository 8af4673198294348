"""Exact-integer form of the bootstrap sums (oracle, NumPy + Python integers) -- TEST INFRASTRUCTURE ONLY.

CPU restatement of what ``salib_b200/csrc/digits.cu`` computes, used to pin the device digit matrix and the
device sums bit for bit.  The sums themselves are those of /root/reference/src/SALib/analyze/sobol.py:209-246
in the multiplicity form:

    psum[c][q] = sum_i w[c][i] * F[i][q]

with the per-row features F in SA_NACC order (include/salib_b200.h):
    A, B, A^2, B^2, A*B, B*(AB_j - A) (:219), (A - AB_j)^2 (:232), BA_j * AB_k for j < k (:242).

Every feature column is scaled by a power of two from its exact maximum, rounded once (half to even) to an
(8*S-2)-bit integer and written as S balanced radix-256 digits; integer sums are then exact.
"""
import numpy as np


def features(Y, Dg, calc_second_order, mean, std):
    """F [N, SA_NACC] from the raw model output Y [N*step] and the statistics the device normalises with."""
    step = 2 * Dg + 2 if calc_second_order else Dg + 2
    M = (np.asarray(Y, dtype=np.float64).reshape(-1, step) - mean) / std    # analyze/sobol.py:141
    A, B, AB = M[:, 0], M[:, step - 1], M[:, 1:1 + Dg]
    cols = [A, B, A * A, B * B, A * B]
    cols += [B * (AB[:, j] - A) for j in range(Dg)]
    cols += [(A - AB[:, j]) * (A - AB[:, j]) for j in range(Dg)]
    if calc_second_order:
        BA = M[:, 1 + Dg:1 + 2 * Dg]
        cols += [BA[:, j] * AB[:, k] for j in range(Dg) for k in range(j + 1, Dg)]
    return np.stack(cols, axis=1)


GUARD_BITS = 7


def guard_count(Y, Dg, calc_second_order, mean, std, vmin, vmax):
    """Heavy-tail guard of the integer engine (digits.cu: feature_max_kernel): the number of normalised values
    within GUARD_BITS bits of the design's largest |x|.  ``vmin``/``vmax``: the global min / max of Y."""
    step = 2 * Dg + 2 if calc_second_order else Dg + 2
    M = (np.asarray(Y, dtype=np.float64).reshape(-1, step) - mean) / std
    thr = np.ldexp(max(abs((vmin - mean) / std), abs((vmax - mean) / std)), -GUARD_BITS)
    return int(np.count_nonzero(np.abs(M) > thr))


def split(F, S):
    """(digits int8 [nfeat*S, N] with row q*S+s = digit s of feature q, inv float64 [nfeat], t int64 [N, nfeat])."""
    N, nfeat = F.shape
    bits = 8 * S - 2
    m = np.max(np.abs(F), axis=0)
    e = np.zeros(nfeat, dtype=np.int64)
    nz = (m > 0) & np.isfinite(m)
    e[nz] = np.frexp(m[nz])[1]                 # m = f * 2^e, 0.5 <= f < 1  ->  m < 2^e
    e = np.clip(e, -900, 900)
    scale = np.ldexp(1.0, (bits - e).astype(np.int64))
    inv = np.ldexp(1.0, (e - bits).astype(np.int64))
    t = np.rint(F * scale).astype(np.int64)    # one rounding, half to even
    digits = np.empty((nfeat * S, N), dtype=np.int8)
    rest = t.copy()
    for s in range(S):
        d = ((rest & 0xFF) ^ 0x80) - 0x80      # low byte, sign-extended: in [-128, 127]
        digits[s::S, :] = d.T.astype(np.int8)
        rest = (rest - d) >> 8
    assert not rest.any(), "value does not fit S balanced digits"
    return digits, inv, t


def sums(w, digits, inv, S):
    """psum [count, nfeat]: exact integer digit sums (what the int8 GEMM accumulates in int32), recombined with
    Python integers and rounded to float64 once, then scaled by the power of two `inv`."""
    w = np.asarray(w, dtype=np.int64)
    acc = w @ digits.astype(np.int64).T        # [count, nfeat*S]; |entries| <= 128 * sum(w) -- exact in int64
    assert np.abs(acc).max(initial=0) < 2 ** 31, "int32 accumulator would overflow"
    count, nfeat = acc.shape[0], acc.shape[1] // S
    out = np.empty((count, nfeat), dtype=np.float64)
    for c in range(count):
        for q in range(nfeat):
            v = 0
            for s in range(S - 1, -1, -1):
                v = v * 256 + int(acc[c, q * S + s])
            out[c, q] = float(v) * inv[q]      # int -> float is correctly rounded; the scaling is exact
    return out


def estimates(psum, N, Dg, calc_second_order):
    """Estimator algebra of analyze/sobol.py:219,232,242-246 on the weighted sums of one resample
    (SA_NACC order) -> (S1 [Dg], ST [Dg], S2 [pairs] or None)."""
    n = float(N)
    m = (psum[0] + psum[1]) / (2.0 * n)
    V = (psum[2] + psum[3]) / (2.0 * n) - m * m
    S1 = (psum[5:5 + Dg] / n) / V
    ST = 0.5 * (psum[5 + Dg:5 + 2 * Dg] / n) / V
    S2 = None
    if calc_second_order:
        mab = psum[4] / n
        pairs = [(j, k) for j in range(Dg) for k in range(j + 1, Dg)]
        S2 = np.array([(psum[5 + 2 * Dg + p] / n - mab) / V - S1[j] - S1[k] for p, (j, k) in enumerate(pairs)])
    return S1, ST, S2


def analyze(Y, Dg, calc_second_order, r, conf_level=0.95, S=7):
    """analyze/sobol.py:141-182 through the exact-integer route: normalise, split the features into digits, sum
    them against the multiplicities of every column of the index matrix ``r`` (and a row of ones for the point
    estimate) with integer arithmetic, apply the estimator algebra, Z * std(ddof=1) over the resamples."""
    from scipy.stats import norm

    Y = np.asarray(Y, dtype=np.float64)
    step = 2 * Dg + 2 if calc_second_order else Dg + 2
    N = Y.size // step
    F = features(Y, Dg, calc_second_order, Y.mean(), Y.std())
    dig, inv, _ = split(F, S)
    R = r.shape[1]
    w = np.stack([np.bincount(r[:, c], minlength=N) for c in range(R)] + [np.ones(N, dtype=np.int64)])
    ps = sums(w, dig, inv, S)
    est = [estimates(ps[c], N, Dg, calc_second_order) for c in range(R + 1)]
    Z = norm.ppf(0.5 + conf_level / 2)
    out = {"S1": est[R][0], "ST": est[R][1],
           "S1_conf_all": np.array([e[0] for e in est[:R]]), "ST_conf_all": np.array([e[1] for e in est[:R]])}
    out["S1_conf"] = Z * out["S1_conf_all"].std(ddof=1, axis=0)
    out["ST_conf"] = Z * out["ST_conf_all"].std(ddof=1, axis=0)
    if calc_second_order:
        iu = np.triu_indices(Dg, 1)
        out["S2"] = np.full((Dg, Dg), np.nan)
        out["S2_conf"] = np.full((Dg, Dg), np.nan)
        out["S2"][iu] = est[R][2]
        out["S2_conf"][iu] = Z * np.array([e[2] for e in est[:R]]).std(ddof=1, axis=0)
    return out
