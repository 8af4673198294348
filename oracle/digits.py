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
