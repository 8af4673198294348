"""Sobol' low-discrepancy sequence, Joe-Kuo direction numbers (oracle, NumPy).

Follows /root/reference/src/SALib/sample/sobol_sequence.py:49-100 (direction-vector
recurrence :62-84, Gray-code stepping :86-89) with the 30-bit word size that
``scipy.stats.qmc.Sobol`` uses at /root/reference/src/SALib/sample/sobol.py:107,136
(values are identical to the reference's 31-bit generator while index < 2**30,
SURVEY F3/F4).  Table: salib_b200/data/joe_kuo_6_21201.npz (numbers of
/root/reference/src/SALib/sample/directions.py, packed by tools/make_joe_kuo_table.py).
"""
import os

import numpy as np

BITS = 30
_TABLE = os.path.join(os.path.dirname(__file__), "..", "salib_b200", "data", "joe_kuo_6_21201.npz")


def load_table():
    z = np.load(_TABLE)
    return z["a"], z["s"], z["m"]


def direction_vectors(dims, bits=BITS):
    """sv[dim, j-1] = V_j  (sobol_sequence.py:62-84), as Python ints -> uint32/uint64."""
    a_all, s_all, m_all = load_table()
    if dims > len(a_all) + 1:
        raise ValueError("Error in Sobol sequence: not enough dimensions")
    sv = np.zeros((dims, bits), dtype=np.uint64)
    for i in range(dims):
        V = [0] * (bits + 1)
        if i == 0:
            for j in range(1, bits + 1):
                V[j] = 1 << (bits - j)
        else:
            a = int(a_all[i - 1])
            s = int(s_all[i - 1])
            m = [0] + [int(v) for v in m_all[i - 1, :s]]
            for j in range(1, min(s, bits) + 1):
                V[j] = m[j] << (bits - j)
            for j in range(s + 1, bits + 1):
                v = V[j - s] ^ (V[j - s] >> s)
                for k in range(1, s):
                    if (a >> (s - 1 - k)) & 1:
                        v ^= V[j - k]
                V[j] = v
        sv[i, :] = V[1:]
    return sv


def _lszb(v):
    """1-based index of the least significant zero bit (sobol_sequence.py:94-100)."""
    idx = 1
    while v & 1:
        v >>= 1
        idx += 1
    return idx


def points(n, dims, skip=0, sv=None, shift=None, bits=BITS):
    """Points skip..skip+n-1 as float64 [n, dims].

    Sequential Gray-code walk exactly like the reference (sobol_sequence.py:86-89): the state
    is advanced one XOR per point from index 0, rows before ``skip`` are discarded
    (saltelli.py:157,165 generate N+skip rows and start reading at ``skip``).
    ``sv``/``shift`` override the table for scrambled sequences.
    """
    if sv is None:
        sv = direction_vectors(dims, bits)
    sv = np.asarray(sv, dtype=np.uint64)
    x = np.zeros(dims, dtype=np.uint64) if shift is None else np.asarray(shift, dtype=np.uint64).copy()
    out = np.empty((n, dims), dtype=np.float64)
    scale = float(2 ** bits)
    total = skip + n
    for k in range(total):
        if k > 0:
            x ^= sv[:, _lszb(k - 1) - 1]
        if k >= skip:
            out[k - skip] = x.astype(np.float64) / scale
    return out


def points_closed_form(n, dims, skip=0, sv=None, shift=None, bits=BITS):
    """Vectorised closed form: x_k = shift ^ XOR_{b in gray(k)} sv[:, b].

    Used for large n where the sequential walk is too slow; checked against ``points``
    in tests/test_oracle.py.
    """
    if sv is None:
        sv = direction_vectors(dims, bits)
    sv = np.asarray(sv, dtype=np.uint64)
    k = np.arange(skip, skip + n, dtype=np.uint64)
    g = k ^ (k >> np.uint64(1))
    x = np.zeros((n, dims), dtype=np.uint64)
    if shift is not None:
        x ^= np.asarray(shift, dtype=np.uint64)[None, :]
    for b in range(bits):
        mask = ((g >> np.uint64(b)) & np.uint64(1)).astype(bool)
        if mask.any():
            x[mask] ^= sv[:, b][None, :]
    return x.astype(np.float64) / float(2 ** bits)


def scipy_scramble_state(dims, seed, bits=BITS):
    """LMS + digital shift state of ``qmc.Sobol(dims, scramble=True, seed=int)`` restated in
    NumPy (scipy 1.18.1 ``_qmc.py`` Sobol._scramble).  Returns (sv, shift)."""
    rng = np.random.default_rng(seed)
    sv0 = direction_vectors(dims, bits)
    bit = rng.integers(2, size=(dims, bits), dtype=np.uint32).astype(np.uint64)
    shift = (bit << np.arange(bits, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint64)
    # numpy draws are big-endian over the bit index: shift = dot(bits, 2**arange(bits))
    ltm = np.tril(rng.integers(2, size=(dims, bits, bits), dtype=np.uint32)).astype(np.uint64)
    idx = np.arange(bits)
    ltm[:, idx, idx] = 1
    # pack row p MSB first
    lsm = (ltm << (np.uint64(bits - 1) - np.arange(bits, dtype=np.uint64))[None, None, :]).sum(axis=2)
    sv = np.zeros_like(sv0)
    for p in range(bits):
        t = lsm[:, p][:, None] & sv0  # [dims, bits]
        par = np.zeros_like(t)
        tt = t.copy()
        while tt.any():
            par ^= tt & np.uint64(1)
            tt >>= np.uint64(1)
        sv |= par << np.uint64(bits - 1 - p)
    return sv.astype(np.uint64), shift
