"""Drop-in for ``SALib.analyze.sobol`` on B200 (reference: src/SALib/analyze/sobol.py:23-206).

``analyze`` keeps the reference signature and returns the same ``ResultDict``
(``S1, S1_conf, ST, ST_conf[, S1_conf_all, ST_conf_all][, S2, S2_conf]``, ``.problem``, bound ``.to_df``).
All arithmetic runs in the CUDA library through the C ABI (include/salib_b200.h):

    moments -> normalise/re-layout -> point estimate (weights = 1)
            -> per batch of resamples: multiplicities -> weighted sums -> estimator algebra
            -> Z * std(ddof=1) over resamples

Weighted sums: every sum is linear in the multiplicities, i.e. all of them together are one GEMM
w[R x N] * F[N x SA_NACC] over per-row features (for second order the pair products BA_j * AB_k are features
too) whose left operand holds small integers.  Two families of engines compute it:
  * "digits" (large problems): F is split error-free into int8 radix-256 digits and the GEMM runs exactly on
    the tensor cores (tcgen05 u8 x s8 -> s32 in TMEM), see csrc/digits.cu;
  * "fp64": FP64-pipe kernels -- the skinny feature GEMV (first/total order, second order while
    Dg <= PAIR_GEMV_MAX_DG) and the 8x8 register-tile kernel, which skips rows of multiplicity 0.

Resample indices.  The reference draws ``r = rng.integers(N, size=(N, R))`` on the host
(:113-117,144).  ``RESAMPLE_MODE``:
  * ``"numpy"``  -- draw exactly that matrix on the host (same seed semantics, ``seed`` falsy -> legacy
                    global RNG), in row blocks that continue one generator stream (bounded host memory),
                    upload it, build multiplicities on the device: results match the reference to ~1e-13;
  * ``"philox"`` -- device Philox4x32-10 keyed by (seed, resample id, draw): statistically
                    equivalent CIs, independent of the number of GPUs, no N*R host array;
  * ``"auto"``   -- numpy whenever a ``seed`` is given and the draw is one the reference itself could make
                    (N*R <= NUMPY_SEEDED_MAX = 2^31 draws: the caller asked for the reference's resamples),
                    numpy for small unseeded problems (N*R <= NUMPY_RESAMPLE_MAX), else philox -- with a
                    warning if a seed was given (the CIs then differ from the reference's for that seed; the
                    reference cannot allocate ``r`` there anyway: 84 GB at N=2^20, R=10^4).  The mode that ran
                    is recorded as ``ResultDict.resample_mode``.

Multi-GPU: when a ``torch.distributed`` job is initialised, ``analyze`` picks the decomposition by itself, like
the reference picks serial vs. pool inside ``analyze`` (:147,183): with the integer engine every rank keeps a
base-row range of Y, walks all resamples over it and the partial sums are reduce-scattered (``shard="rows"``);
with the FP64 engines resamples are sharded by resample id and the per-rank estimates are all-gathered
(``shard="resamples"``).  Every rank returns the full result.
"""
import os
from itertools import combinations
from types import MethodType
from warnings import warn

import numpy as np
import pandas as pd
import torch
from scipy.stats import norm

from .. import _lib, _runtime as rt
from ..util import ResultDict, extract_group_names

__all__ = ["analyze", "analyze_device", "analyze_outputs", "analyze_row_sharded", "combine_moments", "allgather_moments", "first_order", "total_order", "second_order", "separate_output_values",
           "to_df", "Si_to_pandas_dict", "CONST_RESULT_MSG"]

CONST_RESULT_MSG = (
    "Constant values encountered, indicating model evaluations "
    "(or subset of evaluations) produced identical values."
)

RESAMPLE_MODE = os.environ.get("SALIB_B200_RESAMPLE", "auto")
NUMPY_RESAMPLE_MAX = 2 ** 24          # unseeded "auto": N*R index draws made on the host (~0.1 s of NumPy RNG)
NUMPY_SEEDED_MAX = 2 ** 31            # seeded "auto": keep the reference's own resamples up to here (17 GB on device)
NUMPY_BLOCK_DRAWS = 2 ** 24           # host block of the streamed draw (128 MB of int64)
MAX_COUNT_BYTES = 12 << 30            # multiplicity buffer per batch (uint8 [count][Npad]), FP64 engines
# integer engine: resamples go through the GEMM in batches of at most this many multiplicity bytes (digits_batch);
# two buffers alternate, the next batch is drawn (slim Philox kernel, side stream) while the current one is in the
# tensor cores
DIGITS_BATCH_BYTES = int(os.environ.get("SALIB_B200_BATCH_BYTES", 8704 << 20))
_ALGO = {"auto": 0, "fp64": -1, "generic": 1, "first_order": 3, "gemv": 3, "tile8": 4, "dmma": 5, "digits": 6}
# Engines.  "digits": the bootstrap sums as an exact integer GEMM on the tensor cores (csrc/digits.cu);
# "fp64": the FP64-pipe engines (feature GEMV / 8x8 register-tile kernel / generic).  "auto" takes "digits"
# from DIGITS_MIN_WORK multiply-adds on (there it is ~10x faster and its sums carry no accumulation error),
# "fp64" below (launch-latency territory) and whenever the integer engine does not apply (N > 2^24, digit
# matrix over DIGITS_MAX_BYTES, non-finite or constant Y).  SALIB_B200_ALGO overrides "auto".
DEFAULT_ALGO = os.environ.get("SALIB_B200_ALGO", "auto")
# Radix-256 digits per feature value: 7 = 54 bits relative to the column maximum, one more than a float64
# significand.  6 digits are 14 % faster but measurably worse than FP64 accumulation on small indices (headline:
# S1 ~ 3e-5 of the a = 99 factors moves by 1.6e-9 relative, against 2.6e-12 with 7 digits), so 7 is the rule;
# SALIB_B200_DIGITS / DIGITS overrides it.
DIGITS = int(os.environ["SALIB_B200_DIGITS"]) if "SALIB_B200_DIGITS" in os.environ else None


def digits_for(N):
    return DIGITS if DIGITS is not None else 7


ENGINE_NAMES = {1: "fp64-generic", 3: "fp64-gemv", 4: "fp64-tile8", 5: "fp64-dmma", 6: "digits"}
# Heavy-tail guard of the integer engine ("auto" only): the digits of a feature column are relative to the
# column maximum, so when fewer than GUARD_MIN_ROWS base rows carry every value within 7 bits of the design's
# largest |x| (an outlier row), a resample can miss them all and be left with entries whose low bits were cut.
# Such outputs are analysed again with the FP64 engines (guard_trips); ~28+ rows are all absent from a resample
# with probability below 1e-12, i.e. never.  The count comes from the feature kernels (inv[nacc]).
GUARD_MIN_ROWS = 32
GUARD_MISS_PROBABILITY = 1e-12


def guard_trips(guard, step, N):
    """Heavy-tail guard: ``guard`` values lie within 7 bits of the design's largest |x|; they sit in at least
    k = guard / step base rows, and a resample of N draws misses all of them with probability (1 - k/N)^N
    (e^-k for large N).  True when that can happen in practice, i.e. the integer engine's column scaling could
    cost a resample its precision."""
    k = min(float(guard) / step, float(N))
    return (1.0 - k / N) ** N > GUARD_MISS_PROBABILITY


DIGITS_MIN_WORK = 1 << 31
DIGITS_MAX_BYTES = 80 << 30
# feature kernels: SA_NACC maxima + 9 doubles per model output of a base row must fit 200 KB of shared memory
# (digits.cu: feature_tile_rows; second order: Dg <= 207, first/total order only: Dg <= 2320)
DIGITS_MAX_FEATURES = 25600
PAIR_GEMV_MAX_DG = 48                 # second order through the feature GEMV up to this many groups ...
PAIR_GEMV_MAX_BYTES = 48 << 30        # ... while N * SA_NACC doubles stay below this


def _nacc(Dg, c2):
    return 5 + 2 * Dg + (Dg * (Dg - 1) // 2 if c2 else 0)


def _nest(Dg, c2):
    return 2 * Dg + (Dg * (Dg - 1) // 2 if c2 else 0)


def _bcast_from_rank0(t):
    """In a torch.distributed job every rank must use rank 0's draw (seedless calls would otherwise resample
    each row shard from a different bootstrap)."""
    rank, world = rt.world()
    if world > 1:
        import torch.distributed as dist

        dist.broadcast(t, src=0)
    return t


class _Resampler:
    """Produces the multiplicity rows w[c][:] for resamples [c0, c0+count) on the device."""

    def __init__(self, N, R, seed, mode, distributed=True):
        self.N, self.R = N, R
        if mode == "auto":
            if seed and N * R <= NUMPY_SEEDED_MAX:
                mode = "numpy"
            elif N * R <= NUMPY_RESAMPLE_MAX:
                mode = "numpy"
            else:
                mode = "philox"
                if seed:
                    warn(f"num_resamples * N = {N * R} index draws exceed what is drawn on the host "
                         f"({NUMPY_SEEDED_MAX}); resample indices come from the device Philox generator: "
                         "confidence intervals are statistically equivalent to, but not identical with, "
                         "the reference's for this seed.")
        if mode not in ("numpy", "philox"):
            raise ValueError(f"unknown resample mode {mode!r}")
        self.mode = mode
        self.flag = rt.zeros((1,), torch.int32)
        world = rt.world()[1] if distributed else 1
        if mode == "numpy":
            # analyze/sobol.py:113-117,144 -- identical draw, C order [N, R]; row blocks of one generator
            # continue the same stream as a single call (tests/test_oracle.py), so host memory stays bounded
            self.r = rt.empty((N, R), torch.int64)
            if seed or world == 1 or rt.world()[0] == 0:
                rng = np.random.default_rng(seed).integers if seed else np.random.randint
                blk = max(1, NUMPY_BLOCK_DRAWS // max(R, 1))
                for i0 in range(0, N, blk):
                    i1 = min(N, i0 + blk)
                    part = np.ascontiguousarray(rng(N, size=(i1 - i0, R)), dtype=np.int64)
                    self.r[i0:i1].copy_(torch.from_numpy(part), non_blocking=False)
            if not seed and world > 1:
                _bcast_from_rank0(self.r)
        else:
            if seed:
                self.key = int(np.random.default_rng(seed).integers(0, 2 ** 63))
            else:
                key = int(np.random.randint(0, 2 ** 62))
                if world > 1:
                    t = torch.tensor([key], dtype=torch.int64, device=rt.device())
                    key = int(_bcast_from_rank0(t).item())
                self.key = key

    @classmethod
    def from_indices(cls, r):
        self = cls.__new__(cls)
        r = np.asarray(r)
        self.N, self.R = r.shape
        self.mode = "numpy"
        self.flag = rt.zeros((1,), torch.int32)
        self.r = rt.to_device(np.ascontiguousarray(r, dtype=np.int64), torch.int64)
        return self

    def counts(self, c0, count, w, rows=None, slim=False):
        """w[count][pad_rows(nrows)] <- multiplicities of resamples [c0, c0+count) over the row window
        ``rows = (row0, nrows)`` (default: all N rows).  ``slim``: the low-footprint Philox launch that runs
        beside the persistent GEMM kernel."""
        if count <= 0:
            return
        row0, nrows = rows if rows is not None else (0, self.N)
        if self.mode == "numpy":
            _lib.call("sa_counts_from_indices_rows", rt.stream_ptr(), rt.ptr(self.r), self.N, self.R, int(c0),
                      int(count), int(row0), int(nrows), rt.ptr(w), rt.ptr(self.flag))
        else:
            _lib.call("sa_counts_philox_rows_ex", rt.stream_ptr(), self.key, self.N, int(c0), int(count),
                      int(row0), int(nrows), rt.ptr(w), rt.ptr(self.flag), int(bool(slim)))

    @staticmethod
    def raise_for(flag):
        if flag & 2:
            raise ValueError("resample indices out of range [0, N)")
        if flag & 1:
            raise OverflowError("a row was drawn more than 255 times in one resample")

    def check(self):
        self.raise_for(int(self.flag.item()))


def tile_groups(Dg):
    """Warps (tile groups of 32 lane tiles) that carry one resample in the tile kernel."""
    TB = -(-Dg // 8)
    return 1 if TB <= 8 else -(-(TB * (TB + 1) // 2) // 32)


def plan_row_splits(kind, count, N, sms, nacc, groups=1):
    """Row splits so that `count` resamples fill `sms` SMs (at least 32 rows per split).  Pure host logic."""
    max_split = max(1, min(4096, N // 32))
    if kind in (4, 5):   # tile kernels: `groups` warps per resample, 8 warps per block, one block per SM
        blocks = -(-count // 8) * groups
    elif kind == 3:      # feature GEMV kernel (launch_feature_gemv in analyze.cu)
        chunks = 1 if nacc <= 24 else -(-nacc // 16)     # chunk_width() in analyze.cu
        four = nacc > 24 or 8 < nacc <= 16
        if count > 256 and four:      # 4 resamples per lane: 512 per block of 4 warps, 2 blocks/SM
            blocks, resident = -(-count // 512) * chunks, 2
        elif count > 256:             # 2 per lane: 256 per block, 3 blocks/SM
            blocks, resident = -(-count // 256) * chunks, 3
        else:                         # one resample per lane, one block of up to 8 warps
            blocks, resident = chunks, 4
        slots = resident * sms        # aim at three full waves of blocks
        return max(1, min(max_split, -(-3 * slots // blocks) if blocks < 3 * slots else 1))
    else:                # generic kernel: one block per resample
        blocks = count
    return max(1, min(max_split, sms // blocks if blocks <= sms else 1))


def plan_batches(kind, lo, hi, Npad, sms, max_bytes=None, groups=1):
    """Resample batches [c0, c1) covering [lo, hi): bounded by the multiplicity buffer (max_bytes); for the
    one-warp-per-resample tile kernel sized in whole waves (8 * sms), the remainder being its own
    (row-split) launch.  Pure host logic."""
    cap = max(1, (max_bytes or MAX_COUNT_BYTES) // Npad)
    if kind == 6:        # integer GEMM: tiles of 128 resamples, one batch as large as the buffer allows
        cap = cap // 128 * 128 or cap
        return [(c, min(c + cap, hi)) for c in range(lo, hi, cap)]
    wave = max(8, 8 * sms // groups) if kind in (4, 5) else 256 * sms
    out, c = [], lo
    while c < hi:
        left = hi - c
        if left >= wave:
            n = min((cap // wave) * wave, (left // wave) * wave) or min(cap, left)
        else:
            n = min(cap, left)
        out.append((c, c + n))
        c += n
    return out


class SobolAnalyzer:
    """Device state of one ``analyze`` call; also usable directly with a device-resident Y."""

    def __init__(self, Dg, N, calc_second_order, algo="auto", resamples=0, total_rows=None):
        """``total_rows``: base rows of the whole design when this engine holds one row shard of it (a resample
        draws that many rows, which bounds the integer engine's int32 sums)."""
        rt.require_cuda()
        self.Dg, self.N, self.c2 = int(Dg), int(N), bool(calc_second_order)
        self.total_rows = int(total_rows) if total_rows is not None else self.N
        self.step = 2 * Dg + 2 if self.c2 else Dg + 2
        self.nacc, self.nest = _nacc(Dg, self.c2), _nest(Dg, self.c2)
        self.algo = _ALGO[DEFAULT_ALGO if algo == "auto" else algo]
        self.ndigits = digits_for(self.N)
        L = _lib.lib()
        self.auto_digits = False
        if self.algo > 0:
            kind = self.algo
        elif self.algo == 0 and self._digits_pays(int(resamples)):
            kind, self.auto_digits = 6, True
        else:
            kind = self._fp64_kind()
        self.sms = torch.cuda.get_device_properties(rt.device()).multi_processor_count
        self._set_kind(kind)
        self.stride = L.sa_row_stride(self.Dg, int(self.c2))
        self.Npad = int(L.sa_pad_rows(self.N))
        self.stats = rt.zeros((8,), torch.float64)
        self.Yn = None
        self.Yt = None
        self.kernel_events = None   # bench.py sets this to a list to time the accumulate launches

    def _fp64_kind(self):
        """First/total order only: feature-GEMV kernel (any Dg).  Second order: the same GEMV over pair-product
        features while Dg is small (the 8x8 register-tile kernel needs Dg >= 57 to occupy all 32 lanes) and the
        feature matrix fits PAIR_GEMV_MAX_BYTES; else the tile kernel (Dg <= 512); else the generic one."""
        if not self.c2:
            return 3
        feat = 8 * self.N * int(_lib.lib().sa_chunked_feature_doubles(self.nacc))
        if self.Dg <= PAIR_GEMV_MAX_DG and feat <= PAIR_GEMV_MAX_BYTES:
            return 3
        return 4 if self.Dg <= 512 else 1

    def _digits_pays(self, resamples):
        """"auto": is the integer engine applicable, worth it, and does its digit matrix fit next to the
        multiplicity buffer in the memory this process can still get?"""
        L = _lib.lib()
        if not L.sa_digit_available():
            return False
        q_bytes = self.nacc * self.ndigits * int(L.sa_digit_row_pitch(self.N))
        # |digit sum| <= 128 * (draws of one resample) must stay below 2^31
        if not (self.N * max(resamples, 1) * self.nacc >= DIGITS_MIN_WORK and self.total_rows <= 1 << 24
                and q_bytes <= DIGITS_MAX_BYTES and self.nacc + 9 * self.step <= DIGITS_MAX_FEATURES):
            return False
        free, _ = torch.cuda.mem_get_info()
        cached = torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
        w_bytes = min(MAX_COUNT_BYTES, (max(resamples, 1) + 1) * int(L.sa_pad_rows(self.N)))
        c_bytes = 4 * min(max(resamples, 1) + 1, MAX_COUNT_BYTES // int(L.sa_pad_rows(self.N)) + 1) \
            * int(L.sa_digit_ldc(self.nacc, self.ndigits))
        return q_bytes + w_bytes + c_bytes + (1 << 30) <= free + cached

    def _set_kind(self, kind):
        self.kind = kind
        self.groups = tile_groups(self.Dg) if kind == 4 else 1

    # -- normalisation ---------------------------------------------------------------------
    def moments(self, Yd):
        """stats[0:6] = mean, std, min, max of Yd and min, max of its A and B columns."""
        L = _lib.lib()
        work = rt.empty((int(L.sa_moments_workspace_bytes()),), torch.uint8)
        _lib.call("sa_moments_f64", rt.stream_ptr(), rt.ptr(Yd), self.N, self.step, rt.ptr(self.stats),
                  rt.ptr(work), int(work.numel()))

    def load(self, Yd, stats=None):
        """Yd: device float64 [N*step] in the reference order -> stats, Yn (and Yt).  ``stats`` (8 doubles, host
        or device) replaces the moments of Yd -- the statistics of the whole design when Yd is one row shard of
        it.  No host synchronisation: whether Y was degenerate (non-finite / constant, which the integer engine
        cannot scale) is read from ``self.stats`` by the caller together with the results (``degenerate``)."""
        L = _lib.lib()
        if stats is None:
            self.moments(Yd)
        elif isinstance(stats, torch.Tensor):
            self.stats.copy_(stats.reshape(-1)[:8])
        else:
            self.stats.copy_(rt.to_device(np.asarray(stats, dtype=np.float64), torch.float64))
        if self.kind == 6:                   # digit columns of the features (int8, K-major) + their scales
            self.Q = rt.empty((self.nacc * self.ndigits * int(L.sa_digit_row_pitch(self.N)),), torch.int8)
            self.inv = rt.empty((self.nacc + 1,), torch.float64)     # column scales + the heavy-tail guard count
        elif self.kind == 3:                 # feature matrix (first order, or second order with pair products)
            self.Yt = rt.empty((self.N * int(L.sa_chunked_feature_doubles(self.nacc)),), torch.float64)
        else:                                # analysis-layout rows (+ the A and B columns for kinds 4, 5)
            self.Yn = rt.zeros((self.N * self.stride + 16,), torch.float64)
            if self.kind in (4, 5):
                self.Yt = rt.empty((2 * self.Npad,), torch.float64)
        try:
            self.load_normalized(Yd)
        except _lib.SalibB200Error:
            if not self.auto_digits:
                raise
            # the integer engine refused the shape: "auto" continues with the FP64 engines
            self.auto_digits = False
            self._set_kind(self._fp64_kind())
            self.Q = self.inv = None
            return self.load(Yd, stats=self.stats.clone())

    @staticmethod
    def degenerate(stats):
        """Host copy of ``stats`` -> True when the integer engine's result must be discarded: NaN/inf/constant
        outputs keep the reference's NaN propagation through the FP64 engines."""
        st = np.asarray(stats, dtype=np.float64)[:4]
        return not (np.all(np.isfinite(st)) and st[1] > 0.0)

    def digit_sums(self, w, count, rows_alloc=None, draw=None):
        """Integer engine: psum [rows_alloc >= count, nacc] float64 of this engine's rows for the multiplicity rows
        w[count][Npad]; rows beyond ``count`` are zero (padding for the reduce-scatter).  ``draw`` = (key, N_total,
        c0, n, w_next): the GEMM also hosts the Philox draw of resamples [c0, c0+n) over this engine's row window
        into ``w_next`` (the next batch; spare warps of the GEMM's CTAs, see sa_digit_sums_draw_f64)."""
        rows_alloc = int(rows_alloc or count)
        psum = rt.empty((rows_alloc, self.nacc), torch.float64)
        if rows_alloc > count:
            psum[count:].zero_()
        work = rt.empty((int(_lib.lib().sa_digit_sums_workspace_bytes(self.N, self.nacc, self.ndigits,
                                                                       int(count))),), torch.uint8)
        ev = None
        if self.kernel_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        key, n_total, c0, n, w_next = draw if draw is not None else (0, 0, 0, 0, None)
        _lib.call("sa_digit_sums_draw_f64", rt.stream_ptr(), rt.ptr(self.Q), rt.ptr(self.inv), self.N, self.nacc,
                  self.ndigits, rt.ptr(w), self.Npad, int(count), rt.ptr(work), int(work.numel()), rt.ptr(psum),
                  int(key), int(n_total), int(c0), int(n), int(getattr(self, "row0", 0)), self.N, rt.ptr(w_next))
        if ev is not None:
            ev[1].record()
            self.kernel_events.append((int(count), ev[0], ev[1]))
        return psum

    def load_normalized(self, Yd):
        """Re-run only the normalise/re-layout step with the current ``self.stats``."""
        if self.kind == 6:
            work = rt.empty((int(_lib.lib().sa_digit_features_workspace_bytes(self.nacc)),), torch.uint8)
            _lib.call("sa_digit_features_i8", rt.stream_ptr(), rt.ptr(Yd), self.N, self.Dg, int(self.c2),
                      rt.ptr(self.stats), self.ndigits, rt.ptr(self.Q), rt.ptr(self.inv), rt.ptr(work),
                      int(work.numel()))
        elif self.kind == 3:
            _lib.call("sa_normalize_features_f64", rt.stream_ptr(), rt.ptr(Yd), self.N, self.Dg, int(self.c2),
                      rt.ptr(self.stats), rt.ptr(self.Yt))
        else:
            _lib.call("sa_normalize_f64", rt.stream_ptr(), rt.ptr(Yd), self.N, self.Dg, int(self.c2),
                      rt.ptr(self.stats), rt.ptr(self.Yn), rt.ptr(self.Yt))

    # -- work decomposition ----------------------------------------------------------------
    def _nsplit(self, count):
        return plan_row_splits(self.kind, count, self.N, self.sms, self.nacc, self.groups)

    def _partial_sums(self, w, count):
        """psum [nsplit, count, nacc] of this engine's rows; returns (psum, nsplit)."""
        nsplit = 1 if self.kind == 6 else self._nsplit(count)
        psum = rt.empty((nsplit * count * self.nacc,), torch.float64)
        if self.kind == 6:
            if w is None:                    # point estimate: every row once
                w = torch.ones((self.Npad,), dtype=torch.uint8, device=rt.device())
            work = rt.empty((int(_lib.lib().sa_digit_sums_workspace_bytes(self.N, self.nacc, self.ndigits,
                                                                           int(count))),), torch.uint8)
        ev = None
        if self.kernel_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self.kind == 6:
            _lib.call("sa_digit_sums_f64", rt.stream_ptr(), rt.ptr(self.Q), rt.ptr(self.inv), self.N, self.nacc,
                      self.ndigits, rt.ptr(w), self.Npad, int(count), rt.ptr(work), int(work.numel()), rt.ptr(psum))
        else:
            _lib.call("sa_sobol_accumulate_ex_f64", rt.stream_ptr(), rt.ptr(self.Yn), rt.ptr(self.Yt), self.N,
                      self.Dg, int(self.c2), rt.ptr(w), int(count), int(nsplit), rt.ptr(psum), self.kind)
        if ev is not None:
            ev[1].record()
            self.kernel_events.append((int(count), ev[0], ev[1]))
        return psum, nsplit

    def _finalize(self, psum, nsplit, count, n_rows, est_out):
        _lib.call("sa_sobol_finalize_f64", rt.stream_ptr(), rt.ptr(psum), int(n_rows), self.Dg, int(self.c2),
                  int(count), int(nsplit), rt.ptr(self.stats), rt.ptr(est_out))

    def _accumulate(self, w, count, est_out):
        """est_out [count, nest] <- estimates of resamples whose multiplicities are w (None = all ones)."""
        psum, nsplit = self._partial_sums(w, count)
        self._finalize(psum, nsplit, count, self.N, est_out)

    def point(self):
        est = rt.empty((1, self.nest), torch.float64)
        self._accumulate(None, 1, est)
        return est

    def _batches(self, lo, hi):
        return plan_batches(self.kind, lo, hi, self.Npad, self.sms, groups=self.groups)

    def bootstrap(self, resampler, lo, hi):
        """Estimates of resamples [lo, hi) -> device [hi-lo, nest]."""
        return bootstrap_shared([self], resampler, lo, hi)[0]

    def conf(self, est, z):
        """z * std(est, axis=0, ddof=1) -> [nest].  A long estimate matrix is reduced in row slices that fill the
        GPU (one column block per CTA reads est twice from 67 CTAs at the headline: 0.3 ms) and the slices' moments
        are combined (Chan et al.), the same two calls the row-sharded route uses across ranks."""
        R = int(est.shape[0])
        out = rt.empty((self.nest,), torch.float64)
        slices = min(16, R // 512)
        if slices >= 2:
            recs = rt.empty((slices, 1 + 2 * self.nest), torch.float64)
            _lib.call("sa_conf_local_f64", rt.stream_ptr(), rt.ptr(est), R, self.nest, slices, rt.ptr(recs))
            _lib.call("sa_conf_combine_f64", rt.stream_ptr(), rt.ptr(recs), slices, 1 + 2 * self.nest, self.nest, 0,
                      float(z), rt.ptr(out), 0)
        else:
            _lib.call("sa_sobol_conf_f64", rt.stream_ptr(), rt.ptr(est), R, self.nest, float(z), rt.ptr(out))
        return out


class _CountsPrefetch:
    """Multiplicities of the first resample batch, drawn on a side stream while the main stream is still busy
    with Y (host-to-device copy, moments, feature kernels): they depend on the seed only, not on Y."""

    def __init__(self, resampler, batch, targets):
        """batch = (c0, c1); targets = [(row pitch, (row0, nrows) or None)]: one buffer per row shard."""
        self.c0, self.n = batch[0], batch[1] - batch[0]
        # one spare row of ones behind the batch: the point estimate, if the engine takes it along (_ones_row)
        self.ws = [rt.empty(((self.n + 1) * npad,), torch.uint8) for npad, _ in targets]
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(main)          # the cached blocks may still be read by earlier work on the main stream
        with torch.cuda.stream(side):
            for w, (npad, rows) in zip(self.ws, targets):
                resampler.counts(self.c0, self.n, w, rows=rows)
                w[self.n * npad:].fill_(1)
            self.done = torch.cuda.Event()
            self.done.record(side)

    def take(self, c0, n):
        """The buffers if they hold exactly batch [c0, c0+n), after making the current stream wait for them."""
        if (c0, n) != (self.c0, self.n):
            return None
        torch.cuda.current_stream().wait_event(self.done)
        return self.ws


def bootstrap_shared(engines, resampler, lo, hi, prefetch=None, points=None):
    """Estimates of resamples [lo, hi) for several model outputs of the same design (same N, Dg, order):
    the multiplicities of each batch are built once and every output's rows are accumulated against them.
    ``points`` (an empty list): with the integer engine the point estimate -- one multiplicity row of ones --
    rides in the first batch's GEMM instead of reading the digit matrix in a launch of its own; its
    ``[1, nest]`` estimates are appended per engine (the list stays empty if that does not apply)."""
    lead = engines[0]
    ests = [rt.empty((hi - lo, e.nest), torch.float64) for e in engines]
    fuse = points is not None and all(e.kind == 6 for e in engines)
    w = None
    for c0, c1 in lead._batches(lo, hi):
        n = c1 - c0
        ready = prefetch.take(c0, n) if prefetch is not None else None
        if ready is not None:
            w = ready[0]
        else:
            if w is None or w.numel() < (n + 1) * lead.Npad:
                w = rt.empty(((n + 1) * lead.Npad,), torch.uint8)
            resampler.counts(c0, n, w)
            if fuse:
                w[n * lead.Npad:(n + 1) * lead.Npad].fill_(1)
        for e, est in zip(engines, ests):
            if fuse:
                ext = rt.empty((n + 1, e.nest), torch.float64)
                e._accumulate(w, n + 1, ext)
                est[c0 - lo:c1 - lo] = ext[:n]
                points.append(ext[n:n + 1].clone())
            else:
                e._accumulate(w, n, est[c0 - lo:c1 - lo])
        fuse = False
    if prefetch is not None:
        torch.cuda.current_stream().wait_event(prefetch.done)   # never leave the side stream unjoined
    return ests


def allgather_estimates(est_local, R, rank, world):
    """All-gather the per-rank [R_g, nest] estimate blocks (contiguous shares of resample ids, see
    ``_runtime.shard_range``) into the full [R, nest] tensor on every rank."""
    import torch.distributed as dist

    nest = est_local.shape[1]
    per = -(-R // world)
    pad = torch.zeros((per, nest), dtype=est_local.dtype, device=est_local.device)
    pad[: est_local.shape[0]] = est_local
    full = torch.empty((world * per, nest), dtype=est_local.dtype, device=est_local.device)
    dist.all_gather_into_tensor(full, pad)
    parts = []
    for g in range(world):
        lo, hi = rt.shard_range(R, g, world)
        parts.append(full[g * per: g * per + (hi - lo)])
    return torch.cat(parts, dim=0)


def combine_moments(parts):
    """Statistics of the union of row shards from per-shard ``(n_values, stats8)`` (pairwise update of
    Chan, Golub & LeVeque): mean, population std, min, max, min/max over the A and B columns."""
    n = float(sum(c for c, _ in parts))
    mean = sum(c * s[0] for c, s in parts) / n
    m2 = sum(c * (s[1] * s[1] + (s[0] - mean) ** 2) for c, s in parts)
    out = np.zeros(8)
    out[0], out[1] = mean, np.sqrt(m2 / n)
    out[2], out[3] = min(s[2] for _, s in parts), max(s[3] for _, s in parts)
    out[4], out[5] = min(s[4] for _, s in parts), max(s[5] for _, s in parts)
    return out


def allgather_moments(local, world):
    """Every rank's per-shard ``(n_values, stats8)`` list -> the list over all ranks (tiny all-gather;
    ranks may hold different numbers of shards)."""
    if world == 1:
        return list(local)
    import torch.distributed as dist

    n_max = torch.tensor([len(local)], dtype=torch.int64, device=rt.device())
    dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
    n_max = int(n_max.item())
    mine = torch.zeros((n_max, 9), dtype=torch.float64, device=rt.device())
    for i, (c, st) in enumerate(local):
        mine[i, 0] = c
        mine[i, 1:] = torch.as_tensor(np.asarray(st, dtype=np.float64))
    everyone = torch.empty((world * n_max, 9), dtype=torch.float64, device=rt.device())
    dist.all_gather_into_tensor(everyone, mine)
    return [(row[0], row[1:]) for row in everyone.cpu().numpy() if row[0] > 0]


class _Agreement:
    """One small all-reduce (MAX) of quantities every rank must agree on before it issues collectives (engine family,
    batch sizes), started now and read later: the collective and its read-back run on the side stream, so device
    work queued on the main stream in between (the upload of Y, the moments) neither waits for it nor is waited for.
    ``result()`` -> the agreed python ints; identity without a process group."""

    def __init__(self, values_max, world=None):
        self.values = [int(v) for v in values_max]
        self.t = self.event = None
        if world is None:
            world = rt.world()[1]
        if world == 1:
            return
        import torch.distributed as dist

        if not torch.cuda.is_available():    # the gloo tests
            self.t = torch.tensor(self.values, dtype=torch.int64)
            dist.all_reduce(self.t, op=dist.ReduceOp.MAX)
            return
        side = _side_stream()
        with torch.cuda.stream(side):
            t = torch.tensor(self.values, dtype=torch.int64, device=rt.device())
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.t = torch.empty(t.shape, dtype=torch.int64, pin_memory=True)
            self.t.copy_(t, non_blocking=True)
            t.record_stream(side)
            self.event = torch.cuda.Event()
            self.event.record(side)

    def result(self):
        if self.t is None:
            return self.values
        if self.event is not None:
            self.event.synchronize()
        return [int(v) for v in self.t.numpy()]


def _agree(values_max):
    """Start and read an ``_Agreement`` at once."""
    return _Agreement(values_max).result()


class _RankRecords:
    """The ONE collective in front of a row-sharded analyze: every rank contributes a record of 12 doubles -- the
    moments of its rows ``(n_values, stats8)`` (``sa_rank_record_f64`` combines several local shards) and three
    numbers the ranks must agree on before they issue further collectives (engine family, largest shard, spare) --
    in one all-gather.  The moments of the whole design are combined from the gathered records on the device
    (``stats9``); the three numbers are copied to pinned memory behind the collective and reduced with max on the
    host (``result``, the only host wait before the final read-back).  The round-2 first version asked with an
    all-reduce and gathered the moments separately: two waits for the slowest rank instead of one."""

    def __init__(self, parts, k, flags, world):
        self.world = world
        rec = rt.empty((12,), torch.float64)
        _lib.call("sa_rank_record_f64", rt.stream_ptr(), rt.ptr(parts), int(k), float(flags[0]), float(flags[1]),
                  float(flags[2]), rt.ptr(rec))
        import torch.distributed as dist

        self.all = rt.empty((world, 12), torch.float64)
        dist.all_gather_into_tensor(self.all.view(-1), rec)
        self.flags = torch.empty((world, 3), dtype=torch.float64, pin_memory=True)
        self.flags.copy_(self.all[:, 9:], non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    @classmethod
    def decline(cls, world):
        """A rank that cannot take the integer engine (or holds no rows) still joins the collective, so that every
        other rank learns it; returns after the answer is in."""
        self = cls(rt.zeros((1, 9), torch.float64), 1, (1, 0, 0), world)
        self.result()

    def result(self):
        self.event.synchronize()
        f = self.flags.numpy()
        return [int(f[:, j].max()) for j in range(3)]

    def stats9(self):
        out = rt.empty((9,), torch.float64)
        _lib.call("sa_combine_moments_pitch_f64", rt.stream_ptr(), rt.ptr(self.all), self.world, 12, rt.ptr(out))
        return out


DIGITS_BATCH_MIN_MACS = 3.5e12        # ~2 ms of GEMM: amortises the launches of a batch


def digits_batch(R, npad_max, ncols=15043):
    """Resamples per GEMM batch of the integer engine.  About a quarter of the call, so that three of four draws
    hide behind a GEMM (rounded up to whole 256-resample tiles: R = 10^4 runs 4 batches of 2560, not 5 of 2048 --
    every batch re-reads the digit matrix, 15.8 GB at the headline; same-box A/B 106.3-107.9 -> 105.2 ms), but at least DIGITS_BATCH_MIN_MACS of multiply-adds (a batch of B resamples is
    B * npad * ncols of them; small problems stay in one batch) and at most DIGITS_BATCH_BYTES of multiplicities
    (two buffers of that size live at once); whole 256-resample tiles, batches of equal size.  A function of R, the
    largest row shard of the job and the column count only: every rank derives the same batches."""
    npad = max(int(npad_max), 1)
    want = max(R // 4, int(DIGITS_BATCH_MIN_MACS / (npad * max(int(ncols), 1))))
    want = min(-(-want // 256) * 256, DIGITS_BATCH_BYTES // npad)   # a quarter, rounded UP to whole tiles
    cap = max(256, want // 256 * 256)
    if R <= cap:
        return max(R, 1)
    nb = -(-R // cap)
    return min(cap, -(-(-(-R // nb)) // 256) * 256)


def digits_batches(R, npad_max, ncols=15043, first=None):
    """The [c0, c1) resample ranges of the integer engine's GEMM batches.  Equal batches of ``digits_batch``; with
    ``first`` (module default DIGITS_FIRST_BATCH, a multiple of 256 below the batch size) the first batch is that
    short: it is the only one whose multiplicities are drawn by a kernel of their own, in front of the first GEMM
    -- every later batch is drawn inside the previous batch's GEMM."""
    B = digits_batch(R, npad_max, ncols)
    first = DIGITS_FIRST_BATCH if first is None else int(first)
    if not first or R <= B or first >= B:
        return [(c, min(c + B, R)) for c in range(0, R, B)] or [(0, 0)]
    rest = R - first
    nb = -(-rest // B)
    Bs = min(B, -(-(-(-rest // nb)) // 256) * 256)
    return [(0, first)] + [(c, min(c + Bs, R)) for c in range(first, R, Bs)]


DIGITS_FIRST_BATCH = int(os.environ.get("SALIB_B200_FIRST_BATCH", "0"))
PREDRAW_MAX_BYTES = int(float(os.environ.get("SALIB_B200_PREDRAW_GB", "24")) * (1 << 30))


def _free_bytes():
    """Device memory this process can still get: free + cached by the allocator."""
    free, _ = torch.cuda.mem_get_info()
    return free + torch.cuda.memory_reserved() - torch.cuda.memory_allocated()

_SIDE_STREAMS = {}


def _side_stream():
    """Stream of the first batch's multiplicity draw (and of the index-matrix counting kernels), which run beside
    the upload of Y and the feature kernels.  One per device, created once."""
    dev = torch.cuda.current_device()
    if dev not in _SIDE_STREAMS:
        _SIDE_STREAMS[dev] = torch.cuda.Stream()
    return _SIDE_STREAMS[dev]


def batch_share(n, count, world, g):
    """Reduce-scatter bookkeeping of one batch of ``count`` rows (``n`` resamples + possibly the row of ones) over
    ``world`` ranks: rank g keeps rows [lo, lo + chunk) of the batch, of which the first ``valid`` are resamples.
    Pure host logic -> (chunk, lo, valid)."""
    chunk = -(-count // world)
    lo = g * chunk
    return chunk, lo, max(0, min(n - lo, chunk))


class _DigitsRows:
    """The integer engine's analyze over base-row shards -- ONE code path for one GPU (a single shard that holds
    all rows) and for a torch.distributed job (every rank holds a row range).

    Per batch of resamples: multiplicities of the batch over every local row window (Philox or the reference's
    own index matrix) -> exact integer GEMM against the window's digit matrix -> float64 partial sums
    ``[count, nacc]`` -> (multi-GPU) reduce-scatter over resamples, so each rank runs the estimator algebra on its
    share -> estimates.  Two multiplicity buffers alternate: while batch b is in the tensor cores, batch b+1 is
    drawn by spare warps of the same GEMM CTAs (Philox; ``sa_digit_sums_draw_f64``) -- or, for the reference's own
    index matrix, by the counting kernel on a side stream.  The point estimate is the multiplicity row of ones
    appended to the last batch.  The CI is ``Z * std(ddof=1)`` over all
    R estimates: two-pass on one GPU; across ranks two all-reduces of ``nest`` doubles (column sums, then squared
    deviations) replace the all-gather of the estimates.  No host synchronisation before the final read-back.
    """

    def __init__(self, outputs, N, R, rs, npad_max, rank, world, upload=False, after=None):
        """outputs[o] = the engines (kind 6, ``row0`` set) of model output o, one per local row shard; all outputs
        share the row shards (and therefore the multiplicities).  ``upload``: Y is coming from host memory -- the
        GPU idles behind PCIe for about as long as all multiplicities take to draw, so (Philox, memory permitting)
        every batch gets its own buffer and is drawn by the stand-alone kernel on the side stream right now, and
        the GEMMs host nothing (a hosted draw costs a power-capped GEMM ~4 % of its clock).  ``after``: an event
        on the main stream behind which the side stream may start (default: everything queued on the main stream so
        far) -- the caller recorded it before it queued work the draws need not wait for (the moments, the ranks'
        all-gather) and has released no large block since."""
        self.outputs, self.N, self.R, self.rs = outputs, int(N), int(R), rs
        self.rank, self.world = rank, world
        self.shards = outputs[0]
        self.lead = outputs[0][0]
        self.batches = digits_batches(self.R, npad_max, self.lead.nacc * self.lead.ndigits)
        self.nbuf = 2 if len(self.batches) > 1 else 1
        rows_buf = max(c1 - c0 for c0, c1 in self.batches) + 1
        philox = rs is not None and rs.mode == "philox"
        all_bytes = len(self.batches) * rows_buf * sum(e.Npad for e in self.shards)
        self.predrawn = bool(upload and philox and len(self.batches) > 2 and all_bytes <= PREDRAW_MAX_BYTES
                             and all_bytes + (8 << 30) <= _free_bytes())
        if self.predrawn:
            self.nbuf = len(self.batches)
        self.wsets = [[rt.empty((rows_buf * e.Npad,), torch.uint8) for e in self.shards] for _ in range(self.nbuf)]
        self.main = torch.cuda.current_stream()
        self.side = _side_stream()
        if after is not None:                # cached blocks may still be read by earlier work on the main stream
            self.side.wait_event(after)
        else:
            self.side.wait_stream(self.main)
        self.drawn, self.gemm_done = {}, {}
        # Philox draws of batches 1.. are hosted by the previous batch's GEMM (main stream, no events)
        self.hosted = philox and not self.predrawn and not os.environ.get("SALIB_B200_SIDE_DRAW")
        self.draw(0, slim=False)             # the first batch is drawn while Y is uploaded / split into digits
        if self.predrawn:
            for b in range(1, len(self.batches)):
                self.draw(b, slim=False)

    def draw(self, b, slim, after=None):
        c0, c1 = self.batches[b]
        n = c1 - c0
        last = b == len(self.batches) - 1
        with torch.cuda.stream(self.side):
            if after is not None:            # not before the GEMM it is meant to run beside: the draw would take
                self.side.wait_event(after)  # the SMs from the feature kernels on the critical path
            if b >= self.nbuf:               # the GEMM that read this buffer two batches ago
                self.side.wait_event(self.gemm_done[b - self.nbuf])
            for w, e in zip(self.wsets[b % self.nbuf], self.shards):
                if self.rs is not None:
                    self.rs.counts(c0, n, w, rows=(e.row0, e.N), slim=slim)
                if last:                     # the point estimate: every row once
                    w[n * e.Npad:(n + 1) * e.Npad].fill_(1)
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.drawn[b] = ev

    @property
    def draws(self):
        """How the batches after the first got their multiplicities (reported as ``ResultDict.draws``)."""
        return "predrawn" if self.predrawn else "hosted" if self.hosted else "side-stream"

    def abandon(self):
        """Join the side stream before the buffers go back to the allocator (the engine fell back to FP64)."""
        self.main.wait_stream(self.side)

    def run(self, Z, keep):
        """-> per output (point [nest], conf [nest], kept [R, 2Dg] or None) as device/host arrays."""
        import torch.distributed as dist

        lead, world, rank, R = self.lead, self.world, self.rank, self.R
        nacc, nest, Dg = lead.nacc, lead.nest, lead.Dg
        nout = len(self.outputs)
        chunk_max = max(-(-(c1 - c0 + 1) // world) for c0, c1 in self.batches)
        est_loc = [rt.empty((-(-R // world) + len(self.batches) + 2 * chunk_max, nest), torch.float64)
                   for _ in range(nout)]
        # multi-GPU: the record every rank contributes to the ONE collective behind the last batch -- per output
        # (n_local, mean[nest], M2[nest]) of its share of the estimates, the point estimate (zero on every rank but
        # the one whose share holds the row of ones) and the guard count
        stride = 3 * nest + 2
        rec = rt.zeros((nout, stride), torch.float64)
        point = rec[:, 1 + 2 * nest:1 + 3 * nest]
        pos, ids = 0, []
        pending = None

        def finish(item):
            nonlocal pos
            b, n, count, chunk, sums, works = item
            for work in works:
                work.wait()                  # the current stream waits for the collective; the host does not
            _, lo, valid = batch_share(n, count, world, rank)
            for o, engs in enumerate(self.outputs):
                engs[0]._finalize(sums[o], 1, chunk, self.N, est_loc[o][pos:pos + chunk])
                if count > n and lo <= n < lo + chunk:          # the row of ones sits in this rank's share
                    point[o].copy_(est_loc[o][pos + n - lo])
            ids.append((self.batches[b][0] + lo, valid))
            pos += valid

        for b, (c0, c1) in enumerate(self.batches):
            n = c1 - c0
            count = n + (1 if b == len(self.batches) - 1 else 0)
            nxt = b + 1 < len(self.batches)
            if b in self.drawn:
                self.main.wait_event(self.drawn[b])
            if nxt and not self.hosted and not self.predrawn:
                pre = torch.cuda.Event()
                pre.record(self.main)        # fires when this batch's GEMM is next on the main stream
                self.draw(b + 1, slim=True, after=pre)
            chunk = -(-count // world)
            sums = []
            for o, engs in enumerate(self.outputs):
                tot = None
                for i, (e, w) in enumerate(zip(engs, self.wsets[b % self.nbuf])):
                    hosted = None
                    if nxt and self.hosted and o == 0:     # this shard's GEMM draws the shard's next batch
                        n0, n1 = self.batches[b + 1]
                        hosted = (self.rs.key, self.N, n0, n1 - n0, self.wsets[(b + 1) % self.nbuf][i])
                    ps = e.digit_sums(w, count, rows_alloc=chunk * world, draw=hosted)
                    tot = ps if tot is None else tot.add_(ps)
                sums.append(tot)
            if nxt and self.hosted and b + 1 == len(self.batches) - 1:
                n_last = self.batches[b + 1][1] - self.batches[b + 1][0]
                for w, e in zip(self.wsets[(b + 1) % self.nbuf], self.shards):   # the point estimate's row of ones
                    w[n_last * e.Npad:(n_last + 1) * e.Npad].fill_(1)
            ev = torch.cuda.Event()
            ev.record(self.main)
            self.gemm_done[b] = ev
            works = []
            if world > 1:
                # partial sums of all ranks' rows, scattered along the resample axis: rank g keeps rows
                # [g * chunk, (g + 1) * chunk) of the batch
                outs = []
                for o in range(nout):
                    out = rt.empty((chunk, nacc), torch.float64)
                    works.append(dist.reduce_scatter_tensor(out.view(-1), sums[o].view(-1), async_op=True))
                    outs.append(out)
                sums_local = outs
            else:
                sums_local = sums
            if pending is not None:
                finish(pending)              # estimator algebra of the previous batch, behind this batch's GEMM
            pending = (b, n, count, chunk, sums_local, works)
        finish(pending)
        self.main.wait_stream(self.side)     # never leave the side stream unjoined

        # heavy-tail guard: values near the design's largest |x|, over all local shards (and ranks)
        self.guard = torch.stack([torch.stack([e.inv[e.nacc] for e in engs]).sum() for engs in self.outputs])
        confs = []
        if world == 1 or R == 0:
            for o in range(nout):
                confs.append(lead.conf(est_loc[o][:R], Z) if R else
                             torch.full((nest,), float("nan"), dtype=torch.float64, device=rt.device()))
            if world > 1:
                both = torch.cat([point.reshape(-1), self.guard])
                dist.all_reduce(both)
                point = both[:nout * nest].reshape(nout, nest)
                self.guard = both[nout * nest:]
        else:
            # CI = Z * std(ddof=1) over all R estimates without gathering them, point estimates and guard counts:
            # one all-gather of 3 nest + 2 doubles per output and rank, combined on every rank (Chan et al.)
            rec[:, stride - 1] = self.guard
            for o in range(nout):
                _lib.call("sa_conf_local_f64", rt.stream_ptr(), rt.ptr(est_loc[o]), int(pos), nest, 1, rt.ptr(rec[o]))
            recs = rt.empty((world, nout, stride), torch.float64)
            dist.all_gather_into_tensor(recs.view(-1), rec.view(-1))
            extra = rt.empty((nout, nest + 1), torch.float64)
            for o in range(nout):
                conf = rt.empty((nest,), torch.float64)
                _lib.call("sa_conf_combine_f64", rt.stream_ptr(), rt.ptr(recs[0, o]), world, nout * stride, nest,
                          nest + 1, float(Z), rt.ptr(conf), rt.ptr(extra[o]))
                confs.append(conf)
            point = extra[:, :nest]
            self.guard = extra[:, nest].contiguous()
        kept = [None] * nout
        if keep:
            for o in range(nout):
                if world == 1:
                    kept[o] = est_loc[o][:R, :2 * Dg].cpu().numpy()
                    continue
                rows_max = -(-R // world) + len(self.batches)
                mine = torch.zeros((rows_max, 2 * Dg), dtype=torch.float64, device=rt.device())
                mine[:pos] = est_loc[o][:pos, :2 * Dg]
                everyone = torch.empty((world * rows_max, 2 * Dg), dtype=torch.float64, device=rt.device())
                dist.all_gather_into_tensor(everyone, mine)
                everyone = everyone.cpu().numpy().reshape(world, rows_max, 2 * Dg)
                full = np.zeros((R, 2 * Dg))
                for g in range(world):       # the same bookkeeping as finish(), for every rank
                    at = 0
                    for c0, c1 in self.batches:
                        n = c1 - c0
                        _, lo, valid = batch_share(n, n + (1 if c1 == R else 0), world, g)
                        full[c0 + lo:c0 + lo + valid] = everyone[g, at:at + valid]
                        at += valid
                kept[o] = full
        return [(point[o], confs[o], kept[o]) for o in range(nout)]


def analyze_row_sharded(problem, shards, N, calc_second_order=True, num_resamples=100, conf_level=0.95,
                        keep_resamples=False, seed=None, resample=None, r=None, algo="auto", distributed=True,
                        kernel_events=None, auto=False):
    """``analyze`` for a model output partitioned by base-row range ("when Y exceeds one GPU") --
    and the route of the integer engine in general (one GPU = one shard with every row).

    ``shards``: the ``(row0, Y_local)`` pieces this process holds, ``Y_local`` = the ``n_local*step`` outputs
    of base rows ``[row0, row0+n_local)`` (host or CUDA); over all ranks of the torch.distributed job the
    pieces tile ``[0, N)``.  Every rank walks all resamples over its own rows: the multiplicity kernels
    keep the draws that land in the shard's row window, the accumulate kernels produce partial sums
    ``[count, nacc]``, which are added across ranks (integer engine: reduce-scatter, see ``_DigitsRows``; FP64
    engines: all-reduce) before the estimator algebra.  Statistics for the normalisation come from per-shard
    moments combined on the device (``sa_combine_moments_f64``).  Engine family, batch sizes and the seedless
    resample key are agreed across ranks before any collective is issued.
    """
    import torch.distributed as dist

    _, Dg = extract_group_names(problem)
    c2 = bool(calc_second_order)
    step = 2 * Dg + 2 if c2 else Dg + 2
    if not 0 < conf_level < 1:
        raise RuntimeError("Confidence level must be between 0-1.")
    rt.require_cuda()
    rank, world = rt.world() if distributed else (0, 1)
    R_req = np.shape(r)[1] if r is not None else int(num_resamples)

    def build(algo_):
        engs = []
        for row0, Y in shards:
            size = int(Y.numel()) if isinstance(Y, torch.Tensor) else int(np.size(Y))
            if size % step or size == 0:
                raise RuntimeError("row shard size is not a multiple of the Saltelli block length")
            eng = SobolAnalyzer(Dg, size // step, c2, algo=algo_, resamples=R_req, total_rows=N)
            if eng.kind == 6 and N > 1 << 24:
                raise ValueError("the integer engine is exact up to 2^24 base rows per resample; use algo='fp64'")
            eng.kernel_events = kernel_events
            eng.row0 = int(row0)
            engs.append(eng)
        return engs

    # the side stream (multiplicity draws) may start behind this point: nothing queued from here on -- upload,
    # moments, the ranks' all-gather -- is anything a draw has to wait for, and no large block is released before
    # the pipeline allocates its buffers
    entry = torch.cuda.Event()
    entry.record()
    engines = build(algo)
    # every rank must take the same engine family and the same batches: what each rank would do on its
    # own travels in the one all-gather of the moments (``_RankRecords``)
    mine = [0 if all(e.kind == 6 for e in engines) else 1, max(e.Npad for e in engines), len(engines)]
    rows_here = sum(e.N for e in engines)
    if world == 1 and rows_here != N:        # with several ranks the count is checked on the combined moments
        raise RuntimeError(f"row shards hold {rows_here} base rows, expected {N}")
    Z = norm.ppf(0.5 + conf_level / 2)
    upload = not isinstance(shards[0][1], torch.Tensor)
    ncols = engines[0].nacc * engines[0].ndigits

    # moments of the whole design: per-shard records -> (several ranks) one record per rank, all-gathered ->
    # combined on the device
    loaded = []
    parts = rt.zeros((len(engines), 9), torch.float64)
    for i, (eng, (_, Y)) in enumerate(zip(engines, shards)):
        Yd = rt.to_device(Y, torch.float64, non_blocking=True).reshape(-1)
        eng.moments(Yd)
        loaded.append(Yd)
        parts[i, 0].fill_(float(Yd.numel()))     # a fill kernel, not a pageable host-to-device copy
        parts[i, 1:].copy_(eng.stats)
    records = _RankRecords(parts, len(engines), mine, world) if world > 1 else None

    def resampler():
        if r is not None:
            rs_ = _Resampler.from_indices(r)
            if rs_.N != N:
                raise ValueError(f"`r` must have {N} rows")
            return rs_, rs_.R
        R_ = int(num_resamples)
        return (_Resampler(N, R_, seed, resample or RESAMPLE_MODE, distributed=distributed) if R_ > 0 else None), R_

    # host work that needs no answer yet (the collective is in flight): the resampler -- unless it needs a
    # collective of its own (the seedless key is rank 0's, broadcast), which must not be issued before every rank
    # is known to take this route -- and, if this rank would take the integer engine, its pipeline, whose first
    # multiplicity batch starts on the side stream right away
    early = world == 1 or r is not None or bool(seed)
    rs, R, pipe = None, None, None
    if early:
        rs, R = resampler()
        if not mine[0]:
            pipe = _DigitsRows([engines], N, R, rs, mine[1], rank, world, upload=upload, after=entry)

    not_digits, npad_max, _ = records.result() if records is not None else mine
    if not_digits and pipe is not None:
        pipe.abandon()
        pipe = None
    if not_digits and auto:
        return None                          # "auto": some rank cannot take the integer engine -> resample shards
    if not early:
        rs, R = resampler()
    if not_digits and any(e.kind == 6 for e in engines):
        first = engines
        engines = build("fp64")             # some rank (or shard) cannot run the integer engine: nobody does
        for eng, was in zip(engines, first):
            eng.stats.copy_(was.stats)
    digits = not not_digits
    lead = engines[0]

    prefetch = None
    if digits and (pipe is None or pipe.batches != digits_batches(R, npad_max, ncols)):
        if pipe is not None:
            pipe.abandon()                   # a larger shard elsewhere changes the batches: every rank uses the same
        pipe = _DigitsRows([engines], N, R, rs, npad_max, rank, world, upload=upload)
    elif not digits and R > 0 and rs.mode == "philox":
        # device-drawn multiplicities do not depend on Y: start the first batch now, beside the upload of the shards
        first = plan_batches(lead.kind, 0, R, npad_max, lead.sms, groups=lead.groups)[0]
        prefetch = _CountsPrefetch(rs, first, [(e.Npad, (e.row0, e.N)) for e in engines])

    if world > 1 or len(engines) > 1:
        if records is not None:
            stats9 = records.stats9()
        else:
            stats9 = rt.empty((9,), torch.float64)
            _lib.call("sa_combine_moments_f64", rt.stream_ptr(), rt.ptr(parts), int(parts.shape[0]), rt.ptr(stats9))
        for eng, Yd in zip(engines, loaded):
            eng.load(Yd, stats=stats9)
    else:
        stats9 = None
        lead.load(loaded[0], stats=lead.stats)      # the moments are already there
    del loaded

    def again_fp64():
        return analyze_row_sharded(problem, shards, N, calc_second_order=calc_second_order,
                                   num_resamples=num_resamples, conf_level=conf_level,
                                   keep_resamples=keep_resamples, seed=seed, resample=resample, r=r, algo="fp64",
                                   distributed=distributed, kernel_events=kernel_events)

    if digits and not all(e.kind == 6 for e in engines):
        # the feature kernels refused the shape (a function of Dg and the order only, so every rank lands here)
        pipe.abandon()
        return again_fp64()
    if digits:
        (point, conf, kept), = pipe.run(Z, bool(keep_resamples))
        flag = rs.flag if rs is not None else None
    else:
        point, conf, kept, flag = _row_sharded_fp64(engines, N, R, rs, Z, bool(keep_resamples), npad_max, prefetch,
                                                    world)

    # ONE read-back (one synchronisation): results, statistics, row count, flags, guard
    nest = lead.nest
    one = torch.ones((1,), dtype=torch.float64, device=rt.device())
    host = torch.cat([point.reshape(-1)[:nest], conf.reshape(-1)[:nest], lead.stats.reshape(-1)[:8],
                      stats9[8:9] if stats9 is not None else one * float(N * step),
                      flag.to(torch.float64) if flag is not None else one * 0.0,
                      pipe.guard[:1] if digits else one * 0.0]).cpu().numpy()
    point_h, conf_h, stats = host[:nest], host[nest:2 * nest], host[2 * nest:2 * nest + 8]
    total = int(round(float(host[2 * nest + 8])))
    if total != N * step:
        raise RuntimeError(f"row shards hold {total // step} base rows, expected {N}")
    _Resampler.raise_for(int(host[2 * nest + 9]))
    guard = float(host[2 * nest + 10]) if digits else None
    if digits and lead.auto_digits and (SobolAnalyzer.degenerate(stats) or guard_trips(guard, step, N)):
        # non-finite / constant outputs: the reference's behaviour comes from the FP64 engines;
        # an outlier row: the FP64 engines keep every row's own exponent (GUARD_MIN_ROWS)
        return again_fp64()

    const_y = not (stats[4] < stats[5])
    if const_y:
        warn(CONST_RESULT_MSG)
    keep = bool(keep_resamples)
    if keep and kept is None:
        kept = np.zeros((0, 2 * Dg))
    S = _assemble(Dg, c2, keep, point_h, conf_h, kept, const_y)
    S.problem = problem
    S.to_df = MethodType(to_df, S)
    S.rows_here = rows_here
    S.engine = ENGINE_NAMES[lead.kind]
    S.resample_mode = rs.mode if rs is not None else None
    S.digits_guard = guard
    S.draws = pipe.draws if digits else None
    return S


def _row_sharded_fp64(engines, N, R, rs, Z, keep, npad_max, prefetch, world):
    """Row shards through the FP64 engines: partial sums of every batch are all-reduced, then the estimator
    algebra runs on the totals on every rank.  Returns (point, conf, kept, flag) on the device / host."""
    import torch.distributed as dist

    lead = engines[0]

    def totals(ws, count, est_out):
        tot = rt.zeros((count * lead.nacc,), torch.float64)
        red = rt.empty((count * lead.nacc,), torch.float64)
        for eng, w in zip(engines, ws):
            psum, nsplit = eng._partial_sums(w, count)
            _lib.call("sa_reduce_splits_f64", rt.stream_ptr(), rt.ptr(psum), int(count), lead.nacc, int(nsplit),
                      1.0, rt.ptr(red))
            tot += red
        if world > 1:
            dist.all_reduce(tot)
        lead._finalize(tot, 1, count, N, est_out)

    point = rt.empty((1, lead.nest), torch.float64)
    totals([None] * len(engines), 1, point)
    est = None
    if R > 0:
        est = rt.empty((R, lead.nest), torch.float64)
        wbuf = {}
        # batches from quantities every rank shares (the largest shard of the job), not from this rank's shard
        for c0, c1 in plan_batches(lead.kind, 0, R, npad_max, lead.sms, groups=lead.groups):
            n = c1 - c0
            ws = prefetch.take(c0, n) if prefetch is not None else None
            if ws is not None:
                wbuf = dict(enumerate(ws))
            else:
                ws = []
                for i, eng in enumerate(engines):
                    if i not in wbuf or wbuf[i].numel() < (n + 1) * eng.Npad:
                        wbuf[i] = rt.empty(((n + 1) * eng.Npad,), torch.uint8)
                    rs.counts(c0, n, wbuf[i], rows=(eng.row0, eng.N))
                    ws.append(wbuf[i])
            totals(ws, n, est[c0:c1])
        if prefetch is not None:
            torch.cuda.current_stream().wait_event(prefetch.done)
        conf = lead.conf(est, Z)
    else:
        conf = torch.full((lead.nest,), float("nan"), dtype=torch.float64, device=point.device)
    kept = None
    if keep:
        kept = est[:, :2 * lead.Dg].cpu().numpy() if est is not None else None
    return point, conf, kept, (rs.flag if rs is not None else None)


def _assemble(Dg, c2, keep, point, conf, est_all, const_y):
    S = ResultDict((k, np.zeros(Dg)) for k in ("S1", "S1_conf", "ST", "ST_conf"))
    R = 0 if est_all is None else est_all.shape[0]
    if keep:
        S["S1_conf_all"] = np.zeros((R, Dg))
        S["ST_conf_all"] = np.zeros((R, Dg))
    if c2:
        S["S2"] = np.full((Dg, Dg), np.nan)
        S["S2_conf"] = np.full((Dg, Dg), np.nan)
    iu = np.triu_indices(Dg, 1)
    if const_y:
        # analyze/sobol.py:214-217 et al.: every estimator returns 0.0; the CI of a scalar 0.0 is
        # 0 for S1/ST (:157-161) and NaN for S2 (std(ddof=1) of one value, :180-182)
        if c2:
            S["S2"][iu] = 0.0
        return S
    S["S1"][:] = point[:Dg]
    S["ST"][:] = point[Dg:2 * Dg]
    S["S1_conf"][:] = conf[:Dg]
    S["ST_conf"][:] = conf[Dg:2 * Dg]
    if keep:
        S["S1_conf_all"][:] = est_all[:, :Dg]
        S["ST_conf_all"][:] = est_all[:, Dg:2 * Dg]
    if c2:
        S["S2"][iu] = point[2 * Dg:]
        S["S2_conf"][iu] = conf[2 * Dg:]
    return S


def _n_from_size(size, Dg, calc_second_order):
    if calc_second_order and size % (2 * Dg + 2) == 0:
        return size // (2 * Dg + 2)
    if not calc_second_order and size % (Dg + 2) == 0:
        return size // (Dg + 2)
    raise RuntimeError(
        """
        Incorrect number of samples in model output file.
        Confirm that calc_second_order matches option used during sampling."""
    )


def analyze_outputs(problem, Ys, calc_second_order=True, num_resamples=100, conf_level=0.95,
                    keep_resamples=False, seed=None, resample=None, r=None, algo="auto", distributed=True,
                    kernel_events=None):
    """Analyze several model outputs of one design in one pass: ``Ys`` is a list of [N*step] arrays
    (host or CUDA).  All outputs share one resample draw -- with a ``seed`` that is exactly what the
    reference's per-output calls produce (``ProblemSpec.analyze``, util/problem.py:368-371, re-seeds each
    call); without a seed the reference would redraw per output, which is statistically equivalent.
    Returns a list of ResultDicts.  In a torch.distributed job the resamples are sharded by resample id and the
    estimates all-gathered (the row-sharded route of the integer engine is ``analyze_device(shard="rows")``)."""
    _, Dg = extract_group_names(problem)
    sizes = {int(Y.numel()) if isinstance(Y, torch.Tensor) else int(np.size(Y)) for Y in Ys}
    if len(sizes) != 1:
        raise ValueError("all outputs must have the same number of model evaluations")
    N = _n_from_size(sizes.pop(), Dg, calc_second_order)
    if not 0 < conf_level < 1:
        raise RuntimeError("Confidence level must be between 0-1.")

    rt.require_cuda()
    rank, world = rt.world() if distributed else (0, 1)
    R_req = np.shape(r)[1] if r is not None else int(num_resamples)

    def build(algo_):
        engs = []
        for _ in Ys:
            eng = SobolAnalyzer(Dg, N, calc_second_order, algo=algo_, resamples=R_req)
            eng.kernel_events = kernel_events
            eng.row0 = 0
            engs.append(eng)
        return engs

    engines = build(algo)
    if world > 1 and _agree([0 if all(e.kind == 6 for e in engines) else 1])[0] and any(e.kind == 6 for e in engines):
        engines = build("fp64")             # ranks must issue the same collectives: one engine family for all
    Z = norm.ppf(0.5 + conf_level / 2)

    if r is not None:
        rs = _Resampler.from_indices(r)
        if rs.N != N:
            raise ValueError(f"`r` must have {N} rows")
        R = rs.R
    else:
        R = int(num_resamples)
        rs = _Resampler(N, R, seed, resample or RESAMPLE_MODE, distributed=distributed) if R > 0 else None

    def again_fp64():
        return analyze_outputs(problem, Ys, calc_second_order=calc_second_order, num_resamples=num_resamples,
                               conf_level=conf_level, keep_resamples=keep_resamples, seed=seed, resample=resample,
                               r=r, algo="fp64", distributed=distributed, kernel_events=kernel_events)

    keep = bool(keep_resamples)
    # one GPU + integer engine: the pipelined row route with a single shard that holds every row
    pipe = None
    if world == 1 and all(e.kind == 6 for e in engines):
        pipe = _DigitsRows([[e] for e in engines], N, R, rs, engines[0].Npad, 0, 1,
                           upload=not isinstance(Ys[0], torch.Tensor))
    lo, hi = rt.shard_range(R, rank, world)
    # device-drawn multiplicities do not depend on Y: start the first batch now, beside the upload of Y
    prefetch = None
    if pipe is None and hi > lo and rs.mode == "philox":
        prefetch = _CountsPrefetch(rs, engines[0]._batches(lo, hi)[0], [(engines[0].Npad, None)])

    for eng, Y in zip(engines, Ys):
        # multi-GPU: a host Y is uploaded as one shard per rank and all-gathered over NVLink
        Yd = rt.to_device_sharded(Y, torch.float64) if world > 1 else rt.to_device(Y, torch.float64, non_blocking=True)
        eng.load(Yd.reshape(-1))

    if pipe is not None:
        if not all(e.kind == 6 for e in engines):     # the feature kernels refused the shape
            pipe.abandon()
            return again_fp64()
        res = pipe.run(Z, keep)
        points = [p.reshape(1, -1) for p, _, _ in res]
        confs = [c for _, c, _ in res]
        kepts = [k for _, _, k in res]
        guards = pipe.guard
    else:
        points = []
        ests = [None] * len(engines)
        if hi > lo:
            ests = bootstrap_shared(engines, rs, lo, hi, prefetch, points)
        if not points:
            points = [eng.point() for eng in engines]
        if R > 0:
            if hi <= lo:
                ests = bootstrap_shared(engines, rs, lo, hi, prefetch)
            if world > 1:
                ests = [allgather_estimates(e, R, rank, world) for e in ests]
            confs = [eng.conf(e, Z) for eng, e in zip(engines, ests)]
        else:
            confs = [torch.full((eng.nest,), float("nan"), dtype=torch.float64, device=points[0].device)
                     for eng in engines]
        kepts = [None] * len(engines)
        if keep:
            kepts = [est[:, :2 * Dg].cpu().numpy() if est is not None else None for est in ests]

    out = []
    step = 2 * Dg + 2 if calc_second_order else Dg + 2
    # ONE read-back for all outputs: results, statistics, guard counts, the resampler's flag
    nest = engines[0].nest
    zero = torch.zeros((1,), dtype=torch.float64, device=rt.device())
    pieces = []
    for o, (eng, point, conf) in enumerate(zip(engines, points, confs)):
        pieces += [point.reshape(-1)[:nest], conf.reshape(-1)[:nest], eng.stats.reshape(-1)[:8],
                   guards[o:o + 1] if pipe is not None else zero]
    pieces.append(rs.flag.to(torch.float64) if rs is not None else zero)
    host = torch.cat(pieces).cpu().numpy()
    _Resampler.raise_for(int(host[-1]))
    per = 2 * nest + 9
    for o, (eng, kept) in enumerate(zip(engines, kepts)):
        h = host[o * per:(o + 1) * per]
        point_h, conf_h, stats = h[:nest], h[nest:2 * nest], h[2 * nest:2 * nest + 8]
        guard = float(h[2 * nest + 8]) if pipe is not None else None
        if eng.kind == 6 and eng.auto_digits and (SobolAnalyzer.degenerate(stats)
                                                  or (guard is not None and guard_trips(guard, step, N))):
            # non-finite / constant outputs: the reference's behaviour comes from the FP64 engines;
            # an outlier row: the FP64 engines keep every row's own exponent (GUARD_MIN_ROWS)
            return again_fp64()
        # ptp([A;B]) == 0 -- or a constant Y, whose normalisation is 0/0 (tested outcome is zeros)
        const_y = not (stats[4] < stats[5])
        if const_y:
            warn(CONST_RESULT_MSG)
        if keep and kept is None:
            kept = np.zeros((0, 2 * Dg))
        S = _assemble(Dg, bool(calc_second_order), keep, point_h, conf_h, kept, const_y)
        S.problem = problem
        S.to_df = MethodType(to_df, S)
        S.engine = ENGINE_NAMES[eng.kind]
        S.resample_mode = rs.mode if rs is not None else None
        S.digits_guard = guard
        S.draws = pipe.draws if pipe is not None else None
        out.append(S)
    return out


def analyze_device(problem, Y, calc_second_order=True, num_resamples=100, conf_level=0.95,
                   keep_resamples=False, seed=None, resample=None, r=None, algo="auto", distributed=True,
                   kernel_events=None, shard="auto"):
    """``analyze`` with the device knobs exposed.

    Y may be a host ndarray or a CUDA tensor.  ``resample``: "numpy" | "philox" | "auto" (default: module
    ``RESAMPLE_MODE``); ``r``: explicit int [N, R] index matrix (overrides ``resample``/``seed``);
    ``algo``: "auto" | "fp64" | "digits", or one kernel ("generic" | "tile8" | "gemv" | "dmma"); ``distributed``:
    use the initialised torch.distributed job; ``shard``: "resamples" (every rank holds Y, resample ids are split,
    estimates all-gathered), "rows" (every rank keeps only its base-row range of Y and the partial sums of all
    resamples are reduced -- ``analyze_row_sharded``) or "auto": rows when the ranks agree that the integer engine
    applies to their row range (everything it does shrinks with the rows a rank holds), resamples otherwise --
    the strategy switch the reference makes inside ``analyze`` (analyze/sobol.py:147,183).
    """
    rank, world = rt.world() if distributed else (0, 1)
    auto, algo_asked = False, algo
    if shard == "auto":
        shard = "resamples"
        if world > 1:
            _, Dg = extract_group_names(problem)
            size = int(Y.numel()) if isinstance(Y, torch.Tensor) else int(np.size(Y))
            N = _n_from_size(size, Dg, calc_second_order)
            lo, hi = rt.shard_range(N, rank, world)
            R_req = np.shape(r)[1] if r is not None else int(num_resamples)
            eng = SobolAnalyzer(Dg, hi - lo, calc_second_order, algo=algo, resamples=R_req,
                                total_rows=N) if hi > lo else None
            auto = True
            if eng is not None and eng.kind == 6:
                shard, algo = "rows", "digits"   # the ranks' agreement rides in analyze_row_sharded's one collective
            else:
                _RankRecords.decline(world)      # every other rank learns it from the same all-gather
    if shard == "rows":
        _, Dg = extract_group_names(problem)
        size = int(Y.numel()) if isinstance(Y, torch.Tensor) else int(np.size(Y))
        N = _n_from_size(size, Dg, calc_second_order)
        step = size // N
        lo, hi = rt.shard_range(N, rank, world)
        flat = Y.reshape(-1)
        S = analyze_row_sharded(problem, [(lo, flat[lo * step:hi * step])], N,
                                calc_second_order=calc_second_order, num_resamples=num_resamples,
                                conf_level=conf_level, keep_resamples=keep_resamples, seed=seed,
                                resample=resample, r=r, algo=algo, distributed=distributed,
                                kernel_events=kernel_events, auto=auto)
        if S is not None:
            return S
        shard, algo = "resamples", algo_asked    # the ranks did not agree on the integer engine
    if shard != "resamples":
        raise ValueError(f"unknown shard mode {shard!r}")
    return analyze_outputs(problem, [Y], calc_second_order=calc_second_order, num_resamples=num_resamples,
                           conf_level=conf_level, keep_resamples=keep_resamples, seed=seed, resample=resample,
                           r=r, algo=algo, distributed=distributed, kernel_events=kernel_events)[0]


def analyze(
    problem,
    Y,
    calc_second_order=True,
    num_resamples=100,
    conf_level=0.95,
    print_to_console=False,
    parallel=False,
    n_processors=None,
    keep_resamples=False,
    seed=None,
):
    """Sobol' first, second and total-order indices with bootstrap confidence intervals.

    Same contract as ``SALib.analyze.sobol.analyze``.  ``parallel`` / ``n_processors`` (the reference's
    CPU process pool, analyze/sobol.py:183-194) are accepted and ignored: the GPU path is always the
    serial path's arithmetic.  Like the reference, which switches between its serial loop and its pool inside
    this function (:147,183), the decomposition is chosen here: one GPU, or -- when torch.distributed is
    initialised -- base-row shards (integer engine) or resample-id shards (FP64 engines) over the ranks.
    With a ``seed`` the resample indices are the reference's own (``default_rng(seed).integers``) for every
    problem whose index matrix the reference could draw (up to 2^31 draws); see the module docstring.
    """
    S = analyze_device(problem, Y, calc_second_order=calc_second_order, num_resamples=num_resamples,
                       conf_level=conf_level, keep_resamples=keep_resamples, seed=seed)
    if print_to_console:
        for df in S.to_df():
            print(df)
    return S


def separate_output_values(Y, D, N, calc_second_order):
    """(A, B, AB, BA) views of the model output (reference :267-279): Y.reshape(N, step) columns."""
    step = 2 * D + 2 if calc_second_order else D + 2
    M = np.asarray(Y).reshape(N, step)
    return M[:, 0], M[:, step - 1], M[:, 1:1 + D], (M[:, 1 + D:1 + 2 * D] if calc_second_order else None)


def _estimates_from_columns(A, ABs, BAs, B):
    """Run the device estimators on explicit columns (no normalisation): returns the `est` vector of a
    problem with len(ABs) factors; second order iff BAs is given."""
    A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    c2 = BAs is not None
    cols = [A] + [np.asarray(x, dtype=np.float64) for x in ABs] + ([np.asarray(x, dtype=np.float64) for x in BAs] if c2 else []) + [B]
    Y = np.stack(cols, axis=1)
    N, Dg = Y.shape[0], len(ABs)
    eng = SobolAnalyzer(Dg, N, c2)
    Yd = rt.to_device(Y.reshape(-1), torch.float64)
    eng.load(Yd)                                   # moments give the ptp([A;B]) == 0 test ...
    eng.stats[0:2] = torch.tensor([0.0, 1.0], dtype=torch.float64, device=eng.stats.device)
    eng.load_normalized(Yd)                        # ... but the data is used as given: (y - 0) / 1
    est = eng.point()[0].cpu().numpy()
    stats = eng.stats.cpu().numpy()
    if not (stats[4] < stats[5]):
        warn(CONST_RESULT_MSG)
    return est


def _columnwise(fn, arrays):
    """The reference calls the estimators on 1-D arrays and on [N, R] stacks of resampled columns."""
    if np.ndim(arrays[0]) == 1:
        return np.float64(fn(*arrays))
    R = np.shape(arrays[0])[1]
    return np.array([fn(*[np.asarray(a)[:, c] for a in arrays]) for c in range(R)])


def first_order(A, AB, B):
    """mean(B * (AB - A)) / var([A; B]) -- Saltelli et al. 2010 (reference :209-219), on the device."""
    return _columnwise(lambda a, ab, b: _estimates_from_columns(a, [ab], None, b)[0], (A, AB, B))


def total_order(A, AB, B):
    """0.5 * mean((A - AB)^2) / var([A; B]) (reference :222-232), on the device."""
    return _columnwise(lambda a, ab, b: _estimates_from_columns(a, [ab], None, b)[1], (A, AB, B))


def second_order(A, ABj, ABk, BAj, B):
    """mean(BAj * ABk - A * B) / var([A; B]) - S1_j - S1_k (reference :235-246), on the device."""
    def one(a, abj, abk, baj, b):
        return _estimates_from_columns(a, [abj, abk], [baj, np.zeros_like(baj)], b)[4]
    return _columnwise(one, (A, ABj, ABk, BAj, B))


def Si_to_pandas_dict(S_dict):
    """(total, first, (index, second)) dicts ready for DataFrames (reference :365-415)."""
    total = {"ST": S_dict["ST"], "ST_conf": S_dict["ST_conf"]}
    first = {"S1": S_dict["S1"], "S1_conf": S_dict["S1_conf"]}
    idx, second = None, None
    if "S2" in S_dict:
        names, _ = extract_group_names(S_dict.problem)
        idx = list(combinations(names, 2)) if len(names) > 2 else (names,)
        pos = {n: i for i, n in enumerate(names)}
        second = {
            "S2": [S_dict["S2"][pos[i[0]], pos[i[1]]] for i in idx],
            "S2_conf": [S_dict["S2_conf"][pos[i[0]], pos[i[1]]] for i in idx],
        }
    return total, first, (idx, second)


def to_df(self):
    """[total, first(, second)] DataFrames; bound to the ResultDict by ``analyze`` (reference :418-439)."""
    total, first, (idx, second) = Si_to_pandas_dict(self)
    names, _ = extract_group_names(self.problem)
    ret = [pd.DataFrame(total, index=names), pd.DataFrame(first, index=names)]
    if second:
        ret.append(pd.DataFrame(second, index=idx))
    return ret
