/*
 * salib_b200 -- C ABI of the B200-native Sobol'/Saltelli hot path.
 *
 * The reference (SALib) has no FFI: its "plugin" boundary is the pair of Python callables
 *   SALib.sample.sobol.sample(problem, N, ...)        src/SALib/sample/sobol.py:11-189
 *   SALib.analyze.sobol.analyze(problem, Y, ...)      src/SALib/analyze/sobol.py:23-206
 * (and the deprecated alias src/SALib/sample/saltelli.py:12-203).  This header is what a
 * maintainer would bind from inside those two functions (ctypes stub in INTEGRATION.md): every
 * entry point replaces a numbered span of the reference, cited next to it.
 *
 * Conventions
 *   - plain C types only; every `d_` pointer is DEVICE memory owned by the caller, every `h_`
 *     pointer is HOST memory.  The library never allocates or frees result buffers.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises unless stated.
 *   - return value: 0 = ok; < 0 = argument error (SA_ERR_*); > 0 = cudaError_t of the failed
 *     CUDA call.  sa_last_error() returns a thread-local message for the last non-zero status.
 *   - no global mutable state; calls on different streams may run concurrently.
 */
#ifndef SALIB_B200_H
#define SALIB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SA_OK 0
#define SA_ERR_ARG (-1)         /* null pointer / negative size / inconsistent shape      */
#define SA_ERR_RANGE (-2)       /* Sobol' index range beyond 2^30 points or > 21201 dims  */
#define SA_ERR_UNSUPPORTED (-3) /* shape outside what the kernels implement               */
#define SA_ERR_WORKSPACE (-4)   /* workspace too small                                    */

#define SA_SOBOL_BITS 30     /* scipy.stats.qmc.Sobol word size (sample/sobol.py:107)     */
#define SA_SOBOL_MAXDIM 21201 /* new-joe-kuo-6.21201 (sample/directions.py)               */

/* Number of accumulated sums per resample: 5 scalars + 2*Dg (+ Dg*(Dg-1)/2 pair sums). */
#define SA_NACC(Dg, second_order) (5 + 2 * (Dg) + ((second_order) ? (Dg) * ((Dg)-1) / 2 : 0))
/* Number of estimates per resample: S1[Dg], ST[Dg] (, S2[pairs]). */
#define SA_NEST(Dg, second_order) (2 * (Dg) + ((second_order) ? (Dg) * ((Dg)-1) / 2 : 0))

int sa_version(void);
const char *sa_last_error(void);

/* ---- direction numbers (host) -------------------------------------------------------------
 * Replaces the recurrence of sample/sobol_sequence.py:62-84 (== scipy's _initialize_v reached
 * from sample/sobol.py:107).  Inputs are the Joe-Kuo table columns for dimensions 2..dims
 * (a: polynomial bits, s: degree, m: [dims-1][18] initial numbers); output h_sv[dims][30]
 * uint32, h_sv[d][j] = V_{j+1} << (30 - (j+1)).                                              */
int sa_direction_vectors(const uint32_t *h_a, const uint8_t *h_s, const uint32_t *h_m, int dims,
                         uint32_t *h_sv);

/* ---- Saltelli sample: fused Sobol' draw + cross-matrix + uniform scaling ------------------
 * Replaces sample/sobol.py:107-188 (qmc.Sobol -> fast_forward -> random -> the A/AB/BA/B
 * triple loop :149-186 -> util/__init__.py:49-53 scaling).  Base row i (0 <= i < n) uses Sobol'
 * point first_index + i in 2*D dimensions: dims 0..D-1 -> A, D..2D-1 -> B.
 *   d_sv     [2D][30] direction vectors (plain, or LMS-scrambled by the host)
 *   d_shift  [2D] digital shift (zeros when unscrambled), may be NULL
 *   d_group_of_col [D] group index of each column (NULL = ungrouped, Dg must equal D)
 *   d_lb, d_span   [D] lower bound and (ub - lb), both already rounded to f64 on the host
 *   d_dist_code, d_dist_par  NULL for uniform bounds; else [D] codes and [D][4] parameters of the
 *            per-column inverse CDFs (replaces util/__init__.py:126-289; d_lb/d_span then unused):
 *              0 unif      {lb, span}                 1 triang  {loc, scale, c}
 *              2 logunif   {log a, log b - log a}     3 norm    {loc, scale}
 *              4 truncnorm, a < 0   {Phi(a),  Phi(b)-Phi(a), loc, scale}
 *              5 truncnorm, a >= 0  {Phi(-b), Phi(b)-Phi(a), loc, scale}
 *              6 lognorm   {loc, scale}               7 weibull {1/c, scale, loc}
 *   d_out    [(step*n), D] row-major f64, step = 2*Dg+2 (second_order) or Dg+2; 16-byte aligned
 * Uniform values are computed as round(round(x * span) + lb) -- no FMA contraction.  */
int sa_saltelli_sample_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                           uint64_t first_index, uint64_t n, int D, int Dg,
                           const int32_t *d_group_of_col, int second_order, const double *d_lb,
                           const double *d_span, const int32_t *d_dist_code, const double *d_dist_par,
                           double *d_out);

/* In-place scaling of d_X [rows, D] (values in [0,1]) to bounds or per-column distributions; replaces
 * util/__init__.py:25-90 (`scale_samples`, `_scale_samples`, `_nonuniform_scale_samples`) as a stand-alone call. */
int sa_scale_samples_f64(void *stream, double *d_X, uint64_t rows, int D, const double *d_lb,
                         const double *d_span, const int32_t *d_dist_code, const double *d_dist_par);

/* Plain Sobol' points first_index .. first_index+n-1 in `dims` dimensions, d_out [n, dims] row-major, values
 * x * 2^-30.  Replaces sample/sobol_sequence.py:49-91 (`sobol_sequence.sample`, the generator behind
 * sample/saltelli.py:157 and sample/finite_diff.py).                                              */
int sa_sobol_points_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                        uint64_t first_index, uint64_t n, int dims, double *d_out);

/* ---- test functions on the path -----------------------------------------------------------
 * sa_eval_ishigami_f64 replaces test_functions/Ishigami.py:55-59, sa_eval_sobol_g_f64 replaces
 * test_functions/Sobol_G.py:66-86 (X -> Y).  d_flag (int32, may be NULL) is OR-ed with 1 if any
 * value < 0 and 2 if any value > 1 (the reference raises ValueError, Sobol_G.py:68-74).       */
int sa_eval_ishigami_f64(void *stream, const double *d_X, uint64_t rows, double A, double B,
                         double *d_Y);
int sa_eval_sobol_g_f64(void *stream, const double *d_X, uint64_t rows, int D, const double *d_a,
                        const double *d_delta, const double *d_alpha, double *d_Y, int32_t *d_flag);

/* Fused variants: Sobol' indices -> Y without materialising X (same arguments as
 * sa_saltelli_sample_f64; d_Y is [step*n]).  Ishigami requires D == 3.                       */
int sa_saltelli_eval_sobol_g_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                                 uint64_t first_index, uint64_t n, int D, int Dg,
                                 const int32_t *d_group_of_col, int second_order,
                                 const double *d_lb, const double *d_span,
                                 const int32_t *d_dist_code, const double *d_dist_par,
                                 const double *d_a, const double *d_delta, const double *d_alpha,
                                 double *d_Y, int32_t *d_flag);
int sa_saltelli_eval_ishigami_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                                  uint64_t first_index, uint64_t n, int Dg,
                                  const int32_t *d_group_of_col, int second_order,
                                  const double *d_lb, const double *d_span,
                                  const int32_t *d_dist_code, const double *d_dist_par, double A,
                                  double B, double *d_Y);

/* ---- analyze: normalisation ---------------------------------------------------------------
 * sa_moments_f64 replaces Y.mean() / Y.std() of analyze/sobol.py:141 (two-pass, population
 * std).  d_stats[8] = {mean, std, min, max, minAB, maxAB, 0, 0}; minAB/maxAB are taken over
 * the A and B columns only (the ptp test of analyze/sobol.py:214-217).  d_work needs
 * sa_moments_workspace_bytes() bytes.                                                         */
size_t sa_moments_workspace_bytes(void);
int sa_moments_f64(void *stream, const double *d_Y, uint64_t N, int step, double *d_stats,
                   void *d_work, size_t work_bytes);

/* Layout of the normalised model output used by the accumulate kernels ("analysis layout"):
 * row i of length sa_row_stride(Dg, second_order) doubles =
 *   [ AB part | BA part | A B ]   (second order)        [ AB part | A B ]   (first/total only)
 * A part holds its columns in blocks of 8 at a pitch of 10 doubles (80 B): column j sits at
 * (j / 8) * 10 + j % 8, the 2 pad doubles of each block and the columns >= Dg are zero.  The pitch
 * makes the tile kernel's 128-bit shared-memory loads bank-conflict free (5*b mod 8 is a bijection).
 * sa_row_stride = parts * ceil(Dg / 8) * 10 + 2.                                                */
int sa_row_stride(int Dg, int second_order);
/* Replaces (Y - mean)/std and separate_output_values (analyze/sobol.py:141,267-279):
 * d_Y [N*step] reference order [A, AB_0.., (BA_0..,) B] -> d_Yn [N*stride (+8 doubles slack)]
 * in analysis layout; if d_Yt != NULL also writes the auxiliary array the fast kernels read:
 *   second_order != 0: column-major copy of A and B, d_Yt[2][Npad], Npad = sa_pad_rows(N);
 *   second_order == 0: per-row feature matrix, sa_feature_count(Dg) doubles per row in total, stored in
 *     chunks of CW = 8/16/24/32 features as d_Yt[chunk][N][CW]; features in SA_NACC order
 *     { A, B, A^2, B^2, A*B, B*(AB_j - A) (j < Dg), (A - AB_j)^2 (j < Dg) }, zero padded.            */
int sa_feature_count(int Dg);
/* Feature matrix for either order (what algo 3 of sa_sobol_accumulate_ex_f64 consumes): SA_NACC(Dg,
 * second_order) features per row in SA_NACC order -- for second order the list continues with the pair
 * products BA_j * AB_k (j < k) -- chunked as above; sa_chunked_feature_doubles(SA_NACC) doubles per row.
 * With it every sum of analyze/sobol.py:209-246 is one skinny GEMM w[R x N] * F[N x SA_NACC]; the Python
 * side picks it for second order while Dg is small (the 8x8 register-tile kernel leaves lanes idle there).   */
int sa_normalize_features_f64(void *stream, const double *d_Y, uint64_t N, int Dg, int second_order,
                              const double *d_stats, double *d_F);
uint64_t sa_pad_rows(uint64_t N);
int sa_normalize_f64(void *stream, const double *d_Y, uint64_t N, int Dg, int second_order,
                     const double *d_stats, double *d_Yn, double *d_Yt);

/* ---- analyze: resample multiplicities -----------------------------------------------------
 * A bootstrap resample c is the multiset of rows r[:, c] (analyze/sobol.py:144).  The kernels
 * consume it as multiplicities w[c][i] = #{t : r[t, c] == i} (uint8, row pitch sa_pad_rows(N)),
 * which gives the same sums as gathering A[r], AB[r], BA[r], B[r].
 *   sa_counts_from_indices: d_r int64 [N, R] C-order as numpy draws it; resamples
 *                           [c0, c0+count) -> d_w[count][Npad].  d_w must be zeroed by the call
 *                           (it is: the function memsets it on `stream`).
 *   sa_counts_philox:       device Philox4x32-10 draw of N iid uniform indices per resample, as
 *                           multiplicities ~ Multinomial(N, 1/N): a tree of Binomial(n, 1/2) splits
 *                           (popcounts of n random bits) down to windows of <= 16384 rows, then
 *                           uniform draws inside each window, histogrammed in shared memory
 *                           (when 2^L | N; otherwise N direct draws with global atomics).  Counters
 *                           are keyed by (seed, global resample id c0+c, tree node | window, word):
 *                           independent of batching and of how resamples or rows are sharded over GPUs.
 * d_overflow (int32, may be NULL) is set to 1 if any multiplicity exceeded 255.
 * The *_rows forms keep only the draws that land in the row window [row0, row0+nrows) of the N rows and
 * write d_w[count][sa_pad_rows(nrows)] indexed from row0: the multiplicities of one row shard when Y is
 * partitioned by base-row range across GPUs (SURVEY 8(e) "when Y exceeds one GPU").  The draws are the
 * same N per resample whatever the window, so the shards of all GPUs add up to the unsharded w.     */
int sa_counts_from_indices(void *stream, const int64_t *d_r, uint64_t N, uint64_t R, uint64_t c0,
                           uint64_t count, uint8_t *d_w, int32_t *d_overflow);
int sa_counts_philox(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                     uint8_t *d_w, int32_t *d_overflow);
int sa_counts_from_indices_rows(void *stream, const int64_t *d_r, uint64_t N, uint64_t R, uint64_t c0,
                                uint64_t count, uint64_t row0, uint64_t nrows, uint8_t *d_w,
                                int32_t *d_overflow);
int sa_counts_philox_rows(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                          uint64_t row0, uint64_t nrows, uint8_t *d_w, int32_t *d_overflow);
/* slim != 0: the low-footprint launch shape -- one 512-thread block per SM with a single 16 KB window, which fits
 * next to a CTA of the persistent GEMM kernel, so a side stream can draw batch b+1 while batch b is in the tensor
 * cores.  Same multiplicities as the wide launch (the Philox counters do not depend on the launch shape).      */
int sa_counts_philox_rows_ex(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                             uint64_t row0, uint64_t nrows, uint8_t *d_w, int32_t *d_overflow, int slim);

/* ---- analyze: accumulate / finalize -------------------------------------------------------
 * sa_sobol_accumulate_f64 replaces the gathers and reductions inside first_order / total_order
 * / second_order (analyze/sobol.py:209-246) for `count` resamples at once.  For resample c it
 * produces, per row split sp, SA_NACC sums over rows i with weight w = d_w[c][i]:
 *   [0..4]  sum w*A, sum w*B, sum w*A^2, sum w*B^2, sum w*A*B
 *   [5+j]   sum w*B*(AB_j - A)            [5+Dg+j] sum w*(A - AB_j)^2
 *   [5+2Dg+p(j,k)] sum w*BA_j*AB_k, j<k, p = j*Dg - j*(j+1)/2 + (k-j-1)
 * into d_psum[nsplit][count][SA_NACC].  d_w == NULL means weight 1 for every row (the point
 * estimate; count must be 1).  nsplit row ranges of equal size let small `count` fill the GPU.
 * d_Yt is required by the first-order kernel and by the tile kernel (its A/B columns).        */
int sa_sobol_accumulate_f64(void *stream, const double *d_Yn, const double *d_Yt, uint64_t N,
                            int Dg, int second_order, const uint8_t *d_w, uint64_t count,
                            int nsplit, double *d_psum);
/* Same, with an explicit kernel choice for tests: algo 0 = auto, 1 = generic (any shape),
 * 3 = feature-GEMV kernel (first/total order; second order with the pair-feature matrix), 4 = second-order 8x8 register-tile kernel with rows staged in
 * shared memory by cp.async.bulk (what "auto" picks for second order), 5 = EXPERIMENTAL FP64
 * tensor-core (mma.sync m8n8k4.f64) variant of 4, opt-in only (BASELINE rules tensor cores out).  */
int sa_sobol_accumulate_ex_f64(void *stream, const double *d_Yn, const double *d_Yt, uint64_t N,
                               int Dg, int second_order, const uint8_t *d_w, uint64_t count,
                               int nsplit, double *d_psum, int algo);
/* Sums the splits and applies the estimator algebra of analyze/sobol.py:219,232,242-246:
 * d_est[count][SA_NEST] = {S1_j, ST_j, S2_jk}.  If d_stats says ptp([A;B]) == 0 every estimate
 * is 0 (the `return 0.0` branches :214-217,227-230,238-240).                                  */
int sa_sobol_finalize_f64(void *stream, const double *d_psum, uint64_t N, int Dg, int second_order,
                          uint64_t count, int nsplit, const double *d_stats, double *d_est);
/* conf = z * std(d_est[:, col], ddof=1) over R resamples (analyze/sobol.py:159,173,180-182);
 * d_conf[SA_NEST].                                                                            */
int sa_sobol_conf_f64(void *stream, const double *d_est, uint64_t R, int ncols, double z,
                      double *d_conf);
/* The two passes of sa_sobol_conf_f64 for estimates spread over the ranks of a row-sharded analyze (replaces the
 * all-gather of [R, nest] estimates: two all-reduces of nest doubles).  d_mean == NULL: d_out[col] = column sums of
 * the local d_est [R_local, ncols]; else d_out[col] = sum_r (est - mean[col])^2.  sa_conf_finish_f64 scales a
 * vector in place (finish == 0: x *= scale) or turns the all-reduced squared deviations into the CI
 * z * sqrt(x / (R - 1)) (analyze/sobol.py:159,170,180-182: Z * std(ddof=1)).                                   */
int sa_col_moment_f64(void *stream, const double *d_est, uint64_t R, int ncols, const double *d_mean, double *d_out);
int sa_conf_finish_f64(void *stream, double *d_x, int ncols, double scale, double z, uint64_t R, int finish);
/* The same CI with ONE collective: sa_conf_local_f64 reduces the local estimates d_est [R_local, ncols] to the record
 * d_rec = (R_local, mean[ncols], sum of squared deviations from that mean [ncols]) -- or, with slices > 1, to one
 * such record per contiguous row slice (d_rec[slices][1 + 2 ncols]; fills the GPU for a long estimate matrix on one
 * rank, the slices are then combined like ranks); the records of all ranks are
 * all-gathered (k records `pitch` doubles apart, each followed by `nextra` doubles that are simply summed over the
 * records: the point estimates and guard counts of the row-sharded analyze ride along) and sa_conf_combine_f64
 * combines them (Chan/Golub/LeVeque) into d_conf[col] = z * std(all estimates, ddof=1) and d_extra_sum[nextra].   */
int sa_conf_local_f64(void *stream, const double *d_est, uint64_t R, int ncols, int slices, double *d_rec);
int sa_conf_combine_f64(void *stream, const double *d_recs, int k, uint64_t pitch, int ncols, int nextra, double z,
                        double *d_conf, double *d_extra_sum);
/* Mean / population std / min / max of the union of k row shards from per-shard records
 * d_parts[k][9] = (n_values, the 8 statistics of sa_moments_f64) -> d_out[9] = the 8 statistics of the whole
 * design + n_values (replaces Y.mean(), Y.std() of analyze/sobol.py:141 when Y is partitioned over GPUs).  */
int sa_combine_moments_f64(void *stream, const double *d_parts, int k, double *d_out);
/* The same with the k records `pitch` >= 9 doubles apart (records that carry more than the moments).           */
int sa_combine_moments_pitch_f64(void *stream, const double *d_parts, int k, uint64_t pitch, double *d_out);
/* One record per rank for the single collective in front of a row-sharded analyze: d_rec[12] = the rank's k shard
 * records d_parts[k][9] combined into (n_values, 8 statistics), then f0, f1, f2 -- what the ranks must agree on
 * before they issue further collectives (engine family, largest shard, ...; reduced with max on the host after the
 * all-gather).  Replaces the reference's Y.mean()/Y.std() (analyze/sobol.py:141) bookkeeping across GPUs.        */
int sa_rank_record_f64(void *stream, const double *d_parts, int k, double f0, double f1, double f2, double *d_rec);

/* ---- analyze: opt-in exact-integer tensor-core engine ("digits", csrc/digits.cu) ---------------------
 * The sums of sa_sobol_accumulate_f64 by another route: psum[c][q] = sum_i w[c][i] * F[i][q] with integer
 * w is split error-free.  Each feature column F[:, q] (SA_NACC order, the features of
 * sa_normalize_features_f64) is scaled by a power of two taken from its exact max |F|, rounded ONCE to an
 * (8*ndigits-2)-bit integer and stored as `ndigits` balanced radix-256 digits (int8); the digit sums
 * sum_i w_i * d_i are u8 x s8 dot products accumulated exactly in int32 by tcgen05.mma kind::i8 (N <= 2^24),
 * recombined in 128-bit integers and rounded to float64 once.  Never selected automatically (BASELINE's
 * north star rules tensor cores out for the default path); Python: analyze_device(..., algo="digits").
 *   sa_digit_row_pitch(N)   bytes of one digit column (rows of Y padded to 128)
 *   sa_digit_features_i8    d_Y [N*step] reference order + d_stats -> d_Q int8, nfeat*ndigits columns (column
 *                           q*ndigits+s = digit s of feature q) of sa_digit_row_pitch(N) rows, stored K-blocked:
 *                           d_Q[i / 128][column][i % 128]; d_inv [nfeat + 1]: the power of two that maps each
 *                           column's integer sums back, then (entry nfeat) the heavy-tail guard: how many
 *                           normalised values of these rows lie within 7 bits of the design's largest |x|
 *                           (d_stats[2..3] = global min / max).  nfeat = SA_NACC(Dg, second_order).
 *   sa_digit_sums_f64       d_psum [count][nfeat] (the nsplit = 1 layout sa_sobol_finalize_f64 reads) from
 *                           d_w [count][w_pitch] uint8 multiplicities; workspace = int32 digit sums
 *                           [ksplit][count][ldc], ksplit = sa_digit_ksplit(...), ldc = sa_digit_ldc(...).
 *   sa_digit_gemm_i32       the GEMM alone (tests): d_C[ks][count][ldc] = d_w [count][K] * d_Q[ncols][K]^T over
 *                           the ks-th of `ksplit` ranges of K; d_Q row-major with pitch q_pitch, or K-blocked
 *                           as above when q_pitch == 0.                                                     */
int sa_digit_available(void); /* 1: compute capability 10.x and cuTensorMapEncodeTiled exported by the driver */
uint64_t sa_digit_row_pitch(uint64_t N);
uint64_t sa_digit_ldc(int nfeat, int ndigits);
int sa_digit_ksplit(uint64_t N, int nfeat, int ndigits, uint64_t count);
/* work plan of the persistent CTA-pair GEMM (host logic only; tests): out[0..7] = pairs, full_waves, rem_tiles,
 * rsplit, rper, band, epoch, slack */
int sa_digit_pair_plan(uint64_t tiles_m, uint64_t tiles_n, uint64_t ksteps, int max_pairs, int band, uint32_t *out);
size_t sa_digit_features_workspace_bytes(int nfeat);
size_t sa_digit_sums_workspace_bytes(uint64_t N, int nfeat, int ndigits, uint64_t count);
int sa_digit_features_i8(void *stream, const double *d_Y, uint64_t N, int Dg, int second_order,
                         const double *d_stats, int ndigits, int8_t *d_Q, double *d_inv, void *d_work,
                         size_t work_bytes);
int sa_digit_sums_f64(void *stream, const int8_t *d_Q, const double *d_inv, uint64_t N, int nfeat, int ndigits,
                      const uint8_t *d_w, uint64_t w_pitch, uint64_t count, void *d_work, size_t work_bytes,
                      double *d_psum);
/* sa_digit_sums_f64 that also HOSTS the Philox draw of the next batch of multiplicities (draw_count > 0; arguments as
 * sa_counts_philox_rows: d_w_next[draw_count][sa_pad_rows(draw_nrows)] <- resamples [draw_c0, draw_c0 + draw_count)
 * over the row window): 16 spare warps of every CTA of the persistent pair GEMM draw while the tensor cores
 * multiply the current batch -- same stream, no events, and co-residency by construction (a separate kernel only
 * overlaps when the block scheduler and the shared-memory carveout happen to allow it).  Shapes that do not run the
 * persistent pair kernel draw first and multiply after.  d_w_next must not alias d_w.                            */
int sa_digit_sums_draw_f64(void *stream, const int8_t *d_Q, const double *d_inv, uint64_t N, int nfeat, int ndigits,
                           const uint8_t *d_w, uint64_t w_pitch, uint64_t count, void *d_work, size_t work_bytes,
                           double *d_psum, uint64_t draw_seed, uint64_t draw_N, uint64_t draw_c0, uint64_t draw_count,
                           uint64_t draw_row0, uint64_t draw_nrows, uint8_t *d_w_next);
int sa_digit_gemm_i32(void *stream, const uint8_t *d_w, uint64_t w_pitch, uint64_t count, const int8_t *d_Q,
                      uint64_t q_pitch, uint64_t ncols, uint64_t K, int ksplit, int32_t *d_C, uint64_t ldc);

/* ---- DGSM on the same machinery ------------------------------------------------
 * sa_finite_diff_expand_f64: rows of sample/finite_diff.py:76-91 from scaled Sobol' points d_base [n, D]:
 *   d_out [n*(D+1), D]; row 1+j moves column j by base_j*delta, clipped to [d_lo[j], d_hi[j]].
 * sa_dgsm_features_f64: F[chunk][N][CW] with features ((Y_j - Y_base)/(x_j - x_base))^2 (q < D) and their
 *   squares (D <= q < 2D) from the finite_diff matrix d_X and output d_Y (analyze/dgsm.py:105-121);
 *   sa_chunked_feature_doubles(2*D) doubles per row in total.
 * sa_feature_sums_f64: psum[nsplit][count][nfeat] = sum_i w[c][i] * F[i][q] (d_w NULL: weight 1, count 1) --
 *   the bootstrap means of dgsm.py:133-137 for all resamples at once.
 * sa_reduce_splits_f64: d_out[count][nfeat] = scale * sum over splits.                                  */
int sa_chunked_feature_doubles(int nfeat);
int sa_finite_diff_expand_f64(void *stream, const double *d_base, uint64_t n, int D, double delta,
                              const double *d_lo, const double *d_hi, double *d_out);
int sa_dgsm_features_f64(void *stream, const double *d_X, const double *d_Y, uint64_t N, int D, double *d_F);
int sa_feature_sums_f64(void *stream, const double *d_F, uint64_t N, int nfeat, const uint8_t *d_w,
                        uint64_t count, int nsplit, double *d_psum);
/* d_out[q] = sum_i (F[i][q] - d_center[q])^2 for the first ncols columns of the chunked feature matrix: the second
 * pass of np.std(dfdx) in calc_vi_stats (analyze/dgsm.py:103-109), which a one-pass E[x^2] - E[x]^2 cannot
 * reproduce for nearly constant derivatives.                                                                */
int sa_feature_centered_ss_f64(void *stream, const double *d_F, uint64_t N, int nfeat, int ncols,
                               const double *d_center, double *d_out);
int sa_reduce_splits_f64(void *stream, const double *d_psum, uint64_t count, int nfeat, int nsplit,
                         double scale, double *d_out);

/* ---- discrepancy indices --------------------------------------------------------
 * Replaces analyze/discrepancy.py:88-116: X [N, D] and Y [N] are unit-scaled as qmc.scale(reverse=True)
 * does (X by the problem bounds d_lb/d_ub, Y by its own min/max, :92-94), then for every input i the 2-D
 * discrepancy of the projection [X_i, Y] (scipy.stats.qmc.discrepancy, :96-99) is computed -- an N^2 pair
 * sum per input -- and normalised by the sum over inputs (:101); with groups (d_gcol: group index per
 * column, Dg groups) the group mean of the members is returned (:103-111).
 *   method: 0 "WD" wrap-around, 1 "CD" centered, 2 "MD" mixture, 3 "L2-star".
 *   d_out [Dg] normalised indices; d_disc [D] raw discrepancies (may be NULL).
 *   d_flag (int32, zeroed by the caller, may be NULL): bit 0 = a sample lies outside its bounds
 *   ("Sample is out of bounds"), bit 1 = Y is constant or NaN (scipy: inconsistent bounds).           */
size_t sa_discrepancy_workspace_bytes(uint64_t N, int D);
int sa_discrepancy_f64(void *stream, const double *d_X, const double *d_Y, uint64_t N, int D,
                       const double *d_lb, const double *d_ub, int method, int Dg, const int32_t *d_gcol,
                       double *d_out, double *d_disc, void *d_work, size_t work_bytes, int32_t *d_flag);

/* ---- text output ---------------------------------------------------------------
 * The reference's CLI writes samples with np.savetxt(fmt="%.<precision>e") (sample/common_args.py via
 * sample/sobol.py:247-251, saltelli.py, finite_diff.py cli_action).  These two calls produce the same bytes
 * on the device: every value is its exact binary value rounded half-to-even to precision+1 significant
 * digits, as CPython/glibc print it ("nan", "inf", "-inf" for non-finite values).
 *   sa_format_values_f64: d_x [rows, cols] -> d_slots [rows*cols][sa_text_slot_bytes()] (the characters of
 *     each value), d_lens [rows*cols] (their count), d_rowlen [rows] (bytes of each text row including
 *     delimiters and the newline).  0 <= precision <= 17.
 *   sa_pack_text_rows: d_rowoff [rows] = exclusive prefix sum of d_rowlen (the caller scans); writes the rows
 *     "v0<delim>v1...<delim>vn\n" at d_text + d_rowoff[r].                                               */
int sa_text_slot_bytes(void);
int sa_format_values_f64(void *stream, const double *d_x, uint64_t rows, int cols, int precision,
                         uint8_t *d_slots, uint8_t *d_lens, int64_t *d_rowlen);
int sa_pack_text_rows(void *stream, const uint8_t *d_slots, const uint8_t *d_lens, uint64_t rows, int cols,
                      int delimiter, const int64_t *d_rowoff, uint8_t *d_text);

/* ---- text input -----------------------------------------------------------------
 * np.loadtxt as the reference's CLI uses it (analyze/sobol.py:476-491, dgsm.py:151-161, discrepancy.py:131-146):
 * rectangular numeric text -> float64, every token converted as Python's float() does (correctly rounded;
 * "nan", "inf", "infinity" in any case), '#' comments and blank lines skipped.
 *   sa_text_count_newlines: d_blockcount[ceil(len / sa_text_parse_chunk_bytes())] = newlines per chunk.
 *   sa_text_line_starts:    d_blockoff = exclusive prefix sum of d_blockcount (the caller scans);
 *                           d_linestart[l] = byte offset of line l, l < nlines = newlines + 1.
 *   sa_text_parse_lines:    one thread per line.  delimiter 0 = runs of blanks/tabs, else that byte.
 *       mode 0: d_ntok[l] = number of fields on line l (0 for blank / comment lines);
 *       mode 1: d_out[d_rowidx[l]][0..cols) = the converted fields (d_rowidx: index among non-blank lines).
 *   d_flag (int32, zeroed by the caller): bit 0 a field is not a decimal float literal, bit 1 a line has a
 *   different number of fields, bit 2 a field has more than 40 significant digits (not converted).        */
int sa_text_parse_chunk_bytes(void);
int sa_text_count_newlines(void *stream, const uint8_t *d_text, uint64_t len, int64_t *d_blockcount);
int sa_text_line_starts(void *stream, const uint8_t *d_text, uint64_t len, const int64_t *d_blockoff,
                        int64_t *d_linestart);
int sa_text_parse_lines(void *stream, const uint8_t *d_text, uint64_t len, const int64_t *d_linestart,
                        uint64_t nlines, int delimiter, int mode, int32_t *d_ntok, const int64_t *d_rowidx,
                        int cols, double *d_out, int32_t *d_flag);

/* ---- measurement helpers (bench.py only) ----------------------------------------------------
 * sa_launch_count: kernels enqueued by this library since the last reset (bench.py "gpu_launches").
 * sa_microbench_fp64 / sa_microbench_l2_read: the FP64-FMA and L2-read roofline denominators that
 * MEASURED_PEAKS.json lacks.  fp64: flop = 32 * iters * (*h_threads).                */
long long sa_launch_count(int reset);
int sa_microbench_fp64(void *stream, int iters, double *d_out, uint64_t *h_threads);
/* 8x8 register outer-product DFMA mix of the tile kernel, no memory: flop = 128 * iters * (*h_threads). */
int sa_microbench_fp64_outer(void *stream, int iters, double *d_out, uint64_t *h_threads);
int sa_microbench_l2_read(void *stream, const void *d_buf, uint64_t bytes, int reps, double *d_out);
/* int8 tensor-core rate of the digits engine's MMA mix with no operand traffic (ring filled once):
 * d_w [pairs*256][1024] u8, d_Q [512][1024] s8, d_C [pairs*256][512] int32, pairs = SMs / 2;
 * *h_ops = integer operations (2 per multiply-add) of the launch.                                       */
int sa_microbench_i8_tensor(void *stream, const uint8_t *d_w, const int8_t *d_Q, int32_t *d_C, int ksteps,
                            double *h_ops);

#ifdef __cplusplus
}
#endif
#endif /* SALIB_B200_H */
