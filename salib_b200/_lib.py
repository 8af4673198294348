"""ctypes binding of the C ABI declared in include/salib_b200.h.

There is NO CPU fallback: if ``lib/libsalib_b200.so`` is missing or no CUDA device is visible the
product path raises.  (Host-only entry points -- ``sa_direction_vectors``, ``sa_row_stride`` -- work
without a GPU, which is what the ``-m "not gpu"`` tests exercise.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SALIB_B200_LIB: load another build of the same ABI (A/B timing of kernel variants)
LIB_PATH = os.environ.get("SALIB_B200_LIB") or os.path.join(_HERE, "lib", "libsalib_b200.so")

_vp, _u64, _i32, _dbl, _sz = C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/salib_b200.h
SIGNATURES = {
    "sa_version": (C.c_int, []),
    "sa_last_error": (C.c_char_p, []),
    "sa_direction_vectors": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "sa_saltelli_sample_f64": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sa_scale_samples_f64": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp, _vp, _vp]),
    "sa_sobol_points_f64": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _i32, _vp]),
    "sa_eval_ishigami_f64": (C.c_int, [_vp, _vp, _u64, _dbl, _dbl, _vp]),
    "sa_eval_sobol_g_f64": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sa_saltelli_eval_sobol_g_f64": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _i32, _i32, _vp, _i32, _vp, _vp,
                                               _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sa_saltelli_eval_ishigami_f64": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _i32, _vp, _i32, _vp, _vp,
                                                _vp, _vp, _dbl, _dbl, _vp]),
    "sa_discrepancy_workspace_bytes": (_sz, [_u64, _i32]),
    "sa_discrepancy_f64": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sa_text_slot_bytes": (C.c_int, []),
    "sa_format_values_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _vp, _vp, _vp]),
    "sa_pack_text_rows": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _vp, _vp]),
    "sa_text_parse_chunk_bytes": (C.c_int, []),
    "sa_text_count_newlines": (C.c_int, [_vp, _vp, _u64, _vp]),
    "sa_text_line_starts": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "sa_text_parse_lines": (C.c_int, [_vp, _vp, _u64, _vp, _u64, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "sa_normalize_features_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _vp, _vp]),
    "sa_moments_workspace_bytes": (_sz, []),
    "sa_moments_f64": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp, _sz]),
    "sa_row_stride": (C.c_int, [_i32, _i32]),
    "sa_pad_rows": (_u64, [_u64]),
    "sa_feature_count": (C.c_int, [_i32]),
    "sa_normalize_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _vp, _vp, _vp]),
    "sa_counts_from_indices": (C.c_int, [_vp, _vp, _u64, _u64, _u64, _u64, _vp, _vp]),
    "sa_counts_philox": (C.c_int, [_vp, _u64, _u64, _u64, _u64, _vp, _vp]),
    "sa_counts_from_indices_rows": (C.c_int, [_vp, _vp, _u64, _u64, _u64, _u64, _u64, _u64, _vp, _vp]),
    "sa_counts_philox_rows": (C.c_int, [_vp, _u64, _u64, _u64, _u64, _u64, _u64, _vp, _vp]),
    "sa_counts_philox_rows_ex": (C.c_int, [_vp, _u64, _u64, _u64, _u64, _u64, _u64, _vp, _vp, _i32]),
    "sa_sobol_accumulate_f64": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _vp, _u64, _i32, _vp]),
    "sa_sobol_accumulate_ex_f64": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _vp, _u64, _i32, _vp, _i32]),
    "sa_sobol_finalize_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _u64, _i32, _vp, _vp]),
    "sa_sobol_conf_f64": (C.c_int, [_vp, _vp, _u64, _i32, _dbl, _vp]),
    "sa_col_moment_f64": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp]),
    "sa_conf_finish_f64": (C.c_int, [_vp, _vp, _i32, _dbl, _dbl, _u64, _i32]),
    "sa_conf_local_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _vp]),
    "sa_conf_combine_f64": (C.c_int, [_vp, _vp, _i32, _u64, _i32, _i32, _dbl, _vp, _vp]),
    "sa_combine_moments_f64": (C.c_int, [_vp, _vp, _i32, _vp]),
    "sa_combine_moments_pitch_f64": (C.c_int, [_vp, _vp, _i32, _u64, _vp]),
    "sa_rank_record_f64": (C.c_int, [_vp, _vp, _i32, _dbl, _dbl, _dbl, _vp]),
    "sa_chunked_feature_doubles": (C.c_int, [_i32]),
    "sa_finite_diff_expand_f64": (C.c_int, [_vp, _vp, _u64, _i32, _dbl, _vp, _vp, _vp]),
    "sa_dgsm_features_f64": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "sa_feature_sums_f64": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _u64, _i32, _vp]),
    "sa_reduce_splits_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _dbl, _vp]),
    "sa_feature_centered_ss_f64": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _vp, _vp]),
    "sa_digit_available": (C.c_int, []),
    "sa_digit_row_pitch": (_u64, [_u64]),
    "sa_digit_ldc": (_u64, [_i32, _i32]),
    "sa_digit_ksplit": (C.c_int, [_u64, _i32, _i32, _u64]),
    "sa_digit_pair_plan": (C.c_int, [_u64, _u64, _u64, _i32, _i32, _vp]),
    "sa_digit_features_workspace_bytes": (_sz, [_i32]),
    "sa_digit_sums_workspace_bytes": (_sz, [_u64, _i32, _i32, _u64]),
    "sa_digit_features_i8": (C.c_int, [_vp, _vp, _u64, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _sz]),
    "sa_digit_sums_f64": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _vp, _u64, _u64, _vp, _sz, _vp]),
    "sa_digit_sums_draw_f64": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _vp, _u64, _u64, _vp, _sz, _vp,
                                         _u64, _u64, _u64, _u64, _u64, _u64, _vp]),
    "sa_digit_gemm_i32": (C.c_int, [_vp, _vp, _u64, _u64, _vp, _u64, _u64, _u64, _i32, _vp, _u64]),
    "sa_launch_count": (C.c_longlong, [_i32]),
    "sa_microbench_fp64": (C.c_int, [_vp, _i32, _vp, _vp]),
    "sa_microbench_fp64_outer": (C.c_int, [_vp, _i32, _vp, _vp]),
    "sa_microbench_l2_read": (C.c_int, [_vp, _vp, _u64, _i32, _vp]),
    "sa_microbench_i8_tensor": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
}


class SalibB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


_lib = None


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C salib_b200/csrc`). salib_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().sa_last_error().decode("utf-8", "replace")
        kind = "argument error" if status < 0 else "CUDA error"
        raise SalibB200Error(f"{what}: {kind} {status}: {msg}")


def call(name, *args):
    check(getattr(lib(), name)(*args), name)
