"""ctypes binding of ``libadelie_b200.so`` (C ABI in ``include/adelie_b200.h``).

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, every
operator raises.  The library itself is loaded lazily so that importing the package (and
the pure-host helpers) works on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadelie_b200.so")

c_i64 = C.c_int64
c_vp = C.c_void_p


class StateArgs(C.Structure):
    """Mirror of ``ab_state_args`` (include/adelie_b200.h)."""
    _fields_ = [
        ("dtype", C.c_int32),
        ("groups", c_vp), ("group_sizes", c_vp), ("G", c_i64), ("alpha", C.c_double), ("penalty", c_vp),
        ("weights", c_vp), ("X_means", c_vp), ("y_mean", C.c_double), ("y_var", C.c_double), ("resid_sum", C.c_double),
        ("rsq", C.c_double),
        ("resid", c_vp), ("offsets", c_vp), ("eta", c_vp), ("beta0", C.c_double), ("loss_null", C.c_double),
        ("loss_full", C.c_double),
        ("setup_loss_null", C.c_int32), ("irls_max_iters", c_i64), ("irls_tol", C.c_double),
        ("n_classes", c_i64), ("multi_intercept", C.c_int32),
        ("lmda_path", c_vp), ("lmda_path_len", c_i64), ("lmda_max", C.c_double), ("min_ratio", C.c_double),
        ("lmda_path_size", c_i64),
        ("setup_lmda_max", C.c_int32), ("setup_lmda_path", C.c_int32),
        ("max_screen_size", c_i64), ("max_active_size", c_i64), ("pivot_subset_ratio", C.c_double),
        ("pivot_subset_min", c_i64), ("pivot_slack_ratio", C.c_double),
        ("screen_rule", C.c_int32),
        ("max_iters", c_i64), ("tol", C.c_double), ("adev_tol", C.c_double), ("ddev_tol", C.c_double),
        ("newton_tol", C.c_double), ("newton_max_iters", c_i64),
        ("early_exit", C.c_int32), ("intercept", C.c_int32), ("n_threads", c_i64),
        ("screen_set", c_vp), ("screen_set_size", c_i64), ("screen_beta", c_vp), ("screen_beta_size", c_i64),
        ("screen_is_active", c_vp), ("active_set_size", c_i64), ("active_set", c_vp),
        ("lmda", C.c_double), ("grad", c_vp),
    ]


class CovStateArgs(C.Structure):
    """Mirror of ``ab_cov_state_args`` (include/adelie_b200.h)."""
    _fields_ = [
        ("dtype", C.c_int32),
        ("v", c_vp),
        ("groups", c_vp), ("group_sizes", c_vp), ("G", c_i64), ("alpha", C.c_double), ("penalty", c_vp),
        ("lmda_path", c_vp), ("lmda_path_len", c_i64), ("lmda_max", C.c_double), ("min_ratio", C.c_double),
        ("lmda_path_size", c_i64),
        ("setup_lmda_max", C.c_int32), ("setup_lmda_path", C.c_int32),
        ("max_screen_size", c_i64), ("max_active_size", c_i64), ("pivot_subset_ratio", C.c_double),
        ("pivot_subset_min", c_i64), ("pivot_slack_ratio", C.c_double),
        ("screen_rule", C.c_int32),
        ("max_iters", c_i64), ("tol", C.c_double), ("rdev_tol", C.c_double), ("newton_tol", C.c_double),
        ("newton_max_iters", c_i64),
        ("early_exit", C.c_int32), ("n_threads", c_i64),
        ("screen_set", c_vp), ("screen_set_size", c_i64), ("screen_beta", c_vp), ("screen_beta_size", c_i64),
        ("screen_is_active", c_vp), ("active_set_size", c_i64), ("active_set", c_vp),
        ("rsq", C.c_double), ("lmda", C.c_double), ("grad", c_vp),
        ("screen_grad", c_vp),
    ]


GLM_CB2 = C.CFUNCTYPE(C.c_int, c_vp, c_vp, c_vp)                      # gradient / inv_link (ctx, eta, out)
GLM_CB3 = C.CFUNCTYPE(C.c_int, c_vp, c_vp, c_vp, c_vp)                # hessian (ctx, eta, grad, hess)
GLM_CB4 = C.CFUNCTYPE(C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp)          # inv_hessian_gradient (ctx, eta, grad, hess, out)
GLM_CBL = C.CFUNCTYPE(C.c_int, c_vp, c_vp, C.POINTER(C.c_double))     # loss (ctx, eta, out)
GLM_CBF = C.CFUNCTYPE(C.c_int, c_vp, C.POINTER(C.c_double))           # loss_full (ctx, out)


class GlmCallbacks(C.Structure):
    """Mirror of ``ab_glm_callbacks`` (include/adelie_b200.h)."""
    _fields_ = [("ctx", c_vp), ("gradient", GLM_CB2), ("hessian", GLM_CB3), ("inv_hessian_gradient", GLM_CB4), ("loss", GLM_CBL),
                ("loss_full", GLM_CBF), ("inv_link", GLM_CB2)]


EXIT_COND_T = C.CFUNCTYPE(C.c_int, c_vp)
CHECK_SIGNALS_T = C.CFUNCTYPE(C.c_int)

_lib = None

# every symbol declared in include/adelie_b200.h (checked by tests/test_cabi.py)
SYMBOLS = [
    "ab_last_error", "ab_version", "ab_device_count", "ab_set_device", "ab_get_device_info", "ab_mem_info", "ab_device_synchronize", "ab_host_register", "ab_host_unregister", "ab_timer_start", "ab_timer_stop",
    "ab_dist_init", "ab_dist_connect", "ab_dist_allreduce_f64", "ab_dist_info",
    "ab_configs_set", "ab_configs_get",
    "ab_matrix_dense_create", "ab_matrix_dense_alloc", "ab_matrix_dense_fill_normal", "ab_matrix_dense_download",
    "ab_matrix_sparse_create", "ab_matrix_sparse_alloc_random", "ab_matrix_sparse_nnz", "ab_matrix_sparse_download",
    "ab_io_snp_unphased_create", "ab_io_snp_unphased_free", "ab_io_snp_unphased_write", "ab_io_snp_unphased_read", "ab_io_snp_unphased_info",
    "ab_io_snp_unphased_get", "ab_io_snp_unphased_to_dense",
    "ab_matrix_snp_unphased_create", "ab_matrix_snp_unphased_from_calldata", "ab_matrix_snp_unphased_alloc_random", "ab_matrix_snp_unphased_download",
    "ab_matrix_snp_unphased_cache_info", "ab_matrix_standardize_create", "ab_matrix_subset_create",
    "ab_io_snp_phased_ancestry_create", "ab_io_snp_phased_ancestry_free", "ab_io_snp_phased_ancestry_write", "ab_io_snp_phased_ancestry_read",
    "ab_io_snp_phased_ancestry_info", "ab_io_snp_phased_ancestry_get", "ab_io_snp_phased_ancestry_to_dense", "ab_matrix_snp_phased_ancestry_create",
    "ab_matrix_free", "ab_matrix_rows", "ab_matrix_cols", "ab_matrix_cmul", "ab_matrix_ctmul", "ab_matrix_bmul",
    "ab_matrix_btmul", "ab_matrix_mul", "ab_matrix_cov", "ab_matrix_window_gram", "ab_matrix_mul_multi", "ab_matrix_sq_mul", "ab_matrix_sp_tmul",
    "ab_glm_create", "ab_glm_create_callback", "ab_glm_free", "ab_glm_gradient", "ab_glm_hessian", "ab_glm_inv_hessian_gradient", "ab_glm_loss",
    "ab_glm_loss_full", "ab_glm_inv_link",
    "ab_state_create", "ab_state_free", "ab_state_solve", "ab_state_get_scalar", "ab_state_get_vec_f64",
    "ab_state_get_vec_i64", "ab_state_get_betas", "ab_state_get_screen_transform",
    "ab_pin_naive_solve",
    "ab_matrix_cov_dense_create", "ab_matrix_cov_lazy_create", "ab_matrix_cov_free", "ab_matrix_cov_cols", "ab_matrix_cov_bmul",
    "ab_matrix_cov_mul", "ab_matrix_cov_to_dense", "ab_matrix_cov_cache_info",
    "ab_cov_state_create", "ab_cov_state_free", "ab_cov_state_solve", "ab_cov_pin_solve", "ab_cov_state_get_scalar",
    "ab_cov_state_get_vec_f64", "ab_cov_state_get_vec_i64", "ab_cov_state_get_betas", "ab_cov_state_get_screen_transform",
    "ab_bcd_solve", "ab_bcd_root_lower_bound", "ab_bcd_root_upper_bound", "ab_bcd_root_function",
]


def load():
    """Load the CUDA library; raises RuntimeError (never falls back to a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"adelie_b200: CUDA library not built ({LIB_PATH} missing). Run ./build.sh (nvcc, sm_100a). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.ab_last_error.restype = C.c_char_p
    L.ab_matrix_dense_create.argtypes = [C.c_int, c_vp, c_i64, c_i64, C.c_int, c_i64, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_dense_alloc.argtypes = [C.c_int, c_i64, c_i64, C.POINTER(c_vp)]
    L.ab_matrix_dense_fill_normal.argtypes = [c_vp, C.c_uint64, c_i64]
    L.ab_matrix_dense_download.argtypes = [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64]
    L.ab_matrix_free.argtypes = [c_vp]
    L.ab_io_snp_unphased_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(c_vp)]
    L.ab_io_snp_unphased_free.argtypes = [c_vp]
    L.ab_io_snp_unphased_write.argtypes = [c_vp, c_vp, c_i64, c_i64, C.c_char_p, c_vp, c_i64, C.c_int, C.POINTER(C.c_uint64)]
    L.ab_io_snp_unphased_read.argtypes = [c_vp, C.POINTER(C.c_uint64)]
    L.ab_io_snp_unphased_info.argtypes = [c_vp, C.POINTER(C.c_int), C.POINTER(c_i64), C.POINTER(c_i64)]
    L.ab_io_snp_unphased_get.argtypes = [c_vp, C.c_char_p, c_vp]
    L.ab_io_snp_unphased_to_dense.argtypes = [c_vp, C.c_int, c_vp]
    L.ab_matrix_snp_unphased_create.argtypes = [C.c_int, c_vp, c_i64, c_i64, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_snp_unphased_from_calldata.argtypes = [C.c_int, c_vp, c_i64, c_i64, c_vp, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_snp_unphased_alloc_random.argtypes = [C.c_int, c_i64, c_i64, C.c_uint64, c_i64, c_i64, C.c_double, C.c_double, C.c_double, C.POINTER(c_vp)]
    L.ab_matrix_snp_unphased_download.argtypes = [c_vp, c_vp, c_vp]
    L.ab_matrix_snp_unphased_cache_info.argtypes = [c_vp, C.POINTER(c_i64), C.POINTER(c_i64)]
    L.ab_matrix_standardize_create.argtypes = [c_vp, c_vp, c_i64, c_vp, c_i64, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_subset_create.argtypes = [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.POINTER(c_vp)]
    L.ab_io_snp_phased_ancestry_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(c_vp)]
    L.ab_io_snp_phased_ancestry_free.argtypes = [c_vp]
    L.ab_io_snp_phased_ancestry_write.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, C.POINTER(C.c_uint64)]
    L.ab_io_snp_phased_ancestry_read.argtypes = [c_vp, C.POINTER(C.c_uint64)]
    L.ab_io_snp_phased_ancestry_info.argtypes = [c_vp, C.POINTER(C.c_int), C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]
    L.ab_io_snp_phased_ancestry_get.argtypes = [c_vp, C.c_char_p, c_vp]
    L.ab_io_snp_phased_ancestry_to_dense.argtypes = [c_vp, C.c_int, c_vp]
    L.ab_matrix_snp_phased_ancestry_create.argtypes = [C.c_int, c_vp, c_i64, c_i64, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_rows.argtypes = [c_vp, C.POINTER(c_i64)]
    L.ab_matrix_cols.argtypes = [c_vp, C.POINTER(c_i64)]
    L.ab_matrix_cmul.argtypes = [c_vp, c_i64, c_vp, c_vp, C.POINTER(C.c_double)]
    L.ab_matrix_ctmul.argtypes = [c_vp, c_i64, C.c_double, c_vp]
    L.ab_matrix_bmul.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]
    L.ab_matrix_btmul.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp]
    L.ab_matrix_mul.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.ab_matrix_cov.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp]
    L.ab_matrix_mul_multi.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp]
    L.ab_matrix_window_gram.argtypes = [c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp]
    L.ab_matrix_sq_mul.argtypes = [c_vp, c_vp, c_vp]
    L.ab_matrix_sp_tmul.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.ab_glm_create.argtypes = [C.c_int, C.c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.POINTER(c_vp)]
    L.ab_glm_free.argtypes = [c_vp]
    L.ab_glm_create_callback.argtypes = [C.c_int, c_i64, c_i64, C.c_int, C.POINTER(GlmCallbacks), C.POINTER(c_vp)]
    L.ab_glm_gradient.argtypes = [c_vp, c_vp, c_vp]
    L.ab_glm_hessian.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.ab_glm_inv_hessian_gradient.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp]
    L.ab_glm_loss.argtypes = [c_vp, c_vp, C.POINTER(C.c_double)]
    L.ab_glm_loss_full.argtypes = [c_vp, C.POINTER(C.c_double)]
    L.ab_glm_inv_link.argtypes = [c_vp, c_vp, c_vp]
    L.ab_state_create.argtypes = [C.POINTER(StateArgs), c_vp, c_vp, C.POINTER(c_vp)]
    L.ab_state_free.argtypes = [c_vp]
    L.ab_state_solve.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_double)]
    L.ab_pin_naive_solve.argtypes = [c_vp, c_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_double)]
    L.ab_state_get_scalar.argtypes = [c_vp, C.c_char_p, C.POINTER(C.c_double)]
    L.ab_state_get_vec_f64.argtypes = [c_vp, C.c_char_p, c_vp, c_i64, C.POINTER(c_i64)]
    L.ab_state_get_vec_i64.argtypes = [c_vp, C.c_char_p, c_vp, c_i64, C.POINTER(c_i64)]
    L.ab_state_get_betas.argtypes = [c_vp, c_vp, c_vp, c_vp, C.POINTER(c_i64), C.POINTER(c_i64)]
    L.ab_state_get_screen_transform.argtypes = [c_vp, c_i64, c_vp, c_i64, C.POINTER(c_i64)]
    L.ab_matrix_cov_dense_create.argtypes = [C.c_int, c_vp, c_i64, C.c_int, c_i64, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_cov_lazy_create.argtypes = [C.c_int, c_vp, c_i64, c_i64, C.c_int, c_i64, C.c_int, C.POINTER(c_vp)]
    L.ab_matrix_cov_free.argtypes = [c_vp]
    L.ab_matrix_cov_cols.argtypes = [c_vp, C.POINTER(c_i64)]
    L.ab_matrix_cov_bmul.argtypes = [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp]
    L.ab_matrix_cov_mul.argtypes = [c_vp, c_vp, c_vp, c_i64, c_vp]
    L.ab_matrix_cov_to_dense.argtypes = [c_vp, c_i64, c_i64, c_vp]
    L.ab_matrix_cov_cache_info.argtypes = [c_vp, C.POINTER(c_i64)]
    L.ab_cov_state_create.argtypes = [C.POINTER(CovStateArgs), c_vp, C.POINTER(c_vp)]
    L.ab_cov_state_free.argtypes = [c_vp]
    L.ab_cov_state_solve.argtypes = [c_vp, C.c_int, c_vp, c_vp, c_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_double)]
    L.ab_cov_pin_solve.argtypes = [c_vp, c_vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_double)]
    L.ab_cov_state_get_scalar.argtypes = [c_vp, C.c_char_p, C.POINTER(C.c_double)]
    L.ab_cov_state_get_vec_f64.argtypes = [c_vp, C.c_char_p, c_vp, c_i64, C.POINTER(c_i64)]
    L.ab_cov_state_get_vec_i64.argtypes = [c_vp, C.c_char_p, c_vp, c_i64, C.POINTER(c_i64)]
    L.ab_cov_state_get_betas.argtypes = [c_vp, c_vp, c_vp, c_vp, C.POINTER(c_i64), C.POINTER(c_i64)]
    L.ab_cov_state_get_screen_transform.argtypes = [c_vp, c_i64, c_vp, c_i64, C.POINTER(c_i64)]
    L.ab_configs_set.argtypes = [C.c_char_p, C.c_double]
    L.ab_configs_get.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
    L.ab_bcd_solve.argtypes = [C.c_int, c_i64, c_vp, c_vp, C.c_double, C.c_double, C.c_double, c_i64, c_vp, C.POINTER(c_i64)]
    L.ab_bcd_root_lower_bound.argtypes = [c_i64, c_vp, c_vp, C.c_double, C.POINTER(C.c_double)]
    L.ab_bcd_root_upper_bound.argtypes = [c_i64, c_vp, c_vp, C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.ab_bcd_root_function.argtypes = [c_i64, C.c_double, c_vp, c_vp, C.c_double, C.POINTER(C.c_double)]
    L.ab_device_count.argtypes = [C.POINTER(C.c_int)]
    L.ab_set_device.argtypes = [C.c_int]
    L.ab_host_register.argtypes = [c_vp, C.c_size_t]
    L.ab_host_unregister.argtypes = [c_vp]
    L.ab_timer_stop.argtypes = [C.POINTER(C.c_double)]
    L.ab_dist_init.argtypes = [C.c_int, C.c_int, c_vp]
    L.ab_dist_connect.argtypes = [c_vp]
    L.ab_dist_allreduce_f64.argtypes = [c_vp, c_i64]
    L.ab_dist_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ab_mem_info.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.ab_get_device_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    _lib = L
    return L


def check(rc):
    """Raise the library's last error as RuntimeError (adelie_core_error -> RuntimeError, py_adelie_core.cpp)."""
    if rc != 0:
        raise RuntimeError(load().ab_last_error().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def dtype_code(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return 0
    if dtype == np.float64:
        return 1
    raise RuntimeError("adelie_b200: dtype must be float32 or float64.")
