"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of the CPU oracle (oracle/liboracle.so, built from
adelie_oracle.hpp by oracle/Makefile).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module; the
product package ``adelie_b200`` never does.

``grpnet`` below restates the Python-side initialisation of the reference
(adelie/solver.py:621-958) in NumPy and then runs the restated C++ path solver.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import types

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False):
    """Compile oracle/liboracle.so with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


class _PathArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("matrix_kind", C.c_int32),
        ("X", C.c_void_p), ("n", C.c_int64), ("p", C.c_int64), ("ld", C.c_int64),
        ("sp_outer", C.c_void_p), ("sp_inner", C.c_void_p), ("sp_values", C.c_void_p),
        ("K", C.c_int64), ("multi_intercept", C.c_int32),
        ("family", C.c_int32),
        ("y", C.c_void_p), ("weights", C.c_void_p), ("offsets", C.c_void_p),
        ("cox_start", C.c_void_p), ("cox_stop", C.c_void_p), ("cox_status", C.c_void_p), ("cox_strata", C.c_void_p),
        ("cox_efron", C.c_int32),
        ("groups", C.c_void_p), ("group_sizes", C.c_void_p), ("G", C.c_int64), ("penalty", C.c_void_p), ("alpha", C.c_double),
        ("X_means", C.c_void_p), ("y_mean", C.c_double), ("y_var", C.c_double), ("rsq", C.c_double), ("resid_sum", C.c_double),
        ("resid", C.c_void_p), ("grad", C.c_void_p), ("eta", C.c_void_p), ("beta0", C.c_double), ("loss_null", C.c_double),
        ("loss_full", C.c_double), ("setup_loss_null", C.c_int32),
        ("screen_set", C.c_void_p), ("S", C.c_int64), ("screen_beta", C.c_void_p), ("screen_beta_size", C.c_int64),
        ("screen_is_active", C.c_void_p), ("active_set_size", C.c_int64), ("active_set", C.c_void_p),
        ("lmda", C.c_double), ("lmda_max", C.c_double), ("lmda_path", C.c_void_p), ("lmda_path_len", C.c_int64),
        ("setup_lmda_max", C.c_int32), ("setup_lmda_path", C.c_int32),
        ("min_ratio", C.c_double), ("lmda_path_size", C.c_int64), ("max_screen_size", C.c_int64), ("max_active_size", C.c_int64),
        ("pivot_subset_ratio", C.c_double), ("pivot_subset_min", C.c_int64), ("pivot_slack_ratio", C.c_double), ("screen_rule", C.c_int32),
        ("max_iters", C.c_int64), ("tol", C.c_double), ("adev_tol", C.c_double), ("ddev_tol", C.c_double), ("newton_tol", C.c_double),
        ("newton_max_iters", C.c_int64), ("irls_max_iters", C.c_int64), ("irls_tol", C.c_double),
        ("early_exit", C.c_int32), ("intercept", C.c_int32), ("n_threads", C.c_int32),
        ("max_seconds", C.c_double),
    ]


class _PinArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("X", C.c_void_p), ("n", C.c_int64), ("p", C.c_int64), ("ld", C.c_int64),
        ("y_mean", C.c_double), ("y_var", C.c_double),
        ("groups", C.c_void_p), ("group_sizes", C.c_void_p), ("G", C.c_int64), ("alpha", C.c_double), ("penalty", C.c_void_p),
        ("weights", C.c_void_p),
        ("screen_set", C.c_void_p), ("S", C.c_int64),
        ("lmda_path", C.c_void_p), ("L", C.c_int64),
        ("intercept", C.c_int32), ("max_active_size", C.c_int64), ("max_iters", C.c_int64), ("tol", C.c_double),
        ("adev_tol", C.c_double), ("ddev_tol", C.c_double), ("newton_tol", C.c_double), ("newton_max_iters", C.c_int64),
        ("n_threads", C.c_int32),
        ("rsq", C.c_double), ("resid", C.c_void_p), ("resid_sum", C.c_double),
        ("screen_beta", C.c_void_p), ("screen_is_active", C.c_void_p), ("active_set_size", C.c_int64), ("active_set", C.c_void_p),
    ]


class _CovArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("matrix_kind", C.c_int32),
        ("M", C.c_void_p), ("n", C.c_int64), ("p", C.c_int64), ("ld", C.c_int64), ("row_major", C.c_int32),
        ("v", C.c_void_p),
        ("groups", C.c_void_p), ("group_sizes", C.c_void_p), ("G", C.c_int64), ("penalty", C.c_void_p), ("alpha", C.c_double),
        ("screen_set", C.c_void_p), ("S", C.c_int64), ("screen_beta", C.c_void_p), ("screen_beta_size", C.c_int64),
        ("screen_is_active", C.c_void_p), ("active_set_size", C.c_int64), ("active_set", C.c_void_p),
        ("rsq", C.c_double), ("lmda", C.c_double), ("lmda_max", C.c_double), ("grad", C.c_void_p),
        ("screen_grad", C.c_void_p),
        ("lmda_path", C.c_void_p), ("lmda_path_len", C.c_int64), ("setup_lmda_max", C.c_int32), ("setup_lmda_path", C.c_int32),
        ("min_ratio", C.c_double), ("lmda_path_size", C.c_int64), ("max_screen_size", C.c_int64), ("max_active_size", C.c_int64),
        ("pivot_subset_ratio", C.c_double), ("pivot_subset_min", C.c_int64), ("pivot_slack_ratio", C.c_double), ("screen_rule", C.c_int32),
        ("max_iters", C.c_int64), ("tol", C.c_double), ("rdev_tol", C.c_double), ("newton_tol", C.c_double),
        ("newton_max_iters", C.c_int64), ("early_exit", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_path_solve.restype = C.c_void_p
        L.orc_path_solve.argtypes = [C.POINTER(_PathArgs)]
        L.orc_pin_solve.restype = C.c_void_p
        L.orc_pin_solve.argtypes = [C.POINTER(_PinArgs)]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_result_error.restype = C.c_char_p
        L.orc_result_error.argtypes = [C.c_void_p]
        L.orc_result_scalar.restype = C.c_double
        L.orc_result_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_result_vec.restype = C.c_int64
        L.orc_result_vec.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_result_ivec.restype = C.c_int64
        L.orc_result_ivec.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.orc_result_betas.restype = C.c_int64
        L.orc_result_betas.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_dense_cmul.restype = C.c_double
        L.orc_dense_cmul.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dense_ctmul.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_int]
        L.orc_dense_bmul.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dense_btmul.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dense_mul.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_dense_cov.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_glm_eval.restype = C.c_double
        L.orc_glm_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 4
        L.orc_bcd_solve.restype = C.c_int64
        L.orc_bcd_solve.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]
        for f in (L.orc_bcd_root_lower_bound,):
            f.restype = C.c_double
            f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_bcd_root_upper_bound.restype = C.c_double
        L.orc_bcd_root_upper_bound.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.orc_bcd_root_function.restype = C.c_double
        L.orc_bcd_root_function.argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_search_pivot.restype = C.c_int
        L.orc_search_pivot.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_jacobi_eigh.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_config.argtypes = [C.c_char_p, C.c_double]
        L.orc_cov_path_solve.restype = C.c_void_p
        L.orc_cov_path_solve.argtypes = [C.POINTER(_CovArgs)]
        L.orc_cov_pin_solve.restype = C.c_void_p
        L.orc_cov_pin_solve.argtypes = [C.POINTER(_CovArgs)]
        L.orc_cov_matrix_op.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dt(dtype):
    return 0 if np.dtype(dtype) == np.float32 else 1


FAMILIES = {"gaussian_opt": 0, "gaussian": 1, "binomial": 2, "multigaussian": 3, "cox": 4, "poisson": 5, "probit": 6, "multinomial": 7}


def set_config(name, value):
    lib().orc_set_config(name.encode(), float(value))


# --------------------------------------------------------------------------
# result extraction
# --------------------------------------------------------------------------
_VEC_NAMES = ["lmda_path", "lmdas", "devs", "intercepts", "screen_beta", "grad", "abs_grad", "resid", "eta",
              "screen_vars", "screen_X_means", "benchmark_screen", "benchmark_fit_screen", "benchmark_fit_active",
              "benchmark_kkt", "benchmark_invariance", "rsqs", "screen_grad", "benchmark_active", "screen_transforms_flat"]
_IVEC_NAMES = ["screen_set", "screen_begins", "screen_is_active", "active_set", "n_valid_solutions", "active_sizes",
               "screen_sizes"]
_SCALAR_NAMES = ["total_time", "lmda_max", "lmda", "rsq", "resid_sum", "y_mean", "y_var", "loss_null", "loss_full", "beta0",
                 "active_set_size", "n_sweeps", "n_group_updates", "n_irls", "iters"]


def _collect(h, p_cols, dtype):
    L = lib()
    out = types.SimpleNamespace()
    out.error = L.orc_result_error(h).decode()
    for nm in _SCALAR_NAMES:
        v = L.orc_result_scalar(h, nm.encode())
        if not np.isnan(v) or nm in ("loss_null",):
            setattr(out, nm, v)
    for nm in _VEC_NAMES:
        n = L.orc_result_vec(h, nm.encode(), None, 0)
        if n < 0:
            continue
        buf = np.empty(n, dtype=np.float64)
        L.orc_result_vec(h, nm.encode(), _p(buf), n)
        setattr(out, nm, buf if nm.startswith("benchmark") else buf.astype(dtype))
    for nm in _IVEC_NAMES:
        n = L.orc_result_ivec(h, nm.encode(), None, 0)
        if n < 0:
            continue
        buf = np.empty(n, dtype=np.int64)
        L.orc_result_ivec(h, nm.encode(), _p(buf), n)
        setattr(out, nm, buf)
    nnz = L.orc_result_betas(h, None, None, None)
    nl = len(getattr(out, "lmdas", []))
    indptr = np.empty(nl + 1, dtype=np.int64)
    indices = np.empty(nnz, dtype=np.int64)
    values = np.empty(nnz, dtype=np.float64)
    L.orc_result_betas(h, _p(indptr), _p(indices), _p(values))
    out.betas = sp.csr_matrix((values.astype(dtype), indices, indptr), shape=(nl, p_cols))
    L.orc_result_free(h)
    return out


# --------------------------------------------------------------------------
# GLM spec helpers
# --------------------------------------------------------------------------
def glm_spec(family, y, weights=None, dtype=np.float64, **kw):
    """family in {gaussian, binomial, multigaussian, cox, poisson}; weights normalised to sum 1 (adelie/glm.py:46-55)."""
    y = np.ascontiguousarray(y, dtype=dtype)
    n = y.shape[0]
    if weights is None:
        weights = np.full(n, 1 / n, dtype=dtype)
    else:
        weights = np.asarray(weights, dtype=dtype)
        weights = weights / np.sum(weights)
    spec = dict(family=family, y=y, weights=np.ascontiguousarray(weights, dtype=dtype), dtype=dtype, opt=kw.pop("opt", True))
    if family == "cox":
        spec["start"] = np.ascontiguousarray(kw["start"], dtype=dtype)
        spec["stop"] = np.ascontiguousarray(kw["stop"], dtype=dtype)
        spec["status"] = y
        strata = kw.get("strata")
        spec["strata"] = np.zeros(n, dtype=np.int64) if strata is None else np.ascontiguousarray(strata, dtype=np.int64)
        spec["efron"] = kw.get("tie_method", "efron") == "efron"
    return spec


def glm_eval(spec, op, eta=None, grad=None, hess=None):
    """op in gradient|hessian|inv_hessian_gradient|loss|loss_full|inv_link."""
    ops = {"gradient": 0, "hessian": 1, "inv_hessian_gradient": 2, "loss": 3, "loss_full": 4, "inv_link": 5}
    dtype = spec["dtype"]
    y = spec["y"]
    n = y.shape[0]
    K = y.shape[1] if y.ndim == 2 else 1
    fam = FAMILIES[spec["family"]]
    if fam == 0:
        fam = 1
    out = None
    if op in ("gradient", "hessian", "inv_hessian_gradient", "inv_link"):
        out = np.empty(y.shape if spec["family"] != "cox" else (n,), dtype=dtype)
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
    eta, grad, hess = c(eta), c(grad), c(hess)
    r = lib().orc_glm_eval(_dt(dtype), fam, ops[op], n, K, _p(y), _p(spec["weights"]),
                           _p(spec.get("start")), _p(spec.get("stop")), _p(spec.get("status")), _p(spec.get("strata")),
                           int(spec.get("efron", True)), _p(eta), _p(grad), _p(hess), _p(out))
    return out if out is not None else r


# --------------------------------------------------------------------------
# dense matrix operators
# --------------------------------------------------------------------------
class dense:
    """Column-major dense matrix with the reference operator set (adelie/matrix.py:549-680)."""

    def __init__(self, X, n_threads=1):
        self.X = np.asfortranarray(X)
        self.dtype = self.X.dtype
        self.n, self.p = self.X.shape
        self.nt = n_threads

    def _a(self):
        return (_dt(self.dtype), _p(self.X), self.n, self.p, self.n)

    def cmul(self, j, v, w):
        return lib().orc_dense_cmul(*self._a(), j, _p(v), _p(w), self.nt)

    def ctmul(self, j, v, out):
        lib().orc_dense_ctmul(*self._a(), j, float(v), _p(out), self.nt)

    def bmul(self, j, q, v, w, out):
        lib().orc_dense_bmul(*self._a(), j, q, _p(v), _p(w), _p(out), self.nt)

    def btmul(self, j, q, v, out):
        lib().orc_dense_btmul(*self._a(), j, q, _p(v), _p(out), self.nt)

    def mul(self, v, w, out):
        lib().orc_dense_mul(*self._a(), _p(v), _p(w), _p(out), self.nt)

    def cov(self, j, q, sqrt_w, out):
        lib().orc_dense_cov(*self._a(), j, q, _p(sqrt_w), _p(out), self.nt)


# --------------------------------------------------------------------------
# prox
# --------------------------------------------------------------------------
def bcd_solve(quad, linear, l1, l2, tol=1e-12, max_iters=1000, solver="newton_abs", dtype=np.float64):
    L = np.ascontiguousarray(quad, dtype=dtype)
    v = np.ascontiguousarray(linear, dtype=dtype)
    x = np.zeros_like(L)
    it = lib().orc_bcd_solve(_dt(dtype), {"newton": 0, "newton_abs": 1}[solver], L.size, _p(L), _p(v), l1, l2, tol, max_iters, _p(x))
    return dict(beta=x, iters=it)


def root_lower_bound(quad, linear, l1):
    D = np.ascontiguousarray(quad, dtype=np.float64); v = np.ascontiguousarray(linear, dtype=np.float64)
    return lib().orc_bcd_root_lower_bound(D.size, _p(D), _p(v), l1)


def root_upper_bound(quad, linear, l1, zero_tol=1e-14):
    D = np.ascontiguousarray(quad, dtype=np.float64); v = np.ascontiguousarray(linear, dtype=np.float64)
    return lib().orc_bcd_root_upper_bound(D.size, _p(D), _p(v), l1, zero_tol)


def root_function(h, quad, linear, l1):
    D = np.ascontiguousarray(quad, dtype=np.float64); v = np.ascontiguousarray(linear, dtype=np.float64)
    return lib().orc_bcd_root_function(D.size, h, _p(D), _p(v), l1)


def search_pivot(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    mses = np.empty_like(x)
    idx = lib().orc_search_pivot(x.size, _p(x), _p(y), _p(mses))
    return idx, mses


def jacobi_eigh(A):
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    q = A.shape[0]
    D = np.empty(q); V = np.empty((q, q), order="F")
    lib().orc_jacobi_eigh(q, _p(A), _p(D), _p(V))
    return D, V


# --------------------------------------------------------------------------
# pin solve in isolation
# --------------------------------------------------------------------------
def pin_naive_solve(X, y, *, groups, alpha, penalty, weights, screen_set, lmda_path, intercept=True,
                    max_iters=int(1e5), tol=1e-7, adev_tol=0.9, ddev_tol=0, newton_tol=1e-12, newton_max_iters=1000,
                    n_threads=1, screen_beta=None, screen_is_active=None, active_set=None, active_set_size=0,
                    rsq=0.0, resid=None, max_active_size=None):
    """Restates the gaussian_pin_naive wrapper (adelie/state.py:421-720): derives y_mean/y_var/resid, then runs
    CORE/solver/solver_gaussian_pin_naive.hpp:223-401."""
    X = np.asfortranarray(X)
    dtype = X.dtype
    n, p = X.shape
    groups = np.ascontiguousarray(groups, dtype=np.int64)
    G = groups.size
    group_sizes = np.diff(np.concatenate([groups, [p]])).astype(np.int64)
    penalty = np.ascontiguousarray(penalty, dtype=dtype)
    weights = np.ascontiguousarray(weights, dtype=dtype)
    screen_set = np.ascontiguousarray(screen_set, dtype=np.int64)
    lmda_path = np.ascontiguousarray(lmda_path, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    y_mean = float(np.sum(weights * y))
    yc = y - y_mean * intercept
    y_var = float(np.sum(weights * yc ** 2))
    sb_size = int(np.sum(group_sizes[screen_set]))
    if screen_beta is None:
        screen_beta = np.zeros(sb_size, dtype=dtype)
    screen_beta = np.ascontiguousarray(screen_beta, dtype=dtype).copy()
    if screen_is_active is None:
        screen_is_active = np.zeros(screen_set.size, dtype=np.int8)
    screen_is_active = np.ascontiguousarray(screen_is_active, dtype=np.int8).copy()
    act = np.zeros(G, dtype=np.int64)
    if active_set is not None:
        act[:active_set_size] = np.asarray(active_set)[:active_set_size]
    if resid is None:
        resid = yc.copy()
    resid = np.ascontiguousarray(resid, dtype=dtype).copy()
    resid_sum = float(np.sum(weights * resid))
    a = _PinArgs(
        dtype=_dt(dtype), X=_p(X), n=n, p=p, ld=n, y_mean=y_mean, y_var=y_var,
        groups=_p(groups), group_sizes=_p(group_sizes), G=G, alpha=alpha, penalty=_p(penalty), weights=_p(weights),
        screen_set=_p(screen_set), S=screen_set.size, lmda_path=_p(lmda_path), L=lmda_path.size,
        intercept=int(intercept), max_active_size=G if max_active_size is None else max_active_size, max_iters=max_iters,
        tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol, newton_tol=newton_tol, newton_max_iters=newton_max_iters,
        n_threads=n_threads, rsq=rsq, resid=_p(resid), resid_sum=resid_sum,
        screen_beta=_p(screen_beta), screen_is_active=_p(screen_is_active), active_set_size=active_set_size, active_set=_p(act),
    )
    h = lib().orc_pin_solve(C.byref(a))
    out = _collect(h, p, dtype)
    out.resid = resid
    out.screen_beta = screen_beta
    out.screen_is_active = screen_is_active
    out.active_set = act[: a.active_set_size]
    out.active_set_size = a.active_set_size
    out.y_mean, out.y_var = y_mean, y_var
    return out


# --------------------------------------------------------------------------
# grpnet: Python-side initial invariants (adelie/solver.py:621-958) + path
# --------------------------------------------------------------------------
def grpnet(X, glm, *, groups=None, alpha=1.0, penalty=None, offsets=None, lmda_path=None, irls_max_iters=int(1e4),
           irls_tol=1e-7, max_iters=int(1e5), tol=1e-7, adev_tol=0.9, ddev_tol=0.0, newton_tol=1e-12,
           newton_max_iters=1000, n_threads=1, early_exit=True, intercept=True, screen_rule="pivot", min_ratio=1e-2,
           lmda_path_size=100, max_screen_size=None, max_active_size=None, pivot_subset_ratio=0.1, pivot_subset_min=1,
           pivot_slack_ratio=1.25, max_seconds=-1.0):
    dtype = glm["dtype"]
    is_sparse = sp.issparse(X)
    if is_sparse:
        Xs = sp.csc_matrix(X, dtype=dtype)
        Xs.sort_indices()
        n, p = Xs.shape
        outer = Xs.indptr.astype(np.int32); inner = Xs.indices.astype(np.int32); vals = np.ascontiguousarray(Xs.data, dtype=dtype)
        Xmul = lambda v, w: np.asarray(Xs.T @ (v * w)).ravel().astype(dtype)
    else:
        Xd = np.asfortranarray(X, dtype=dtype)
        n, p = Xd.shape
        Xmul = lambda v, w: (Xd.T @ (v * w)).astype(dtype)
    y = glm["y"]; weights = glm["weights"]
    family = glm["family"]
    is_multi = family in ("multigaussian", "multinomial")
    K = y.shape[1] if is_multi else 1
    is_opt = family in ("gaussian", "multigaussian") and glm.get("opt", True)
    if offsets is None:
        offsets = np.zeros(y.shape, dtype=dtype)
    offsets = np.ascontiguousarray(offsets, dtype=dtype)
    if lmda_path is not None:
        lmda_path = np.array(np.flip(np.sort(lmda_path)), dtype=dtype)
    if groups is None:
        groups = np.arange(p, dtype=np.int64)
    groups = np.asarray(groups, dtype=np.int64)

    a = _PathArgs()
    keep = []   # keep arrays alive
    def P(arr):
        keep.append(arr); return _p(arr)

    if is_multi:
        groups = groups * K                                                     # solver.py:705
        if intercept:
            groups = np.concatenate([np.arange(K), K + groups]).astype(np.int64)
        group_sizes = np.diff(np.concatenate([groups, [(p + intercept) * K]])).astype(np.int64)
        if penalty is None:
            penalty = np.sqrt(group_sizes).astype(dtype)
            if intercept:
                penalty[:K] = 0
        else:
            penalty = np.asarray(penalty, dtype=dtype)
            if intercept:
                penalty = np.concatenate([np.zeros(K), penalty]).astype(dtype)
        p_aug = (p + intercept) * K
        def Xaug_mul(v, w):      # v, w flattened (n*K,) row-major
            V = (v * w).reshape(n, K)
            gx = np.stack([Xmul(np.ascontiguousarray(V[:, l]), np.ones(n, dtype=dtype)) for l in range(K)], axis=1).ravel()
            if intercept:
                return np.concatenate([V.sum(axis=0), gx]).astype(dtype)
            return gx.astype(dtype)
    else:
        group_sizes = np.diff(np.concatenate([groups, [p]])).astype(np.int64)
        if penalty is None:
            penalty = np.sqrt(group_sizes).astype(dtype)
        penalty = np.asarray(penalty, dtype=dtype)
        p_aug = p
    G = groups.size
    screen_set = np.arange(G)[(penalty <= 0) | (alpha <= 0)].astype(np.int64)
    screen_beta = np.zeros(int(np.sum(group_sizes[screen_set])), dtype=dtype)
    screen_is_active = np.ones(screen_set.shape[0], dtype=np.int8)
    active_set_size = screen_set.shape[0]
    active_set = np.zeros(G, dtype=np.int64)
    active_set[:active_set_size] = np.arange(active_set_size)

    ones = np.ones(n, dtype=dtype)
    if is_opt:
        a.family = 0
        if is_multi:                                                            # solver.py:766-816
            wms = weights / K
            X_means = np.repeat(Xmul(ones, wms), K)
            if intercept:
                X_means = np.concatenate([np.full(K, 1 / K), X_means]).astype(dtype)
            y_off = y - offsets
            y_var = np.sum(wms[:, None] * y_off ** 2)
            if intercept:
                y_off_c = y_off - (y_off.T @ weights)[None]
                yc_var = np.sum(wms[:, None] * y_off_c ** 2)
                rsq = yc_var - y_var
                y_var = yc_var
            else:
                rsq = 0
            resid = np.ascontiguousarray(y_off.ravel(), dtype=dtype)
            resid_sum = np.sum(wms[:, None] * y_off)
            grad = Xaug_mul(resid, np.repeat(wms, K))
            y_mean = 0.0
        else:                                                                   # solver.py:887-904
            X_means = Xmul(ones, weights)
            y_off = y - offsets
            y_mean = np.sum(y_off * weights)
            yc = y_off - y_mean if intercept else y_off
            y_var = np.sum(weights * yc ** 2)
            rsq = 0
            resid = np.ascontiguousarray(yc, dtype=dtype)
            resid_sum = np.sum(weights * resid)
            grad = Xmul(resid, weights)
        a.X_means = P(np.ascontiguousarray(X_means, dtype=dtype)); a.y_mean = float(y_mean); a.y_var = float(y_var)
        a.rsq = float(rsq); a.resid_sum = float(resid_sum)
        a.resid = P(resid); a.grad = P(np.ascontiguousarray(grad, dtype=dtype))
    else:                                                                       # solver.py:926-950 / 818-846
        a.family = FAMILIES[family]
        eta = np.ascontiguousarray(offsets, dtype=dtype)
        resid = glm_eval(glm, "gradient", eta=eta)
        if is_multi:
            grad = Xaug_mul(resid.ravel(), np.ones(n * K, dtype=dtype))
        else:
            grad = Xmul(resid, ones)
        a.eta = P(np.ascontiguousarray(eta.ravel(), dtype=dtype)); a.resid = P(np.ascontiguousarray(resid.ravel(), dtype=dtype))
        a.grad = P(np.ascontiguousarray(grad, dtype=dtype))
        a.beta0 = 0.0; a.loss_null = 0.0; a.setup_loss_null = 1
        a.loss_full = float(glm_eval(glm, "loss_full"))
        a.offsets = P(np.ascontiguousarray(offsets.ravel(), dtype=dtype))

    a.dtype = _dt(dtype); a.matrix_kind = 1 if is_sparse else 0
    if is_sparse:
        a.sp_outer = P(outer); a.sp_inner = P(inner); a.sp_values = P(vals)
    else:
        a.X = P(Xd)
    a.n = n; a.p = p; a.ld = n
    a.K = K; a.multi_intercept = int(intercept and is_multi)
    a.y = P(y); a.weights = P(weights)
    if family == "cox":
        a.cox_start = P(glm["start"]); a.cox_stop = P(glm["stop"]); a.cox_status = P(glm["status"]); a.cox_strata = P(glm["strata"])
        a.cox_efron = int(glm["efron"])
    a.groups = P(groups); a.group_sizes = P(group_sizes); a.G = G; a.penalty = P(penalty); a.alpha = alpha
    a.screen_set = P(screen_set); a.S = screen_set.size; a.screen_beta = P(screen_beta); a.screen_beta_size = screen_beta.size
    a.screen_is_active = P(screen_is_active); a.active_set_size = active_set_size; a.active_set = P(active_set)
    a.lmda = np.inf; a.lmda_max = -1.0
    a.setup_lmda_max = 1
    if lmda_path is None:
        a.setup_lmda_path = 1; a.lmda_path = None; a.lmda_path_len = 0
    else:
        a.setup_lmda_path = 0; a.lmda_path = P(lmda_path); a.lmda_path_len = lmda_path.size
    a.min_ratio = min_ratio; a.lmda_path_size = lmda_path_size
    a.max_screen_size = G if max_screen_size is None else min(max_screen_size, G)
    a.max_active_size = G if max_active_size is None else min(max_active_size, G)
    a.pivot_subset_ratio = pivot_subset_ratio; a.pivot_subset_min = pivot_subset_min; a.pivot_slack_ratio = pivot_slack_ratio
    a.screen_rule = {"strong": 0, "pivot": 1}[screen_rule]
    a.max_iters = max_iters; a.tol = tol; a.adev_tol = adev_tol; a.ddev_tol = ddev_tol; a.newton_tol = newton_tol
    a.newton_max_iters = newton_max_iters; a.irls_max_iters = irls_max_iters; a.irls_tol = irls_tol
    a.early_exit = int(early_exit); a.intercept = int(intercept); a.n_threads = n_threads
    a.max_seconds = max_seconds

    h = lib().orc_path_solve(C.byref(a))
    out = _collect(h, p_aug, dtype)
    out.groups = groups; out.group_sizes = group_sizes; out.penalty = penalty
    if is_multi:                                                                # tidy: solver_multigaussian_naive.hpp:31-43
        B = out.betas.tocsc()
        if intercept:
            out.intercepts = np.asarray(B[:, :K].todense()).astype(dtype)
            out.betas = sp.csr_matrix(B[:, K:])
        else:
            out.intercepts = np.zeros((out.betas.shape[0], K), dtype=dtype)
    return out


# --------------------------------------------------------------------------
# covariance method (oracle/cov_oracle.hpp): MatrixCov operators, gaussian_pin_cov, gaussian_cov
# --------------------------------------------------------------------------
class _cov_matrix:
    """MatrixCovDense / MatrixCovLazyCov restated (CORE/matrix/matrix_cov_{dense,lazy_cov}.ipp).  The lazy cache lives for one call."""

    def __init__(self, mat, kind):
        mat = np.asarray(mat)
        if not (mat.flags.c_contiguous or mat.flags.f_contiguous):
            mat = np.asfortranarray(mat)
        self.mat = mat
        self.kind = kind
        self.dtype = mat.dtype
        self.row_major = bool(mat.flags.c_contiguous and not (mat.flags.f_contiguous and mat.shape[0] != 1 and mat.shape[1] != 1))
        self.n, self.p = (mat.shape if kind == 1 else (0, mat.shape[1]))
        self.ld = mat.shape[1] if self.row_major else mat.shape[0]

    def cols(self):
        return self.p

    def _op(self, op, subset, indices, values, i0, q, out):
        subset = None if subset is None else np.ascontiguousarray(subset, dtype=np.int64)
        indices = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
        values = None if values is None else np.ascontiguousarray(values, dtype=self.dtype)
        lib().orc_cov_matrix_op(_dt(self.dtype), self.kind, _p(self.mat), self.n, self.p, self.ld, int(self.row_major), op,
                                _p(subset), 0 if subset is None else subset.size, _p(indices), _p(values),
                                0 if indices is None else indices.size, i0, q, _p(out))

    def bmul(self, subset, indices, values, out):
        self._op(0, subset, indices, values, 0, 0, out)

    def mul(self, indices, values, out):
        self._op(1, None, indices, values, 0, 0, out)

    def to_dense(self, i, q, out):
        tmp = np.empty(q * q, dtype=self.dtype)
        self._op(2, None, None, None, i, q, tmp)
        out[...] = tmp.reshape(q, q, order="F")


def cov_dense(A):
    return _cov_matrix(A, 0)


def cov_lazy(X):
    return _cov_matrix(X, 1)


def _cov_args(A, keep, *, groups, alpha, penalty, max_iters, tol, rdev_tol, newton_tol, newton_max_iters,
              max_active_size=None, max_screen_size=None):
    if isinstance(A, np.ndarray):
        A = cov_dense(A)
    dtype = A.dtype
    p = A.cols()
    def P(arr):
        keep.append(arr); return _p(arr)
    groups = np.arange(p, dtype=np.int64) if groups is None else np.ascontiguousarray(groups, dtype=np.int64)
    G = groups.size
    group_sizes = np.diff(np.concatenate([groups, [p]])).astype(np.int64)
    penalty = np.sqrt(group_sizes).astype(dtype) if penalty is None else np.ascontiguousarray(penalty, dtype=dtype)
    a = _CovArgs()
    a.dtype = _dt(dtype); a.matrix_kind = A.kind; a.M = P(A.mat); a.n = A.n; a.p = p; a.ld = A.ld; a.row_major = int(A.row_major)
    a.groups = P(groups); a.group_sizes = P(group_sizes); a.G = G; a.penalty = P(penalty); a.alpha = alpha
    a.max_iters = max_iters; a.tol = tol; a.rdev_tol = rdev_tol; a.newton_tol = newton_tol; a.newton_max_iters = newton_max_iters
    a.max_screen_size = G if max_screen_size is None else min(max_screen_size, G)
    a.max_active_size = G if max_active_size is None else min(max_active_size, G)
    a.pivot_subset_ratio = 0.1; a.pivot_subset_min = 1; a.pivot_slack_ratio = 1.25; a.screen_rule = 1
    a.min_ratio = 1e-2; a.lmda_path_size = 100
    return A, a, P, groups, group_sizes, penalty, dtype, p, G


def gaussian_pin_cov(A, *, groups, alpha, penalty, screen_set, lmda_path, rsq, screen_beta, screen_grad, screen_is_active,
                     active_set_size, active_set, max_active_size=None, max_iters=int(1e5), tol=1e-7, rdev_tol=1e-4,
                     newton_tol=1e-12, newton_max_iters=1000):
    """StateGaussianPinCov.solve: the gaussian_pin_cov wrapper (adelie/state.py:739-1000) + pin::cov::solve
    (CORE/solver/solver_gaussian_pin_cov.hpp:529-777)."""
    keep = []
    A, a, P, groups, group_sizes, penalty, dtype, p, G = _cov_args(
        A, keep, groups=groups, alpha=alpha, penalty=penalty, max_iters=max_iters, tol=tol, rdev_tol=rdev_tol,
        newton_tol=newton_tol, newton_max_iters=newton_max_iters, max_active_size=max_active_size)
    screen_set = np.ascontiguousarray(screen_set, dtype=np.int64)
    act = np.zeros(G, dtype=np.int64)
    act[:active_set_size] = np.asarray(active_set)[:active_set_size]
    sb = np.ascontiguousarray(screen_beta, dtype=dtype)
    a.screen_set = P(screen_set); a.S = screen_set.size; a.screen_beta = P(sb); a.screen_beta_size = sb.size
    a.screen_is_active = P(np.ascontiguousarray(screen_is_active, dtype=np.int8)); a.active_set_size = active_set_size; a.active_set = P(act)
    a.rsq = rsq; a.lmda = np.inf; a.lmda_max = -1.0
    a.screen_grad = P(np.ascontiguousarray(screen_grad, dtype=dtype))
    lp = np.ascontiguousarray(lmda_path, dtype=dtype)
    a.lmda_path = P(lp); a.lmda_path_len = lp.size
    h = lib().orc_cov_pin_solve(C.byref(a))
    out = _collect(h, p, dtype)
    out.screen_is_active = out.screen_is_active.astype(bool)
    return out


def gaussian_cov(A, v, *, groups=None, alpha=1.0, penalty=None, lmda_path=None, max_iters=int(1e5), tol=1e-7, rdev_tol=1e-3,
                 newton_tol=1e-12, newton_max_iters=1000, early_exit=True, screen_rule="pivot", min_ratio=1e-2, lmda_path_size=100,
                 max_screen_size=None, max_active_size=None, pivot_subset_ratio=0.1, pivot_subset_min=1, pivot_slack_ratio=1.25):
    """adelie.solver.gaussian_cov (adelie/solver.py:39-352: initial invariants) + gaussian::cov::solve
    (CORE/solver/solver_gaussian_cov.hpp:359-457)."""
    keep = []
    A, a, P, groups, group_sizes, penalty, dtype, p, G = _cov_args(
        A, keep, groups=groups, alpha=alpha, penalty=penalty, max_iters=max_iters, tol=tol, rdev_tol=rdev_tol,
        newton_tol=newton_tol, newton_max_iters=newton_max_iters, max_active_size=max_active_size, max_screen_size=max_screen_size)
    if lmda_path is not None:
        lmda_path = np.array(np.flip(np.sort(lmda_path)), dtype=dtype)
    v = np.ascontiguousarray(v, dtype=dtype)
    screen_set = np.arange(G)[(penalty <= 0) | (alpha <= 0)].astype(np.int64)
    screen_beta = np.zeros(int(np.sum(group_sizes[screen_set])), dtype=dtype)
    screen_is_active = np.ones(screen_set.shape[0], dtype=np.int8)
    active_set = np.zeros(G, dtype=np.int64)
    active_set[:screen_set.size] = np.arange(screen_set.size)
    grad = v.copy()                                                               # v - A @ 0
    a.v = P(v); a.grad = P(grad)
    a.screen_set = P(screen_set); a.S = screen_set.size; a.screen_beta = P(screen_beta); a.screen_beta_size = screen_beta.size
    a.screen_is_active = P(screen_is_active); a.active_set_size = screen_set.size; a.active_set = P(active_set)
    a.rsq = 0.0; a.lmda = np.inf; a.lmda_max = -1.0; a.setup_lmda_max = 1
    if lmda_path is None:
        a.setup_lmda_path = 1; a.lmda_path = None; a.lmda_path_len = 0
    else:
        a.setup_lmda_path = 0; a.lmda_path = P(lmda_path); a.lmda_path_len = lmda_path.size
    a.min_ratio = min_ratio; a.lmda_path_size = lmda_path_size
    a.pivot_subset_ratio = pivot_subset_ratio; a.pivot_subset_min = pivot_subset_min; a.pivot_slack_ratio = pivot_slack_ratio
    a.screen_rule = {"strong": 0, "pivot": 1}[screen_rule]
    a.early_exit = int(early_exit)
    h = lib().orc_cov_path_solve(C.byref(a))
    out = _collect(h, p, dtype)
    out.groups = groups; out.group_sizes = group_sizes; out.penalty = penalty
    out.screen_is_active = out.screen_is_active.astype(bool)
    return out
