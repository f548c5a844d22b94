"""State objects (reference: adelie/state.py; core classes adelie/src/py_state.cpp).

A state owns the NumPy buffers it was built from (``_name`` attributes, as the reference) and a
core handle (``ab_state``) that lives on the device.  ``solve()`` follows the reference contract
(adelie/state.py:157-176, py_state.cpp:62-145): the state is *copied* (a fresh core state is built
from the stored inputs), the copy is solved and returned as a new state object with ``error`` and
``total_time`` attached; solver errors never raise.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np
import scipy.sparse

from . import _lib
from . import glm as _glm
from . import matrix as _matrix

logger = logging.getLogger("adelie_b200")

_VEC_F = ["lmda_path", "screen_beta", "grad", "abs_grad", "devs", "lmdas", "X_means", "screen_X_means", "screen_vars",
          "resid", "eta", "sweep_stats", "benchmark_screen", "benchmark_fit_screen", "benchmark_fit_active", "benchmark_kkt",
          "benchmark_invariance", "launch_cols", "launch_sweeps", "launch_ms"]
_VEC_I = ["screen_set", "screen_begins", "screen_is_active", "active_set", "n_valid_solutions", "active_sizes", "screen_sizes"]
_SCALARS = ["lmda_max", "lmda", "rsq", "resid_sum", "y_mean", "y_var", "loss_null", "loss_full", "beta0", "active_set_size",
            "n_sweeps", "n_group_updates", "n_col_updates", "n_irls", "n_pin_solves", "n_kernel_launches", "time_sweep_kernel", "sweep_ncta",
            "sweep_stages", "sweep_smem_bytes", "sweep_staged", "sweep_threads", "sweep_batch", "n_panels_built", "n_batched_launches"]


def _render_inputs(*, groups, lmda_max, lmda_path, lmda_path_size, max_screen_size, max_active_size, dtype):
    """adelie/state.py:1385-1418 (_render_gaussian_naive_inputs)"""
    G = groups.shape[0]
    if max_screen_size is None:
        max_screen_size = G
    if max_active_size is None:
        max_active_size = G
    max_screen_size = int(np.minimum(max_screen_size, G))
    max_active_size = int(np.minimum(max_active_size, G))
    setup_lmda_max = lmda_max is None
    setup_lmda_path = lmda_path is None
    if setup_lmda_max:
        lmda_max = -1
    if setup_lmda_path:
        lmda_path = np.empty(0, dtype=dtype)
    else:
        lmda_path_size = len(lmda_path)
    return max_screen_size, max_active_size, lmda_path_size, setup_lmda_max, setup_lmda_path, lmda_max, lmda_path


class base:
    """Common wrapper (adelie/state.py:79-176)."""
    _is_multi = False

    def _build_core(self):
        a = _lib.StateArgs()
        keep = self._fill_args(a)
        h = C.c_void_p()
        X = self._X
        glm_handle = self._glm._core() if self._use_glm else None
        _lib.check(_lib.load().ab_state_create(C.byref(a), X._core(), glm_handle, C.byref(h)))
        del keep
        return h

    def _core(self):
        """The core handle; (re)built from the stored inputs when ``solve()`` handed the previous one to the returned state."""
        if self._handle is None:
            if not self.__dict__.get("_pristine", True):
                raise RuntimeError("adelie_b200: this solved state was closed; its results are gone.")
            self._handle = self._build_core()
        return self._handle

    def close(self):
        """Frees the core state's device buffers now (extension; otherwise freed when the last reference is dropped)."""
        h, self._handle = self.__dict__.get("_handle"), None
        if h is not None:
            try:
                _lib.load().ab_state_free(h)
            except Exception:
                pass

    def __del__(self):
        self.close()

    # ---- output accessors ------------------------------------------------------------------
    def _scalar(self, name):
        out = C.c_double()
        _lib.check(_lib.load().ab_state_get_scalar(self._core(), name.encode(), C.byref(out)))
        return out.value

    def _vec_f(self, name, dtype=None):
        L = _lib.load()
        n = C.c_int64()
        _lib.check(L.ab_state_get_vec_f64(self._core(), name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.float64)
        _lib.check(L.ab_state_get_vec_f64(self._core(), name.encode(), _lib.ptr(buf), n.value, C.byref(n)))
        if name.startswith("benchmark") or name == "sweep_stats":
            return buf
        return buf.astype(self._dtype if dtype is None else dtype)

    def _vec_i(self, name):
        L = _lib.load()
        n = C.c_int64()
        _lib.check(L.ab_state_get_vec_i64(self._core(), name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.int64)
        _lib.check(L.ab_state_get_vec_i64(self._core(), name.encode(), _lib.ptr(buf), n.value, C.byref(n)))
        return buf

    def __getattr__(self, name):
        # only called when normal lookup fails: resolve core fields lazily
        if name.startswith("_"):
            raise AttributeError(name)
        if name in _SCALARS:
            v = self._scalar(name)
            return int(v) if name in ("active_set_size", "n_sweeps", "n_group_updates", "n_col_updates", "n_irls", "n_pin_solves",
                                      "n_kernel_launches", "sweep_ncta", "sweep_stages", "sweep_smem_bytes",
                                      "sweep_staged", "sweep_threads", "sweep_batch", "n_panels_built", "n_batched_launches") else self._dtype(v) if name not in ("time_sweep_kernel",) else v
        if name.startswith("t_"):
            return self._scalar(name)
        if name in _VEC_F:
            return self._vec_f(name)
        if name in _VEC_I:
            v = self._vec_i(name)
            if name == "screen_is_active":
                return v.astype(bool)
            if name in ("n_valid_solutions", "active_sizes", "screen_sizes"):
                return v.astype(np.int32)
            return v
        raise AttributeError(name)

    @property
    def X(self):
        return self._X

    @property
    def groups(self):
        return self._groups

    @property
    def group_sizes(self):
        return self._group_sizes

    @property
    def penalty(self):
        return self._penalty

    @property
    def alpha(self):
        return self._cfg["alpha"]

    @property
    def screen_hashset(self):
        return set(self.screen_set.tolist())

    @property
    def screen_transforms(self):
        L = _lib.load()
        out = []
        for i in range(self.screen_set.shape[0]):
            n = C.c_int64()
            _lib.check(L.ab_state_get_screen_transform(self._core(), i, None, 0, C.byref(n)))
            buf = np.empty(n.value, dtype=np.float64)
            _lib.check(L.ab_state_get_screen_transform(self._core(), i, _lib.ptr(buf), n.value, C.byref(n)))
            gs = int(round(np.sqrt(n.value)))
            out.append(buf.astype(self._dtype).reshape(gs, gs))
        return out

    @property
    def betas(self):
        """(L, p) scipy CSR with int64 indices (py_state.cpp:9-60)."""
        L = _lib.load()
        nnz, nl = C.c_int64(), C.c_int64()
        _lib.check(L.ab_state_get_betas(self._core(), None, None, None, C.byref(nnz), C.byref(nl)))
        indptr = np.empty(nl.value + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int64)
        values = np.empty(nnz.value, dtype=np.float64)
        _lib.check(L.ab_state_get_betas(self._core(), _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(values), C.byref(nnz), C.byref(nl)))
        p = self._p_cols
        B = scipy.sparse.csr_matrix((values.astype(self._dtype), indices, indptr), shape=(nl.value, p))
        B.indices = B.indices.astype(np.int64); B.indptr = B.indptr.astype(np.int64)     # int64 like the reference (py_state.cpp:9-60)
        return B

    @property
    def intercepts(self):
        return self._vec_f("intercepts")

    def __getstate__(self):
        raise RuntimeError("adelie_b200 states hold device memory and cannot be pickled.")

    # ---- solve -----------------------------------------------------------------------------
    def solve(self, progress_bar: bool = True, exit_cond=None):
        new = self._clone()
        L = _lib.load()
        err = C.create_string_buffer(4096)
        total = C.c_double()
        cb_exit = None
        if exit_cond is not None:
            cb_exit = _lib.EXIT_COND_T(lambda ctx: int(bool(exit_cond(new))))
        # interrupt polling (py_state.cpp:70-74): the core calls back once per pin solve; a pending KeyboardInterrupt (or any
        # exception raised by a Python signal handler) stops the solve and is re-raised here
        pending = []
        def _poll():
            try:
                C.pythonapi.PyErr_CheckSignals()
            except BaseException as e:          # noqa: BLE001 -- whatever the signal handler raised
                pending.append(e)
                return 1
            return 0
        cb_sig = _lib.CHECK_SIGNALS_T(_poll)
        rc = L.ab_state_solve(new._handle, int(progress_bar), C.cast(cb_exit, C.c_void_p) if cb_exit else None, None,
                              C.cast(cb_sig, C.c_void_p), err, len(err), C.byref(total))
        if pending:
            raise pending[0]
        _lib.check(rc)
        new.error = err.value.decode()
        new.total_time = total.value
        if new.error != "":                                       # adelie/state.py:160-166
            if new.error.startswith("adelie_core solver: "):
                logger.error(RuntimeError(new.error))
            else:
                logger.warning(RuntimeError(new.error))
        return new

    def _clone(self):
        """The copy ``solve`` works on (py_state.cpp:62-145 takes the state by value).  The core state of ``self`` is pristine -- every
        accessor is read-only -- so it is *moved* into the copy instead of building a second one (screen-derived Grams, uploads); ``self``
        rebuilds its own lazily if it is read again."""
        obj = object.__new__(type(self))
        for k, v in self.__dict__.items():
            if k != "_handle":
                setattr(obj, k, v)
        obj._handle = None
        if self.__dict__.get("_pristine", True):
            obj._handle, self._handle = self._handle, None
        if obj._handle is None:
            obj._handle = obj._build_core()      # a solved state re-solves from its stored inputs; its own results stay
        obj._pristine = False
        return obj

    def check(self, method=None, logger=logger):
        return


class _Naive(base):
    """Shared construction of the four naive states."""
    def __init__(self, *, X, glm_obj, use_glm, dtype, cfg, arrays, p_cols):
        self._X = X
        self._glm = glm_obj
        self._use_glm = use_glm
        self._dtype = np.dtype(dtype).type
        self._cfg = cfg
        self._p_cols = p_cols
        for k, v in arrays.items():
            setattr(self, "_" + k, v)
        self._handle = None
        self._handle = self._build_core()

    def _fill_args(self, a):
        c = self._cfg
        dtype = self._dtype
        a.dtype = _lib.dtype_code(dtype)
        a.groups = _lib.ptr(self._groups); a.group_sizes = _lib.ptr(self._group_sizes); a.G = self._groups.shape[0]
        a.alpha = c["alpha"]; a.penalty = _lib.ptr(self._penalty)
        if not self._use_glm:
            a.weights = _lib.ptr(self._weights); a.X_means = _lib.ptr(self._X_means)
            a.y_mean = c["y_mean"]; a.y_var = c["y_var"]; a.resid_sum = c["resid_sum"]; a.rsq = c["rsq"]
        else:
            a.offsets = _lib.ptr(self._offsets); a.eta = _lib.ptr(self._eta)
            a.beta0 = c["beta0"]; a.loss_full = c["loss_full"]
            a.setup_loss_null = int(c["loss_null"] is None)
            a.loss_null = 0.0 if c["loss_null"] is None else c["loss_null"]
            a.irls_max_iters = c["irls_max_iters"]; a.irls_tol = c["irls_tol"]
        a.resid = _lib.ptr(self._resid)
        a.n_classes = c.get("n_classes", 1); a.multi_intercept = int(c.get("multi_intercept", False))
        a.lmda_path = _lib.ptr(self._lmda_path); a.lmda_path_len = self._lmda_path.shape[0]
        a.lmda_max = c["lmda_max"]; a.min_ratio = c["min_ratio"]; a.lmda_path_size = c["lmda_path_size"]
        a.setup_lmda_max = int(c["setup_lmda_max"]); a.setup_lmda_path = int(c["setup_lmda_path"])
        a.max_screen_size = c["max_screen_size"]; a.max_active_size = c["max_active_size"]
        a.pivot_subset_ratio = c["pivot_subset_ratio"]; a.pivot_subset_min = c["pivot_subset_min"]
        a.pivot_slack_ratio = c["pivot_slack_ratio"]
        rules = {"strong": 0, "pivot": 1}
        if c["screen_rule"] not in rules:
            raise RuntimeError("adelie_core: Invalid screen rule type: " + str(c["screen_rule"]))
        a.screen_rule = rules[c["screen_rule"]]
        a.max_iters = c["max_iters"]; a.tol = c["tol"]; a.adev_tol = c["adev_tol"]; a.ddev_tol = c["ddev_tol"]
        a.newton_tol = c["newton_tol"]; a.newton_max_iters = c["newton_max_iters"]
        a.early_exit = int(c["early_exit"]); a.intercept = int(c["intercept"]); a.n_threads = c["n_threads"]
        a.screen_set = _lib.ptr(self._screen_set); a.screen_set_size = self._screen_set.shape[0]
        a.screen_beta = _lib.ptr(self._screen_beta); a.screen_beta_size = self._screen_beta.shape[0]
        a.screen_is_active = _lib.ptr(self._screen_is_active_i8); a.active_set_size = c["active_set_size"]
        a.active_set = _lib.ptr(self._active_set)
        a.lmda = c["lmda"]; a.grad = _lib.ptr(self._grad)
        return None

    # static configuration read-back (Appendix B of SURVEY.md)
    def __getattr__(self, name):
        cfg = self.__dict__.get("_cfg", {})
        if name in ("min_ratio", "lmda_path_size", "max_screen_size", "max_active_size", "pivot_subset_ratio",
                    "pivot_subset_min", "pivot_slack_ratio", "screen_rule", "max_iters", "tol", "adev_tol", "ddev_tol",
                    "newton_tol", "newton_max_iters", "early_exit", "setup_lmda_max", "setup_lmda_path", "intercept",
                    "n_threads", "irls_max_iters", "irls_tol", "n_classes", "multi_intercept") and name in cfg:
            return cfg[name]
        return base.__getattr__(self, name)

    @property
    def weights(self):
        return self._glm.weights

    @property
    def offsets(self):
        return self._offsets

    @property
    def constraints(self):
        return [None] * self._groups.shape[0]

    @property
    def dual_groups(self):
        return np.zeros(self._groups.shape[0], dtype=int)


def _common_arrays(*, groups, group_sizes, penalty, lmda_path, screen_set, screen_beta, screen_is_active, active_set, grad,
                   resid, dtype):
    G = groups.shape[0]
    act = np.zeros(G, dtype=np.int64)
    active_set = np.asarray(active_set)
    act[: min(G, active_set.shape[0])] = active_set[:G]
    return dict(
        groups=np.array(groups, copy=True, dtype=np.int64),
        group_sizes=np.array(group_sizes, copy=True, dtype=np.int64),
        penalty=np.array(penalty, copy=True, dtype=dtype),
        lmda_path=np.ascontiguousarray(lmda_path, dtype=dtype),
        screen_set=np.ascontiguousarray(screen_set, dtype=np.int64),
        screen_beta=np.ascontiguousarray(screen_beta, dtype=dtype),
        screen_is_active_i8=np.ascontiguousarray(screen_is_active, dtype=np.int8),
        active_set=act,
        grad=np.ascontiguousarray(grad, dtype=dtype),
        resid=np.ascontiguousarray(np.asarray(resid).ravel(), dtype=dtype),
    )


def _check_constraints(constraints):
    if constraints is not None and any(c is not None for c in constraints):
        raise RuntimeError("adelie_b200: constraints are out of scope for the B200 path (pass constraints=None).")


def gaussian_naive(*, X, y, X_means, y_mean, y_var, resid, resid_sum, constraints, groups, group_sizes, alpha, penalty,
                   weights, offsets, screen_set, screen_beta, screen_is_active, active_set_size, active_set, rsq, lmda, grad,
                   lmda_path=None, lmda_max=None, max_iters=int(1e5), tol=1e-7, adev_tol=0.9, ddev_tol=0, newton_tol=1e-12,
                   newton_max_iters=1000, n_threads=1, early_exit=True, intercept=True, screen_rule="pivot", min_ratio=1e-2,
                   lmda_path_size=100, max_screen_size=None, max_active_size=None, pivot_subset_ratio=0.1, pivot_subset_min=1,
                   pivot_slack_ratio=1.25):
    """Gaussian naive-method state (adelie/state.py:1677-2024; core StateGaussianNaive)."""
    _check_constraints(constraints)
    if isinstance(X, np.ndarray):
        X = _matrix.dense(X, method="naive", n_threads=n_threads)
    dtype = X.dtype
    groups = np.asarray(groups)
    (max_screen_size, max_active_size, lmda_path_size, setup_lmda_max, setup_lmda_path, lmda_max, lmda_path) = _render_inputs(
        groups=groups, lmda_max=lmda_max, lmda_path=lmda_path, lmda_path_size=lmda_path_size,
        max_screen_size=max_screen_size, max_active_size=max_active_size, dtype=dtype)
    glm_obj = _glm.gaussian(y=np.asarray(y, dtype=dtype), weights=weights, dtype=dtype)
    arrays = _common_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=lmda_path, screen_set=screen_set,
                            screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set, grad=grad,
                            resid=resid, dtype=dtype)
    arrays["weights"] = glm_obj.weights
    arrays["X_means"] = np.array(X_means, copy=True, dtype=dtype)
    arrays["offsets"] = np.array(offsets, copy=True, dtype=dtype)
    cfg = dict(alpha=float(alpha), y_mean=float(y_mean), y_var=float(y_var), resid_sum=float(resid_sum), rsq=float(rsq),
               lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size), setup_lmda_max=setup_lmda_max,
               setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size, max_active_size=max_active_size,
               pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
               screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, intercept=intercept,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float(lmda))
    return _Naive(X=X, glm_obj=glm_obj, use_glm=False, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=X.cols())


def glm_naive(*, X, glm, constraints, groups, group_sizes, alpha, penalty, offsets, screen_set, screen_beta, screen_is_active,
              active_set_size, active_set, beta0, lmda, grad, eta, resid, loss_full, loss_null=None, lmda_path=None,
              lmda_max=None, irls_max_iters=int(1e4), irls_tol=1e-7, max_iters=int(1e5), tol=1e-7, adev_tol=0.9, ddev_tol=0,
              newton_tol=1e-12, newton_max_iters=1000, n_threads=1, early_exit=True, intercept=True, screen_rule="pivot",
              min_ratio=1e-2, lmda_path_size=100, max_screen_size=None, max_active_size=None, pivot_subset_ratio=0.1,
              pivot_subset_min=1, pivot_slack_ratio=1.25):
    """GLM naive-method state (adelie/state.py:2407-2753; core StateGlmNaive)."""
    _check_constraints(constraints)
    if isinstance(X, np.ndarray):
        X = _matrix.dense(X, method="naive", n_threads=n_threads)
    dtype = X.dtype
    groups = np.asarray(groups)
    (max_screen_size, max_active_size, lmda_path_size, setup_lmda_max, setup_lmda_path, lmda_max, lmda_path) = _render_inputs(
        groups=groups, lmda_max=lmda_max, lmda_path=lmda_path, lmda_path_size=lmda_path_size,
        max_screen_size=max_screen_size, max_active_size=max_active_size, dtype=dtype)
    arrays = _common_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=lmda_path, screen_set=screen_set,
                            screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set, grad=grad,
                            resid=resid, dtype=dtype)
    arrays["offsets"] = np.array(np.asarray(offsets).ravel(), copy=True, dtype=dtype)
    arrays["eta"] = np.array(np.asarray(eta).ravel(), copy=True, dtype=dtype)
    cfg = dict(alpha=float(alpha), beta0=float(beta0), loss_null=None if loss_null is None else float(loss_null),
               loss_full=float(loss_full), irls_max_iters=int(irls_max_iters), irls_tol=irls_tol,
               lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size), setup_lmda_max=setup_lmda_max,
               setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size, max_active_size=max_active_size,
               pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
               screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, intercept=intercept,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float(lmda))
    return _Naive(X=X, glm_obj=glm, use_glm=True, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=X.cols())


class _MultiNaive(_Naive):
    """Multi-response states: the core solves the reformulated single-response problem; ``betas`` / ``intercepts`` are tidied as
    solver_multigaussian_naive.hpp:31-43 / solver_multiglm_naive.hpp:205-224 do (first K coefficients -> (L, K) intercepts)."""
    _is_multi = True

    def _raw_betas(self):
        return base.betas.fget(self)

    @property
    def betas(self):
        B = self._raw_betas()
        K = self._cfg["n_classes"]
        if not self._cfg["multi_intercept"]:
            return B
        B = scipy.sparse.csr_matrix(B.tocsc()[:, K:])
        B.indices = B.indices.astype(np.int64); B.indptr = B.indptr.astype(np.int64)
        return B

    @property
    def intercepts(self):
        K = self._cfg["n_classes"]
        B = self._raw_betas()
        if not self._cfg["multi_intercept"]:
            return np.zeros((B.shape[0], K), dtype=self._dtype)
        return np.asarray(B.tocsc()[:, :K].todense()).astype(self._dtype).reshape(B.shape[0], K)

    @property
    def X_expanded(self):
        return self._X_expanded


def _render_multi_inputs(*, X, offsets, intercept, n_threads, dtype):
    """adelie/state.py:1100-1125"""
    offsets = np.asarray(offsets, order="C", dtype=dtype)
    n, K = offsets.shape
    Xe = _matrix.kronecker_eye(X, K, n_threads=n_threads)
    if intercept:
        Xe = _matrix.concatenate([_matrix.kronecker_eye(np.ones((n, 1), dtype=dtype), K, n_threads=n_threads), Xe], axis=1,
                                 n_threads=n_threads)
    return Xe, offsets


def multigaussian_naive(*, X, y, X_means, y_var, resid, resid_sum, constraints, groups, group_sizes, alpha, penalty, weights,
                        offsets, screen_set, screen_beta, screen_is_active, active_set_size, active_set, rsq, lmda, grad,
                        lmda_path=None, lmda_max=None, max_iters=int(1e5), tol=1e-7, adev_tol=0.9, ddev_tol=0, newton_tol=1e-12,
                        newton_max_iters=1000, n_threads=1, early_exit=True, intercept=True, screen_rule="pivot", min_ratio=1e-2,
                        lmda_path_size=100, max_screen_size=None, max_active_size=None, pivot_subset_ratio=0.1,
                        pivot_subset_min=1, pivot_slack_ratio=1.25):
    """Multi-response Gaussian naive state (adelie/state.py:2027-2391; core StateMultiGaussianNaive)."""
    _check_constraints(constraints)
    if isinstance(X, np.ndarray):
        X = _matrix.dense(X, method="naive", n_threads=n_threads)
    dtype = X.dtype
    groups = np.asarray(groups)
    (max_screen_size, max_active_size, lmda_path_size, setup_lmda_max, setup_lmda_path, lmda_max, lmda_path) = _render_inputs(
        groups=groups, lmda_max=lmda_max, lmda_path=lmda_path, lmda_path_size=lmda_path_size,
        max_screen_size=max_screen_size, max_active_size=max_active_size, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    K = y.shape[-1]
    Xe, offsets = _render_multi_inputs(X=X, offsets=offsets, intercept=intercept, n_threads=n_threads, dtype=dtype)
    assert np.asarray(X_means).shape[0] == Xe.cols(), "X_means must have the same length as the number of columns of X after reshaping."
    glm_obj = _glm.multigaussian(y=y, weights=weights, dtype=dtype)
    arrays = _common_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=lmda_path, screen_set=screen_set,
                            screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set, grad=grad,
                            resid=resid, dtype=dtype)
    arrays["weights"] = np.ascontiguousarray(np.repeat(glm_obj.weights, K) / K, dtype=dtype)      # state.py:2315
    arrays["X_means"] = np.array(X_means, copy=True, dtype=dtype)
    arrays["offsets"] = np.array(offsets, copy=True, dtype=dtype)
    arrays["X_expanded"] = Xe
    # not the actual y_mean: a value that yields the right loss_null / loss_full (state.py:2339-2343)
    y_mean = np.linalg.norm(np.asarray(_dist_sum(np.sum(glm_obj.weights[:, None] * (y - offsets), axis=0))) / K)
    cfg = dict(alpha=float(alpha), y_mean=float(y_mean), y_var=float(y_var), resid_sum=float(resid_sum), rsq=float(rsq),
               lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size), setup_lmda_max=setup_lmda_max,
               setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size, max_active_size=max_active_size,
               pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
               screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, intercept=False,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float(lmda), n_classes=int(K),
               multi_intercept=bool(intercept))
    return _MultiNaive(X=X, glm_obj=glm_obj, use_glm=False, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=Xe.cols())


def _dist_sum(v):
    from . import dist as _dist
    return _dist.allreduce(v)


def multiglm_naive(*, X, glm, constraints, groups, group_sizes, alpha, penalty, offsets, screen_set, screen_beta,
                   screen_is_active, active_set_size, active_set, lmda, grad, eta, resid, loss_full, loss_null=None,
                   lmda_path=None, lmda_max=None, irls_max_iters=int(1e4), irls_tol=1e-7, max_iters=int(1e5), tol=1e-7,
                   adev_tol=0.9, ddev_tol=0, newton_tol=1e-12, newton_max_iters=1000, n_threads=1, early_exit=True,
                   intercept=True, screen_rule="pivot", min_ratio=1e-2, lmda_path_size=100, max_screen_size=None,
                   max_active_size=None, pivot_subset_ratio=0.1, pivot_subset_min=1, pivot_slack_ratio=1.25):
    """Multi-response GLM naive state (adelie/state.py:2756-3100; core StateMultiGlmNaive)."""
    _check_constraints(constraints)
    if isinstance(X, np.ndarray):
        X = _matrix.dense(X, method="naive", n_threads=n_threads)
    dtype = X.dtype
    groups = np.asarray(groups)
    (max_screen_size, max_active_size, lmda_path_size, setup_lmda_max, setup_lmda_path, lmda_max, lmda_path) = _render_inputs(
        groups=groups, lmda_max=lmda_max, lmda_path=lmda_path, lmda_path_size=lmda_path_size,
        max_screen_size=max_screen_size, max_active_size=max_active_size, dtype=dtype)
    K = glm.y.shape[-1]
    Xe, offsets = _render_multi_inputs(X=X, offsets=offsets, intercept=intercept, n_threads=n_threads, dtype=dtype)
    arrays = _common_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=lmda_path, screen_set=screen_set,
                            screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set, grad=grad,
                            resid=resid, dtype=dtype)
    arrays["offsets"] = np.array(np.asarray(offsets).ravel(), copy=True, dtype=dtype)
    arrays["eta"] = np.array(np.asarray(eta).ravel(), copy=True, dtype=dtype)
    arrays["X_expanded"] = Xe
    cfg = dict(alpha=float(alpha), beta0=0.0, loss_null=None if loss_null is None else float(loss_null),
               loss_full=float(loss_full), irls_max_iters=int(irls_max_iters), irls_tol=irls_tol,
               lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size), setup_lmda_max=setup_lmda_max,
               setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size, max_active_size=max_active_size,
               pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
               screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, intercept=False,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float(lmda), n_classes=int(K),
               multi_intercept=bool(intercept))
    return _MultiNaive(X=X, glm_obj=glm, use_glm=True, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=Xe.cols())
