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

_VEC_F = ["rsqs", "lmda_path", "screen_beta", "grad", "abs_grad", "devs", "lmdas", "X_means", "screen_X_means", "screen_vars",
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
    _api = "ab_state"        # prefix of the C-ABI entry points behind this state ("ab_cov_state" for the covariance-method states)

    def _fn(self, suffix):
        return getattr(_lib.load(), self._api + suffix)

    def _build_core(self):
        a = _lib.StateArgs()
        keep = self._fill_args(a)
        h = C.c_void_p()
        X = self.__dict__.get("_X_core") or self._X      # sparse multi-response states solve on the expanded CSC matrix
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
                self._fn("_free")(h)
            except Exception:
                pass

    def __del__(self):
        self.close()

    # ---- output accessors ------------------------------------------------------------------
    def _scalar(self, name):
        out = C.c_double()
        _lib.check(self._fn("_get_scalar")(self._core(), name.encode(), C.byref(out)))
        return out.value

    def _vec_f(self, name, dtype=None):
        L = _lib.load()
        n = C.c_int64()
        _lib.check(self._fn("_get_vec_f64")(self._core(), name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.float64)
        _lib.check(self._fn("_get_vec_f64")(self._core(), name.encode(), _lib.ptr(buf), n.value, C.byref(n)))
        if name.startswith("benchmark") or name == "sweep_stats":
            return buf
        return buf.astype(self._dtype if dtype is None else dtype)

    def _vec_i(self, name):
        L = _lib.load()
        n = C.c_int64()
        _lib.check(self._fn("_get_vec_i64")(self._core(), name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.int64)
        _lib.check(self._fn("_get_vec_i64")(self._core(), name.encode(), _lib.ptr(buf), n.value, C.byref(n)))
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
            _lib.check(self._fn("_get_screen_transform")(self._core(), i, None, 0, C.byref(n)))
            buf = np.empty(n.value, dtype=np.float64)
            _lib.check(self._fn("_get_screen_transform")(self._core(), i, _lib.ptr(buf), n.value, C.byref(n)))
            gs = int(round(np.sqrt(n.value)))
            out.append(buf.astype(self._dtype).reshape(gs, gs))
        return out

    @property
    def betas(self):
        """(L, p) scipy CSR with int64 indices (py_state.cpp:9-60)."""
        L = _lib.load()
        nnz, nl = C.c_int64(), C.c_int64()
        _lib.check(self._fn("_get_betas")(self._core(), None, None, None, C.byref(nnz), C.byref(nl)))
        indptr = np.empty(nl.value + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int64)
        values = np.empty(nnz.value, dtype=np.float64)
        _lib.check(self._fn("_get_betas")(self._core(), _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(values), C.byref(nnz), C.byref(nl)))
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
        rc = self._fn("_solve")(new._handle, int(progress_bar), C.cast(cb_exit, C.c_void_p) if cb_exit else None, None,
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

    def _check(self, passed, msg, method, logger):
        """adelie/state.py:84-90: every check is logged; method="assert" also asserts."""
        if passed:
            logger.info(msg)
        else:
            logger.error(msg)
            if method == "assert":
                assert False, msg

    def _check_layout(self, method, logger, p):
        """Index-structure invariants shared by every state (adelie/state.py:1434-1548, 186-300): groups / group_sizes / penalty /
        screen_set / screen_begins / screen_beta sizes and dtypes."""
        c = lambda ok, msg: self._check(bool(ok), msg, method, logger)
        groups, gsz, S = self.groups, self.group_sizes, self.screen_set
        G = len(groups)
        c(np.all((0 <= groups) & (groups <= p)), "check groups is in [0, p)")
        c(len(groups) == len(np.unique(groups)), "check groups has unique values")
        c(groups.dtype == np.dtype("int"), "check groups dtype is int")
        c(len(gsz) == G, "check groups and group_sizes have same length")
        c(np.sum(gsz) == p, "check sum of group_sizes is p")
        c(np.all((0 < gsz) & (gsz <= p)), "check group_sizes is in (0, p]")
        c(gsz.dtype == np.dtype("int"), "check group_sizes dtype is int")
        c(np.array_equal(groups, np.cumsum(np.concatenate([[0], gsz]))[:-1]), "check groups and group_sizes consistency")
        c(np.all(self.penalty >= 0), "check penalty is non-negative")
        c(len(self.penalty) == G, "check penalty and groups have same length")
        c(np.all((0 <= S) & (S < G)), "check screen_set is a subset of [0, G)")
        c(len(S) == len(np.unique(S)), "check screen_set has unique values")
        c(S.dtype == np.dtype("int"), "check screen_set dtype is int")
        begins = np.cumsum(np.concatenate([[0], gsz[S]]).astype(int))
        sb = self.screen_begins
        c(np.array_equal(sb, begins[:-1]), "check screen_begins is [0, g1, g2, ...] where gi is the group size of (i-1)th screen group.")
        c(sb.dtype == np.dtype("int"), "check screen_begins dtype is int")
        return begins[-1]

    def check(self, method=None, logger=logger):
        """Checks consistency of the members; every check is logged, ``method="assert"`` also asserts (adelie/state.py:92-116).
        Subclasses re-derive every invariant of their state from X, y and the coefficients."""
        self._check_layout(method, logger, self._p_cols)


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
        if c.get("core_expanded", False):      # [kron(1, I_K) | kron(X, I_K)] was materialised (sparse X): the core sees a single response
            a.n_classes = 1; a.multi_intercept = 0
        else:
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

    def check(self, method=None, logger=logger):
        """Re-derives every invariant of a Gaussian naive state from X, y and ``screen_beta`` (adelie/state.py:1422-1674): layout,
        screen_is_active vs the non-zero blocks, rsq, grad, abs_grad, resid, resid_sum, screen_X_means, and for every screen group that
        V^T (X_g^T W X_g - xbar xbar^T) V is diagonal with diagonal screen_vars.  float32 states are compared at float32 resolution
        (the reference's np.allclose defaults assume float64).  GLM / multi-response states check the layout only."""
        WS = self._check_layout(method, logger, self._p_cols)
        if self._use_glm or self._is_multi or _dist_active():
            return
        c = lambda ok, msg: self._check(bool(ok), msg, method, logger)
        f32 = np.dtype(self._dtype) == np.float32
        close = (lambda a, b: np.allclose(a, b, rtol=2e-4, atol=2e-5)) if f32 else np.allclose
        X, w = self._X, np.asarray(self.weights, dtype=self._dtype)
        n, p = X.rows(), X.cols()
        groups, gsz, S, sb = self.groups, self.group_sizes, self.screen_set, self.screen_begins
        beta = self.screen_beta
        c(np.all(w >= 0), "check weights is non-negative")
        c(np.allclose(np.sum(w, dtype=np.float64), 1), "check weights sum to 1")
        c(len(beta) == WS, "check screen_beta size")
        nnz = np.array([i for i in range(len(S)) if np.any(beta[sb[i]:sb[i] + gsz[S[i]]] != 0)], dtype=int)
        c(np.all(self.screen_is_active[nnz]), "check screen_is_active is only active on non-zeros of screen_beta")
        yc = np.asarray(self._glm.y, dtype=np.float64) - np.asarray(self._offsets, dtype=np.float64)
        w64 = w.astype(np.float64)
        if self.intercept:
            yc = yc - np.sum(yc * w64)
        Xbeta = np.zeros(n, dtype=self._dtype)
        cols = []
        for i in range(len(S)):
            g, gs = int(groups[S[i]]), int(gsz[S[i]])
            cols.append(np.arange(g, g + gs))
            X.btmul(g, gs, np.ascontiguousarray(beta[sb[i]:sb[i] + gs]), Xbeta)
        cols = np.concatenate(cols).astype(int) if cols else np.zeros(0, dtype=int)
        resid = yc - Xbeta
        grad = np.empty(p, dtype=self._dtype)
        X.mul(np.ascontiguousarray(resid, dtype=self._dtype), w, grad)
        grad = grad.astype(np.float64)
        X_means = self.X_means.astype(np.float64)
        if self.intercept:
            grad -= X_means * np.sum(w64 * resid)
        sXm = self.screen_X_means.astype(np.float64)
        WXcbeta = w64 * (Xbeta - (sXm @ beta if self.intercept else 0.0))      # the reference subtracts the means unconditionally (:1573)
        pos = w64 > 0
        expected = 2 * np.sum(yc * WXcbeta) - np.sum(WXcbeta[pos] ** 2 / w64[pos])
        c(close(self.rsq, expected), "check rsq")
        c(close(self.grad, grad), "check grad")
        lmda = 1e35 if np.isinf(self.lmda) else float(self.lmda)
        gc_ = grad.copy()
        for i in range(len(S)):
            g, gs = int(groups[S[i]]), int(gsz[S[i]])
            gc_[g:g + gs] -= lmda * (1 - self.alpha) * float(self.penalty[S[i]]) * beta[sb[i]:sb[i] + gs]
        abs_grad = np.sqrt(np.add.reduceat(gc_ ** 2, groups)) if len(groups) else np.zeros(0)
        c((self.lmda_max == -1) or close(self.abs_grad, abs_grad), "check abs_grad")
        c(close(self.resid, resid), "check resid")
        c(close(self.resid_sum, np.sum(w64 * resid)) or abs(self.resid_sum - np.sum(w64 * resid)) < (1e-5 if f32 else 1e-10), "check resid_sum")
        c(close(sXm, X_means[cols]), "check screen_X_means")
        sqrt_w = np.sqrt(w)
        sv, st = self.screen_vars, self.screen_transforms
        c(len(sv) == WS and np.all(sv >= 0), "check screen_vars size and sign")
        c(len(st) == len(S), "check screen_transforms size")
        for i in range(len(S)):
            g, gs = int(groups[S[i]]), int(gsz[S[i]])
            C_ = np.empty((gs, gs), dtype=self._dtype, order="F")
            X.cov(g, gs, sqrt_w, C_)
            C_ = C_.astype(np.float64)
            if self.intercept:
                C_ -= np.outer(X_means[g:g + gs], X_means[g:g + gs])
            V = st[i].astype(np.float64)
            D = V.T @ C_ @ V
            scale = max(1.0, float(np.max(np.abs(np.diag(C_))))) if gs else 1.0
            tol_d = (2e-4 if f32 else 1e-8) * scale
            c(np.allclose(np.maximum(np.diag(D), 0), sv[sb[i]:sb[i] + gs], rtol=2e-4 if f32 else 1e-5, atol=tol_d), f"check screen_vars[{sb[i]}:{sb[i]}+{gs}]")
            np.fill_diagonal(D, 0)
            c(np.all(np.abs(D) <= tol_d), "check VT Xi V is nearly 0 after zeroing the diagonal")

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


def _expand_sparse_multi(X, K, intercept, n_threads):
    """Sparse X in a multi-response problem: [kron(1, I_K) | kron(X, I_K)] (adelie/state.py:1100-1125) is itself a sparse matrix with K
    times the non-zeros.  The reference wraps X in generic kronecker_eye / concatenate views (matrix_naive_kronecker_eye.ipp:29-352); the
    sparse device kernels are single-response, so the expanded CSC matrix is built once and the core state runs with n_classes = 1 on it
    (same problem, same iterates: the multi-response layout is exactly this reformulation)."""
    import scipy.sparse as sp
    M = X._mat
    n = M.shape[0]
    blocks = []
    if intercept:
        blocks.append(sp.kron(np.ones((n, 1), dtype=M.dtype), sp.identity(K, dtype=M.dtype, format="csc"), format="csc"))
    blocks.append(sp.kron(M, sp.identity(K, dtype=M.dtype, format="csc"), format="csc"))
    Xs = sp.hstack(blocks, format="csc") if len(blocks) > 1 else blocks[0]
    Xs = sp.csc_matrix(Xs, dtype=M.dtype)
    Xs.sum_duplicates(); Xs.sort_indices()
    return _matrix.sparse(Xs, n_threads=n_threads)


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
    expanded = isinstance(X, _matrix._Sparse)
    if expanded:
        arrays["X_core"] = _expand_sparse_multi(X, K, intercept, n_threads)
    # not the actual y_mean: a value that yields the right loss_null / loss_full (state.py:2339-2343)
    y_mean = np.linalg.norm(np.asarray(_dist_sum(np.sum(glm_obj.weights[:, None] * (y - offsets), axis=0))) / K)
    cfg = dict(alpha=float(alpha), y_mean=float(y_mean), y_var=float(y_var), resid_sum=float(resid_sum), rsq=float(rsq),
               lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size), setup_lmda_max=setup_lmda_max,
               setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size, max_active_size=max_active_size,
               pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
               screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, intercept=False,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float(lmda), n_classes=int(K),
               multi_intercept=bool(intercept), core_expanded=expanded)
    return _MultiNaive(X=X, glm_obj=glm_obj, use_glm=False, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=Xe.cols())


def _dist_sum(v):
    from . import dist as _dist
    return _dist.allreduce(v)


def _dist_active():
    from . import dist as _dist
    return _dist.is_active()


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
    expanded = isinstance(X, _matrix._Sparse)
    if expanded:
        arrays["X_core"] = _expand_sparse_multi(X, K, intercept, n_threads)
        if loss_null is None:
            # the null model of a multi-response GLM fits the K intercepts (update_loss_null of solver_multiglm_naive.hpp:99-186); the
            # single-response core on the expanded matrix would take the loss at the offsets instead, so it is computed here, by the
            # multi-response core on a one-column zero matrix (the null model does not involve X)
            from .solver import grpnet as _grpnet
            null = _grpnet(np.zeros((glm.y.shape[0], 1), dtype=dtype, order="F"), glm, offsets=np.asarray(offsets).reshape(glm.y.shape),
                           lmda_path_size=1, intercept=intercept, irls_max_iters=irls_max_iters, irls_tol=irls_tol, early_exit=False,
                           progress_bar=False)
            if null.error != "":
                raise RuntimeError(null.error)
            loss_null = float(null.loss_null)
            null.close()
    cfg = dict(alpha=float(alpha), beta0=0.0, loss_null=None if loss_null is None else float(loss_null),
               loss_full=float(loss_full), irls_max_iters=int(irls_max_iters), irls_tol=irls_tol,
               lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size), setup_lmda_max=setup_lmda_max,
               setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size, max_active_size=max_active_size,
               pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
               screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, intercept=False,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float(lmda), n_classes=int(K),
               multi_intercept=bool(intercept), core_expanded=expanded)
    return _MultiNaive(X=X, glm_obj=glm, use_glm=True, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=Xe.cols())


class _Pin(_Naive):
    """Gaussian pin state (adelie/state.py:179-420 gaussian_pin_base / gaussian_pin_naive_base + :421-720; core StateGaussianPinNaive,
    adelie/src/py_state.cpp:389-411): solves its own ``lmda_path`` on a FIXED screen set (pin::naive::solve)."""

    def solve(self, progress_bar: bool = False, exit_cond=None):
        new = self._clone()
        L = _lib.load()
        err = C.create_string_buffer(4096)
        total = C.c_double()
        pending = []
        def _poll():
            try:
                C.pythonapi.PyErr_CheckSignals()
            except BaseException as e:          # noqa: BLE001
                pending.append(e)
                return 1
            return 0
        cb_sig = _lib.CHECK_SIGNALS_T(_poll)
        rc = L.ab_pin_naive_solve(new._handle, C.cast(cb_sig, C.c_void_p), err, len(err), C.byref(total))
        if pending:
            raise pending[0]
        _lib.check(rc)
        new.error = err.value.decode()
        new.total_time = total.value
        if new.error != "":
            (logger.error if new.error.startswith("adelie_core solver: ") else logger.warning)(RuntimeError(new.error))
        return new

    @property
    def rsqs(self):
        return self._vec_f("rsqs")

    @property
    def iters(self):
        return self.n_sweeps

    @property
    def benchmark_screen(self):
        return self._vec_f("benchmark_fit_screen")

    @property
    def benchmark_active(self):
        return self._vec_f("benchmark_fit_active")

    @property
    def active_order(self):
        a = self.active_set[: self.active_set_size]
        return np.argsort(self.groups[self.screen_set[a]], kind="stable").astype(int)

    @property
    def active_begins(self):
        a = self.active_set[: self.active_set_size]
        return np.cumsum(np.concatenate([[0], self.group_sizes[self.screen_set[a]]]).astype(int))[:-1]

    def check(self, method=None, logger=logger):
        """adelie/state.py:180-420: layout, lmda_path, screen_is_active vs active_set, active_begins / active_order, output shapes."""
        WS = self._check_layout(method, logger, self._p_cols)
        c = lambda ok, msg: self._check(bool(ok), msg, method, logger)
        S = len(self.screen_set)
        sv = self.screen_vars
        c(len(sv) == WS, "check screen_vars size")
        c(np.all(sv >= 0), "check screen_vars is non-negative")
        c(len(self.screen_transforms) == S, "check screen_transforms size")
        c(np.all(self.lmda_path >= 0), "check lmda_path is non-negative")
        a = self.active_set[: self.active_set_size]
        c(np.array_equal(np.arange(S)[self.screen_is_active], np.sort(a)), "check screen_is_active is consistent with active_set")
        c(self.screen_is_active.dtype == np.dtype("bool"), "check screen_is_active dtype is bool")
        c(np.all((0 <= a) & (a < S)), "check active_set is in [0, S)")
        c(len(a) == len(np.unique(a)), "check active_set is unique")
        c(a.dtype == np.dtype("int"), "check active_set dtype is int")
        order = self.groups[self.screen_set[a[self.active_order]]]
        c(np.array_equal(order, np.sort(order)), "check active_order orders active_set such that groups is ordered")
        B = self.betas
        c(B.shape[0] <= self.lmda_path.shape[0], "check betas rows is no more than the number of lmda_path")
        c(isinstance(B, scipy.sparse.csr_matrix), "check betas type")
        c(B.shape[1] == self._p_cols, "check betas shape")
        c(self.rsqs.shape == (B.shape[0],), "check rsqs shape")
        c(np.all(self.rsqs >= 0), "check rsqs is non-negative")
        c(self.lmdas.shape == (B.shape[0],), "check lmdas shape")
        c(self.resid.shape[0] == self._X.rows(), "check resid shape")


def gaussian_pin_naive(*, X, y_mean, y_var, constraints, groups, alpha, penalty, weights, screen_set, lmda_path, rsq, resid, screen_beta,
                       screen_is_active, active_set_size, active_set, intercept=True, max_active_size=None, max_iters=int(1e5), tol=1e-7,
                       adev_tol=0.9, ddev_tol=0, newton_tol=1e-12, newton_max_iters=1000, n_threads=1):
    """Gaussian pin naive-method state (adelie/state.py:421-720; core StateGaussianPinNaive{32,64}).  As in the reference wrapper the
    column means come from one ``X.mul`` pass; the screen groups' Grams and eigendecompositions are computed on the device when the
    core state is built; ``tol`` is used unscaled (the path driver passes ``tol * y_var``, solver_gaussian_naive.hpp:314)."""
    _check_constraints(constraints)
    if not isinstance(X, _matrix.MatrixNaiveBase):
        raise ValueError("X must be an instance of MatrixNaiveBase32 or MatrixNaiveBase64.")
    dtype = X.dtype
    n, p = X.rows(), X.cols()
    groups = np.asarray(groups)
    G = groups.shape[0]
    group_sizes = np.concatenate([groups, [p]], dtype=int)
    group_sizes = group_sizes[1:] - group_sizes[:-1]
    weights = np.array(weights, copy=True, dtype=dtype)
    glm_obj = _glm.gaussian(y=np.zeros(n, dtype=dtype), weights=weights, dtype=dtype)     # carrier of the weights only (the pin state has no y)
    X_means = np.empty(p, dtype=dtype)
    X.mul(np.ones(n, dtype=dtype), glm_obj.weights, X_means)
    X_means = np.asarray(_dist_sum(X_means), dtype=dtype)
    resid = np.array(resid, copy=True, dtype=dtype)
    arrays = _common_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=np.array(lmda_path, copy=True, dtype=dtype),
                            screen_set=screen_set, screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set,
                            grad=np.zeros(p, dtype=dtype), resid=resid, dtype=dtype)
    arrays["weights"] = glm_obj.weights
    arrays["X_means"] = X_means
    arrays["offsets"] = np.zeros(n, dtype=dtype)
    max_active_size = G if max_active_size is None else int(np.minimum(max_active_size, G))
    cfg = dict(alpha=float(alpha), y_mean=float(y_mean), y_var=float(y_var), resid_sum=float(_dist_sum(np.sum(glm_obj.weights * resid))),
               rsq=float(rsq), lmda_max=-1.0, min_ratio=1e-2, lmda_path_size=len(arrays["lmda_path"]), setup_lmda_max=False,
               setup_lmda_path=False, max_screen_size=G, max_active_size=max_active_size, pivot_subset_ratio=0.1, pivot_subset_min=1,
               pivot_slack_ratio=1.25, screen_rule="pivot", max_iters=int(max_iters), tol=tol, adev_tol=adev_tol, ddev_tol=ddev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=False, intercept=intercept,
               n_threads=int(n_threads), active_set_size=int(active_set_size), lmda=float("inf"))
    return _Pin(X=X, glm_obj=glm_obj, use_glm=False, dtype=dtype, cfg=cfg, arrays=arrays, p_cols=p)


# ------------------------------------------------------------------------------------------------------------------
# covariance method: StateGaussianCov / StateGaussianPinCov
# ------------------------------------------------------------------------------------------------------------------
_COV_VEC_F = ["sweep_stats", "rsqs", "lmda_path", "screen_beta", "screen_grad", "screen_vars", "grad", "abs_grad", "v", "devs", "lmdas",
              "benchmark_screen", "benchmark_fit_screen", "benchmark_fit_active", "benchmark_kkt", "benchmark_invariance"]
_COV_VEC_I = ["screen_set", "screen_begins", "screen_is_active", "active_set", "screen_subset_order", "screen_subset_ordered",
              "n_valid_solutions", "active_sizes", "screen_sizes"]
_COV_INT = ("active_set_size", "n_sweeps", "n_group_updates", "n_col_updates", "n_pin_solves", "n_kernel_launches", "cov_cluster",
            "cov_smem_bytes")
_COV_SCALARS = ["lmda_max", "lmda", "rsq", "time_sweep_kernel"] + list(_COV_INT)


class _Cov(base):
    """Gaussian covariance-method state (adelie/state.py:1128-1420 gaussian_cov; core StateGaussianCov{32,64}, solve =
    gaussian::cov::solve, CORE/solver/solver_gaussian_cov.hpp:359-457)."""
    _api = "ab_cov_state"

    def __init__(self, *, A, dtype, cfg, arrays):
        self._A = A
        self._dtype = np.dtype(dtype).type
        self._cfg = cfg
        self._p_cols = A.cols()
        for k, v in arrays.items():
            setattr(self, "_" + k, v)
        self._handle = None
        self._handle = self._build_core()

    def _build_core(self):
        c = self._cfg
        a = _lib.CovStateArgs()
        a.dtype = _lib.dtype_code(self._dtype)
        a.v = _lib.ptr(self._v)
        a.groups = _lib.ptr(self._groups); a.group_sizes = _lib.ptr(self._group_sizes); a.G = self._groups.shape[0]
        a.alpha = c["alpha"]; a.penalty = _lib.ptr(self._penalty)
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
        a.max_iters = c["max_iters"]; a.tol = c["tol"]; a.rdev_tol = c["rdev_tol"]
        a.newton_tol = c["newton_tol"]; a.newton_max_iters = c["newton_max_iters"]
        a.early_exit = int(c["early_exit"]); a.n_threads = c["n_threads"]
        a.screen_set = _lib.ptr(self._screen_set); a.screen_set_size = self._screen_set.shape[0]
        a.screen_beta = _lib.ptr(self._screen_beta); a.screen_beta_size = self._screen_beta.shape[0]
        a.screen_is_active = _lib.ptr(self._screen_is_active_i8); a.active_set_size = c["active_set_size"]
        a.active_set = _lib.ptr(self._active_set)
        a.rsq = c["rsq"]; a.lmda = c["lmda"]; a.grad = _lib.ptr(self._grad)
        a.screen_grad = _lib.ptr(self.__dict__.get("_screen_grad_in"))
        h = C.c_void_p()
        _lib.check(_lib.load().ab_cov_state_create(C.byref(a), self._A._core(), C.byref(h)))
        return h

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        cfg = self.__dict__.get("_cfg", {})
        if name in ("min_ratio", "lmda_path_size", "max_screen_size", "max_active_size", "pivot_subset_ratio", "pivot_subset_min",
                    "pivot_slack_ratio", "screen_rule", "max_iters", "tol", "rdev_tol", "newton_tol", "newton_max_iters", "early_exit",
                    "setup_lmda_max", "setup_lmda_path", "n_threads") and name in cfg:
            return cfg[name]
        if name in _COV_SCALARS:
            v = self._scalar(name)
            return int(v) if name in _COV_INT else (v if name == "time_sweep_kernel" else self._dtype(v))
        if name in _COV_VEC_F:
            return self._vec_f(name)
        if name in _COV_VEC_I:
            v = self._vec_i(name)
            if name == "screen_is_active":
                return v.astype(bool)
            if name in ("n_valid_solutions", "active_sizes", "screen_sizes"):
                return v.astype(np.int32)
            return v
        raise AttributeError(name)

    @property
    def A(self):
        return self._A

    @property
    def X(self):
        raise AttributeError("covariance-method states have no X")

    @property
    def v(self):
        return self._v

    @property
    def constraints(self):
        return None

    @property
    def dual_groups(self):
        return np.zeros(len(self._groups), dtype=int)

    def check(self, method=None, logger=logger):
        """The reference's gaussian_cov state inherits the no-op base check (adelie/state.py:92-116); here the index layout plus the
        covariance-method invariants grad = v - A beta and screen_grad = grad on the screen values are re-derived."""
        p = self._p_cols
        vs = self._check_layout(method, logger, p)
        c = lambda ok, msg: self._check(bool(ok), msg, method, logger)
        c(isinstance(self._A, (_matrix.MatrixCovBase32, _matrix.MatrixCovBase64)), "check A type")
        sb = self.screen_beta
        c(len(sb) == vs, "check screen_beta size")
        S, gsz, groups = self.screen_set, self.group_sizes, self.groups
        subset = np.concatenate([np.arange(groups[i], groups[i] + gsz[i]) for i in S]) if len(S) else np.zeros(0, dtype=int)
        order = np.argsort(subset)
        out = np.empty(p, dtype=self._dtype)
        self._A.mul(subset[order], np.ascontiguousarray(sb[order], dtype=self._dtype), out)
        grad = self._v - out
        tol = 1e-10 if self._dtype == np.float64 else 1e-3
        c(np.allclose(self.grad, grad, rtol=tol, atol=tol * max(1.0, float(np.max(np.abs(self._v), initial=0)))), "check grad = v - A beta")


class _PinCov(_Cov):
    """Gaussian pin covariance-method state (adelie/state.py:723-1000; core StateGaussianPinCov{32,64}, solve = gaussian::pin::cov::solve,
    CORE/solver/solver_gaussian_pin_cov.hpp:529-725): solves its own ``lmda_path`` on a FIXED screen set."""

    def solve(self, progress_bar: bool = False, exit_cond=None):
        new = self._clone()
        err = C.create_string_buffer(4096)
        total = C.c_double()
        pending = []
        def _poll():
            try:
                C.pythonapi.PyErr_CheckSignals()
            except BaseException as e:          # noqa: BLE001
                pending.append(e)
                return 1
            return 0
        cb_sig = _lib.CHECK_SIGNALS_T(_poll)
        rc = _lib.load().ab_cov_pin_solve(new._handle, C.cast(cb_sig, C.c_void_p), err, len(err), C.byref(total))
        if pending:
            raise pending[0]
        _lib.check(rc)
        new.error = err.value.decode()
        new.total_time = total.value
        if new.error != "":
            (logger.error if new.error.startswith("adelie_core solver: ") else logger.warning)(RuntimeError(new.error))
        return new

    @property
    def iters(self):
        return self.n_sweeps

    @property
    def benchmark_screen(self):
        return self._vec_f("benchmark_fit_screen")

    @property
    def benchmark_active(self):
        return self._vec_f("benchmark_fit_active")

    def check(self, method=None, logger=logger):
        vs = self._check_layout(method, logger, self._p_cols)
        c = lambda ok, msg: self._check(bool(ok), msg, method, logger)
        c(isinstance(self._A, (_matrix.MatrixCovBase32, _matrix.MatrixCovBase64)), "check A type")
        c(len(self.screen_beta) == vs and len(self.screen_grad) == vs, "check screen_beta / screen_grad size")


def _cov_arrays(*, groups, group_sizes, penalty, lmda_path, screen_set, screen_beta, screen_is_active, active_set, grad, v, dtype):
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
        v=np.array(v, copy=True, dtype=dtype),
    )


def gaussian_cov(*, A, v, constraints, groups, group_sizes, alpha, penalty, screen_set, screen_beta, screen_is_active, active_set_size,
                 active_set, rsq, lmda, grad, lmda_path=None, lmda_max=None, max_iters=int(1e5), tol=1e-7, rdev_tol=1e-4, newton_tol=1e-12,
                 newton_max_iters=1000, n_threads=1, early_exit=True, screen_rule="pivot", min_ratio=1e-2, lmda_path_size=100,
                 max_screen_size=None, max_active_size=None, pivot_subset_ratio=0.1, pivot_subset_min=1, pivot_slack_ratio=1.25):
    """Gaussian covariance-method state (adelie/state.py:1128-1420; core StateGaussianCov{32,64})."""
    _check_constraints(constraints)
    if isinstance(A, np.ndarray):
        A = _matrix.dense(A, method="cov", n_threads=n_threads)
    if not isinstance(A, (_matrix.MatrixCovBase32, _matrix.MatrixCovBase64)):
        raise ValueError("A must be an instance of MatrixCovBase32, MatrixCovBase64, or np.ndarray.")
    dtype = A.dtype
    groups = np.asarray(groups)
    if np.asarray(v).shape != (A.cols(),):
        raise RuntimeError("adelie_core: v must be (p,) where A is (p, p).")          # state_gaussian_cov.ipp:11-13
    if np.asarray(grad).shape != (A.cols(),):
        raise RuntimeError("adelie_core: grad must be (p,) where A is (p, p).")
    (max_screen_size, max_active_size, lmda_path_size, setup_lmda_max, setup_lmda_path, lmda_max, lmda_path) = _render_inputs(
        groups=groups, lmda_max=lmda_max, lmda_path=lmda_path, lmda_path_size=lmda_path_size,
        max_screen_size=max_screen_size, max_active_size=max_active_size, dtype=dtype)
    arrays = _cov_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=lmda_path, screen_set=screen_set,
                         screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set, grad=grad, v=v, dtype=dtype)
    cfg = dict(alpha=float(alpha), rsq=float(rsq), lmda_max=float(lmda_max), min_ratio=min_ratio, lmda_path_size=int(lmda_path_size),
               setup_lmda_max=setup_lmda_max, setup_lmda_path=setup_lmda_path, max_screen_size=max_screen_size,
               max_active_size=max_active_size, pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min,
               pivot_slack_ratio=pivot_slack_ratio, screen_rule=screen_rule, max_iters=int(max_iters), tol=tol, rdev_tol=rdev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=early_exit, n_threads=int(n_threads),
               active_set_size=int(active_set_size), lmda=float(lmda))
    return _Cov(A=A, dtype=dtype, cfg=cfg, arrays=arrays)


def gaussian_pin_cov(*, A, constraints, groups, alpha, penalty, screen_set, lmda_path, rsq, screen_beta, screen_grad, screen_is_active,
                     active_set_size, active_set, max_active_size=None, max_iters=int(1e5), tol=1e-7, rdev_tol=1e-4, newton_tol=1e-12,
                     newton_max_iters=1000, n_threads=1):
    """Gaussian pin covariance-method state (adelie/state.py:739-1000; core StateGaussianPinCov{32,64}).  ``screen_vars``,
    ``screen_transforms`` and ``screen_subset_order`` are derived from the diagonal blocks of ``A`` when the core state is built
    (the reference wrapper does the same with ``A.to_dense`` + ``numpy.linalg.eigh``, :912-936)."""
    _check_constraints(constraints)
    if not isinstance(A, (_matrix.MatrixCovBase32, _matrix.MatrixCovBase64)):
        raise ValueError("A must be an instance of MatrixCovBase32 or MatrixCovBase64.")
    dtype = A.dtype
    p = A.cols()
    groups = np.asarray(groups)
    G = groups.shape[0]
    group_sizes = np.concatenate([groups, [p]], dtype=int)
    group_sizes = group_sizes[1:] - group_sizes[:-1]
    arrays = _cov_arrays(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda_path=np.array(lmda_path, copy=True, dtype=dtype),
                         screen_set=screen_set, screen_beta=screen_beta, screen_is_active=screen_is_active, active_set=active_set,
                         grad=np.zeros(p, dtype=dtype), v=np.zeros(p, dtype=dtype), dtype=dtype)
    arrays["screen_grad_in"] = np.array(screen_grad, copy=True, dtype=dtype)
    max_active_size = G if max_active_size is None else int(np.minimum(max_active_size, G))
    cfg = dict(alpha=float(alpha), rsq=float(rsq), lmda_max=-1.0, min_ratio=1e-2, lmda_path_size=len(arrays["lmda_path"]),
               setup_lmda_max=False, setup_lmda_path=False, max_screen_size=G, max_active_size=max_active_size, pivot_subset_ratio=0.1,
               pivot_subset_min=1, pivot_slack_ratio=1.25, screen_rule="pivot", max_iters=int(max_iters), tol=tol, rdev_tol=rdev_tol,
               newton_tol=newton_tol, newton_max_iters=int(newton_max_iters), early_exit=False, n_threads=int(n_threads),
               active_set_size=int(active_set_size), lmda=float("inf"))
    return _PinCov(A=A, dtype=dtype, cfg=cfg, arrays=arrays)
