"""GLM family objects (reference: adelie/glm.py:33-538, adelie/src/py_glm.cpp:101-234).

Each object keeps host copies of ``y`` / ``weights`` (the attribute surface the callers read)
and a device-resident core object; ``gradient`` / ``hessian`` / ``inv_hessian_gradient`` /
``loss`` / ``loss_full`` / ``inv_link`` run coalesced CUDA kernels (csrc/glm.cuh).
"""
from __future__ import annotations

import ctypes as C
from typing import Union

import numpy as np

from . import _lib
from . import dist as _dist

_FAMILY = {"gaussian": 1, "binomial_logit": 2, "multigaussian": 3, "cox": 4, "poisson": 5, "binomial_probit": 6, "multinomial": 7}


def _coerce_dtype(y, dtype):
    """adelie/glm.py:13-30"""
    if dtype is None:
        valid = [np.float32, np.float64]
        ok = [y.dtype == d for d in valid]
        if not any(ok):
            raise RuntimeError("y must be of type numpy.float32 or numpy.float64 if dtype is None.")
        dtype = valid[int(np.argmax(ok))]
    else:
        y = np.asarray(y, dtype=dtype)
    return y, np.dtype(dtype).type


class GlmBase:
    """Common machinery of the single-response families (GlmBase32/64)."""
    is_multi = False

    def _init_common(self, name, y, weights, dtype, K=1, **extra):
        self.name = name
        self.dtype = dtype
        self.y = np.array(y, copy=True, dtype=dtype)
        n = self.y.shape[0]
        if weights is not None:
            weights = np.asarray(weights)
            if weights.shape != (n,):
                raise RuntimeError("y and weights must have same length." if K == 1 else "y rows and weights must have same length.")
            ws = _dist.allreduce(float(np.sum(weights)))      # row-sharded: the weights of ALL ranks sum to one
            if not np.allclose(ws, 1):
                weights = weights / ws
        else:
            weights = np.full(n, 1 / _dist.allreduce(float(n)), dtype=dtype)
        self.weights = np.array(weights, copy=True, dtype=dtype)
        self._K = K
        self._n = n
        self._extra = extra
        self._handle = None

    # the device object is created lazily so that constructing a GLM does not need a GPU
    def _core(self):
        if self._handle is None:
            L = _lib.load()
            h = C.c_void_p()
            e = self._extra
            _lib.check(L.ab_glm_create(
                _lib.dtype_code(self.dtype), _FAMILY[self.name], self._n, self._K, _lib.ptr(self.y), _lib.ptr(self.weights),
                _lib.ptr(e.get("start")), _lib.ptr(e.get("stop")), _lib.ptr(e.get("strata")), int(e.get("efron", 1)), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().ab_glm_free(self._handle)
        except Exception:
            pass

    def _chk(self, what, **arrs):
        shape = self.y.shape
        for k, a in arrs.items():
            if a.shape != shape:
                sizes = ", ".join(f"{kk}={aa.size}" for kk, aa in arrs.items())
                raise RuntimeError(f"adelie_core: {what}() is given inconsistent inputs! (weights={self.weights.size}, y={self.y.size}, {sizes})")
            if a.dtype != self.dtype or not a.flags.c_contiguous:
                raise TypeError(f"{what}(): arguments must be C-contiguous arrays of dtype {np.dtype(self.dtype).name}")

    def gradient(self, eta, grad):
        self._chk("gradient", eta=eta, grad=grad)
        _lib.check(_lib.load().ab_glm_gradient(self._core(), _lib.ptr(eta), _lib.ptr(grad)))

    def hessian(self, eta, grad, hess):
        self._chk("hessian", eta=eta, grad=grad, hess=hess)
        _lib.check(_lib.load().ab_glm_hessian(self._core(), _lib.ptr(eta), _lib.ptr(grad), _lib.ptr(hess)))

    def inv_hessian_gradient(self, eta, grad, hess, inv_hess_grad):
        self._chk("inv_hessian_grad", eta=eta, grad=grad, hess=hess, inv_hess_grad=inv_hess_grad)
        _lib.check(_lib.load().ab_glm_inv_hessian_gradient(self._core(), _lib.ptr(eta), _lib.ptr(grad), _lib.ptr(hess), _lib.ptr(inv_hess_grad)))

    def loss(self, eta):
        self._chk("loss", eta=eta)
        out = C.c_double()
        _lib.check(_lib.load().ab_glm_loss(self._core(), _lib.ptr(eta), C.byref(out)))
        return self.dtype(out.value)

    def loss_full(self):
        out = C.c_double()
        _lib.check(_lib.load().ab_glm_loss_full(self._core(), C.byref(out)))
        return self.dtype(out.value)

    def inv_link(self, eta, out):
        self._chk("inv_link", eta=eta, out=out)
        _lib.check(_lib.load().ab_glm_inv_link(self._core(), _lib.ptr(eta), _lib.ptr(out)))


class GlmMultiBase(GlmBase):
    is_multi = True



class _UserGlm(GlmBase):
    """A GLM defined in Python (reference: subclass ``adelie.glm.GlmBase64`` / ``GlmMultiBase64`` and override the virtuals; the pybind
    trampolines PyGlmBase / PyGlmMultiBase, adelie/src/py_glm.cpp:8-92, 240-330).  The subclass implements ``gradient(eta, grad)``,
    ``hessian(eta, grad, hess)``, ``loss(eta)``, ``loss_full()`` and optionally ``inv_hessian_gradient`` / ``inv_link`` on NumPy arrays; the
    device solver calls them once per IRLS iteration through ``ab_glm_create_callback`` (eta / grad / hess travel through pinned host
    memory), the coordinate descent itself stays on the device."""
    _user_dtype = None
    opt = False

    def __init__(self, name, y, weights):
        y = np.asarray(y)
        dtype = self._user_dtype
        K = y.shape[1] if self.is_multi else 1
        self._init_common(str(name), np.asarray(y, dtype=dtype), weights, dtype, K=K)
        self._cb_keep = None

    # the reference's pure virtuals
    def gradient(self, eta, grad):
        raise NotImplementedError("gradient() must be implemented by the user-defined GLM.")

    def hessian(self, eta, grad, hess):
        raise NotImplementedError("hessian() must be implemented by the user-defined GLM.")

    def loss(self, eta):
        raise NotImplementedError("loss() must be implemented by the user-defined GLM.")

    def loss_full(self):
        raise NotImplementedError("loss_full() must be implemented by the user-defined GLM.")

    def inv_hessian_gradient(self, eta, grad, hess, inv_hess_grad):
        """adelie_core/glm/glm_base.ipp:25-36"""
        from . import configs as _configs
        hmin = _configs._DEFAULTS["hessian_min"]
        inv_hess_grad[...] = grad / (np.maximum(hess, 0) + hmin * (hess <= 0))

    def inv_link(self, eta, out):
        raise NotImplementedError("inv_link() is not implemented by the user-defined GLM.")

    def _core(self):
        if self._handle is None:
            shape = self.y.shape
            n_el = int(np.prod(shape))
            dt = np.dtype(self.dtype)
            cty = C.c_float if dt == np.float32 else C.c_double
            self._cb_errors = []

            def view(p, writable=True):
                a = np.ctypeslib.as_array(C.cast(p, C.POINTER(cty)), shape=(n_el,)).reshape(shape)
                return a

            def guard(fn):
                def wrapped(*args):
                    try:
                        fn(*args)
                        return 0
                    except BaseException as e:      # noqa: BLE001 -- reported through the solver's error string, re-raised by solve()
                        self._cb_errors.append(e)
                        return 1
                return wrapped

            def _loss(ctx, eta, out):
                out[0] = float(self.loss(view(eta)))

            def _loss_full(ctx, out):
                out[0] = float(self.loss_full())
            own_ihg = type(self).inv_hessian_gradient is not _UserGlm.inv_hessian_gradient
            own_inv_link = type(self).inv_link is not _UserGlm.inv_link
            cb = _lib.GlmCallbacks(
                None,
                _lib.GLM_CB2(guard(lambda ctx, eta, grad: self.gradient(view(eta), view(grad)))),
                _lib.GLM_CB3(guard(lambda ctx, eta, grad, hess: self.hessian(view(eta), view(grad), view(hess)))),
                _lib.GLM_CB4(guard(lambda ctx, eta, grad, hess, out: self.inv_hessian_gradient(view(eta), view(grad), view(hess), view(out))))
                if own_ihg else _lib.GLM_CB4(),
                _lib.GLM_CBL(guard(_loss)),
                _lib.GLM_CBF(guard(_loss_full)),
                _lib.GLM_CB2(guard(lambda ctx, eta, out: self.inv_link(view(eta), view(out)))) if own_inv_link else _lib.GLM_CB2(),
            )
            self._cb_keep = cb                       # the function pointers must outlive the core object
            h = C.c_void_p()
            _lib.check(_lib.load().ab_glm_create_callback(_lib.dtype_code(self.dtype), shape[0], self._K, int(self.is_multi), C.byref(cb), C.byref(h)))
            self._handle = h
        return self._handle


class GlmBase64(_UserGlm):
    _user_dtype = np.float64


class GlmBase32(_UserGlm):
    _user_dtype = np.float32


class GlmMultiBase64(_UserGlm):
    _user_dtype = np.float64
    is_multi = True


class GlmMultiBase32(_UserGlm):
    _user_dtype = np.float32
    is_multi = True


class _Gaussian(GlmBase):
    def __init__(self, y, weights, dtype, opt):
        if y.ndim != 1:
            raise RuntimeError("y must be 1-dimensional.")
        self.opt = opt
        self._init_common("gaussian", y, weights, dtype)

    def reweight(self, weights=None):
        return gaussian(y=self.y, weights=self.weights if weights is None else weights, dtype=self.dtype, opt=self.opt)


class _Binomial(GlmBase):
    def __init__(self, y, weights, dtype, link="logit"):
        if y.ndim != 1:
            raise RuntimeError("y must be 1-dimensional.")
        self.link = link
        self._init_common("binomial_" + link, y, weights, dtype)

    def reweight(self, weights=None):
        return binomial(y=self.y, weights=self.weights if weights is None else weights, link=self.link, dtype=self.dtype)


class _Poisson(GlmBase):
    def __init__(self, y, weights, dtype):
        if y.ndim != 1:
            raise RuntimeError("y must be 1-dimensional.")
        self._init_common("poisson", y, weights, dtype)

    def reweight(self, weights=None):
        return poisson(y=self.y, weights=self.weights if weights is None else weights, dtype=self.dtype)


class _MultiGaussian(GlmMultiBase):
    def __init__(self, y, weights, dtype, opt):
        if y.ndim != 2:
            raise RuntimeError("y must be 2-dimensional.")
        self.opt = opt
        y = np.ascontiguousarray(y)
        self._init_common("multigaussian", y, weights, dtype, K=y.shape[1])

    def reweight(self, weights=None):
        return multigaussian(y=self.y, weights=self.weights if weights is None else weights, dtype=self.dtype, opt=self.opt)


class _Multinomial(GlmMultiBase):
    opt = False

    def __init__(self, y, weights, dtype):
        if y.ndim != 2:
            raise RuntimeError("y must be 2-dimensional.")
        if y.shape[1] <= 1:
            raise RuntimeError("adelie_core: y must have at least 2 columns (classes).")
        y = np.ascontiguousarray(y)
        self._init_common("multinomial", y, weights, dtype, K=y.shape[1])

    def reweight(self, weights=None):
        return multinomial(y=self.y, weights=self.weights if weights is None else weights, dtype=self.dtype)


class _Cox(GlmBase):
    def __init__(self, start, stop, status, strata, weights, tie_method, dtype):
        if status.ndim != 1:
            raise RuntimeError("y must be 1-dimensional.")
        n = status.shape[0]
        self.start = np.array(start, copy=True, dtype=dtype)
        self.stop = np.array(stop, copy=True, dtype=dtype)
        self.status = np.array(status, copy=True, dtype=dtype)
        self.strata = np.zeros(n, dtype=np.int64) if strata is None else np.array(strata, copy=True, dtype=np.int64)
        self.tie_method = tie_method
        for nm, a in (("start", self.start), ("stop", self.stop), ("strata", self.strata)):
            if a.shape != (n,):
                raise RuntimeError(f"adelie_core: {nm} must be (n,) where status is (n,).")
        self._init_common("cox", status, weights, dtype, start=self.start, stop=self.stop, strata=self.strata,
                          efron=int(tie_method == "efron"))

    def reweight(self, weights=None):
        return cox(start=self.start, stop=self.stop, status=self.status, strata=self.strata,
                   weights=self.weights if weights is None else weights, tie_method=self.tie_method, dtype=self.dtype)


def gaussian(y: np.ndarray, *, weights: np.ndarray = None, dtype: Union[np.float32, np.float64] = None, opt: bool = True):
    """Gaussian family (adelie/glm.py:374-453; CORE/glm/glm_gaussian.ipp:17-64)."""
    y, dtype = _coerce_dtype(np.asarray(y), dtype)
    return _Gaussian(y, weights, dtype, opt)


def binomial(y: np.ndarray, *, weights: np.ndarray = None, link: str = "logit", dtype: Union[np.float32, np.float64] = None):
    """Binomial family, logit or probit link (adelie/glm.py:83-196; CORE/glm/glm_binomial.ipp:47-98, probit :100-190)."""
    if link not in ("logit", "probit"):
        raise RuntimeError("link must be one of 'logit', 'probit'.")
    y, dtype = _coerce_dtype(np.asarray(y), dtype)
    return _Binomial(y, weights, dtype, link)


def poisson(y: np.ndarray, *, weights: np.ndarray = None, dtype: Union[np.float32, np.float64] = None):
    """Poisson family, log link (adelie/glm.py ``poisson``; CORE/glm/glm_poisson.ipp:7-66) -- SURVEY 8f rank 4."""
    y, dtype = _coerce_dtype(np.asarray(y), dtype)
    return _Poisson(y, weights, dtype)


def multigaussian(y: np.ndarray, *, weights: np.ndarray = None, dtype: Union[np.float32, np.float64] = None, opt: bool = True):
    """MultiGaussian family (adelie/glm.py:456-538; CORE/glm/glm_multigaussian.ipp:17-68)."""
    y, dtype = _coerce_dtype(np.asarray(y), dtype)
    return _MultiGaussian(y, weights, dtype, opt)


def multinomial(y: np.ndarray, *, weights: np.ndarray = None, dtype: Union[np.float32, np.float64] = None):
    """Multinomial family (adelie/glm.py ``multinomial``; CORE/glm/glm_multinomial.ipp:6-132) -- SURVEY 8f rank 4."""
    y, dtype = _coerce_dtype(np.asarray(y), dtype)
    return _Multinomial(y, weights, dtype)


def cox(start: np.ndarray, stop: np.ndarray, status: np.ndarray, *, strata: np.ndarray = None, weights: np.ndarray = None,
        tie_method: str = "efron", dtype: Union[np.float32, np.float64] = None):
    """Cox family (adelie/glm.py:199-371; CORE/glm/glm_cox.ipp:356-750)."""
    status, dtype = _coerce_dtype(np.asarray(status), dtype)
    return _Cox(np.asarray(start), np.asarray(stop), status, strata, weights, tie_method, dtype)
