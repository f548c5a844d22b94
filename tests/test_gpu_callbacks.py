"""GPU: user-defined GLMs and matrices across the boundary (the reference's pybind trampolines PyGlmBase / PyGlmMultiBase,
adelie/src/py_glm.cpp:8-92, 240-330, and PyMatrixNaiveBase, py_matrix.cpp:627-825).

A GLM written in NumPy as a subclass of ``ad.glm.GlmBase64`` / ``GlmMultiBase64`` (formulas of the reference's own test classes,
tests/test_glm.py GlmTestBinomialLogit / GlmTestMultinomial) must give the path of the built-in device family; a matrix written as a
subclass of ``ad.matrix.MatrixNaiveBase64`` must give the path of the dense matrix it represents; exceptions raised inside a callback
come back in the state's error string like every other solver error (py_state.cpp:83-90)."""
import numpy as np
import pytest
from scipy.special import xlogy

import adelie_b200 as ad

pytestmark = pytest.mark.gpu


class MyBinomial(ad.glm.GlmBase64):
    def __init__(self, y, weights, fail_after=None):
        ad.glm.GlmBase64.__init__(self, "my_binomial", y, weights)
        self.calls = 0
        self.fail_after = fail_after

    def gradient(self, eta, grad):
        self.calls += 1
        if self.fail_after is not None and self.calls > self.fail_after:
            raise ValueError("boom")
        grad[...] = self.weights * (self.y - 1 / (1 + np.exp(-eta)))

    def hessian(self, eta, grad, hess):
        p = 1 / (1 + np.exp(-eta))
        hess[...] = self.weights * p * (1 - p)

    def loss(self, eta):
        return np.sum(self.weights * (-self.y * eta + np.log1p(np.exp(eta))))

    def loss_full(self):
        return -np.sum(self.weights * (xlogy(self.y, self.y) + xlogy(1 - self.y, 1 - self.y)))


class MyMultinomial(ad.glm.GlmMultiBase64):
    def __init__(self, y, weights):
        ad.glm.GlmMultiBase64.__init__(self, "my_multinomial", y, weights)

    def _mu(self, eta):
        mu = np.exp(eta - eta.max(axis=-1, keepdims=True))
        return mu / np.sum(mu, axis=-1)[:, None]

    def gradient(self, eta, grad):
        K = self.y.shape[-1]
        grad[...] = self.weights[:, None] * (self.y - self._mu(eta)) / K

    def hessian(self, eta, grad, hess):
        K = self.y.shape[-1]
        mu = self._mu(eta)
        hess[...] = 2 * self.weights[:, None] / K * mu * (1 - mu)

    def loss(self, eta):
        K = self.y.shape[-1]
        m = eta.max(axis=-1)
        A = m + np.log(np.sum(np.exp(eta - m[:, None]), axis=-1))
        return np.sum(self.weights * (-np.sum(self.y * eta, axis=-1) + A)) / K

    def loss_full(self):
        K = self.y.shape[-1]
        return -np.sum(self.weights * np.sum(xlogy(self.y, self.y), axis=-1)) / K


def _data(n=300, p=40, seed=0):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    beta = np.zeros(p); beta[:5] = rng.standard_normal(5)
    eta = X @ beta
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    w = rng.uniform(1, 2, n); w /= w.sum()
    return X, y, w


def test_user_defined_glm_matches_builtin():
    X, y, w = _data()
    groups = np.arange(0, 40, 4)
    kw = dict(groups=groups, alpha=0.7, tol=1e-10, irls_tol=1e-10, lmda_path_size=15, min_ratio=0.05, early_exit=False, progress_bar=False)
    ref = ad.grpnet(X, ad.glm.binomial(y, weights=w), **kw)
    g = MyBinomial(y, w)
    st = ad.grpnet(X, g, **kw)
    assert ref.error == "" and st.error == "" and g.calls > 15
    np.testing.assert_allclose(st.lmdas, ref.lmdas, rtol=1e-10)
    np.testing.assert_allclose(st.betas.toarray(), ref.betas.toarray(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(st.intercepts, ref.intercepts, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(st.devs, ref.devs, rtol=1e-7, atol=1e-10)
    # the default inv_hessian_gradient of the base class (glm_base.ipp:25-36) vs the device one
    eta = np.linspace(-2, 2, y.size); grad = np.empty_like(eta); hess = np.empty_like(eta); a = np.empty_like(eta); b = np.empty_like(eta)
    g.gradient(eta, grad); g.hessian(eta, grad, hess); g.inv_hessian_gradient(eta, grad, hess, a)
    ad.glm.binomial(y, weights=w).inv_hessian_gradient(eta, grad, hess, b)
    np.testing.assert_allclose(a, b, rtol=1e-12)


def test_user_defined_glm_exception_is_a_solver_error():
    X, y, w = _data()
    kw = dict(groups=np.arange(0, 40, 4), lmda_path_size=15, min_ratio=0.05, early_exit=False, progress_bar=False)
    clean = MyBinomial(y, w)
    full = ad.grpnet(X, clean, **kw)
    assert full.error == "" and len(full.lmdas) == 15
    g = MyBinomial(y, w, fail_after=(2 * clean.calls) // 3)
    st = ad.grpnet(X, g, **kw)
    assert st.error.startswith("adelie_core solver: user-defined GLM: gradient() raised")
    assert isinstance(g._cb_errors[0], ValueError)
    L = len(st.lmdas)
    assert 0 < L < 15                             # valid up to the last solved lambda (py_state.cpp:83-90)
    np.testing.assert_allclose(st.betas.toarray(), full.betas.toarray()[:L], rtol=1e-10, atol=1e-12)


def test_user_defined_multi_glm_matches_builtin():
    rng = np.random.default_rng(1)
    n, p, K = 200, 12, 3
    X = np.asfortranarray(rng.standard_normal((n, p)))
    B = np.zeros((p, K)); B[:3] = rng.standard_normal((3, K))
    P = np.exp(X @ B); P /= P.sum(axis=1, keepdims=True)
    y = np.array([rng.multinomial(1, P[i]) for i in range(n)], dtype=np.float64)
    w = np.full(n, 1 / n)
    kw = dict(alpha=0.8, tol=1e-10, irls_tol=1e-10, lmda_path_size=10, min_ratio=0.1, early_exit=False, progress_bar=False)
    ref = ad.grpnet(X, ad.glm.multinomial(y, weights=w), **kw)
    st = ad.grpnet(X, MyMultinomial(y, w), **kw)
    assert ref.error == "" and st.error == ""
    np.testing.assert_allclose(st.betas.toarray(), ref.betas.toarray(), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(st.intercepts, ref.intercepts, rtol=1e-6, atol=1e-8)


class ScaledColumns(ad.matrix.MatrixNaiveBase64):
    """X = Z diag(s) without ever forming it on the host side of the API."""
    def __init__(self, Z, s):
        ad.matrix.MatrixNaiveBase64.__init__(self, n_threads=1)
        self.Z, self.s = Z, s

    def rows(self):
        return self.Z.shape[0]

    def cols(self):
        return self.Z.shape[1]

    def ctmul(self, j, v, out):
        out += v * self.s[j] * self.Z[:, j]

    def mul(self, v, weights, out):               # an operator the user chose to implement in NumPy
        out[...] = self.s * (self.Z.T @ (v * weights))


def test_user_defined_matrix_matches_dense():
    X, y, w = _data(seed=2)
    s = np.linspace(0.5, 2.0, X.shape[1])
    M = ScaledColumns(X, s)
    D = np.asfortranarray(X * s[None])
    kw = dict(groups=np.arange(0, 40, 4), tol=1e-10, lmda_path_size=15, min_ratio=0.05, early_exit=False, progress_bar=False)
    ref = ad.grpnet(D, ad.glm.gaussian(y, weights=w), **kw)
    st = ad.grpnet(M, ad.glm.gaussian(y, weights=w), **kw)
    assert ref.error == "" and st.error == ""
    np.testing.assert_allclose(st.betas.toarray(), ref.betas.toarray(), rtol=1e-8, atol=1e-10)
    # inherited operators run on the materialised device copy
    v = np.random.default_rng(0).standard_normal(X.shape[0]); out = np.empty(4)
    M.bmul(4, 4, v, w, out)
    np.testing.assert_allclose(out, D[:, 4:8].T @ (v * w), atol=1e-12)
    np.testing.assert_allclose(M @ np.ones(40), D @ np.ones(40), atol=1e-10)
    M.close()
