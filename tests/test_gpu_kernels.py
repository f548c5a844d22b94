"""GPU parity of the standalone operators through the C ABI: dense matrix ops (reference run_naive,
tests/test_matrix.py:251-411; atol 1e-4 f32 / 1e-12 f64), GLM families vs the reference-generated golden vectors and the
oracle (tests/test_glm.py), the device prox vs the oracle (tests/test_bcd.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "glm_golden.npz"))


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,p", [(1, 1), (2, 3), (57, 23), (1000, 40), (5001, 17)])
@pytest.mark.parametrize("order", ["F", "C"])
def test_dense_run_naive(dtype, atol, n, p, order):
    rng = np.random.RandomState(0)
    Xh = rng.normal(size=(n, p)).astype(dtype)
    Xh = np.asfortranarray(Xh) if order == "F" else np.ascontiguousarray(Xh)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = ad.matrix.dense(Xh)
    O = orc.dense(Xh)
    atol = atol * max(1.0, n / 100)
    rtol = 2e-5 if dtype == np.float32 else 1e-10          # the float32 oracle accumulates in float32 (as the reference does)
    assert M.shape == (n, p)
    v = rng.normal(size=n).astype(dtype); w = rng.uniform(size=n).astype(dtype)
    for j in range(0, p, max(1, p // 5)):
        assert abs(float(M.cmul(j, v, w)) - O.cmul(j, v, w)) < atol
        out = rng.normal(size=n).astype(dtype); ref = out.copy()
        M.ctmul(j, 0.37, out); O.ctmul(j, 0.37, ref)
        np.testing.assert_allclose(out, ref, atol=atol, rtol=rtol)
    for j, q in [(0, 1), (0, p), (p // 2, p - p // 2), (max(0, p - 3), min(3, p))]:
        o = np.empty(q, dtype=dtype); r = np.empty(q, dtype=dtype)
        M.bmul(j, q, v, w, o); O.bmul(j, q, v, w, r)
        np.testing.assert_allclose(o, r, atol=atol, rtol=rtol)
        vv = rng.normal(size=q).astype(dtype)
        o2 = rng.normal(size=n).astype(dtype); r2 = o2.copy()
        M.btmul(j, q, vv, o2); O.btmul(j, q, vv, r2)
        np.testing.assert_allclose(o2, r2, atol=atol, rtol=rtol)
        C = np.empty((q, q), dtype=dtype, order="F"); Cr = np.empty((q, q), dtype=dtype, order="F")
        M.cov(j, q, np.sqrt(w), C); O.cov(j, q, np.sqrt(w), Cr)
        np.testing.assert_allclose(C, Cr, atol=atol, rtol=rtol)
    o = np.empty(p, dtype=dtype); r = np.empty(p, dtype=dtype)
    M.mul(v, w, o); O.mul(v, w, r)
    np.testing.assert_allclose(o, r, atol=atol, rtol=rtol)
    sq = np.empty(p, dtype=dtype); M.sq_mul(w, sq)
    np.testing.assert_allclose(sq, (Xh.astype(np.float64) ** 2).T @ w, atol=atol, rtol=rtol)
    # sugar: X @ v, X.T @ v, sparse rhs
    b = rng.normal(size=p).astype(dtype)
    np.testing.assert_allclose(M @ b, Xh @ b, atol=atol * 10)
    np.testing.assert_allclose(M.T @ v, Xh.T @ v, atol=atol * 10)
    S = sp.random(3, p, density=0.5, random_state=rng, format="csr", dtype=dtype)
    out = np.empty((3, n), dtype=dtype); M.sp_tmul(S, out)
    np.testing.assert_allclose(out, (S @ Xh.T), atol=atol * 10)


def test_dense_shape_errors():
    M = ad.matrix.dense(np.asfortranarray(np.zeros((4, 3))))
    with pytest.raises(RuntimeError, match="cmul"):
        M.cmul(3, np.zeros(4), np.zeros(4))
    with pytest.raises(RuntimeError, match="mul"):
        M.mul(np.zeros(3), np.zeros(4), np.zeros(3))
    with pytest.raises(RuntimeError, match="cov"):
        M.cov(2, 2, np.zeros(4), np.zeros((2, 2), order="F"))


def _glm_check(model, pre, dtype, rtol, atol):
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    eta = c(G[pre + "eta"])
    grad = np.empty_like(eta); model.gradient(eta, grad)
    np.testing.assert_allclose(grad, G[pre + "grad"], rtol=rtol, atol=atol)
    hess = np.empty_like(eta); model.hessian(eta, c(G[pre + "grad"]), hess)
    np.testing.assert_allclose(hess, G[pre + "hess"], rtol=rtol * 10, atol=atol)
    ihg = np.empty_like(eta); model.inv_hessian_gradient(eta, c(G[pre + "grad"]), c(G[pre + "hess"]), ihg)
    if dtype == np.float64:
        np.testing.assert_allclose(ihg, G[pre + "inv_hess_grad"], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(model.loss(eta), G[pre + "loss"], rtol=rtol * 10, atol=atol)
    np.testing.assert_allclose(model.loss_full(), G[pre + "loss_full"], rtol=rtol * 10, atol=atol)
    inv = np.empty_like(eta); model.inv_link(eta, inv)
    np.testing.assert_allclose(inv, G[pre + "inv_link"], rtol=rtol, atol=atol)


@pytest.mark.parametrize("dtype,rtol,atol", [(np.float64, 1e-10, 1e-12), (np.float32, 1e-4, 1e-6)])
@pytest.mark.parametrize("n", [1, 2, 5, 10, 20, 100])
def test_glm_families_vs_reference_golden(dtype, rtol, atol, n):
    pre = f"gaussian_{n}_"
    _glm_check(ad.glm.gaussian(G[pre + "y"], weights=G[pre + "w"], dtype=dtype), pre, dtype, rtol, atol)
    for binary in (0, 1):
        pre = f"binomial_{n}_{binary}_"
        _glm_check(ad.glm.binomial(G[pre + "y"], weights=G[pre + "w"], dtype=dtype), pre, dtype, rtol, atol)
    for K in (1, 2, 3, 4):
        pre = f"multigaussian_{n}_{K}_"
        _glm_check(ad.glm.multigaussian(G[pre + "y"], weights=G[pre + "w"], dtype=dtype), pre, dtype, rtol, atol)
    for K in (2, 3, 4):
        for binary in (0, 1):
            pre = f"multinomial_{n}_{K}_{binary}_"
            _glm_check(ad.glm.multinomial(G[pre + "y"], weights=G[pre + "w"], dtype=dtype), pre, dtype, rtol * 5, atol * 5)
    for binary in (0, 1):
        pre = f"probit_{n}_{binary}_"
        _glm_check(ad.glm.binomial(G[pre + "y"], weights=G[pre + "w"], link="probit", dtype=dtype), pre, dtype, rtol * 5, atol * 5)
    pre = f"poisson_{n}_"
    _glm_check(ad.glm.poisson(G[pre + "y"], weights=G[pre + "w"], dtype=dtype), pre, dtype, rtol, atol)


def test_glm_large_vs_oracle():
    rng = np.random.RandomState(1)
    n = 100_003
    y = rng.binomial(1, 0.4, n).astype(np.float64); w = rng.uniform(size=n); w /= w.sum()
    eta = rng.normal(size=n) * 3
    m = ad.glm.binomial(y, weights=w); spec = orc.glm_spec("binomial", y, w)
    g = np.empty(n); m.gradient(eta, g)
    go = orc.glm_eval(spec, "gradient", eta=eta)
    np.testing.assert_allclose(g, go, rtol=1e-12, atol=1e-18)
    h = np.empty(n); m.hessian(eta, g, h)
    np.testing.assert_allclose(h, orc.glm_eval(spec, "hessian", eta=eta, grad=go), rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(m.loss(eta), orc.glm_eval(spec, "loss", eta=eta), rtol=1e-12)
    np.testing.assert_allclose(m.loss_full(), orc.glm_eval(spec, "loss_full"), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("p", [1, 2, 10, 33, 100])
@pytest.mark.parametrize("sparsity", [0.0, 0.5])
@pytest.mark.parametrize("l2", [0.0, 1e-2])
@pytest.mark.parametrize("solver", ["newton", "newton_abs"])
def test_bcd_solve_vs_oracle(p, sparsity, l2, solver):
    rng = np.random.RandomState(p)
    quad = rng.uniform(0.05, 1, p); quad[rng.choice(p, int(sparsity * p), replace=False)] *= 1e-3
    linear = np.sqrt(quad) * rng.normal(size=p)
    l1 = 0.4 * np.linalg.norm(linear)
    out = ad.bcd.solve(quad=quad, linear=linear, l1=l1, l2=l2, solver=solver)
    ref = orc.bcd_solve(quad, linear, l1, l2, solver=solver)
    np.testing.assert_allclose(out["beta"], ref["beta"], rtol=1e-9, atol=1e-12)
    x = out["beta"]
    if np.linalg.norm(x) > 0:                                         # stationarity of the prox objective
        np.testing.assert_allclose((quad + l2) * x + l1 * x / np.linalg.norm(x), linear, atol=1e-9)
    assert np.all(ad.bcd.solve(quad=quad, linear=linear, l1=2 * np.linalg.norm(linear), l2=l2, solver=solver)["beta"] == 0)
    lo = ad.bcd.root_lower_bound(quad=quad + l2, linear=linear, l1=l1)
    hi = ad.bcd.root_upper_bound(quad=quad + l2, linear=linear, l1=l1, zero_tol=0.0)
    assert abs(lo - orc.root_lower_bound(quad + l2, linear, l1)) < 1e-10
    assert abs(hi - orc.root_upper_bound(quad + l2, linear, l1, 0.0)) < 1e-9 * max(1, hi)
    assert ad.bcd.root_function(lo, D=quad + l2, v=linear, l1=l1) >= -1e-9
    assert ad.bcd.root_function(hi, D=quad + l2, v=linear, l1=l1) <= 1e-9


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_poisson_path_vs_oracle(alpha):
    """the poisson family through the generic IRLS driver."""
    data = ad.data.dense(800, 30, 10, glm="poisson", seed=6)
    y = data["glm"].y
    kw = dict(groups=data["groups"], penalty=data["penalty"], alpha=alpha, tol=1e-13, irls_tol=1e-11, early_exit=False, lmda_path_size=10, min_ratio=0.2)
    st = ad.grpnet(data["X"], ad.glm.poisson(y), progress_bar=False, **kw)
    ref = orc.grpnet(data["X"], orc.glm_spec("poisson", y), **kw)
    assert st.error == "" and ref.error == "", (st.error, ref.error)
    B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
    assert np.max(np.abs(B - Br)) <= 1e-6 * np.max(np.abs(Br))
    np.testing.assert_allclose(st.intercepts, ref.intercepts, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(st.devs, ref.devs, rtol=1e-6, atol=1e-8)


def test_probit_path_vs_oracle():
    data = ad.data.dense(800, 30, 10, glm="binomial", seed=8)
    y = data["glm"].y
    kw = dict(groups=data["groups"], penalty=data["penalty"], alpha=0.7, tol=1e-13, irls_tol=1e-11, early_exit=False, lmda_path_size=10, min_ratio=0.2)
    st = ad.grpnet(data["X"], ad.glm.binomial(y, link="probit"), progress_bar=False, **kw)
    ref = orc.grpnet(data["X"], orc.glm_spec("probit", y), **kw)
    assert st.error == "" and ref.error == "", (st.error, ref.error)
    B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
    assert np.max(np.abs(B - Br)) <= 1e-6 * np.max(np.abs(Br))
    np.testing.assert_allclose(st.devs, ref.devs, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("K,intercept", [(3, True), (4, False)])
def test_multinomial_path_vs_oracle(K, intercept):
    """the multinomial family through the multi-response IRLS driver (solver_multiglm_naive.hpp)."""
    rng = np.random.default_rng(K)
    n, p = 600, 20
    X = np.asfortranarray(rng.normal(size=(n, p)))
    Bt = np.zeros((p, K)); Bt[:4] = rng.normal(size=(4, K))
    eta = X @ Bt; P = np.exp(eta - eta.max(1, keepdims=True)); P /= P.sum(1, keepdims=True)
    Y = np.array([rng.multinomial(1, pp) for pp in P]).astype(np.float64)
    kw = dict(tol=1e-13, irls_tol=1e-11, early_exit=False, lmda_path_size=8, min_ratio=0.3, intercept=intercept)
    st = ad.grpnet(X, ad.glm.multinomial(Y), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("multinomial", Y), **kw)
    assert st.error == "" and ref.error == "", (st.error, ref.error)
    B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
    assert B.shape == Br.shape == (len(st.lmdas), p * K)
    assert np.max(np.abs(B - Br)) <= 1e-6 * np.max(np.abs(Br))
    np.testing.assert_allclose(np.asarray(st.intercepts), np.asarray(ref.intercepts), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(st.devs, ref.devs, rtol=1e-6, atol=1e-8)
