"""GPU parity tests of the multi-response path (multigaussian_naive / multiglm_naive states; the kron(X, I_K) + intercept-column
layout of adelie/solver.py:699-846) against the CPU oracle, plus the kronecker_eye / concatenate operator front-ends against
dense NumPy (the reference's own test style, tests/test_matrix.py:251-411)."""
import numpy as np
import pytest
import scipy.sparse as sp

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    scale = np.max(np.abs(b)) if np.size(b) else 0.0
    return np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0)


def _multi_data(n, p, K, seed, dtype):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p)), dtype=dtype)
    B = np.zeros((p, K)); sup = rng.choice(p, max(2, p // 8), replace=False)
    B[sup] = rng.standard_normal((sup.size, K))
    Y = (X @ B + 0.7 + rng.standard_normal((n, K))).astype(dtype)
    return X, np.ascontiguousarray(Y)


def _compare(st, ref, rtol):
    assert st.error == "", st.error
    assert ref.error == "", ref.error
    assert len(st.lmdas) == len(ref.lmdas)
    np.testing.assert_allclose(st.lmdas, ref.lmdas, rtol=rtol)
    B = np.asarray(st.betas.todense()); Br = np.asarray(ref.betas.todense())
    assert B.shape == Br.shape
    assert _rel(B, Br) <= rtol, _rel(B, Br)
    I = np.asarray(st.intercepts); Ir = np.asarray(ref.intercepts)
    assert I.shape == Ir.shape
    assert _rel(I, Ir) <= rtol, _rel(I, Ir)
    np.testing.assert_allclose(st.devs, ref.devs, rtol=10 * rtol, atol=10 * rtol)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-6), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,p,K,alpha,intercept", [
    (300, 40, 3, 1.0, True),
    (300, 40, 4, 0.6, False),
    (2000, 64, 8, 1.0, True),            # K = 8 groups of size 8 (the snp_unphased K=8 layout of config 5)
    (257, 30, 2, 0.9, True),             # ragged row count (pad rows)
])
def test_multigaussian_path_vs_oracle(dtype, rtol, n, p, K, alpha, intercept):
    X, Y = _multi_data(n, p, K, 11, dtype)
    tol = 1e-12 if dtype == np.float64 else 1e-10        # float32: converge past sqrt(tol) ~ 1e-4 so that rounding is what is compared
    newton_tol = 1e-12 if dtype == np.float64 else 1e-6
    kw = dict(alpha=alpha, intercept=intercept, tol=tol, early_exit=False, lmda_path_size=20, min_ratio=0.05, newton_tol=newton_tol)
    st = ad.grpnet(X, ad.glm.multigaussian(Y, dtype=dtype), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("multigaussian", Y, dtype=dtype), **kw)
    _compare(st, ref, rtol)
    assert st.intercepts.shape == (len(st.lmdas), K)
    assert st.betas.shape == (len(st.lmdas), p * K)


def test_multigaussian_feature_groups():
    """groups over features (3 features per group) -> groups of 3 * K coefficients spanning several X columns."""
    n, p, K = 400, 30, 3
    X, Y = _multi_data(n, p, K, 5, np.float64)
    groups = np.arange(0, p, 3)
    kw = dict(groups=groups, tol=1e-12, early_exit=False, lmda_path_size=15, min_ratio=0.1)
    st = ad.grpnet(X, ad.glm.multigaussian(Y), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("multigaussian", Y), **kw)
    _compare(st, ref, 1e-6)


@pytest.mark.parametrize("intercept", [True, False])
def test_multigaussian_irls_matches_oracle(intercept):
    """multigaussian(opt=False) runs the multi-GLM IRLS driver (solver_multiglm_naive.hpp)."""
    n, p, K = 300, 25, 3
    X, Y = _multi_data(n, p, K, 9, np.float64)
    kw = dict(tol=1e-12, irls_tol=1e-12, early_exit=False, lmda_path_size=12, min_ratio=0.1, intercept=intercept)
    st = ad.grpnet(X, ad.glm.multigaussian(Y, opt=False), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("multigaussian", Y, opt=False), **kw)
    _compare(st, ref, 1e-6)
    opt = ad.grpnet(X, ad.glm.multigaussian(Y, opt=True), progress_bar=False, **kw)
    assert _rel(np.asarray(st.betas.todense()), np.asarray(opt.betas.todense())) < 1e-6


def test_multi_equals_separate_ridge_free_lasso_when_K1_like():
    """With alpha = 1 and K = 1-column-per-class independence broken only by the group norm: sanity-check the KKT conditions of
    the multi-response solution directly (group soft-threshold stationarity) instead of trusting only the oracle."""
    n, p, K = 500, 20, 3
    X, Y = _multi_data(n, p, K, 21, np.float64)
    st = ad.grpnet(X, ad.glm.multigaussian(Y), progress_bar=False, tol=1e-14, early_exit=False, lmda_path_size=10, min_ratio=0.2)
    assert st.error == ""
    B = np.asarray(st.betas.todense()); I = np.asarray(st.intercepts)
    for l in range(len(st.lmdas)):
        Bl = B[l].reshape(p, K)
        R = Y - X @ Bl - I[l][None]
        grad = X.T @ R / (n * K)                 # weights 1/n, every term / K (glm_multigaussian.ipp)
        lam = st.lmdas[l] * np.sqrt(K)           # penalty sqrt(group size)
        nrm = np.linalg.norm(grad, axis=1)
        act = np.linalg.norm(Bl, axis=1) > 0
        assert np.all(nrm[~act] <= lam * (1 + 1e-6))
        np.testing.assert_allclose(nrm[act], lam, rtol=1e-5)
        np.testing.assert_allclose(R.mean(axis=0), 0, atol=1e-8)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_kronecker_concatenate_operators(dtype):
    rng = np.random.default_rng(0)
    n, p, K = 50, 7, 3
    Xh = np.asfortranarray(rng.standard_normal((n, p)), dtype=dtype)
    A = ad.matrix.concatenate([ad.matrix.kronecker_eye(np.ones((n, 1), dtype=dtype), K), ad.matrix.kronecker_eye(Xh, K)], axis=1)
    D = np.hstack([np.kron(np.ones((n, 1)), np.eye(K)), np.kron(Xh, np.eye(K))]).astype(dtype)
    assert A.shape == D.shape
    rt = 1e-10 if dtype == np.float64 else 2e-5
    v = rng.standard_normal(n * K).astype(dtype); w = rng.uniform(0.5, 1, n * K).astype(dtype)
    out = np.empty(D.shape[1], dtype=dtype); A.mul(v, w, out)
    np.testing.assert_allclose(out, D.T @ (v * w), rtol=rt, atol=rt)
    for j in [0, K - 1, K, K + 4, D.shape[1] - 1]:
        np.testing.assert_allclose(A.cmul(j, v, w), D[:, j] @ (v * w), rtol=rt, atol=rt)
        o = rng.standard_normal(n * K).astype(dtype); o2 = o.copy()
        A.ctmul(j, 1.5, o)
        np.testing.assert_allclose(o, o2 + 1.5 * D[:, j], rtol=rt, atol=rt)
    j, q = K + 2, 5
    ob = np.empty(q, dtype=dtype); A.bmul(j, q, v, w, ob)
    np.testing.assert_allclose(ob, D[:, j:j + q].T @ (v * w), rtol=rt, atol=rt)
    vv = rng.standard_normal(q).astype(dtype); o = np.zeros(n * K, dtype=dtype); A.btmul(j, q, vv, o)
    np.testing.assert_allclose(o, D[:, j:j + q] @ vv, rtol=rt, atol=rt)
    C = np.empty((q, q), dtype=dtype, order="F"); A.cov(j, q, np.sqrt(w), C)
    np.testing.assert_allclose(C, D[:, j:j + q].T @ (w[:, None] * D[:, j:j + q]), rtol=rt, atol=rt)
    sq = np.empty(D.shape[1], dtype=dtype); A.sq_mul(w, sq)
    np.testing.assert_allclose(sq, (D ** 2).T @ w, rtol=rt, atol=rt)
    S = sp.random(4, D.shape[1], density=0.3, random_state=1, format="csr", dtype=dtype)
    o = np.empty((4, n * K), dtype=dtype); A.sp_tmul(S, o)
    np.testing.assert_allclose(o, (S @ D.T), rtol=rt, atol=rt)
