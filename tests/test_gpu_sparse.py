"""GPU parity tests of the sparse CSC matrix (reference MatrixNaiveSparse, matrix_naive_sparse.ipp):
operators against dense NumPy as the reference's run_naive does (T/test_matrix.py:251-411; atol 1e-14 f64 / 1e-4 f32),
and the path solver on a sparse matrix against (i) the same problem held as a dense device matrix (the reference's
special-matrix-vs-dense equivalence, T/test_solver.py:652-818) and (ii) the CPU oracle's sparse implementation."""
import numpy as np
import pytest
import scipy.sparse as sp

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    s = np.max(np.abs(b)) if np.size(b) else 0.0
    return np.max(np.abs(a - b)) / (s if s > 0 else 1.0)


def _rand_csc(n, p, density, dtype, seed):
    rng = np.random.default_rng(seed)
    M = sp.random(n, p, density=density, format="csc", dtype=np.float64, random_state=rng, data_rvs=rng.standard_normal)
    return M.astype(dtype)


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 1e-4)])      # the reference's own float32 operator tolerance (T/test_matrix.py: atol 1e-4)
@pytest.mark.parametrize("n,p,density", [(200, 50, 0.3), (1500, 120, 0.05), (64, 7, 1.0)])
def test_sparse_operators_vs_numpy(dtype, atol, n, p, density):
    M = _rand_csc(n, p, density, dtype, seed=n + p)
    D = np.asarray(M.todense())
    X = ad.matrix.sparse(M)
    rng = np.random.default_rng(1)
    v = rng.normal(size=n).astype(dtype); w = rng.uniform(0.1, 1, size=n).astype(dtype)
    out = np.empty(p, dtype=dtype)
    X.mul(v, w, out)
    np.testing.assert_allclose(out, D.T @ (v * w), rtol=atol, atol=atol)
    X.sq_mul(w, out)
    np.testing.assert_allclose(out, (D ** 2).T @ w, rtol=atol, atol=atol)
    for j, q in [(0, 1), (p // 3, min(5, p - p // 3)), (p - 1, 1), (0, p if p <= 10 else 9)]:
        o = np.empty(q, dtype=dtype)
        X.bmul(j, q, v, w, o)
        np.testing.assert_allclose(o, D[:, j:j + q].T @ (v * w), rtol=atol, atol=atol)
        assert abs(X.cmul(j, v, w) - D[:, j] @ (v * w)) <= atol * (1 + abs(D[:, j] @ (v * w)))
        vv = rng.normal(size=q).astype(dtype)
        acc = rng.normal(size=n).astype(dtype); expect = acc + D[:, j:j + q] @ vv
        X.btmul(j, q, vv, acc)
        np.testing.assert_allclose(acc, expect, rtol=atol, atol=atol)
        acc = np.zeros(n, dtype=dtype)
        X.ctmul(j, 1.5, acc)
        np.testing.assert_allclose(acc, 1.5 * D[:, j], rtol=atol, atol=atol)
        C = np.empty((q, q), dtype=dtype, order="F")
        X.cov(j, q, np.sqrt(w), C)
        np.testing.assert_allclose(C, D[:, j:j + q].T @ (w[:, None] * D[:, j:j + q]), rtol=atol, atol=atol)
    with pytest.raises(RuntimeError, match="bmul"):
        X.bmul(p - 1, 2, v, w, np.empty(2, dtype=dtype))


def test_sparse_constructor_contract():
    with pytest.raises(RuntimeError, match="scipy sparse"):
        ad.matrix.sparse(np.zeros((3, 3)))
    M = sp.csr_matrix(np.eye(4))
    with pytest.warns(UserWarning, match="CSC"):
        X = ad.matrix.sparse(M)
    assert X.shape == (4, 4)
    with pytest.raises(RuntimeError, match="numpy.float32 or numpy.float64"):
        ad.matrix.sparse(sp.csc_matrix(np.eye(3, dtype=np.int32)))


def _problem(n, p, density, dtype, seed):
    M = _rand_csc(n, p, density, dtype, seed)
    rng = np.random.default_rng(seed + 1)
    beta = np.zeros(p); supp = rng.choice(p, max(1, p // 10), replace=False); beta[supp] = rng.normal(size=supp.size)
    eta = M @ beta
    y = (eta + (np.linalg.norm(beta) * 0.5 + 0.1) * rng.normal(size=n)).astype(dtype)
    return M, y


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-6), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,p,density,sizes,alpha,intercept", [
    (2000, 150, 0.05, None, 1.0, True),                                   # lasso (config-4 shape in small)
    (1200, 60, 0.2, [1, 4, 10, 3, 12, 5, 2, 8, 15], 1.0, True),           # groups (pair joins in the Gram)
    (1200, 60, 0.2, [6] * 10, 0.5, False),
])
def test_sparse_path_vs_dense_and_oracle(dtype, rtol, n, p, density, sizes, alpha, intercept):
    M, y = _problem(n, p, density, dtype, seed=7)
    groups = np.arange(p) if sizes is None else np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(int)
    tol = 1e-12 if dtype == np.float64 else 1e-7
    nt = 1e-12 if dtype == np.float64 else 1e-5
    kw = dict(groups=groups, alpha=alpha, intercept=intercept, tol=tol, newton_tol=nt, early_exit=False, lmda_path_size=20, min_ratio=0.05)
    st = ad.grpnet(ad.matrix.sparse(M), ad.glm.gaussian(y, dtype=dtype), progress_bar=False, **kw)
    dn = ad.grpnet(np.asfortranarray(M.todense(), dtype=dtype), ad.glm.gaussian(y, dtype=dtype), progress_bar=False, **kw)
    assert st.error == "" and dn.error == "", (st.error, dn.error)
    B, Bd = np.asarray(st.betas.todense()), np.asarray(dn.betas.todense())
    np.testing.assert_allclose(st.lmdas, dn.lmdas, rtol=rtol)
    assert _rel(B, Bd) <= rtol, _rel(B, Bd)
    assert _rel(np.asarray(st.intercepts), np.asarray(dn.intercepts)) <= rtol
    o = orc.grpnet(M, orc.glm_spec("gaussian", y, dtype=dtype), **kw)
    assert o.error == ""
    assert _rel(B, np.asarray(o.betas.todense())) <= rtol
    np.testing.assert_allclose(st.devs, o.devs, rtol=10 * rtol, atol=10 * rtol)


def test_sparse_binomial_path_vs_dense():
    n, p = 1500, 40
    M = _rand_csc(n, p, 0.2, np.float64, seed=3)
    rng = np.random.default_rng(4)
    beta = np.zeros(p); beta[:5] = rng.normal(size=5)
    eta = M @ beta
    y = rng.binomial(1, 1 / (1 + np.exp(-eta))).astype(np.float64)
    kw = dict(alpha=0.5, tol=1e-12, irls_tol=1e-10, early_exit=False, lmda_path_size=10, min_ratio=0.2, progress_bar=False)
    st = ad.grpnet(ad.matrix.sparse(M), ad.glm.binomial(y), **kw)
    dn = ad.grpnet(np.asfortranarray(M.todense()), ad.glm.binomial(y), **kw)
    assert st.error == "" and dn.error == ""
    assert _rel(np.asarray(st.betas.todense()), np.asarray(dn.betas.todense())) <= 1e-6


def test_device_generated_sparse_matrix():
    X = ad.matrix.sparse_device_random(5000, 64, 37, dtype=np.float32, seed=5)
    M = X.to_host()
    assert M.shape == (5000, 64) and M.nnz == 64 * 37
    assert np.all(np.diff(M.indptr) == 37) and M.has_sorted_indices
    for j in range(0, 64, 9):
        idx = M.indices[M.indptr[j]:M.indptr[j + 1]]
        assert np.all(np.diff(idx) > 0)
    assert abs(M.data.mean()) < 0.1 and abs(M.data.std() - 1) < 0.1
    v = np.ones(5000, dtype=np.float32); out = np.empty(64, dtype=np.float32)
    X.mul(v, v, out)
    np.testing.assert_allclose(out, np.asarray(M.sum(axis=0)).ravel(), atol=1e-3)


@pytest.mark.parametrize("intercept", [True, False])
@pytest.mark.parametrize("family", ["multigaussian", "multinomial"])
def test_sparse_multi_response_vs_dense(family, intercept):
    """Multi-response on sparse X (reference: the generic kronecker_eye / concatenate views work on any matrix,
    matrix_naive_kronecker_eye.ipp:29-352): the expanded CSC matrix [kron(1, I_K) | kron(X, I_K)] runs on the single-response sparse
    kernels; the path must equal the one of the dense matrix on the multi-response kernels."""
    n, p, K = 600, 30, 3
    M = _rand_csc(n, p, 0.25, np.float64, seed=5)
    rng = np.random.default_rng(6)
    B = np.zeros((p, K)); B[:4] = rng.normal(size=(4, K))
    eta = np.asarray(M @ B)
    if family == "multigaussian":
        y = eta + rng.normal(size=(n, K))
        glm = lambda: ad.glm.multigaussian(y)
        kw = dict(tol=1e-12)
    else:
        P = np.exp(eta - eta.max(axis=1, keepdims=True)); P /= P.sum(axis=1, keepdims=True)
        y = np.array([rng.multinomial(1, P[i]) for i in range(n)], dtype=np.float64)
        glm = lambda: ad.glm.multinomial(y)
        kw = dict(tol=1e-12, irls_tol=1e-10)
    kw.update(alpha=0.7, intercept=intercept, groups=np.arange(0, p, 3), early_exit=False, lmda_path_size=10, min_ratio=0.2, progress_bar=False)
    st = ad.grpnet(ad.matrix.sparse(M), glm(), **kw)
    dn = ad.grpnet(np.asfortranarray(M.todense()), glm(), **kw)
    assert st.error == "" and dn.error == ""
    assert st.betas.shape == dn.betas.shape == (10, p * K)
    np.testing.assert_allclose(st.lmdas, dn.lmdas, rtol=1e-9)
    assert _rel(np.asarray(st.betas.todense()), np.asarray(dn.betas.todense())) <= 1e-6
    np.testing.assert_allclose(st.intercepts, dn.intercepts, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(st.devs, dn.devs, rtol=1e-6, atol=1e-9)
