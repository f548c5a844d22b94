"""CPU: small building blocks of the oracle against NumPy / brute force (reference tests/test_bcd.py:7-189,
tests/test_matrix.py:251-411, tests/test_optimization.py search_pivot)."""
import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import oracle as orc


@pytest.mark.parametrize("q", [1, 2, 5, 10, 37])
def test_jacobi_eigh(q):
    rng = np.random.RandomState(q)
    A = rng.normal(size=(q, q)); A = A @ A.T
    D, V = orc.jacobi_eigh(A)
    np.testing.assert_allclose(D, np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(V @ np.diag(D) @ V.T, A, atol=1e-10)
    np.testing.assert_allclose(V.T @ V, np.eye(q), atol=1e-12)


@pytest.mark.parametrize("p", [1, 10, 100])
@pytest.mark.parametrize("sparsity", [0.1, 0.5, 0.9])
def test_root_bounds(p, sparsity):                       # tests/test_bcd.py:7-82
    rng = np.random.RandomState(p)
    quad = rng.uniform(0, 1, p); quad[rng.choice(p, int(sparsity * p), replace=False)] = 0
    linear = np.sqrt(quad) * rng.normal(size=p) + 1e-3 * rng.normal(size=p) * (quad > 0)
    l1 = 0.5 * np.linalg.norm(linear)
    if l1 <= 0 or np.all(quad == 0):
        return
    lo = orc.root_lower_bound(quad, linear, l1)
    assert orc.root_function(lo, quad, linear, l1) >= -1e-10
    hi = orc.root_upper_bound(quad, linear, l1, 0.0)
    assert orc.root_function(hi, quad, linear, l1) <= 1e-10


@pytest.mark.parametrize("p", [1, 5, 20])
@pytest.mark.parametrize("l2", [0.0, 1e-2])
@pytest.mark.parametrize("solver", ["newton", "newton_abs"])
def test_bcd_solve_is_minimiser(p, l2, solver):          # tests/test_bcd.py:83-150 (objective rule, cvxpy replaced by scipy)
    rng = np.random.RandomState(10 * p + int(l2 > 0))
    quad = rng.uniform(0.1, 1, p)
    linear = np.sqrt(quad) * rng.normal(size=p)
    l1 = 0.3 * np.linalg.norm(linear)
    out = orc.bcd_solve(quad, linear, l1, l2, solver=solver)
    x = out["beta"]
    obj = lambda b: 0.5 * np.sum(quad * b * b) - linear @ b + l1 * np.linalg.norm(b) + 0.5 * l2 * np.sum(b * b)
    best = min((minimize(obj, x0, method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-14, maxiter=20000)).fun
                for x0 in (x + 1e-3 * rng.normal(size=p), np.zeros(p) + 1e-3)), default=np.inf)
    assert obj(x) <= best + 1e-9
    # stationarity: (L + l2) x + l1 x/||x|| = v
    if np.linalg.norm(x) > 0:
        np.testing.assert_allclose((quad + l2) * x + l1 * x / np.linalg.norm(x), linear, atol=1e-9)
    else:
        assert np.linalg.norm(linear) <= l1 + 1e-12


def test_bcd_edge_cases():
    quad = np.array([1.0, 2.0, 3.0]); v = np.array([0.1, -0.2, 0.05])
    assert np.all(orc.bcd_solve(quad, v, 1.0, 0.0)["beta"] == 0)                      # ||v|| <= l1
    np.testing.assert_allclose(orc.bcd_solve(quad, v, 0.0, 0.5)["beta"], v / (quad + 0.5))   # l1 == 0


@pytest.mark.parametrize("n", [1, 2, 10, 50])
def test_search_pivot(n):
    rng = np.random.RandomState(n)
    x = np.arange(n, dtype=float)
    y = np.sort(rng.uniform(size=n)) ** 3
    idx, mses = orc.search_pivot(x, y)
    # brute force: regress y on (x_i - x)_+ with intercept, pick the best i >= 1 (search_pivot.hpp:7-62)
    best, arg = np.inf, 0
    for i in range(1, n):
        t = np.maximum(x[i] - x, 0)
        tc, yc = t - t.mean(), y - y.mean()
        vt = tc @ tc
        mse = -(tc @ yc) ** 2 / vt if vt > 0 else np.nan
        if mse < best:
            best, arg = mse, i
    assert idx == (arg if n > 1 else 0)


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
def test_dense_ops(dtype, atol):                          # run_naive, tests/test_matrix.py:251-411
    rng = np.random.RandomState(0)
    n, p = 57, 23
    X = np.asfortranarray(rng.normal(size=(n, p)).astype(dtype))
    M = orc.dense(X, n_threads=2)
    orc.set_config("min_bytes", 20)                       # force the threaded code paths (tests/test_matrix.py:9-11)
    try:
        v = rng.normal(size=n).astype(dtype); w = rng.uniform(size=n).astype(dtype)
        for j in range(p):
            assert abs(M.cmul(j, v, w) - X[:, j] @ (v * w)) < atol * 10
        out = np.zeros(n, dtype=dtype); M.ctmul(3, 0.7, out)
        np.testing.assert_allclose(out, 0.7 * X[:, 3], atol=atol)
        for j, q in [(0, 1), (2, 5), (10, 13), (0, p)]:
            o = np.empty(q, dtype=dtype); M.bmul(j, q, v, w, o)
            np.testing.assert_allclose(o, X[:, j:j + q].T @ (v * w), atol=atol * 10)
            vv = rng.normal(size=q).astype(dtype); o2 = np.ones(n, dtype=dtype); M.btmul(j, q, vv, o2)
            np.testing.assert_allclose(o2, 1 + X[:, j:j + q] @ vv, atol=atol * 10)
            C = np.empty((q, q), dtype=dtype, order="F"); M.cov(j, q, np.sqrt(w), C)
            np.testing.assert_allclose(C, X[:, j:j + q].T @ (w[:, None] * X[:, j:j + q]), atol=atol * 10)
        o = np.empty(p, dtype=dtype); M.mul(v, w, o)
        np.testing.assert_allclose(o, X.T @ (v * w), atol=atol * 10)
    finally:
        orc.set_config("min_bytes", 1 << 17)
