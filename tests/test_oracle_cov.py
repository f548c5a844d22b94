"""CPU: the covariance-method oracle (oracle/cov_oracle.hpp) pinned the way the reference pins its own covariance solver.

* MatrixCovDense / MatrixCovLazyCov operators vs NumPy (the reference's tests/test_matrix.py `run_cov` pattern: bmul / mul / to_dense
  against the dense matrix);
* `test_gaussian_cov` of the reference (tests/test_solver.py:983-1026): gaussian_cov(A = X^T X / n, v = X^T y / n) on the lambdas of
  grpnet(X, y, intercept=False) gives the same coefficients -- replayed against the naive oracle path, which is itself pinned by the
  reference's own state.check (tests/test_reference_state_check.py);
* `test_solve_gaussian_pin_cov` of the reference (:536-596): random fixed screen set, dense and lazy_cov matrices, solve + warm start
  at 0.8 * the last lambda; cvxpy is absent, so the solutions are checked against the KKT conditions of the problem restricted to
  the screen set.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from cov_data import create_data_gaussian_pin_cov, kkt_cov


@pytest.mark.parametrize("dtype, atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("kind", ["dense", "lazy"])
def test_cov_matrix_ops(kind, order, dtype, atol):
    rng = np.random.default_rng(0)
    n, p = 30, 17
    X = np.array(rng.standard_normal((n, p)) / np.sqrt(n), dtype=dtype, order=order)
    A = (X.T.astype(np.float64) @ X.astype(np.float64))
    M = orc.cov_lazy(X) if kind == "lazy" else orc.cov_dense(np.array(A, dtype=dtype, order=order))
    assert M.cols() == p
    for _ in range(3):
        k = int(rng.integers(1, p))
        indices = np.sort(rng.choice(p, k, replace=False))
        values = rng.standard_normal(k).astype(dtype)
        subset = np.sort(rng.choice(p, int(rng.integers(1, p)), replace=False))
        out = np.empty(subset.size, dtype=dtype)
        M.bmul(subset, indices, values, out)
        np.testing.assert_allclose(out, (values.astype(np.float64) @ A[indices][:, subset]), atol=atol)
        out = np.empty(p, dtype=dtype)
        M.mul(indices, values, out)
        np.testing.assert_allclose(out, values.astype(np.float64) @ A[indices], atol=atol)
        i0 = int(rng.integers(0, p - 3)); q = int(rng.integers(1, p - i0))
        blk = np.empty((q, q), dtype=dtype, order="F")
        M.to_dense(i0, q, blk)
        np.testing.assert_allclose(blk, A[i0:i0 + q, i0:i0 + q], atol=atol)


@pytest.mark.parametrize("n, p, G", [[10, 50, 10], [40, 13, 7], [100, 60, 60], [200, 120, 30]])
@pytest.mark.parametrize("alpha", [1.0, 0.6])
def test_gaussian_cov_equals_naive(n, p, G, alpha):
    """tests/test_solver.py:983-1026 of the reference."""
    rng = np.random.default_rng(n + p)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    beta = np.zeros(p); beta[rng.choice(p, max(2, p // 10), replace=False)] = rng.standard_normal(max(2, p // 10))
    y = X @ beta + rng.standard_normal(n)
    groups = np.sort(np.concatenate([[0], rng.choice(np.arange(1, p), size=G - 1, replace=False)])).astype(int)
    glm = orc.glm_spec("gaussian", y, dtype=np.float64)
    sn = orc.grpnet(X, glm, groups=groups, alpha=alpha, intercept=False, adev_tol=0.2 if n < p else 0.6, tol=1e-12)
    assert sn.error == "" and len(sn.lmdas) > 3
    A = np.asfortranarray(X.T @ X) / n
    v = X.T @ y / n
    sc = orc.gaussian_cov(A, v, groups=groups, alpha=alpha, lmda_path=sn.lmdas, tol=1e-12, early_exit=False)
    assert sc.error == ""
    np.testing.assert_allclose(sc.lmdas, sn.lmdas)
    np.testing.assert_allclose(sn.betas.toarray(), sc.betas.toarray(), rtol=1e-6, atol=1e-7)
    gs = np.diff(np.concatenate([groups, [p]]))
    kkt_cov(A, v, groups, gs, np.sqrt(gs), alpha, sc.betas.toarray(), sc.lmdas, atol=1e-5)
    # lazy_cov of X / sqrt(n) is the same problem
    sl = orc.gaussian_cov(orc.cov_lazy(np.asfortranarray(X / np.sqrt(n))), v, groups=groups, alpha=alpha, lmda_path=sn.lmdas, tol=1e-12, early_exit=False)
    np.testing.assert_allclose(sl.betas.toarray(), sc.betas.toarray(), rtol=1e-7, atol=1e-9)
    # generated path + early exit on the relative deviance change (cov::early_exit, solver_gaussian_cov.hpp:186-203)
    s2 = orc.gaussian_cov(A, v, groups=groups, alpha=alpha, tol=1e-12, rdev_tol=1e-2)
    assert s2.error == "" and 2 <= len(s2.lmdas) <= 100
    np.testing.assert_allclose(s2.lmda_max, sn.lmda_max, rtol=1e-10)
    d = s2.devs
    assert np.all(np.diff(d) >= -1e-12)
    if len(d) < 100:
        assert d[-1] - d[-2] <= 1e-2 * d[-1]


@pytest.mark.parametrize("n, p, G, S", [[10, 4, 2, 2], [10, 100, 10, 2], [10, 100, 20, 13], [100, 23, 4, 3], [100, 100, 50, 20]])
def test_solve_gaussian_pin_cov(n, p, G, S):
    """tests/test_solver.py:536-596 of the reference."""
    args, ex = create_data_gaussian_pin_cov(n, p, G, S)
    sols = []
    for A in (orc.cov_dense(np.asfortranarray(ex["A"])), orc.cov_lazy(ex["WsqrtX"])):
        a = {k: v for k, v in args.items() if k != "constraints"}
        st = orc.gaussian_pin_cov(A, **a, tol=1e-12)
        assert st.error == ""
        kkt_cov(ex["A"], ex["v"], args["groups"], ex["group_sizes"], args["penalty"], args["alpha"], st.betas.toarray(), st.lmdas,
                restrict=args["screen_set"])
        a2 = dict(a)
        a2.update(lmda_path=[st.lmdas[-1] * 0.8], rsq=st.rsq, screen_beta=st.screen_beta, screen_grad=st.screen_grad,
                  screen_is_active=st.screen_is_active, active_set_size=int(st.active_set_size), active_set=st.active_set)
        st2 = orc.gaussian_pin_cov(A, **a2, tol=1e-12)
        assert st2.error == ""
        kkt_cov(ex["A"], ex["v"], args["groups"], ex["group_sizes"], args["penalty"], args["alpha"], st2.betas.toarray(), st2.lmdas,
                restrict=args["screen_set"])
        sols.append((st.betas.toarray(), st2.betas.toarray()))
    np.testing.assert_allclose(sols[0][0], sols[1][0], atol=1e-9)
    np.testing.assert_allclose(sols[0][1], sols[1][1], atol=1e-9)
