"""GPU: the covariance-method solver through the C ABI against the CPU oracle (oracle/cov_oracle.hpp).

* adelie.matrix.dense(method="cov") / adelie.matrix.lazy_cov operators vs NumPy (reference tests/test_matrix.py: atol 1e-14-class for
  float64, 1e-4 for float32) and the reference's error strings;
* state.gaussian_pin_cov mirroring tests/test_solver.py:536-596 (fixed random screen set, dense and lazy_cov, warm start at 0.8 * the
  last lambda) vs the oracle's pin::cov::solve: 1e-6 rel float64 / 1e-4 float32, with every cluster size of the device kernel;
* solver.gaussian_cov mirroring tests/test_solver.py:983-1026: equals grpnet(intercept=False) on the same lambdas, equals the oracle
  path, KKT; a screen set large enough for the automatic multi-CTA cluster.
"""
import numpy as np
import pytest

import adelie_b200 as ad
from oracle import oracle as orc
from cov_data import create_data_gaussian_pin_cov, kkt_cov

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_cluster():
    yield
    ad.configs.set_configs("cov_cluster", None)


@pytest.mark.parametrize("dtype, atol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("kind", ["dense", "lazy"])
def test_cov_matrix_ops(kind, order, dtype, atol):
    rng = np.random.default_rng(0)
    n, p = 300, 70
    X = np.array(rng.standard_normal((n, p)) / np.sqrt(n), dtype=dtype, order=order)
    A = X.T.astype(np.float64) @ X.astype(np.float64)
    M = ad.matrix.lazy_cov(X, n_threads=2) if kind == "lazy" else ad.matrix.dense(np.array(A, dtype=dtype, order=order), method="cov", n_threads=2)
    assert isinstance(M, ad.matrix.MatrixCovBase64 if dtype == np.float64 else ad.matrix.MatrixCovBase32)
    assert M.cols() == p and M.rows() == p and M.shape == (p, p)
    for _ in range(4):
        k = int(rng.integers(1, p))
        indices = np.sort(rng.choice(p, k, replace=False))
        values = rng.standard_normal(k).astype(dtype)
        subset = np.sort(rng.choice(p, int(rng.integers(1, p)), replace=False))
        out = np.empty(subset.size, dtype=dtype)
        M.bmul(subset, indices, values, out)
        np.testing.assert_allclose(out, values.astype(np.float64) @ A[indices][:, subset], atol=atol)
        out = np.empty(p, dtype=dtype)
        M.mul(indices, values, out)
        np.testing.assert_allclose(out, values.astype(np.float64) @ A[indices], atol=atol)
        i0 = int(rng.integers(0, p - 3)); q = int(rng.integers(1, p - i0))
        blk = np.empty((q, q), dtype=dtype, order="F")
        M.to_dense(i0, q, blk)
        np.testing.assert_allclose(blk, A[i0:i0 + q, i0:i0 + q], atol=atol)
    if kind == "lazy":
        assert 0 < M.cached_rows() <= p
    # argument checks (matrix_cov_base.hpp:66-131)
    with pytest.raises(RuntimeError, match="bmul\\(\\) is given inconsistent inputs"):
        M.bmul(np.arange(3), np.arange(2), np.ones(2, dtype=dtype), np.empty(4, dtype=dtype))
    with pytest.raises(RuntimeError, match="mul\\(\\) is given inconsistent inputs"):
        M.mul(np.arange(2), np.ones(3, dtype=dtype), np.empty(p, dtype=dtype))
    with pytest.raises(RuntimeError, match="to_dense\\(\\) is given inconsistent inputs"):
        M.to_dense(p - 1, 3, np.empty((3, 3), dtype=dtype, order="F"))
    with pytest.raises(RuntimeError, match="mat must be \\(p, p\\)"):
        ad.matrix.dense(np.zeros((3, 4), dtype=dtype), method="cov")._core()
    M.close()


def _pin_args(args):
    return {k: v for k, v in args.items()}


@pytest.mark.parametrize("cluster", [1, 2, 8])
@pytest.mark.parametrize("n, p, G, S", [[10, 4, 2, 2], [10, 100, 10, 2], [10, 100, 20, 13], [100, 23, 4, 3], [100, 100, 50, 20]])
def test_solve_gaussian_pin_cov(n, p, G, S, cluster):
    """tests/test_solver.py:536-596 of the reference, against the oracle."""
    ad.configs.set_configs("cov_cluster", cluster)
    args, ex = create_data_gaussian_pin_cov(n, p, G, S)
    oa = {k: v for k, v in args.items() if k != "constraints"}
    ref = orc.gaussian_pin_cov(orc.cov_dense(np.asfortranarray(ex["A"])), **oa, tol=1e-12)
    assert ref.error == ""
    for A in (ad.matrix.dense(np.asfortranarray(ex["A"]), method="cov", n_threads=3), ad.matrix.lazy_cov(ex["WsqrtX"], n_threads=3)):
        state = ad.state.gaussian_pin_cov(A=A, **args, tol=1e-12)
        state.check(method="assert")
        state = state.solve()
        assert state.error == ""
        assert state.cov_cluster == cluster
        state.check(method="assert")
        np.testing.assert_allclose(state.lmdas, ref.lmdas)
        np.testing.assert_allclose(state.betas.toarray(), ref.betas.toarray(), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(state.rsqs, ref.rsqs, rtol=1e-6, atol=1e-10)
        np.testing.assert_allclose(state.screen_grad, ref.screen_grad, rtol=1e-6, atol=1e-9)
        np.testing.assert_array_equal(state.screen_is_active, ref.screen_is_active)
        assert state.active_set_size == ref.active_set_size
        np.testing.assert_array_equal(state.active_set[: state.active_set_size], ref.active_set)
        assert state.iters == ref.iters
        kkt_cov(ex["A"], ex["v"], args["groups"], ex["group_sizes"], args["penalty"], args["alpha"], state.betas.toarray(), state.lmdas,
                restrict=args["screen_set"])
        # warm start at 0.8 * the last lambda from the solved state (:586-596)
        a2 = dict(args)
        a2.update(lmda_path=[state.lmdas[-1] * 0.8], rsq=state.rsq, screen_beta=state.screen_beta, screen_grad=state.screen_grad,
                  screen_is_active=state.screen_is_active, active_set_size=state.active_set_size, active_set=state.active_set)
        st2 = ad.state.gaussian_pin_cov(A=A, **a2, tol=1e-12).solve()
        assert st2.error == ""
        o2 = {k: v for k, v in a2.items() if k != "constraints"}
        ref2 = orc.gaussian_pin_cov(orc.cov_dense(np.asfortranarray(ex["A"])), **o2, tol=1e-12)
        np.testing.assert_allclose(st2.betas.toarray(), ref2.betas.toarray(), rtol=1e-6, atol=1e-9)
        kkt_cov(ex["A"], ex["v"], args["groups"], ex["group_sizes"], args["penalty"], args["alpha"], st2.betas.toarray(), st2.lmdas,
                restrict=args["screen_set"])


def _problem(n, p, G, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    k = max(2, p // 10)
    beta = np.zeros(p); beta[rng.choice(p, k, replace=False)] = rng.standard_normal(k)
    y = X @ beta + rng.standard_normal(n)
    groups = np.sort(np.concatenate([[0], rng.choice(np.arange(1, p), size=G - 1, replace=False)])).astype(int)
    A = np.asfortranarray(X.T @ X) / n
    v = X.T @ y / n
    return X.astype(dtype), y.astype(dtype), groups, A.astype(dtype), v.astype(dtype)


@pytest.mark.parametrize("n, p, G", [[10, 50, 10], [40, 13, 7], [200, 120, 30]])
@pytest.mark.parametrize("alpha", [1.0, 0.6])
def test_gaussian_cov_vs_naive_and_oracle(n, p, G, alpha):
    """tests/test_solver.py:983-1026 of the reference: gaussian_cov == grpnet(intercept=False) on the same lambdas."""
    X, y, groups, A, v = _problem(n, p, G, n + p)
    sn = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, alpha=alpha, intercept=False, adev_tol=0.2 if n < p else 0.6, tol=1e-12, progress_bar=False)
    assert sn.error == "" and len(sn.lmdas) > 3
    sc = ad.gaussian_cov(A=A, v=v, groups=groups, alpha=alpha, lmda_path=sn.lmdas, tol=1e-12, early_exit=False, check_state=True, progress_bar=False)
    assert sc.error == ""
    sc.check(method="assert")
    np.testing.assert_allclose(sn.betas.toarray(), sc.betas.toarray(), rtol=1e-6, atol=1e-7)
    so = orc.gaussian_cov(A, v, groups=groups, alpha=alpha, lmda_path=sn.lmdas, tol=1e-12, early_exit=False)
    np.testing.assert_allclose(sc.betas.toarray(), so.betas.toarray(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(sc.devs, so.devs, rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(sc.grad, so.grad, rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(sc.screen_set, so.screen_set)
    np.testing.assert_allclose(sc.intercepts, 0)
    gs = np.diff(np.concatenate([groups, [p]]))
    kkt_cov(A, v, groups, gs, np.sqrt(gs), alpha, sc.betas.toarray(), sc.lmdas, atol=1e-5)
    # the generated path: lmda_max, early exit on the relative deviance change, strong rule
    s2 = ad.gaussian_cov(A=A, v=v, groups=groups, alpha=alpha, tol=1e-12, rdev_tol=1e-2, progress_bar=False)
    o2 = orc.gaussian_cov(A, v, groups=groups, alpha=alpha, tol=1e-12, rdev_tol=1e-2)
    assert s2.error == "" and len(s2.lmdas) == len(o2.lmdas)
    np.testing.assert_allclose(s2.lmda_max, o2.lmda_max, rtol=1e-12)
    np.testing.assert_allclose(s2.betas.toarray(), o2.betas.toarray(), rtol=1e-6, atol=1e-9)
    s3 = ad.gaussian_cov(A=A, v=v, groups=groups, alpha=alpha, tol=1e-12, screen_rule="strong", lmda_path_size=20, early_exit=False, progress_bar=False)
    o3 = orc.gaussian_cov(A, v, groups=groups, alpha=alpha, tol=1e-12, screen_rule="strong", lmda_path_size=20, early_exit=False)
    np.testing.assert_allclose(s3.betas.toarray(), o3.betas.toarray(), rtol=1e-6, atol=1e-9)
    # lazy_cov of X / sqrt(n) is the same problem
    sl = ad.gaussian_cov(A=ad.matrix.lazy_cov(np.asfortranarray(X / np.sqrt(n))), v=v, groups=groups, alpha=alpha, lmda_path=sn.lmdas, tol=1e-12,
                         early_exit=False, progress_bar=False)
    np.testing.assert_allclose(sl.betas.toarray(), sc.betas.toarray(), rtol=1e-6, atol=1e-9)


def test_gaussian_cov_float32_and_large_screen_set():
    """float32 at 1e-4 rel; a screen set beyond 1024 values so that the automatic cluster has several CTAs (slices over DSMEM)."""
    n, p, G = 3000, 2400, 600
    X, y, groups, A, v = _problem(n, p, G, 7)
    lm = None
    for dtype, rtol in ((np.float64, 1e-6), (np.float32, 1e-4)):
        nt = 1e-12 if dtype == np.float64 else 1e-6
        so = orc.gaussian_cov(A.astype(dtype), v.astype(dtype), groups=groups, alpha=0.9, tol=1e-10, newton_tol=nt, lmda_path=lm, lmda_path_size=40, min_ratio=0.05, early_exit=False)
        assert so.error == ""
        lm = so.lmdas if lm is None else lm
        sc = ad.gaussian_cov(A=A.astype(dtype), v=v.astype(dtype), groups=groups, alpha=0.9, tol=1e-10, newton_tol=nt, lmda_path=lm, early_exit=False, progress_bar=False)
        assert sc.error == ""
        assert sc.cov_cluster >= 2 and len(sc.screen_beta) > 1024
        Bc, Bo = sc.betas.toarray(), so.betas.toarray()
        scale = np.max(np.abs(Bo), axis=1, keepdims=True) + 1e-30
        assert np.max(np.abs(Bc - Bo) / scale) < rtol
        np.testing.assert_allclose(sc.devs, so.devs, rtol=rtol)


def test_gaussian_cov_errors_and_warm_start():
    X, y, groups, A, v = _problem(60, 40, 12, 3)
    st = ad.gaussian_cov(A=A, v=v, groups=groups, lmda_path_size=30, early_exit=False, progress_bar=False)
    assert st.error == "" and len(st.lmdas) == 30
    # warm start: continue down a longer path from the returned state (adelie/solver.py:203-215)
    lm2 = st.lmdas[-1] * np.array([0.9, 0.8])
    st2 = ad.gaussian_cov(A=A, v=v, groups=groups, lmda_path=lm2, early_exit=False, warm_start=st, progress_bar=False)
    ref = orc.gaussian_cov(A, v, groups=groups, lmda_path=np.concatenate([st.lmdas, lm2]), early_exit=False)
    np.testing.assert_allclose(st2.betas.toarray(), ref.betas.toarray()[-2:], rtol=1e-5, atol=1e-7)
    # solver errors are returned, not raised (py_state.cpp:83-90)
    s3 = ad.gaussian_cov(A=A, v=v, groups=groups, max_iters=3, early_exit=False, progress_bar=False)
    assert s3.error.startswith("adelie_core solver: max coordinate descents reached")
    s4 = ad.gaussian_cov(A=A, v=v, groups=groups, max_active_size=1, early_exit=False, progress_bar=False)
    assert "Maximum number of active groups reached" in s4.error
    with pytest.raises(RuntimeError, match="v must be \\(p,\\) where A is \\(p, p\\)"):
        ad.state.gaussian_cov(A=A, v=v[:-1], constraints=None, groups=groups, group_sizes=np.diff(np.concatenate([groups, [40]])), alpha=1,
                              penalty=np.ones(12), screen_set=np.zeros(0, dtype=int), screen_beta=np.zeros(0), screen_is_active=np.zeros(0, dtype=bool),
                              active_set_size=0, active_set=np.zeros(12, dtype=int), rsq=0, lmda=np.inf, grad=v)
    # exit_cond is honoured
    calls = []
    s5 = ad.gaussian_cov(A=A, v=v, groups=groups, early_exit=False, exit_cond=lambda s: calls.append(1) or len(calls) >= 3, progress_bar=False)
    assert len(s5.lmdas) <= 4


def test_block_diag_and_sparse_cov_matrices():
    """adelie.matrix.block_diag(method="cov") and adelie.matrix.sparse(method="cov") (matrix_cov_block_diag.ipp, matrix_cov_sparse.ipp):
    operators vs the dense block matrix, and a path on the block-diagonal problem equals the two independent paths."""
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    blocks = []
    for q in (7, 12, 5):
        Z = rng.standard_normal((40, q))
        blocks.append(np.asfortranarray(Z.T @ Z / 40))
    D = np.asfortranarray(sp.block_diag(blocks).toarray())
    p = D.shape[0]
    for M in (ad.matrix.block_diag([blocks[0], ad.matrix.dense(blocks[1], method="cov"), blocks[2]], method="cov"),
              ad.matrix.sparse(sp.csc_matrix(D), method="cov")):
        assert isinstance(M, ad.matrix.MatrixCovBase64) and M.cols() == p
        idx = np.array([1, 2, 9, 20]); vals = rng.standard_normal(4); sub = np.array([0, 2, 8, 9, 23])
        out = np.empty(sub.size); M.bmul(sub, idx, vals, out)
        np.testing.assert_allclose(out, vals @ D[idx][:, sub], atol=1e-12)
        out = np.empty(p); M.mul(idx, vals, out)
        np.testing.assert_allclose(out, vals @ D[idx], atol=1e-12)
        blk = np.empty((6, 6), order="F"); M.to_dense(5, 6, blk)
        np.testing.assert_allclose(blk, D[5:11, 5:11], atol=1e-12)
    v = rng.standard_normal(p)
    lm = np.array([0.5, 0.3, 0.2, 0.1])
    full = ad.gaussian_cov(A=ad.matrix.block_diag(blocks, method="cov"), v=v, lmda_path=lm, tol=1e-12, early_exit=False, progress_bar=False)
    assert full.error == ""
    off = 0
    for Bk in blocks:
        q = Bk.shape[0]
        part = ad.gaussian_cov(A=Bk, v=v[off:off + q], lmda_path=lm, tol=1e-12, early_exit=False, progress_bar=False)
        np.testing.assert_allclose(full.betas.toarray()[:, off:off + q], part.betas.toarray(), rtol=1e-6, atol=1e-9)
        off += q


def test_gaussian_cov_large_groups():
    """Groups above 32 columns take the generic prox (record read from global memory, shared-memory reductions)."""
    n, p = 300, 120
    rng = np.random.default_rng(13)
    X = rng.standard_normal((n, p)); y = X[:, :8] @ rng.standard_normal(8) + rng.standard_normal(n)
    A = np.asfortranarray(X.T @ X) / n; v = X.T @ y / n
    groups = np.array([0, 40, 50, 110])
    kw = dict(groups=groups, alpha=0.9, tol=1e-12, lmda_path_size=12, min_ratio=0.05, early_exit=False)
    for cluster in (1, 4):
        ad.configs.set_configs("cov_cluster", cluster)
        st = ad.gaussian_cov(A=A, v=v, progress_bar=False, **kw)
        so = orc.gaussian_cov(A, v, **kw)
        assert st.error == "" and so.error == ""
        np.testing.assert_allclose(st.betas.toarray(), so.betas.toarray(), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(st.devs, so.devs, rtol=1e-6, atol=1e-10)
