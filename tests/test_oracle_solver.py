"""CPU: pins the oracle's path solver (oracle/adelie_oracle.hpp) with checks that do not share code with it:
  * scikit-learn's coordinate-descent Lasso / ElasticNet (independent solver) on the lasso configs,
  * NumPy KKT residuals of the group elastic net (Gaussian and binomial),
  * the acceptance rule of the reference's tests (tests/test_solver.py:408-466): objective not worse than an
    independent proximal-gradient solution,
  * invariants of the returned state (reference adelie/state.py:1421-1674: rsq, resid, grad)."""
import numpy as np
import pytest
from sklearn.linear_model import ElasticNet, Lasso

from oracle import oracle as orc


def make(n, p, G, seed, equal=False, glm="gaussian"):
    rng = np.random.RandomState(seed)
    if equal:
        groups = (p // G) * np.arange(G)
    else:
        groups = np.sort(np.concatenate([[0], rng.choice(np.arange(1, p), size=G - 1, replace=False)])).astype(int)
    gs = np.diff(np.concatenate([groups, [p]]))
    X = np.asfortranarray(rng.normal(size=(n, p)))
    beta = np.zeros(p)
    nz = rng.choice(p, max(1, p // 10), replace=False)
    beta[nz] = rng.normal(size=nz.size)
    eta = X @ beta
    if glm == "gaussian":
        y = eta + np.linalg.norm(beta) * rng.normal(size=n)
    else:
        y = rng.binomial(1, 1 / (1 + np.exp(-eta / max(np.linalg.norm(beta), 1e-12)))).astype(float)
    w = rng.uniform(0.5, 1.5, n); w /= w.sum()
    return X, y, w, groups, gs


def kkt_residual(X, resid_w, beta, lmda, alpha, groups, gs, penalty):
    """max violation of the group elastic-net stationarity conditions given the weighted residual (negative loss gradient)."""
    g = X.T @ resid_w
    worst = 0.0
    for j, q, pk in zip(groups, gs, penalty):
        b = beta[j:j + q]; gg = g[j:j + q]
        bn = np.linalg.norm(b)
        if bn > 0:
            worst = max(worst, np.max(np.abs(gg - lmda * pk * (alpha * b / bn + (1 - alpha) * b))))
        else:
            worst = max(worst, max(0.0, np.linalg.norm(gg) - lmda * alpha * pk))
    return worst


def test_lasso_matches_sklearn_config1():
    """BASELINE config 1: Gaussian lasso n=1000 p=500, 100-lambda path."""
    rng = np.random.RandomState(0)
    n, p = 1000, 500
    X = np.asfortranarray(rng.normal(size=(n, p)))
    beta = np.zeros(p); idx = rng.choice(p, 25, replace=False); beta[idx] = rng.normal(size=25)
    y = X @ beta + np.linalg.norm(beta) * rng.normal(size=n)
    st = orc.grpnet(X, orc.glm_spec("gaussian", y), tol=1e-18, early_exit=False)
    assert st.error == "" and len(st.lmdas) == 100
    assert np.all(np.diff(st.lmdas) < 0) and abs(st.lmdas[-1] / st.lmdas[0] - 1e-2) < 1e-10
    for l in [1, 10, 40, 70]:
        m = Lasso(alpha=st.lmdas[l], fit_intercept=True, tol=1e-15, max_iter=200000).fit(X, y)
        b = np.asarray(st.betas[l].todense()).ravel()
        # sklearn stops on a duality-gap criterion: agreement to ~1e-6 is its accuracy, the KKT check below is the tight one
        assert np.max(np.abs(b - m.coef_)) < 2e-6
        assert abs(st.intercepts[l] - m.intercept_) < 2e-6
        w = np.full(n, 1 / n)
        r = y - X @ b - st.intercepts[l]
        ones = np.ones(p)
        assert kkt_residual(X, w * r, b, st.lmdas[l], 1.0, np.arange(p), ones.astype(int), ones) < 1e-7


def test_elastic_net_matches_sklearn():
    X, y, _, _, _ = make(400, 60, 60, 1)
    alpha = 0.6
    st = orc.grpnet(X, orc.glm_spec("gaussian", y), alpha=alpha, tol=1e-18, early_exit=False, lmda_path_size=20, min_ratio=0.05)
    assert st.error == ""
    for l in [3, 10, 19]:
        m = ElasticNet(alpha=st.lmdas[l], l1_ratio=alpha, fit_intercept=True, tol=1e-15, max_iter=200000).fit(X, y)
        b = np.asarray(st.betas[l].todense()).ravel()
        assert np.max(np.abs(b - m.coef_)) < 2e-6


@pytest.mark.parametrize("alpha,intercept", [(1.0, True), (0.5, True), (0.3, False)])
def test_group_path_kkt_and_invariants(alpha, intercept):
    X, y, w, groups, gs = make(300, 80, 15, 2)
    penalty = np.sqrt(gs).astype(float)
    st = orc.grpnet(X, orc.glm_spec("gaussian", y, w), groups=groups, alpha=alpha, penalty=penalty, intercept=intercept,
                    tol=1e-18, early_exit=False, lmda_path_size=25, min_ratio=0.05)
    assert st.error == "" and len(st.lmdas) == 25
    yc_var = np.sum(w * (y - intercept * np.sum(w * y)) ** 2)
    for l in range(0, 25, 4):
        b = np.asarray(st.betas[l].todense()).ravel()
        r = y - X @ b - st.intercepts[l]
        assert kkt_residual(X, w * r, b, st.lmdas[l], alpha, groups, gs, penalty) < 1e-7
        if intercept:
            assert abs(np.sum(w * r)) < 1e-9                       # intercept stationarity
        # devs = rsq / y_var = 1 - ||r||_W^2 / ||y_c||_W^2
        assert abs(st.devs[l] - (1 - np.sum(w * r ** 2) / yc_var)) < 1e-8
    # state invariants at the last lambda (adelie/state.py:1421-1674)
    b = np.asarray(st.betas[-1].todense()).ravel()
    yc = y - np.sum(w * y) * intercept
    np.testing.assert_allclose(st.resid, yc - X @ b, atol=1e-9)
    Xm = X.T @ w
    g_expected = X.T @ (w * st.resid) - intercept * np.sum(w * st.resid) * Xm
    np.testing.assert_allclose(st.grad, g_expected, atol=1e-9)


def test_group_lasso_objective_vs_proximal_gradient():
    """Acceptance rule of the reference (tests/test_solver.py:444-466): objective <= independent solver's * (1 + eps)."""
    X, y, w, groups, gs = make(200, 40, 8, 3)
    penalty = np.sqrt(gs).astype(float)
    st = orc.grpnet(X, orc.glm_spec("gaussian", y, w), groups=groups, penalty=penalty, tol=1e-18, early_exit=False,
                    lmda_path_size=10, min_ratio=0.1)
    l = 9; lam = st.lmdas[l]
    b_or = np.asarray(st.betas[l].todense()).ravel(); b0_or = st.intercepts[l]

    def objective(b, b0):
        r = y - X @ b - b0
        return 0.5 * np.sum(w * r ** 2) + lam * sum(pk * np.linalg.norm(b[j:j + q]) for j, q, pk in zip(groups, gs, penalty))

    # independent solver: proximal gradient (ISTA with group soft-thresholding) on centred data
    Xc = X - (X.T @ w)[None]; yc = y - np.sum(w * y)
    Lip = np.linalg.eigvalsh(Xc.T @ (w[:, None] * Xc)).max()
    b = np.zeros(X.shape[1])
    for _ in range(20000):
        g = -Xc.T @ (w * (yc - Xc @ b))
        z = b - g / Lip
        for j, q, pk in zip(groups, gs, penalty):
            zn = np.linalg.norm(z[j:j + q])
            z[j:j + q] *= max(0.0, 1 - lam * pk / Lip / max(zn, 1e-300))
        if np.max(np.abs(z - b)) < 1e-15:
            b = z; break
        b = z
    b0 = np.sum(w * (y - X @ b))
    assert objective(b_or, b0_or) <= objective(b, b0) * (1 + 1e-10)
    assert np.max(np.abs(b - b_or)) < 1e-6


def test_binomial_path_kkt():
    X, y, w, groups, gs = make(400, 50, 10, 4, glm="binomial")
    penalty = np.sqrt(gs).astype(float)
    alpha = 0.5
    st = orc.grpnet(X, orc.glm_spec("binomial", y, w), groups=groups, alpha=alpha, penalty=penalty, tol=1e-18, irls_tol=1e-12,
                    early_exit=False, lmda_path_size=12, min_ratio=0.1)
    assert st.error == "" and len(st.lmdas) == 12
    for l in range(0, 12, 3):
        b = np.asarray(st.betas[l].todense()).ravel()
        eta = X @ b + st.intercepts[l]
        r = w * (y - 1 / (1 + np.exp(-eta)))
        assert kkt_residual(X, r, b, st.lmdas[l], alpha, groups, gs, penalty) < 1e-6
        assert abs(np.sum(r)) < 1e-7
    assert np.all(np.diff(st.devs) > -1e-12)


def test_pin_solve_and_warm_start():
    """reference tests/test_solver.py:483-532: pin solve on a fixed screen set, then continue from the solution."""
    X, y, w, groups, gs = make(150, 30, 6, 5)
    penalty = np.sqrt(gs).astype(float)
    screen_set = np.array([0, 2, 3, 5])
    lmda_path = np.array([0.3, 0.1, 0.05])
    out = orc.pin_naive_solve(X, y, groups=groups, alpha=1.0, penalty=penalty, weights=w, screen_set=screen_set, lmda_path=lmda_path,
                              tol=1e-18)
    assert out.error == ""
    assert len(out.lmdas) == 3
    # restricted problem KKT on the screen groups
    cols = np.concatenate([np.arange(groups[g], groups[g] + gs[g]) for g in screen_set])
    b = np.asarray(out.betas[-1].todense()).ravel()
    assert np.all(b[np.setdiff1d(np.arange(X.shape[1]), cols)] == 0)
    r = y - X @ b - out.intercepts[-1]
    sub_groups = groups[screen_set]; sub_gs = gs[screen_set]; sub_pen = penalty[screen_set]
    assert kkt_residual(X, w * r, b, lmda_path[-1], 1.0, sub_groups, sub_gs, sub_pen) < 1e-7
    # warm start: one more lambda from the returned screen_beta / active set must reproduce a cold solve
    out2 = orc.pin_naive_solve(X, y, groups=groups, alpha=1.0, penalty=penalty, weights=w, screen_set=screen_set,
                               lmda_path=np.array([0.04]), tol=1e-18, screen_beta=out.screen_beta, screen_is_active=out.screen_is_active,
                               active_set=out.active_set, active_set_size=out.active_set_size, rsq=out.rsq, resid=out.resid)
    cold = orc.pin_naive_solve(X, y, groups=groups, alpha=1.0, penalty=penalty, weights=w, screen_set=screen_set,
                               lmda_path=np.array([0.04]), tol=1e-18)
    np.testing.assert_allclose(np.asarray(out2.betas.todense()), np.asarray(cold.betas.todense()), atol=1e-8)


def test_float32_path_close_to_float64():
    X, y, w, groups, gs = make(300, 40, 8, 6)
    kw = dict(groups=groups, tol=1e-7, newton_tol=1e-5, early_exit=False, lmda_path_size=10, min_ratio=0.1)
    a = orc.grpnet(X, orc.glm_spec("gaussian", y, w, dtype=np.float64), **kw)
    b = orc.grpnet(X.astype(np.float32), orc.glm_spec("gaussian", y, w, dtype=np.float32), **kw)
    assert a.error == "" and b.error == ""
    Ba, Bb = np.asarray(a.betas.todense()), np.asarray(b.betas.todense())
    assert np.max(np.abs(Ba - Bb)) / np.max(np.abs(Ba)) < 1e-3


def test_sparse_and_multigaussian_run():
    import scipy.sparse as sp
    rng = np.random.RandomState(7)
    X = sp.random(200, 40, density=0.1, random_state=rng, format="csc")
    y = rng.normal(size=200)
    st = orc.grpnet(X, orc.glm_spec("gaussian", y), tol=1e-18, early_exit=False, lmda_path_size=8, min_ratio=0.1)
    dense = orc.grpnet(np.asfortranarray(X.toarray()), orc.glm_spec("gaussian", y), tol=1e-18, early_exit=False, lmda_path_size=8, min_ratio=0.1)
    assert st.error == "" and dense.error == ""
    np.testing.assert_allclose(np.asarray(st.betas.todense()), np.asarray(dense.betas.todense()), atol=1e-9)
    # multigaussian: K responses, group = feature x K classes; check block KKT in (n,K) form
    n, p, K = 150, 12, 3
    Xd = np.asfortranarray(rng.normal(size=(n, p))); Y = rng.normal(size=(n, K)) + Xd[:, :2] @ rng.normal(size=(2, K))
    mg = orc.grpnet(Xd, orc.glm_spec("multigaussian", Y), tol=1e-18, early_exit=False, lmda_path_size=6, min_ratio=0.2)
    assert mg.error == ""
    l = 5
    B = np.asarray(mg.betas[l].todense()).reshape(p, K)
    R = Y - Xd @ B - mg.intercepts[l][None]
    Gm = Xd.T @ (R / n) / K                       # gradient blocks, weights w/K
    lam = mg.lmdas[l]
    for j in range(p):
        bn = np.linalg.norm(B[j])
        pk = np.sqrt(K)
        if bn > 0:
            assert np.max(np.abs(Gm[j] - lam * pk * B[j] / bn)) < 1e-7
        else:
            assert np.linalg.norm(Gm[j]) <= lam * pk + 1e-9
