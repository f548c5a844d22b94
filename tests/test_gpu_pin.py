"""GPU: the Gaussian pin state in isolation (StateGaussianPinNaive), mirroring the reference's tests/test_solver.py:483-532
(`test_solve_gaussian_pin_naive`): a random fixed screen set, the lambda path of `create_data_gaussian(pin=True)` (:232-335), solve,
then a warm start at 0.8 * the last lambda from the solved state.  `state.check(method="assert")` runs before and after each solve
as `run_solve_gaussian` (:474-480) does; cvxpy is absent here, so the solution is compared with the CPU oracle's restatement of
pin::naive::solve (1e-6 rel, float64) and with the KKT conditions of the problem restricted to the screen set."""
import numpy as np
import pytest

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def create_data_gaussian_pin(n, p, G, S, intercept=True, alpha=1, sparsity=0.95, seed=0, min_ratio=0.5, n_lmdas=20):
    """tests/test_solver.py:214-335 of the reference with pin=True, method="naive" (same draws in the same order)."""
    np.random.seed(seed)
    groups = np.sort(np.concatenate([[0], np.random.choice(np.arange(1, p), size=G - 1, replace=False)])).astype(int)
    group_sizes = np.diff(np.concatenate([groups, [p]])).astype(int)
    X = np.random.normal(0, 1, (n, p))
    beta = np.random.normal(0, 1, p)
    beta[np.random.choice(p, int(sparsity * p), replace=False)] = 0
    y = X @ beta + np.random.normal(0, 1, n)
    X /= np.sqrt(n); y /= np.sqrt(n)
    penalty = np.random.uniform(0, 1, G)
    penalty[np.random.choice(G, int(0.05 * G), replace=False)] = 0
    penalty /= np.linalg.norm(penalty) / np.sqrt(p)
    weights = np.random.uniform(1, 2, n)
    weights /= np.sum(weights)
    X_means = np.sum(weights[:, None] * X, axis=0)
    X_c = X - intercept * X_means[None]
    y_mean = np.sum(weights * y)
    y_c = y - intercept * y_mean
    y_var = np.sum(weights * y_c ** 2)
    grad = X_c.T @ (weights * y_c)
    screen_set = np.random.choice(G, S, replace=False)
    abs_grad = np.array([np.linalg.norm(grad[g:g + gs]) for g, gs in zip(groups, group_sizes)])
    nz = penalty > 0
    lmda_max = np.max(abs_grad[nz] / (alpha * penalty[nz]))
    lmda_path = lmda_max * min_ratio ** (np.arange(n_lmdas) / (n_lmdas - 1))
    return dict(X=np.asfortranarray(X), y=y, constraints=None, groups=groups, alpha=alpha, penalty=penalty, weights=weights, rsq=0,
                intercept=intercept, active_set_size=0, active_set=np.zeros(G, dtype=int), lmda_path=lmda_path, y_mean=y_mean, resid=y_c,
                y_var=y_var, screen_set=screen_set, screen_is_active=np.zeros(S, dtype=bool), screen_beta=np.zeros(np.sum(group_sizes[screen_set]))), group_sizes


def kkt_on_screen_set(args, group_sizes, state):
    """Stationarity of min 1/2 ||y_c - X_c b||_W^2 + lmda sum_g p_g ||b_g|| over the screen groups only, at every solved lambda."""
    X, y, w = args["X"], args["y"], args["weights"]
    Xc = X - (X.T @ w)[None]; yc = y - np.sum(w * y)
    B = state.betas.toarray()
    for l, lmda in enumerate(state.lmdas):
        grad = Xc.T @ (w * (yc - Xc @ B[l]))
        for i in args["screen_set"]:
            g, gs, pen = args["groups"][i], group_sizes[i], args["penalty"][i]
            ng, nb = np.linalg.norm(grad[g:g + gs]), np.linalg.norm(B[l, g:g + gs])
            if nb > 0:
                np.testing.assert_allclose(grad[g:g + gs], lmda * pen * B[l, g:g + gs] / nb, atol=2e-5 * max(1.0, lmda))
            else:
                assert ng <= lmda * pen * (1 + 1e-6) + 1e-6
        off = np.setdiff1d(np.arange(len(args["groups"])), args["screen_set"])
        for i in off:
            assert not np.any(B[l, args["groups"][i]:args["groups"][i] + group_sizes[i]])       # pinned at zero


@pytest.mark.parametrize("n, p, G, S", [[10, 4, 2, 2], [10, 100, 10, 2], [10, 100, 20, 13], [100, 23, 4, 3], [100, 100, 50, 20], [3000, 120, 30, 12]])
def test_solve_gaussian_pin_naive(n, p, G, S):
    args, group_sizes = create_data_gaussian_pin(n, p, G, S)
    Xm = ad.matrix.dense(args["X"], method="naive", n_threads=2)
    a = dict(args); a["X"] = Xm; a.pop("y")
    st0 = ad.state.gaussian_pin_naive(**a, tol=1e-7)
    st0.check(method="assert")
    st = st0.solve()
    assert st.error == ""
    st.check(method="assert")
    okw = dict(groups=args["groups"], alpha=args["alpha"], penalty=args["penalty"], weights=args["weights"], screen_set=args["screen_set"],
               lmda_path=args["lmda_path"], tol=1e-7)
    ref = orc.pin_naive_solve(args["X"], args["y"], **okw)
    assert ref.error == "" and len(st.lmdas) == len(ref.lmdas) and st.betas.shape == (len(ref.lmdas), p)
    Br = ref.betas.toarray()
    scale = max(np.max(np.abs(Br)), 1e-12)
    assert np.max(np.abs(st.betas.toarray() - Br)) <= 1e-6 * scale
    np.testing.assert_allclose(st.intercepts, ref.intercepts, atol=1e-6 * max(1.0, np.max(np.abs(ref.intercepts))))
    np.testing.assert_allclose(st.rsqs, ref.rsqs, rtol=1e-6, atol=1e-12)
    assert st.iters == int(ref.iters) and st.active_set_size == int(ref.active_set_size)
    kkt_on_screen_set(args, group_sizes, st)
    # warm start at 0.8 * the last lambda from the solved state (tests/test_solver.py:520-532)
    a2 = dict(a)
    a2.update(lmda_path=[st.lmdas[-1] * 0.8], rsq=st.rsq, resid=st.resid, screen_beta=st.screen_beta, screen_is_active=st.screen_is_active,
              active_set_size=st.active_set_size, active_set=st.active_set)
    w0 = ad.state.gaussian_pin_naive(**a2, tol=1e-7)
    w0.check(method="assert")
    ws = w0.solve()
    assert ws.error == ""
    ws.check(method="assert")
    ref2 = orc.pin_naive_solve(args["X"], args["y"], **dict(okw, lmda_path=np.array([ref.lmdas[-1] * 0.8]), rsq=ref.rsq, resid=ref.resid,
                               screen_beta=ref.screen_beta, screen_is_active=ref.screen_is_active, active_set=ref.active_set,
                               active_set_size=int(ref.active_set_size)))
    assert np.max(np.abs(ws.betas.toarray() - ref2.betas.toarray())) <= 1e-6 * max(np.max(np.abs(ref2.betas.toarray())), 1e-12)
    kkt_on_screen_set(args, group_sizes, ws)


def test_pin_state_errors_and_limits():
    args, _ = create_data_gaussian_pin(100, 100, 50, 20)
    Xm = ad.matrix.dense(args["X"], method="naive")
    a = dict(args); a["X"] = Xm; a.pop("y")
    with pytest.raises(ValueError):
        ad.state.gaussian_pin_naive(**dict(a, X=args["X"]))                 # a raw ndarray is not a matrix object (state.py:581-588)
    st = ad.state.gaussian_pin_naive(**a, max_iters=1).solve()
    assert st.error.startswith("adelie_core solver: max coordinate descents reached at lambda index:")
    st = ad.state.gaussian_pin_naive(**a, max_active_size=1).solve()
    assert "Maximum number of active groups reached" in st.error or st.active_set_size <= 1
    st = ad.state.gaussian_pin_naive(**a, adev_tol=0.0).solve()             # rsq >= 0 * y_var after the first lambda: stops there (:398)
    assert len(st.lmdas) == 1
