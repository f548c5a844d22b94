"""GPU: state / API behaviour of the drop-in surface (reference tests/test_state.py:55,179; tests/test_solver.py:633-649
warm start; adelie/state.py:157-176 solve contract)."""
import numpy as np
import pytest

import adelie_b200 as ad

pytestmark = pytest.mark.gpu


def _data(seed=0, n=400, p=50, G=10):
    d = ad.data.dense(n, p, G, seed=seed)
    return d["X"], d["glm"].y, d["groups"], d["penalty"]


def test_state_roundtrip_and_surface():
    X, y, groups, penalty = _data()
    Xm = ad.matrix.dense(X)
    st = ad.grpnet(Xm, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, lmda_path_size=12, min_ratio=0.1)
    assert id(st.X) == id(Xm)                                   # tests/test_state.py:55
    assert st.error == "" and st.total_time > 0
    L = len(st.lmdas)
    assert st.betas.shape == (L, X.shape[1]) and st.betas.indices.dtype == np.int64
    assert st.intercepts.shape == (L,) and st.devs.shape == (L,)
    assert np.array_equal(st.groups, groups) and st.alpha == 1 and st.intercept is True
    assert st.screen_set.dtype == np.int64 and st.screen_is_active.dtype == bool
    assert len(st.screen_begins) == len(st.screen_set) == len(st.screen_transforms)
    assert st.screen_beta.shape[0] == st.screen_vars.shape[0] == st.screen_X_means.shape[0]
    assert st.active_set_size <= len(st.screen_set) and st.resid.shape == (X.shape[0],) and st.grad.shape == (X.shape[1],)
    assert len(st.benchmark_fit_active) == len(st.n_valid_solutions) >= L - 1
    # invariants of the final state (adelie/state.py:1421-1674): resid, grad, rsq
    b = np.asarray(st.betas[-1].todense()).ravel()
    w = np.full(X.shape[0], 1 / X.shape[0])
    yc = y - np.sum(w * y)
    np.testing.assert_allclose(st.resid, yc - X @ b, atol=1e-8)
    Xc = X - (X.T @ w)[None]
    np.testing.assert_allclose(st.grad, Xc.T @ (w * st.resid), atol=1e-8)
    assert abs(st.rsq - (np.sum(w * yc ** 2) - np.sum(w * (st.resid - np.sum(w * st.resid)) ** 2))) < 1e-8
    # eigen-decomposition of each screen block: V diag(A) V^T == centred weighted Gram
    for i, g in enumerate(st.screen_set):
        j, q = groups[g], st.group_sizes[g]
        V = st.screen_transforms[i]; A = st.screen_vars[st.screen_begins[i]:st.screen_begins[i] + q]
        np.testing.assert_allclose(V @ np.diag(A) @ V.T, Xc[:, j:j + q].T @ (w[:, None] * Xc[:, j:j + q]), atol=1e-9)


def test_solve_does_not_mutate_input_state_and_warm_start():
    X, y, groups, penalty = _data(seed=1)
    kw = dict(groups=groups, penalty=penalty, progress_bar=False, tol=1e-12, early_exit=False)
    full = ad.grpnet(X, ad.glm.gaussian(y), lmda_path_size=20, min_ratio=0.05, **kw)
    half = ad.grpnet(X, ad.glm.gaussian(y), lmda_path=full.lmdas[:10], **kw)
    np.testing.assert_allclose(np.asarray(half.betas.todense()), np.asarray(full.betas[:10].todense()), atol=1e-8)
    n_before = len(half.lmdas)
    rest = ad.grpnet(X, ad.glm.gaussian(y), lmda_path=full.lmdas[10:], warm_start=half, **kw)   # tests/test_solver.py:633-649
    assert len(half.lmdas) == n_before                          # the warm-start state itself is untouched
    np.testing.assert_allclose(np.asarray(rest.betas.todense()), np.asarray(full.betas[10:].todense()), atol=1e-7)
    np.testing.assert_allclose(rest.intercepts, full.intercepts[10:], atol=1e-7)


def test_solver_errors_are_returned_not_raised():
    X, y, groups, penalty = _data(seed=2)
    st = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, max_iters=3, early_exit=False)
    assert st.error.startswith("adelie_core solver: max coordinate descents reached")
    assert len(st.lmdas) < 100                                   # valid up to the last solved lambda
    st = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, max_screen_size=2, early_exit=False)
    assert "maximum screen set size reached" in st.error
    st = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, max_active_size=1, early_exit=False)
    assert "Maximum number of active groups reached" in st.error
    with pytest.raises(RuntimeError, match="alpha must be in"):  # ctor validation errors DO propagate
        ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, alpha=1.5, progress_bar=False)


def test_early_exit_and_exit_cond():
    X, y, groups, penalty = _data(seed=3)
    st = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, adev_tol=0.3)
    assert st.devs[-1] >= 0.3 and (len(st.devs) == 1 or st.devs[-2] < 0.3)
    st = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, early_exit=False,
                   exit_cond=lambda s: len(s.lmdas) >= 5)
    assert len(st.lmdas) == 5


def test_strong_rule_and_user_path_match_pivot():
    X, y, groups, penalty = _data(seed=4)
    kw = dict(groups=groups, penalty=penalty, progress_bar=False, tol=1e-12, early_exit=False, lmda_path_size=15, min_ratio=0.1)
    a = ad.grpnet(X, ad.glm.gaussian(y), screen_rule="pivot", **kw)
    b = ad.grpnet(X, ad.glm.gaussian(y), screen_rule="strong", **kw)
    np.testing.assert_allclose(np.asarray(a.betas.todense()), np.asarray(b.betas.todense()), atol=1e-8)
    big = np.concatenate([[a.lmda_max * 2, a.lmda_max * 1.5], a.lmdas[:5]])
    c = ad.grpnet(X, ad.glm.gaussian(y), groups=groups, penalty=penalty, progress_bar=False, tol=1e-12, early_exit=False, lmda_path=big)
    assert len(c.lmdas) == 7 and c.betas[:2].nnz == 0           # lambdas above lmda_max give the null model
    np.testing.assert_allclose(np.asarray(c.betas[2:].todense()), np.asarray(a.betas[:5].todense()), atol=1e-8)


def test_property_path_at_scale():
    """Size-independent properties at a size the oracle would take long on: deviance is monotone along the path, the KKT
    conditions hold for every solved lambda (float32, multi-CTA staged kernel)."""
    n, p, gs = 60_000, 600, 6
    X = ad.matrix.dense_device_normal(n, p, dtype=np.float32, seed=5)
    rng = np.random.default_rng(0)
    beta = np.zeros(p, dtype=np.float32); beta[rng.choice(p, 30, replace=False)] = rng.normal(size=30)
    y = (X @ beta + np.linalg.norm(beta) * rng.normal(size=n)).astype(np.float32)
    groups = np.arange(0, p, gs)
    st = ad.grpnet(X, ad.glm.gaussian(y, dtype=np.float32), groups=groups, progress_bar=False, early_exit=False, lmda_path_size=25,
                   min_ratio=0.05, newton_tol=1e-6)
    assert st.error == "" and len(st.lmdas) == 25 and st.sweep_ncta > 1
    assert np.all(np.diff(st.devs) > -1e-5)
    Xh = X.to_host().astype(np.float64)
    w = 1.0 / n
    for l in (5, 15, 24):
        b = np.asarray(st.betas[l].todense()).ravel().astype(np.float64)
        r = y - Xh @ b - st.intercepts[l]
        g = Xh.T @ (w * r)
        for j in groups:
            gn = np.linalg.norm(g[j:j + gs]); bn = np.linalg.norm(b[j:j + gs])
            lam_p = st.lmdas[l] * np.sqrt(gs)
            if bn > 0:
                assert np.max(np.abs(g[j:j + gs] - lam_p * b[j:j + gs] / bn)) < 2e-3 * lam_p
            else:
                assert gn <= lam_p * (1 + 1e-3)


def _free_hbm():
    import ctypes as C
    from adelie_b200 import _lib
    fr, tot = C.c_size_t(), C.c_size_t()
    _lib.check(_lib.load().ab_device_synchronize())
    _lib.check(_lib.load().ab_mem_info(C.byref(fr), C.byref(tot)))
    return fr.value


def test_no_hbm_leak_over_repeated_grpnet_calls():
    """30 x grpnet(ndarray): every call uploads its own 160 MB copy of X; free HBM must stay flat (round-1 leak: the copies were only
    released by the cyclic GC).  Automatic garbage collection is switched off to make the test deterministic."""
    import gc
    rng = np.random.default_rng(0)
    n, p = 20000, 2000
    X = np.asfortranarray(rng.standard_normal((n, p), dtype=np.float32))
    y = (X[:, :5] @ np.ones(5, dtype=np.float32) + rng.standard_normal(n, dtype=np.float32)).astype(np.float32)
    kw = dict(groups=np.arange(0, p, 10), progress_bar=False, lmda_path_size=5, min_ratio=0.5, newton_tol=1e-6)
    ad.grpnet(X, ad.glm.gaussian(y, dtype=np.float32), **kw)            # warm the memory pool
    gc.collect(); gc.disable()
    try:
        free0 = _free_hbm()
        for _ in range(30):
            st = ad.grpnet(X, ad.glm.gaussian(y, dtype=np.float32), **kw)
            assert st.error == ""
        del st
        free1 = _free_hbm()
    finally:
        gc.enable()
    assert free0 - free1 < (1 << 30), f"HBM leak: {(free0 - free1) / 2**20:.0f} MiB over 30 solves"


def test_solve_moves_the_core_state_and_the_input_state_stays_usable():
    X, y, groups, penalty = _data()
    kw = dict(groups=groups, penalty=penalty, progress_bar=False, lmda_path_size=6, min_ratio=0.3)
    ref = ad.grpnet(X, ad.glm.gaussian(y), **kw)
    # an unsolved state: solve() works on a copy; the pristine core is moved into the copy, not built twice ...
    w = np.full(X.shape[0], 1 / X.shape[0]); Xm = ad.matrix.dense(X)
    yc = y - np.sum(w * y); grad = np.empty(X.shape[1]); Xm.mul(yc, w, grad); xm = np.empty(X.shape[1]); Xm.mul(np.ones_like(yc), w, xm)
    G = len(groups)
    st0 = ad.state.gaussian_naive(X=Xm, y=y, X_means=xm, y_mean=np.sum(w * y), y_var=np.sum(w * yc ** 2), resid=yc, resid_sum=np.sum(w * yc),
                                  constraints=None, groups=groups, group_sizes=np.diff(np.append(groups, X.shape[1])), alpha=1, penalty=penalty,
                                  weights=w, offsets=np.zeros_like(y), screen_set=np.zeros(0, dtype=int), screen_beta=np.zeros(0),
                                  screen_is_active=np.zeros(0, dtype=bool), active_set_size=0, active_set=np.zeros(G, dtype=int), rsq=0,
                                  lmda=np.inf, grad=grad, lmda_path_size=6, min_ratio=0.3)
    assert st0._handle is not None
    st1 = st0.solve(progress_bar=False)
    assert st1.error == "" and st0._handle is None
    np.testing.assert_allclose(st1.betas.toarray(), ref.betas.toarray(), atol=1e-12)
    assert len(st0.lmdas) == 0 and st0._handle is not None        # ... and rebuilt lazily from the stored inputs when read again
    # a solved state keeps its results when it is solved again
    st2 = st1.solve(progress_bar=False)
    assert len(st1.lmdas) == len(st2.lmdas) == len(ref.lmdas)
    st1.close()
    with pytest.raises(RuntimeError, match="closed"):
        st1.betas
