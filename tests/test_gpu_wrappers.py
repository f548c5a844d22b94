"""GPU tests of the `matrix.standardize` / `matrix.subset` wrappers (reference MatrixNaiveStandardize / MatrixNaiveCSubset /
MatrixNaiveRSubset): every operator against dense NumPy like the reference's run_naive (T/test_matrix.py:251-411, 662-710, 413-480), the
path solver on the wrapped matrix against the oracle on the NumPy equivalent, and the constructor error contract."""
import numpy as np
import pytest

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _run_naive(cX, X, dtype, atol):
    n, p = X.shape
    assert cX.shape == (n, p)
    rng = np.random.default_rng(n + p)
    v = rng.normal(size=n).astype(dtype); w = rng.uniform(0, 1, size=n).astype(dtype)
    out = np.empty(p, dtype=dtype)
    cX.mul(v, w, out); np.testing.assert_allclose(out, X.T @ (v * w), atol=atol * n)
    cX.sq_mul(w, out); np.testing.assert_allclose(out, (X ** 2).T @ w, atol=atol * n)
    for j, q in [(0, 1), (p // 2, min(4, p - p // 2)), (p - 1, 1)]:
        o = np.empty(q, dtype=dtype); cX.bmul(j, q, v, w, o)
        np.testing.assert_allclose(o, X[:, j:j + q].T @ (v * w), atol=atol * n)
        vv = rng.normal(size=q).astype(dtype); acc = rng.normal(size=n).astype(dtype); exp = acc + X[:, j:j + q] @ vv
        cX.btmul(j, q, vv, acc); np.testing.assert_allclose(acc, exp, atol=atol * 20)
        C = np.empty((q, q), dtype=dtype, order="F"); cX.cov(j, q, np.sqrt(w), C)
        np.testing.assert_allclose(C, X[:, j:j + q].T @ (w[:, None] * X[:, j:j + q]), atol=atol * n)


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-13), (np.float32, 2e-5)])
@pytest.mark.parametrize("n,p,ddof", [(100, 20, 0), (333, 7, 1), (2000, 64, 0)])
def test_standardize_vs_numpy(dtype, atol, n, p, ddof):
    rng = np.random.default_rng(0)
    Z = np.asfortranarray(rng.normal(2.0, 3.0, (n, p)), dtype=dtype)
    cX = ad.matrix.standardize(ad.matrix.dense(Z), ddof=ddof)
    c = Z.mean(axis=0); s = np.sqrt(((Z - c) ** 2).sum(axis=0) / (n - ddof))
    np.testing.assert_allclose(cX._centers, c, rtol=1e-5 if dtype == np.float32 else 1e-12)
    np.testing.assert_allclose(cX._scales, s, rtol=1e-4 if dtype == np.float32 else 1e-12)
    X = ((Z - cX._centers) / cX._scales).astype(dtype)
    _run_naive(cX, X, dtype, atol)
    np.testing.assert_allclose(ad.matrix.standardize(Z, ddof=ddof), (Z - c) / s, rtol=1e-4 if dtype == np.float32 else 1e-12, atol=1e-5)
    # user-given centers / scales
    c2 = rng.normal(size=p).astype(dtype); s2 = rng.uniform(0.5, 2, size=p).astype(dtype)
    _run_naive(ad.matrix.standardize(ad.matrix.dense(Z), centers=c2, scales=s2), ((Z - c2) / s2).astype(dtype), dtype, atol)


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-13), (np.float32, 2e-5)])
def test_subset_vs_numpy(dtype, atol):
    rng = np.random.default_rng(1)
    n, p = 500, 40
    Z = np.asfortranarray(rng.normal(size=(n, p)), dtype=dtype)
    M = ad.matrix.dense(Z)
    rows = rng.choice(n, 123, replace=False); cols = rng.choice(p, 11, replace=False)
    _run_naive(ad.matrix.subset(M, rows, axis=0), Z[rows], dtype, atol)
    _run_naive(ad.matrix.subset(M, cols, axis=1), Z[:, cols], dtype, atol)
    _run_naive(M[rows], Z[rows], dtype, atol)
    _run_naive(M[:, cols], Z[:, cols], dtype, atol)
    _run_naive(M[10:200:3, cols], Z[10:200:3][:, cols], dtype, atol)
    mask = rng.uniform(size=n) < 0.3
    _run_naive(M[mask], Z[mask], dtype, atol)
    assert np.array_equal(ad.matrix.subset(Z, rows, axis=0), Z[rows])


def test_wrapper_errors():
    Z = np.asfortranarray(np.random.default_rng(2).normal(size=(50, 6)))
    M = ad.matrix.dense(Z)
    with pytest.raises(RuntimeError, match="centers must be \\(p,\\)"):
        ad.matrix.standardize(M, centers=np.zeros(5), scales=np.ones(6))
    with pytest.raises(RuntimeError, match="scales must be \\(p,\\)"):
        ad.matrix.standardize(M, centers=np.zeros(6), scales=np.ones(7))
    with pytest.raises(RuntimeError, match="n_threads must be >= 1"):
        ad.matrix.standardize(M, n_threads=0)
    with pytest.raises(RuntimeError, match="subset must be non-empty"):
        ad.matrix.subset(M, np.array([], dtype=int))
    with pytest.raises(RuntimeError, match="unique values in the range \\[0, p\\)"):
        ad.matrix.subset(M, np.array([1, 6]), axis=1)
    with pytest.raises(RuntimeError, match="unique values in the range \\[0, n\\)"):
        ad.matrix.subset(M, np.array([3, 3]), axis=0)
    with pytest.raises(ValueError):
        M[[1, 2], [0, 1]]


def test_grpnet_on_standardized_and_subset_matrix():
    data = ad.data.dense(1500, 60, 12, seed=5)
    Z = data["X"] * 3.0 + 1.5
    y = data["glm"].y
    rows = np.sort(np.random.default_rng(0).choice(1500, 1000, replace=False))
    cX = ad.matrix.standardize(ad.matrix.dense(Z)[rows])
    Xn = ad.matrix.standardize(np.asfortranarray(Z[rows]))
    kw = dict(groups=data["groups"], penalty=data["penalty"], tol=1e-13, early_exit=False, lmda_path_size=12, min_ratio=0.1)
    st = ad.grpnet(cX, ad.glm.gaussian(y[rows]), progress_bar=False, **kw)
    ref = orc.grpnet(Xn, orc.glm_spec("gaussian", y[rows]), **kw)
    assert st.error == "" and ref.error == ""
    B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
    assert np.max(np.abs(B - Br)) <= 1e-6 * np.max(np.abs(Br))
    np.testing.assert_allclose(st.intercepts, ref.intercepts, rtol=1e-6, atol=1e-9)


def _cv_reference(X, y, family, n_folds, seed, min_ratio, L, kw):
    """The reference's CV procedure (adelie/cv.py:247-325) restated on the CPU oracle + NumPy losses (the checker)."""
    n = X.shape[0]
    np.random.seed(seed)
    order = np.random.choice(n, n, replace=False)
    fold_size, remaining = divmod(n, n_folds)
    w = np.full(n, 1 / n)

    def loss(eta, ww):
        if family == "gaussian":
            return np.sum(ww * (0.5 * eta ** 2 - y * eta))
        return np.sum(ww * (np.logaddexp(0, eta) - y * eta))
    lm0 = orc.grpnet(X, orc.glm_spec(family, y, w), lmda_path_size=1, **kw).lmda_max
    full = lm0 * np.logspace(0, np.log10(min_ratio), L)
    out = np.empty((n_folds, L))
    for fold in range(n_folds):
        begin = (fold_size + 1) * min(fold, remaining) + max(fold - remaining, 0) * fold_size
        held = order[begin:begin + fold_size + (fold < remaining)]
        wf = w.copy(); wf[held] = 0; ws = wf.sum(); wf /= ws
        spec = orc.glm_spec(family, y, wf)
        lmf = orc.grpnet(X, spec, lmda_path_size=1, **kw).lmda_max
        cur = lmf * np.logspace(0, np.log10(min_ratio), L)
        aug = np.sort(np.concatenate([full, cur[cur > full[0]]]))[::-1]
        st = orc.grpnet(X, spec, lmda_path=aug, early_exit=False, **kw)
        B = np.asarray(st.betas.todense()); b0 = np.asarray(st.intercepts); lm = np.asarray(st.lmdas)
        for i, l in enumerate(full):
            k = int(np.argmin(np.abs(lm - l)))                 # the common grid is part of the augmented path
            eta = X @ B[k] + b0[k]
            out[fold, i] = (loss(eta, w) - ws * loss(eta, wf)) / w[held].sum()
    return full, out


@pytest.mark.parametrize("family", ["gaussian", "binomial"])
def test_cv_grpnet_vs_oracle_procedure(family):
    data = ad.data.dense(600, 30, 10, glm=family, seed=7)
    X, y = data["X"], data["glm"].y
    kw = dict(groups=data["groups"], penalty=data["penalty"], tol=1e-12)
    if family == "binomial":
        kw["irls_tol"] = 1e-10
    glm = ad.glm.gaussian(y) if family == "gaussian" else ad.glm.binomial(y)
    res = ad.cv_grpnet(X, glm, n_folds=4, seed=3, min_ratio=0.2, lmda_path_size=12, **kw)
    full, ref = _cv_reference(X, y, family, 4, 3, 0.2, 12, kw)
    np.testing.assert_allclose(res.lmdas, full, rtol=1e-9)
    np.testing.assert_allclose(res.losses, ref, rtol=1e-5, atol=1e-8)
    assert res.best_idx == int(np.argmin(ref.mean(axis=0)))
    np.testing.assert_allclose(res.avg_losses, ref.mean(axis=0), rtol=1e-5, atol=1e-8)


def test_cv_grpnet_multigaussian_vs_oracle_procedure():
    """Multi-response CV (adelie/cv.py:92 is_multi): the K-fold procedure on a multigaussian problem against the same procedure on the oracle."""
    rng = np.random.default_rng(4)
    n, p, K, n_folds, L, min_ratio = 500, 20, 3, 4, 10, 0.2
    X = np.asfortranarray(rng.normal(size=(n, p)))
    Bt = np.zeros((p, K)); Bt[:4] = rng.normal(size=(4, K))
    Y = np.ascontiguousarray(X @ Bt + 0.5 + rng.normal(size=(n, K)))
    kw = dict(tol=1e-13)
    res = ad.cv_grpnet(X, ad.glm.multigaussian(Y), n_folds=n_folds, seed=5, min_ratio=min_ratio, lmda_path_size=L, **kw)
    # the procedure restated on the oracle
    np.random.seed(5)
    order = np.random.choice(n, n, replace=False)
    fold_size, remaining = divmod(n, n_folds)
    w = np.full(n, 1 / n)
    loss = lambda eta, ww: np.sum(ww[:, None] * (0.5 * eta ** 2 - Y * eta)) / K          # glm_multigaussian.ipp:17-68
    lm0 = orc.grpnet(X, orc.glm_spec("multigaussian", Y, w), lmda_path_size=1, **kw).lmda_max
    full = lm0 * np.logspace(0, np.log10(min_ratio), L)
    ref = np.empty((n_folds, L))
    for fold in range(n_folds):
        begin = (fold_size + 1) * min(fold, remaining) + max(fold - remaining, 0) * fold_size
        held = order[begin:begin + fold_size + (fold < remaining)]
        wf = w.copy(); wf[held] = 0; ws = wf.sum(); wf /= ws
        spec = orc.glm_spec("multigaussian", Y, wf)
        lmf = orc.grpnet(X, spec, lmda_path_size=1, **kw).lmda_max
        cur = lmf * np.logspace(0, np.log10(min_ratio), L)
        aug = np.sort(np.concatenate([full, cur[cur > full[0]]]))[::-1]
        st = orc.grpnet(X, spec, lmda_path=aug, early_exit=False, **kw)
        B = np.asarray(st.betas.todense()); b0 = np.asarray(st.intercepts); lm = np.asarray(st.lmdas)
        for i, l in enumerate(full):
            k = int(np.argmin(np.abs(lm - l)))
            eta = X @ B[k].reshape(p, K) + b0[k][None]
            ref[fold, i] = (loss(eta, w) - ws * loss(eta, wf)) / w[held].sum()
    np.testing.assert_allclose(res.lmdas, full, rtol=1e-9)
    np.testing.assert_allclose(res.losses, ref, rtol=1e-5, atol=1e-8)
    assert res.best_idx == int(np.argmin(ref.mean(axis=0)))
    # sklearn front-end: multinomial score uses the last lambda's prediction
    labels = rng.integers(0, K, size=n); onehot = np.eye(K)[labels]
    est = ad.sklearn.GroupElasticNet(solver="grpnet", family="multinomial").fit(X, onehot, lmda_path_size=5, min_ratio=0.5)
    sc = est.score(X, onehot)
    last = np.argmax(np.asarray(ad.diagnostic.predict(X, est.coef_, est.intercept_))[-1], axis=-1)
    assert sc == float(np.mean(last == labels))


@pytest.mark.parametrize("dtype,atol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_standardize_snp_unphased_view(dtype, atol):
    """standardize(snp_unphased): a view on the packed genotypes (nothing materialised), every operator and the path vs the NumPy equivalent"""
    from oracle import snp_oracle as so
    data = ad.data.snp_unphased(1200, 45, seed=4, sparsity=0.8)
    cd = data["X"]
    imp = np.sum(np.where(cd > 0, cd, 0), axis=0) / np.maximum(np.sum(cd >= 0, axis=0), 1)
    base = ad.matrix.snp_unphased_from_calldata(cd, imp, dtype=dtype)
    D = so.dense_equivalent(cd, imp, dtype)
    c = D.mean(axis=0).astype(dtype); s = D.std(axis=0).astype(dtype)
    cX = ad.matrix.standardize(base, centers=c, scales=s)
    X = ((D - c) / s).astype(dtype)
    _run_naive(cX, X, dtype, atol)
    assert base.cache_info()[0] == 0                      # the view has its own decoded-column cache; the base decoded nothing
    # default centers / scales follow the reference's snp_unphased quirk: mean() == 0, var() == 1  ->  identity
    ident = ad.matrix.standardize(base)
    np.testing.assert_array_equal(ident._centers, 0); np.testing.assert_array_equal(ident._scales, 1)
    with pytest.raises(RuntimeError, match="already a standardized"):
        ad.matrix.standardize(cX, centers=c, scales=s)
    if dtype == np.float64:
        y = data["glm"].y
        kw = dict(tol=1e-13, early_exit=False, lmda_path_size=10, min_ratio=0.1)
        st = ad.grpnet(cX, ad.glm.gaussian(y), progress_bar=False, **kw)
        ref = orc.grpnet(X, orc.glm_spec("gaussian", y), **kw)
        assert st.error == "" and ref.error == ""
        B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
        assert np.max(np.abs(B - Br)) <= 1e-6 * np.max(np.abs(Br))


def test_sklearn_group_elastic_net():
    """adelie/sklearn.py GroupElasticNet over grpnet / cv_grpnet on the device."""
    from adelie_b200.sklearn import GroupElasticNet
    data = ad.data.dense(500, 20, 5, seed=1)
    X, y = data["X"], data["glm"].y
    m = GroupElasticNet().fit(X, y, groups=data["groups"], lmda_path_size=10, min_ratio=0.1, early_exit=False)
    ref = orc.grpnet(X, orc.glm_spec("gaussian", y), groups=data["groups"], lmda_path_size=10, min_ratio=0.1, early_exit=False)
    np.testing.assert_allclose(np.asarray(m.coef_.todense()), np.asarray(ref.betas.todense()), rtol=1e-4, atol=1e-6)
    pred = m.predict(X)
    np.testing.assert_allclose(pred, np.asarray(ref.betas.todense()) @ X.T + np.asarray(ref.intercepts)[:, None], rtol=1e-4, atol=1e-5)
    assert 0.0 < m.score(X, y) <= 1.0
    mc = GroupElasticNet(solver="cv_grpnet").fit(X, y, groups=data["groups"], n_folds=3, seed=0, lmda_path_size=8, min_ratio=0.2)
    assert mc.coef_.shape == (1, 20) and mc.lambda_.shape == (1,) and mc.predict(X).shape == (500,)
    db = ad.data.dense(400, 15, 15, glm="binomial", seed=2)
    mb = GroupElasticNet(family="binomial").fit(db["X"], db["glm"].y, lmda_path_size=6, min_ratio=0.3, early_exit=False)
    proba = mb.predict_proba(db["X"])
    assert proba.shape == (6, 400, 2) and np.allclose(proba.sum(axis=-1), 1)
    assert set(np.unique(mb.predict(db["X"]))) <= {0, 1}
    rng = np.random.default_rng(0)
    Y = np.eye(3)[rng.integers(0, 3, 300)]
    mm = GroupElasticNet(family="multinomial").fit(db["X"][:300], Y, lmda_path_size=4, min_ratio=0.5, early_exit=False)
    pm = mm.predict_proba(db["X"][:300])
    assert pm.shape == (4, 300, 3) and np.allclose(pm.sum(axis=-1), 1)
    with pytest.raises(RuntimeError, match="not been fitted"):
        GroupElasticNet().predict(X)


def test_diagnostic_objective_matches_numpy():
    """adelie/diagnostic.py:124-276: loss on the device + penalty, and the reference's acceptance rule objective(ours) <= (1 + eps) objective(oracle)."""
    from adelie_b200.diagnostic import objective
    data = ad.data.dense(400, 24, 6, seed=3)
    X, y = data["X"], data["glm"].y
    kw = dict(groups=data["groups"], penalty=data["penalty"], alpha=0.6, lmda_path_size=8, min_ratio=0.2, early_exit=False)
    st = ad.grpnet(X, ad.glm.gaussian(y), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("gaussian", y), **kw)
    okw = dict(groups=data["groups"], alpha=0.6, penalty=data["penalty"])
    o1 = objective(X, ad.glm.gaussian(y), st.betas, st.intercepts, st.lmdas, **okw)
    o2 = objective(X, ad.glm.gaussian(y), ref.betas, np.asarray(ref.intercepts), np.asarray(ref.lmdas), **okw)
    w = np.full(400, 1 / 400)
    B = np.asarray(st.betas.todense())
    exp = np.array([0.5 * np.sum(w * (y - X @ B[l] - st.intercepts[l]) ** 2) for l in range(B.shape[0])])
    gs = data["group_sizes"]
    pen = np.array([sum(pk * (0.6 * np.linalg.norm(b[g:g + s_]) + 0.2 * np.linalg.norm(b[g:g + s_]) ** 2) for g, s_, pk in zip(data["groups"], gs, data["penalty"])) for b in B])
    np.testing.assert_allclose(o1, exp + st.lmdas * pen, rtol=1e-9, atol=1e-12)
    assert np.all(o1 <= o2 * (1 + 1e-6) + 1e-12)
