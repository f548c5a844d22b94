"""GPU parity tests of the path solver: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Tolerances: 1e-6 relative for float64, 1e-4 relative for float32 (north_star), with
both sides run at tight convergence tolerance (see SURVEY.md section 7 'Parity definition')."""
import numpy as np
import pytest

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    scale = np.max(np.abs(b)) if np.size(b) else 0.0
    return np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0)


def _compare_paths(st, ref, rtol):
    assert st.error == "", st.error
    assert ref.error == "", ref.error
    L = min(len(st.lmdas), len(ref.lmdas))
    assert L == len(ref.lmdas) == len(st.lmdas)
    np.testing.assert_allclose(st.lmdas, ref.lmdas, rtol=rtol)
    B = np.asarray(st.betas.todense()); Br = np.asarray(ref.betas.todense())
    assert _rel(B, Br) <= rtol, _rel(B, Br)
    assert _rel(np.asarray(st.intercepts), np.asarray(ref.intercepts)) <= rtol
    np.testing.assert_allclose(st.devs, ref.devs, rtol=10 * rtol, atol=10 * rtol)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-6), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,p,G,alpha,intercept", [
    (1000, 500, 500, 1.0, True),        # config 1: Gaussian lasso n=1000 p=500
    (300, 120, 25, 1.0, True),          # random unequal groups
    (300, 120, 25, 0.5, False),
    (5000, 60, 12, 0.8, True),
])
def test_gaussian_path_vs_oracle(dtype, rtol, n, p, G, alpha, intercept):
    data = ad.data.dense(n, p, G, seed=3)
    X = np.asfortranarray(data["X"], dtype=dtype)
    y = data["glm"].y.astype(dtype)
    tol = 1e-12 if dtype == np.float64 else 1e-7
    # float32 cannot resolve |phi(h)| <= 1e-12 (the reference's newton_tol default): its own templates would hit
    # newton_max_iters; use a float-resolvable tolerance on both sides.
    newton_tol = 1e-12 if dtype == np.float64 else 1e-5
    kw = dict(groups=data["groups"], alpha=alpha, penalty=data["penalty"].astype(dtype), intercept=intercept, tol=tol,
              early_exit=False, lmda_path_size=30, min_ratio=0.05, newton_tol=newton_tol)
    st = ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("gaussian", y, dtype=dtype), **kw)
    _compare_paths(st, ref, rtol)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-6), (np.float32, 1e-4)])
@pytest.mark.parametrize("n,p,G,alpha", [(400, 60, 60, 1.0), (600, 90, 18, 0.5)])
def test_binomial_path_vs_oracle(dtype, rtol, n, p, G, alpha):
    data = ad.data.dense(n, p, G, glm="binomial", seed=5)
    X = np.asfortranarray(data["X"], dtype=dtype)
    y = data["glm"].y.astype(dtype)
    # the sweep stops when max_g sum(A dbeta^2) / gs < tol: coefficients are only accurate to ~sqrt(tol).  For a 1e-4 comparison of two
    # float32 implementations both sides converge to 1e-10 / 1e-9 (float32 resolves that), so that rounding, not the stopping rule, is compared
    tol = 1e-12 if dtype == np.float64 else 1e-10
    newton_tol = 1e-12 if dtype == np.float64 else 1e-6
    irls_tol = 1e-10 if dtype == np.float64 else 1e-9
    kw = dict(groups=data["groups"], alpha=alpha, penalty=data["penalty"].astype(dtype), tol=tol, irls_tol=irls_tol,
              early_exit=False, lmda_path_size=20, min_ratio=0.1, newton_tol=newton_tol)
    st = ad.grpnet(X, ad.glm.binomial(y, dtype=dtype), progress_bar=False, **kw)
    ref = orc.grpnet(X, orc.glm_spec("binomial", y, dtype=dtype), **kw)
    _compare_paths(st, ref, rtol)


def test_gaussian_glm_irls_matches_opt():
    """gaussian(opt=False) goes through the IRLS driver and must reproduce the optimized Gaussian path."""
    data = ad.data.dense(500, 80, 20, seed=7)
    X = data["X"]; y = data["glm"].y
    kw = dict(groups=data["groups"], penalty=data["penalty"], tol=1e-12, irls_tol=1e-12, early_exit=False, lmda_path_size=15,
              min_ratio=0.1, progress_bar=False)
    a = ad.grpnet(X, ad.glm.gaussian(y, opt=True), **kw)
    b = ad.grpnet(X, ad.glm.gaussian(y, opt=False), **kw)
    assert a.error == "" and b.error == ""
    np.testing.assert_allclose(a.lmdas, b.lmdas, rtol=1e-8)
    assert _rel(np.asarray(b.betas.todense()), np.asarray(a.betas.todense())) < 1e-6
    np.testing.assert_allclose(a.intercepts, b.intercepts, rtol=1e-6, atol=1e-9)


def test_glm_fused_means_equals_separate_pass():
    """The IRLS-weighted column means of the screen groups out of the diagonal-Gram pass (Configs::glm_fuse_means, default) give the path of
    the separate GEMV pass (solver_glm_naive.hpp:361-372 computes them per IRLS iteration), dense and snp_unphased, float64 and float32."""
    import adelie_b200 as ad
    rng = np.random.default_rng(21)
    n, p = 3000, 60
    X = np.asfortranarray(rng.standard_normal((n, p)) + 0.5)
    beta = np.zeros(p); beta[:6] = rng.standard_normal(6)
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(X @ beta - 1.0)))).astype(np.float64)
    w = rng.uniform(1, 2, n); w /= w.sum()
    for dtype, rtol in ((np.float64, 1e-9), (np.float32, 1e-4)):
        kw = dict(groups=np.arange(0, p, 5), alpha=0.6, tol=1e-10 if dtype == np.float64 else 1e-7, irls_tol=1e-10 if dtype == np.float64 else 1e-7,
                  newton_tol=1e-12 if dtype == np.float64 else 1e-6, lmda_path_size=12, min_ratio=0.05, early_exit=False, progress_bar=False)
        Xd = np.asfortranarray(X, dtype=dtype)
        out = []
        for fuse in (1, 0):
            ad.configs.set_configs("glm_fuse_means", fuse)
            try:
                st = ad.grpnet(Xd, ad.glm.binomial(y.astype(dtype), weights=w.astype(dtype)), **kw)
            finally:
                ad.configs.set_configs("glm_fuse_means", None)
            assert st.error == ""
            out.append((st.betas.toarray(), st.intercepts, st.devs))
        scale = np.max(np.abs(out[1][0])) + 1e-30
        assert np.max(np.abs(out[0][0] - out[1][0])) / scale <= rtol
        np.testing.assert_allclose(out[0][1], out[1][1], rtol=10 * rtol, atol=10 * rtol)
        np.testing.assert_allclose(out[0][2], out[1][2], rtol=10 * rtol, atol=10 * rtol)
