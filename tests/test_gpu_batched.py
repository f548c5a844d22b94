"""GPU parity tests of the batched look-ahead sweep kernel (csrc/sweep_batched.cuh) through the C ABI:
  * against the per-group kernel (``sweep_batch = 1``), which follows the reference's update order literally;
  * against the CPU oracle (1e-6 rel float64 / 1e-4 rel float32, north_star);
covering every code path of the control warp: lasso (gs = 1), register-resident prox (2 <= gs <= 12), shared-memory prox
(12 < gs <= 32), mixed sizes in one batch, partial batches, the relaunch when the active list outgrows its Gram panels,
elastic net, no intercept, one CTA and many CTAs, every batch size."""
import numpy as np
import pytest

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    s = np.max(np.abs(b)) if np.size(b) else 0.0
    return np.max(np.abs(a - b)) / (s if s > 0 else 1.0)


@pytest.fixture(autouse=True)
def _restore_configs():
    yield
    ad.set_configs("sweep_batch", None)
    ad.set_configs("sweep_ctas", None)


def _problem(n, p, groups, dtype, seed=3):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(size=(n, p)), dtype=dtype)
    beta = np.zeros(p)
    supp = rng.choice(p, max(1, p // 8), replace=False)
    beta[supp] = rng.normal(size=supp.size)
    y = (X @ beta + np.linalg.norm(beta) * rng.normal(size=n)).astype(dtype)
    gsz = np.diff(np.concatenate([groups, [p]]))
    return X, y, np.sqrt(gsz).astype(dtype)


def _solve(X, y, groups, penalty, dtype, *, batch, ctas, alpha=1.0, intercept=True, L=16, min_ratio=0.05, use_oracle=False):
    tol = 1e-12 if dtype == np.float64 else 1e-7
    nt = 1e-12 if dtype == np.float64 else 1e-5
    kw = dict(groups=groups, alpha=alpha, penalty=penalty, intercept=intercept, tol=tol, early_exit=False, lmda_path_size=L,
              min_ratio=min_ratio, newton_tol=nt)
    if use_oracle:
        return orc.grpnet(X, orc.glm_spec("gaussian", y, dtype=dtype), **kw)
    ad.set_configs("sweep_batch", batch)
    ad.set_configs("sweep_ctas", ctas)
    return ad.grpnet(X, ad.glm.gaussian(y, dtype=dtype), progress_bar=False, **kw)


def _groups_from_sizes(sizes):
    return np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(int)


CASES = [
    # name, rows per CTA (kept small enough that the tiles of the widest group fit the shared-memory ring in float64), group sizes
    ("lasso", 1500, [1] * 80),
    ("gs10", 1024, [10] * 14),
    ("small_mixed", 1024, [1, 3, 8, 2, 12, 5, 1, 1, 7, 4, 9, 12, 6, 2, 10, 11]),
    ("gs20", 256, [20] * 8),
    ("big_mixed", 64, [1, 20, 5, 32, 13, 2, 16, 10, 25, 1, 12, 18]),
]


@pytest.mark.parametrize("dtype,rtol_kernel,rtol_oracle", [(np.float64, 1e-9, 1e-6), (np.float32, 5e-5, 1e-4)])
@pytest.mark.parametrize("name,n,sizes", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("ctas", [1, 4])
def test_batched_kernel_matches_per_group_kernel_and_oracle(dtype, rtol_kernel, rtol_oracle, name, n, sizes, ctas):
    groups = _groups_from_sizes(sizes); p = int(np.sum(sizes))
    X, y, pen = _problem(n * ctas, p, groups, dtype)
    ref = _solve(X, y, groups, pen, dtype, batch=1, ctas=ctas)
    st = _solve(X, y, groups, pen, dtype, batch=0, ctas=ctas)
    assert ref.error == "" and st.error == "", (ref.error, st.error)
    assert ref.sweep_batch == 1 and ref.n_batched_launches == 0
    assert st.sweep_batch > 1 and st.n_batched_launches >= st.n_pin_solves, "the batched kernel was not the one that ran"
    B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
    assert _rel(B, Br) <= rtol_kernel, _rel(B, Br)
    np.testing.assert_allclose(st.lmdas, ref.lmdas, rtol=1e-6)
    assert abs(st.n_sweeps - ref.n_sweeps) <= max(2, 0.02 * ref.n_sweeps)
    o = _solve(X, y, groups, pen, dtype, batch=0, ctas=ctas, use_oracle=True)
    assert o.error == ""
    assert _rel(B, np.asarray(o.betas.todense())) <= rtol_oracle
    assert _rel(np.asarray(st.intercepts), np.asarray(o.intercepts)) <= rtol_oracle
    np.testing.assert_allclose(st.devs, o.devs, rtol=10 * rtol_oracle, atol=10 * rtol_oracle)


@pytest.mark.parametrize("batch", [2, 3, 4, 5, 6, 8])
def test_every_batch_size(batch):
    sizes = [10] * 23                     # 23 groups: partial last batch for every batch size
    groups = _groups_from_sizes(sizes); p = int(np.sum(sizes))
    X, y, pen = _problem(1800, p, groups, np.float64, seed=11)
    ref = _solve(X, y, groups, pen, np.float64, batch=1, ctas=3)
    st = _solve(X, y, groups, pen, np.float64, batch=batch, ctas=3)
    assert st.error == "" and 2 <= st.sweep_batch <= min(batch, 6)      # (the planner may shrink the batch to fit shared memory)
    assert _rel(np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())) <= 1e-9
    assert st.n_sweeps == ref.n_sweeps


@pytest.mark.parametrize("alpha,intercept", [(0.5, True), (1.0, False), (0.2, False)])
def test_elastic_net_and_no_intercept(alpha, intercept):
    sizes = [4, 10, 10, 1, 7, 10, 3, 10, 10, 2, 9, 10]
    groups = _groups_from_sizes(sizes); p = int(np.sum(sizes))
    X, y, pen = _problem(2400, p, groups, np.float64, seed=5)
    st = _solve(X, y, groups, pen, np.float64, batch=0, ctas=4, alpha=alpha, intercept=intercept)
    o = _solve(X, y, groups, pen, np.float64, batch=0, ctas=4, alpha=alpha, intercept=intercept, use_oracle=True)
    assert st.error == "" and o.error == "" and st.sweep_batch > 1
    assert _rel(np.asarray(st.betas.todense()), np.asarray(o.betas.todense())) <= 1e-6
    assert _rel(np.asarray(st.intercepts), np.asarray(o.intercepts)) <= 1e-6


def test_full_grid_kkt_and_residual_invariants_fp32():
    """Many CTAs (one per SM), float32, a size where every warp role is busy: the converged path satisfies the KKT conditions of
    the group lasso and the maintained residual equals y - X beta - intercept."""
    n, gs, G = 150_000, 10, 48
    sizes = [gs] * G
    groups = _groups_from_sizes(sizes); p = gs * G
    X, y, pen = _problem(n, p, groups, np.float32, seed=9)
    st = _solve(X, y, groups, pen, np.float32, batch=0, ctas=0, L=12, min_ratio=0.1)
    assert st.error == "" and st.sweep_batch > 1 and st.sweep_ncta > 100
    w = np.full(n, 1.0 / n)
    Xd = X.astype(np.float64); yd = y.astype(np.float64)
    B = np.asarray(st.betas.todense(), dtype=np.float64)
    for li in (len(st.lmdas) // 2, len(st.lmdas) - 1):
        b = B[li]; lam = float(st.lmdas[li])
        r = yd - Xd @ b - float(st.intercepts[li])
        g = Xd.T @ (w * r)
        for k in range(G):
            gk = g[groups[k]:groups[k] + gs]; bk = b[groups[k]:groups[k] + gs]
            if np.any(bk != 0):
                np.testing.assert_allclose(gk, lam * pen[k] * bk / np.linalg.norm(bk), atol=2e-4 * lam * pen[k] + 1e-6)
            else:
                assert np.linalg.norm(gk) <= lam * pen[k] * (1 + 1e-3) + 1e-6
    r_dev = np.asarray(st.resid, dtype=np.float64)
    r_ref = yd - (yd @ w) - Xd @ B[-1]                        # the kernel's residual is centred by y_mean, intercept kept apart
    np.testing.assert_allclose(r_dev, r_ref, atol=5e-4 * np.max(np.abs(r_ref)))


# ------------------------------------------------------------------------------------------------------------------------------
# GLM / IRLS on the batched kernel: every Gram panel is rebuilt for the weights of each IRLS iteration (tensor-core panel kernel
# for float32, TF32 operands).  The fixed point of the sweep does not depend on the panels' precision: paths must match the per-group
# kernel (glm_batched = 0) and the oracle at the same tolerances as everything else.
# ------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def _restore_glm_configs():
    yield
    for k in ("glm_batched", "panel_tc", "sweep_ctas", "sweep_batch"):
        ad.set_configs(k, None)


def _glm_problem(n, p, gs, dtype, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(size=(n, p)), dtype=dtype)
    beta = np.zeros(p); supp = rng.choice(p, max(1, p // 8), replace=False); beta[supp] = rng.normal(size=supp.size)
    eta = X @ beta / np.linalg.norm(beta)
    y = (rng.uniform(size=n) < 1 / (1 + np.exp(-eta))).astype(dtype)
    groups = np.arange(0, p, gs)
    return X, y, groups


@pytest.mark.parametrize("dtype,mode,tc,tol_k,tol_o", [(np.float64, 2, 0, 1e-8, 1e-6), (np.float32, 1, 1, 1e-4, 1e-4), (np.float32, 1, 0, 1e-4, 1e-4)])
@pytest.mark.parametrize("n,p,gs,ctas", [(4096, 120, 10, 4), (3000, 64, 1, 2), (4000, 96, 6, 2)])
def test_irls_on_the_batched_kernel_matches_per_group_kernel_and_oracle(_restore_glm_configs, dtype, mode, tc, tol_k, tol_o, n, p, gs, ctas):
    X, y, groups = _glm_problem(n, p, gs, dtype, seed=n + gs)
    f64 = dtype == np.float64
    kw = dict(groups=groups, alpha=0.5, tol=1e-12 if f64 else 1e-7, irls_tol=1e-10 if f64 else 1e-7, newton_tol=1e-12 if f64 else 1e-6,
              early_exit=False, lmda_path_size=10, min_ratio=0.1)
    ad.set_configs("sweep_ctas", ctas)
    ad.set_configs("glm_batched", 0)
    ref = ad.grpnet(X, ad.glm.binomial(y, dtype=dtype), progress_bar=False, **kw)
    ad.set_configs("glm_batched", mode); ad.set_configs("panel_tc", tc)
    st = ad.grpnet(X, ad.glm.binomial(y, dtype=dtype), progress_bar=False, **kw)
    assert ref.error == "" and st.error == "", (ref.error, st.error)
    assert ref.n_batched_launches == 0 and st.n_batched_launches >= st.n_irls > 0, "IRLS did not run on the batched kernel"
    B, Br = st.betas.toarray(), ref.betas.toarray()
    assert _rel(B, Br) <= tol_k, _rel(B, Br)
    o = orc.grpnet(X, orc.glm_spec("binomial", y, dtype=dtype), **kw)
    assert o.error == ""
    assert _rel(B, o.betas.toarray()) <= tol_o, _rel(B, o.betas.toarray())
    assert _rel(np.asarray(st.intercepts), np.asarray(o.intercepts)) <= tol_o
    np.testing.assert_allclose(st.devs, o.devs, rtol=10 * tol_o, atol=10 * tol_o)


@pytest.mark.parametrize("family", ["gaussian", "binomial"])
def test_full_grid_fp32_against_the_oracle(_restore_glm_configs, family):
    """> 100 CTAs (one per SM), float32, compared with the ORACLE (not just properties): Gaussian on the incremental panels, binomial
    on the per-IRLS-iteration tensor-core panels.  A short path keeps the CPU oracle at a few seconds."""
    n, gs, G = 160_000, 10, 24
    p = gs * G
    rng = np.random.default_rng(21)
    X = np.asfortranarray(rng.standard_normal((n, p), dtype=np.float32))
    beta = np.zeros(p); beta[rng.choice(p, 30, replace=False)] = rng.normal(size=30)
    eta = X @ beta
    if family == "gaussian":
        y = (eta + np.linalg.norm(beta) * rng.normal(size=n)).astype(np.float32)
        mk, spec = ad.glm.gaussian, "gaussian"
    else:
        y = (rng.uniform(size=n) < 1 / (1 + np.exp(-eta / np.linalg.norm(beta)))).astype(np.float32)
        mk, spec = ad.glm.binomial, "binomial"
    kw = dict(groups=np.arange(0, p, gs), alpha=1.0 if family == "gaussian" else 0.5, tol=1e-7, irls_tol=1e-7, newton_tol=1e-6, early_exit=False,
              lmda_path_size=8, min_ratio=0.2)
    st = ad.grpnet(X, mk(y, dtype=np.float32), progress_bar=False, **kw)
    assert st.error == "" and st.sweep_ncta > 100 and st.sweep_batch > 1 and st.n_batched_launches > 0
    o = orc.grpnet(X, orc.glm_spec(spec, y, dtype=np.float32), n_threads=8, **kw)
    assert o.error == ""
    assert _rel(st.betas.toarray(), o.betas.toarray()) <= 1e-4
    # the intercept is ~0 here (balanced classes): it is compared on the scale of the coefficients (both are linear-predictor units)
    scale = max(np.max(np.abs(o.intercepts)), np.max(np.abs(o.betas.toarray())))
    assert np.max(np.abs(np.asarray(st.intercepts) - np.asarray(o.intercepts))) <= 1e-4 * scale


def test_tall_row_tiles_narrow_the_ring_items():
    """3392 rows per CTA (n = 1M on 2 GPUs): the default 5-column ring item no longer fits shared memory next to the residual tile and
    the panel slots; the planner narrows the item to 2 columns instead of falling back to the per-group kernel."""
    sizes = [10] * 14
    groups = _groups_from_sizes(sizes); p = int(np.sum(sizes))
    X, y, pen = _problem(3392 * 2, p, groups, np.float32, seed=13)
    ref = _solve(X, y, groups, pen, np.float32, batch=1, ctas=2)
    st = _solve(X, y, groups, pen, np.float32, batch=0, ctas=2)
    assert st.error == "" and st.sweep_batch > 1 and st.n_batched_launches > 0
    assert _rel(st.betas.toarray(), ref.betas.toarray()) <= 5e-5
    o = _solve(X, y, groups, pen, np.float32, batch=0, ctas=2, use_oracle=True)
    assert _rel(st.betas.toarray(), o.betas.toarray()) <= 1e-4
