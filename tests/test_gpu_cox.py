"""GPU parity tests of the Cox family (reference GlmCox, adelie_core/glm/glm_cox.ipp): the device
pipeline of gathers + segmented scans against (i) the golden vectors produced by the reference's own O(n^2) NumPy test classes
(tests/golden/make_golden.py from T/test_glm.py:298-661), (ii) the CPU oracle at multi-block sizes with heavy ties, strata, start
times and zero weights, and (iii) the cox path through grpnet against the oracle's path."""
import os

import numpy as np
import pytest

import adelie_b200 as ad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "glm_golden.npz"))


@pytest.mark.parametrize("dtype,rtol,atol", [(np.float64, 1e-9, 1e-10), (np.float32, 1e-4, 1e-5)])
@pytest.mark.parametrize("tie", ["efron", "breslow"])
@pytest.mark.parametrize("n", [1, 2, 5, 10, 20, 100])
def test_cox_vs_reference_golden(dtype, rtol, atol, tie, n):
    pre = f"cox_{n}_{tie}_"
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    m = ad.glm.cox(start=G[pre + "start"], stop=G[pre + "stop"], status=G[pre + "status"], strata=G[pre + "strata"], weights=G[pre + "w"],
                   tie_method=tie, dtype=dtype)
    eta = c(G[pre + "eta"])
    grad = np.empty_like(eta); m.gradient(eta, grad)
    np.testing.assert_allclose(grad, G[pre + "grad"], rtol=rtol, atol=atol)
    hess = np.empty_like(eta); m.hessian(eta, c(G[pre + "grad"]), hess)
    np.testing.assert_allclose(hess, G[pre + "hess"], rtol=rtol * 10, atol=atol)
    np.testing.assert_allclose(m.loss(eta), G[pre + "loss"], rtol=rtol * 10, atol=atol)
    np.testing.assert_allclose(m.loss_full(), G[pre + "loss_full"], rtol=rtol * 10, atol=atol)
    inv = np.empty_like(eta); m.inv_link(eta, inv)
    np.testing.assert_allclose(inv, G[pre + "inv_link"], rtol=rtol, atol=atol)


def _surv(n, rng, n_strata, tied):
    start = rng.exponential(1.0, n)
    stop = start + 0.05 + rng.exponential(1.0, n)
    if tied:                                          # heavy ties, also between start and stop times
        start = np.round(start * 8) / 8
        stop = np.maximum(np.round(stop * 8) / 8, start + 0.125)
    status = (rng.uniform(size=n) < 0.6).astype(np.float64)
    w = rng.uniform(0.1, 1.0, n); w[rng.uniform(size=n) < 0.05] = 0.0; w /= w.sum()
    strata = rng.integers(0, n_strata, n).astype(np.int64) if n_strata > 1 else None
    return start, stop, status, w, strata


@pytest.mark.parametrize("tie", ["efron", "breslow"])
@pytest.mark.parametrize("n,n_strata,tied", [(1500, 1, True), (5003, 7, True), (40_000, 3, False), (100_003, 1, True), (3000, 2900, True)])
def test_cox_large_vs_oracle(n, n_strata, tied, tie):
    """multi-block segmented scans (n > 1024), strata boundaries inside blocks, tie groups spanning blocks, singleton strata"""
    rng = np.random.default_rng(n + n_strata)
    start, stop, status, w, strata = _surv(n, rng, n_strata, tied)
    m = ad.glm.cox(start=start, stop=stop, status=status, strata=strata, weights=w, tie_method=tie)
    spec = orc.glm_spec("cox", status, w, start=start, stop=stop, strata=strata, tie_method=tie)
    eta = rng.normal(size=n)
    g = np.empty(n); m.gradient(eta, g)
    go = orc.glm_eval(spec, "gradient", eta=eta)
    scale = np.max(np.abs(go))
    assert np.max(np.abs(g - go)) <= 1e-9 * scale
    h = np.empty(n); m.hessian(eta, go, h)
    ho = orc.glm_eval(spec, "hessian", eta=eta, grad=go)
    assert np.max(np.abs(h - ho)) <= 1e-8 * np.max(np.abs(ho))
    np.testing.assert_allclose(m.loss(eta), orc.glm_eval(spec, "loss", eta=eta), rtol=1e-9)
    np.testing.assert_allclose(m.loss_full(), orc.glm_eval(spec, "loss_full"), rtol=1e-9)
    assert abs(np.sum(g)) <= 1e-9 * n * scale                    # the Cox score sums to zero within every stratum


def test_cox_loss_shift_is_per_stratum():
    """a stratum whose linear predictors sit 800 below another's must not underflow (glm_cox.ipp:474 shifts by the pack's own max)"""
    rng = np.random.default_rng(3)
    n = 400
    start, stop, status, w, _ = _surv(n, rng, 1, False)
    strata = (np.arange(n) % 2).astype(np.int64)
    eta = rng.normal(size=n) + np.where(strata == 0, 300.0, -500.0)
    m = ad.glm.cox(start=start, stop=stop, status=status, strata=strata, weights=w)
    spec = orc.glm_spec("cox", status, w, start=start, stop=stop, strata=strata)
    lo = orc.glm_eval(spec, "loss", eta=eta)
    assert np.isfinite(lo)
    np.testing.assert_allclose(m.loss(eta), lo, rtol=1e-10)


@pytest.mark.parametrize("tie,alpha", [("efron", 1.0), ("breslow", 0.5)])
def test_cox_path_vs_oracle(tie, alpha):
    n, p, Gn = 1200, 40, 10
    data = ad.data.dense(n, p, Gn, glm="cox", seed=4)
    glm = data["glm"]
    rng = np.random.default_rng(0)
    stop = np.round(glm.stop * 20) / 20 + 1.0          # ties
    strata = rng.integers(0, 3, n).astype(np.int64)
    m = ad.glm.cox(start=glm.start, stop=stop, status=glm.status, strata=strata, tie_method=tie)
    spec = orc.glm_spec("cox", glm.status, None, start=glm.start, stop=stop, strata=strata, tie_method=tie)
    kw = dict(groups=data["groups"], penalty=data["penalty"], alpha=alpha, tol=1e-13, irls_tol=1e-11, early_exit=False, lmda_path_size=10,
              min_ratio=0.2, intercept=False)
    st = ad.grpnet(data["X"], m, progress_bar=False, **kw)
    ref = orc.grpnet(data["X"], spec, **kw)
    assert st.error == "" and ref.error == "", (st.error, ref.error)
    assert len(st.lmdas) == len(ref.lmdas)
    np.testing.assert_allclose(st.lmdas, ref.lmdas, rtol=1e-9)
    B, Br = np.asarray(st.betas.todense()), np.asarray(ref.betas.todense())
    assert np.max(np.abs(B - Br)) <= 1e-6 * np.max(np.abs(Br))
    np.testing.assert_allclose(st.devs, ref.devs, rtol=1e-6, atol=1e-8)
