"""CPU: the oracle's GLM families against the golden vectors produced by the REFERENCE's own NumPy test classes
(tests/golden/make_golden.py; reference tests/test_glm.py:114-732)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "glm_golden.npz"))
SIZES = [1, 2, 5, 10, 20, 100]


def _check(spec, pre, atol=1e-12):
    eta = G[pre + "eta"]
    grad = orc.glm_eval(spec, "gradient", eta=eta)
    np.testing.assert_allclose(grad, G[pre + "grad"], atol=atol, rtol=1e-10)
    hess = orc.glm_eval(spec, "hessian", eta=eta, grad=G[pre + "grad"])
    np.testing.assert_allclose(hess, G[pre + "hess"], atol=atol, rtol=1e-9)
    ihg = orc.glm_eval(spec, "inv_hessian_gradient", eta=eta, grad=G[pre + "grad"], hess=G[pre + "hess"])
    np.testing.assert_allclose(ihg, G[pre + "inv_hess_grad"], rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(orc.glm_eval(spec, "loss", eta=eta), G[pre + "loss"], rtol=1e-10, atol=atol)
    np.testing.assert_allclose(orc.glm_eval(spec, "loss_full"), G[pre + "loss_full"], rtol=1e-10, atol=atol)
    np.testing.assert_allclose(orc.glm_eval(spec, "inv_link", eta=eta), G[pre + "inv_link"], rtol=1e-12, atol=atol)


@pytest.mark.parametrize("n", SIZES)
def test_gaussian(n):
    pre = f"gaussian_{n}_"
    _check(orc.glm_spec("gaussian", G[pre + "y"], G[pre + "w"]), pre)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("binary", [0, 1])
def test_binomial(n, binary):
    pre = f"binomial_{n}_{binary}_"
    _check(orc.glm_spec("binomial", G[pre + "y"], G[pre + "w"]), pre)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("tie", ["efron", "breslow"])
def test_cox(n, tie):
    pre = f"cox_{n}_{tie}_"
    spec = orc.glm_spec("cox", G[pre + "status"], G[pre + "w"], start=G[pre + "start"], stop=G[pre + "stop"],
                        strata=G[pre + "strata"], tie_method=tie)
    _check(spec, pre, atol=1e-10)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_multigaussian(n, K):
    pre = f"multigaussian_{n}_{K}_"
    _check(orc.glm_spec("multigaussian", G[pre + "y"], G[pre + "w"]), pre)


@pytest.mark.parametrize("n", SIZES)
def test_poisson(n):
    pre = f"poisson_{n}_"
    _check(orc.glm_spec("poisson", G[pre + "y"], G[pre + "w"]), pre)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("binary", [0, 1])
def test_binomial_probit(n, binary):
    pre = f"probit_{n}_{binary}_"
    _check(orc.glm_spec("probit", G[pre + "y"], G[pre + "w"]), pre)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("K", [2, 3, 4])
@pytest.mark.parametrize("binary", [0, 1])
def test_multinomial(n, K, binary):
    pre = f"multinomial_{n}_{K}_{binary}_"
    _check(orc.glm_spec("multinomial", G[pre + "y"], G[pre + "w"]), pre)
