"""Generates tests/golden/glm_golden.npz from the REFERENCE's own NumPy restatements of the GLM families.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference's compiled core cannot be built here (Eigen is not vendored), but its test-suite carries independent,
pure-NumPy definitions of every family on the hot path (tests/test_glm.py: GlmTestGaussian :114-132,
GlmTestBinomialLogit :158-181, GlmTestCoxPack :458-593 / GlmTestCox :596-661 with exact O(n^2) at-risk sums,
GlmTestMultiGaussian :711-732, GlmTestPoisson :252-272).  We import those classes from the reference tree (with the compiled `adelie` package
stubbed out), feed them the seeded inputs of the reference's own test functions, and commit the outputs as golden vectors.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/tests/test_glm.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glm_golden.npz")


def load_reference_test_classes():
    # stub the compiled package: the NumPy classes only need configs.Configs.hessian_min
    adelie = types.ModuleType("adelie")
    glm = types.ModuleType("adelie.glm"); configs = types.ModuleType("adelie.configs"); core = types.ModuleType("adelie.adelie_core")
    class Configs: hessian_min = 1e-24
    configs.Configs = Configs
    # GlmTestCoxPack calls two compiled helpers; stand-ins use the O(n^2) NumPy definitions the reference's own tests hold
    # for them (test_cox_nnz_event_ties_sum :394-420, test_cox_scale :423-455).
    class GlmCoxPack64:
        @staticmethod
        def _nnz_event_ties_sum(a, t, status, w, out):
            exp = np.sum((t[None] == t[:, None]) * ((w != 0) * status * a)[None], axis=-1).astype(float)
            exp[(status == 0) | (w == 0)] = 0
            out[...] = exp
        @staticmethod
        def _scale(t, status, w, method, out):
            n = t.shape[0]
            exp = np.zeros(n)
            if method == "efron":
                mask = (status != 0) & (w != 0)
                ta = t[mask]
                if ta.size:
                    _, counts = np.unique(ta, return_counts=True)
                    exp[mask] = np.concatenate([np.arange(c, dtype=float) / c for c in counts])
            out[...] = exp
    core.glm = types.SimpleNamespace(GlmCoxPack64=GlmCoxPack64)
    adelie.glm = glm; adelie.configs = configs; adelie.adelie_core = core
    sys.modules.update({"adelie": adelie, "adelie.glm": glm, "adelie.configs": configs, "adelie.adelie_core": core})
    ns = {"__name__": "ref_test_glm"}
    with open(REF) as f:
        src = f.read()
    exec(compile(src, REF, "exec"), ns)
    return ns


def evaluate(model, eta, shape):
    grad = np.empty(shape); hess = np.empty(shape); ihg = np.empty(shape); inv = np.empty(shape)
    model.gradient(eta, grad)
    model.hessian(eta, grad, hess)
    model.inv_hessian_gradient(eta, grad, hess, ihg)
    model.inv_link(eta, inv)
    return dict(grad=grad, hess=hess, inv_hess_grad=ihg, loss=np.float64(model.loss(eta)), loss_full=np.float64(model.loss_full()), inv_link=inv)


def weights_like_reference(n):
    w = np.random.uniform(0, 1, n)
    w[np.random.binomial(1, 0.2, n).astype(bool)] = 0
    w[0] = 1
    w /= np.sum(w)
    return w


def main():
    ns = load_reference_test_classes()
    out = {}
    sizes = [1, 2, 5, 10, 20, 100]
    for n in sizes:                                   # test_gaussian (tests/test_glm.py:135-150)
        np.random.seed(0)
        y = np.random.normal(0, 1, n); w = weights_like_reference(n)
        eta = np.random.normal(0, 1, n)
        r = evaluate(ns["GlmTestGaussian"](y=y, weights=w), eta, (n,))
        out.update({f"gaussian_{n}_{k}": v for k, v in dict(y=y, w=w, eta=eta, **r).items()})
    for n in sizes:                                   # test_binomial (logit; :222-246), binary and fractional responses
        for binary in (True, False):
            np.random.seed(0)
            y = np.random.binomial(1, 0.5, n).astype(float) if binary else np.random.uniform(0, 1, n)
            w = weights_like_reference(n)
            eta = np.random.normal(0, 1, n)
            r = evaluate(ns["GlmTestBinomialLogit"](y=y, weights=w), eta, (n,))
            out.update({f"binomial_{n}_{int(binary)}_{k}": v for k, v in dict(y=y, w=w, eta=eta, **r).items()})
            r = evaluate(ns["GlmTestBinomialProbit"](y=y, weights=w), eta, (n,))          # probit link (GlmTestBinomialProbit :184-217)
            out.update({f"probit_{n}_{int(binary)}_{k}": v for k, v in dict(y=y, w=w, eta=eta, **r).items()})
    for n in sizes:                                   # test_cox (:663-705): discrete times => ties, 3 strata, zero weights
        for tie in ("efron", "breslow"):
            np.random.seed(0)
            s = np.random.choice(20, n).astype(float)
            t = 1 + s + np.random.choice(20, n)
            d = np.random.binomial(1, 0.5, n).astype(float)
            w = weights_like_reference(n)
            strata = np.random.choice(min(n, 3), n)
            eta = np.random.normal(0, 1, n)
            r = evaluate(ns["GlmTestCox"](start=s, stop=t, status=d, strata=strata, weights=w, tie_method=tie), eta, (n,))
            out.update({f"cox_{n}_{tie}_{k}": v for k, v in dict(start=s, stop=t, status=d, strata=strata, w=w, eta=eta, **r).items()})
    for n in sizes:                                   # test_multigaussian (:736-754)
        for K in (1, 2, 3, 4):
            np.random.seed(0)
            y = np.random.normal(0, 1, (n, K)); w = weights_like_reference(n)
            eta = np.random.normal(0, 1, (n, K))
            r = evaluate(ns["GlmTestMultiGaussian"](y=y, weights=w), eta, (n, K))
            out.update({f"multigaussian_{n}_{K}_{k}": v for k, v in dict(y=y, w=w, eta=eta, **r).items()})
    for n in sizes:                                   # test_multinomial (:790-803): one-hot and Dirichlet responses
        for K in (2, 3, 4):
            for binary in (True, False):
                np.random.seed(0)
                y = np.random.multinomial(1, np.full(K, 1 / K), n).astype(float) if binary else np.random.dirichlet(np.ones(K), n)
                w = weights_like_reference(n)
                eta = np.random.normal(0, 1, (n, K))
                r = evaluate(ns["GlmTestMultinomial"](y=y, weights=w), eta, (n, K))
                out.update({f"multinomial_{n}_{K}_{int(binary)}_{k}": v for k, v in dict(y=y, w=w, eta=eta, **r).items()})
    for n in sizes:                                   # test_poisson (:275-290)
        np.random.seed(0)
        y = np.random.poisson(1, n).astype(float); w = weights_like_reference(n)
        eta = np.random.normal(0, 1, n)
        r = evaluate(ns["GlmTestPoisson"](y=y, weights=w), eta, (n,))
        out.update({f"poisson_{n}_{k}": v for k, v in dict(y=y, w=w, eta=eta, **r).items()})
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
