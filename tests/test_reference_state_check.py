"""CPU, build container only (needs /root/reference; skipped on the GPU box): pins the solver oracle to reference-held code.

The reference's compiled core cannot be built here (Eigen is not vendored), but its Python wrapper carries a pure-NumPy re-derivation
of every invariant of a solved state -- `gaussian_naive_base.check` (adelie/state.py:1421-1674: rsq, grad, abs_grad, resid, resid_sum,
screen_X_means, screen_vars / screen_transforms against X, y and screen_beta) and `gaussian_pin_base.check` (:179-420).  We import
adelie/state.py from the reference tree with the compiled package stubbed out and run the reference's OWN `check(method="assert")`
on states produced by the oracle (the restatement under oracle/), on the path driver and on the pin solver.  A corrupted state must
fail the same check.  intercept=True only: the reference's rsq formula (:1573-1577) subtracts the column means unconditionally and
its own tests call check() only on intercept=True states."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from oracle import oracle as orc

REF_ROOT = "/root/reference/adelie"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_ROOT, "state.py")), reason="reference tree not present")


def load_reference_state_module():
    """adelie.state from the reference tree; adelie.adelie_core / matrix / glm / constraint are stubs (only class names are needed
    at import time), adelie.logger is the reference's own."""
    saved = {k: v for k, v in sys.modules.items() if k == "adelie" or k.startswith("adelie.")}
    pkg = types.ModuleType("adelie"); pkg.__path__ = []
    core = types.ModuleType("adelie.adelie_core")
    mat = types.ModuleType("adelie.matrix"); glm = types.ModuleType("adelie.glm"); con = types.ModuleType("adelie.constraint")
    for nm in ("MatrixConstraintBase32", "MatrixConstraintBase64", "MatrixCovBase32", "MatrixCovBase64", "MatrixNaiveBase32", "MatrixNaiveBase64"):
        setattr(mat, nm, type(nm, (), {}))
    for nm in ("GlmBase32", "GlmBase64", "GlmMultiBase32", "GlmMultiBase64"):
        setattr(glm, nm, type(nm, (), {}))
    for nm in ("ConstraintBase32", "ConstraintBase64"):
        setattr(con, nm, type(nm, (), {}))
    sys.modules.update({"adelie": pkg, "adelie.adelie_core": core, "adelie.matrix": mat, "adelie.glm": glm, "adelie.constraint": con})
    try:
        for name in ("logger", "state"):
            spec = importlib.util.spec_from_file_location("adelie." + name, os.path.join(REF_ROOT, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules["adelie." + name] = mod
            setattr(pkg, name, mod)
            spec.loader.exec_module(mod)
        return sys.modules["adelie.state"], mat
    finally:
        for k in [k for k in sys.modules if k == "adelie" or k.startswith("adelie.")]:
            del sys.modules[k]
        sys.modules.update(saved)


REF_STATE, REF_MATRIX = load_reference_state_module() if os.path.exists(os.path.join(REF_ROOT, "state.py")) else (None, None)


def numpy_matrix(X):
    """A MatrixNaiveBase64 (isinstance-checked by the reference) with the NumPy semantics of matrix_naive_base.hpp:57-143."""
    class M(REF_MATRIX.MatrixNaiveBase64):
        def rows(self): return X.shape[0]
        def cols(self): return X.shape[1]
        def btmul(self, j, q, v, out): out += X[:, j:j + q] @ v
        def mul(self, v, w, out): out[...] = X.T @ (v * w)
        def cov(self, j, q, sqrt_w, out):
            Y = sqrt_w[:, None] * X[:, j:j + q]
            out[...] = Y.T @ Y
    return M()


class Silent:
    def info(self, *a, **k): pass
    def warning(self, *a, **k): pass
    def error(self, *a, **k): pass


def wrap_path_state(ref, X, y, w, groups, group_sizes, penalty, alpha, intercept):
    s = types.SimpleNamespace()
    s.X = numpy_matrix(X); s._glm = types.SimpleNamespace(y=y); s.weights = w; s.intercept = intercept
    s.groups = groups.astype(int); s.group_sizes = group_sizes.astype(int); s.penalty = penalty; s.alpha = alpha
    s.constraints = [None] * len(groups)
    s.screen_set = ref.screen_set.astype(int); s.screen_begins = ref.screen_begins.astype(int); s.screen_beta = ref.screen_beta
    s.screen_is_active = ref.screen_is_active.astype(bool)
    s.rsq = ref.rsq; s.lmda = ref.lmda; s.lmda_max = ref.lmda_max; s.resid = ref.resid; s.resid_sum = ref.resid_sum
    s.grad = ref.grad; s.abs_grad = ref.abs_grad; s.X_means = X.T @ w
    s.screen_X_means = ref.screen_X_means; s.screen_vars = ref.screen_vars
    flat, st, o = ref.screen_transforms_flat, [], 0
    for i in s.screen_set:
        gs = int(group_sizes[i]); st.append(flat[o:o + gs * gs].reshape(gs, gs)); o += gs * gs
    s.screen_transforms = st
    s._check = lambda passed, msg, method, logger, *a, **k: REF_STATE.base._check(s, passed, msg, method, logger, *a, **k)
    return s


@pytest.mark.parametrize("n, p, G, alpha, intercept, seed", [(120, 40, 40, 1.0, True, 0), (200, 60, 12, 0.6, True, 1), (150, 48, 8, 1.0, True, 2),
                                                              (300, 90, 30, 0.3, True, 3)])
def test_oracle_path_state_passes_the_reference_check(n, p, G, alpha, intercept, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(size=(n, p)))
    groups = np.sort(np.concatenate([[0], rng.choice(np.arange(1, p), size=G - 1, replace=False)])).astype(np.int64)
    group_sizes = np.diff(np.concatenate([groups, [p]])).astype(np.int64)
    beta = np.where(rng.uniform(size=p) < 0.2, rng.normal(size=p), 0)
    y = X @ beta + 0.4 + rng.normal(size=n)
    w = rng.uniform(1, 2, n); w /= w.sum()
    penalty = rng.uniform(0.5, 1.5, G)
    ref = orc.grpnet(X, orc.glm_spec("gaussian", y, weights=w), groups=groups, alpha=alpha, penalty=penalty, intercept=intercept,
                     tol=1e-14, early_exit=False, lmda_path_size=15, min_ratio=0.05)
    assert ref.error == "" and len(ref.lmdas) == 15
    s = wrap_path_state(ref, X, y, w, groups, group_sizes, penalty, alpha, intercept)
    REF_STATE.gaussian_naive_base.check(s, method="assert", logger=Silent())            # the reference's own invariant checker
    # the checker is not vacuous: corrupt one invariant at a time
    for field, delta in (("rsq", 1e-3), ("resid_sum", 1e-3)):
        bad = wrap_path_state(ref, X, y, w, groups, group_sizes, penalty, alpha, intercept)
        setattr(bad, field, getattr(bad, field) + delta)
        with pytest.raises(AssertionError):
            REF_STATE.gaussian_naive_base.check(bad, method="assert", logger=Silent())
    bad = wrap_path_state(ref, X, y, w, groups, group_sizes, penalty, alpha, intercept)
    bad.screen_beta = bad.screen_beta.copy(); bad.screen_beta[np.flatnonzero(bad.screen_beta)[0]] *= 1.01
    with pytest.raises(AssertionError):
        REF_STATE.gaussian_naive_base.check(bad, method="assert", logger=Silent())


@pytest.mark.parametrize("n, p, G, S", [[10, 100, 20, 13], [100, 23, 4, 3], [100, 100, 50, 20]])
def test_oracle_pin_state_passes_the_reference_check(n, p, G, S):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_pin import create_data_gaussian_pin                                   # the reference's own generator (:214-335)
    args, group_sizes = create_data_gaussian_pin(n, p, G, S)
    ref = orc.pin_naive_solve(args["X"], args["y"], groups=args["groups"], alpha=args["alpha"], penalty=args["penalty"], weights=args["weights"],
                              screen_set=args["screen_set"], lmda_path=args["lmda_path"], tol=1e-7)
    assert ref.error == ""
    s = object.__new__(type("PinDuck", (REF_STATE.gaussian_pin_naive_base,), {}))          # the checker calls super().check()
    s.X = numpy_matrix(args["X"]); s.groups = args["groups"].astype(int); s.group_sizes = group_sizes.astype(int); s.penalty = args["penalty"]
    s.screen_set = args["screen_set"].astype(int); s.screen_begins = ref.screen_begins.astype(int); s.screen_vars = ref.screen_vars
    flat, st, o = ref.screen_transforms_flat, [], 0
    for i in s.screen_set:
        gs = int(group_sizes[i]); st.append(flat[o:o + gs * gs].reshape(gs, gs)); o += gs * gs
    s.screen_transforms = st; s.lmda_path = args["lmda_path"]
    s.active_set_size = int(ref.active_set_size); s.active_set = np.asarray(ref.active_set).astype(int)
    s.screen_is_active = np.asarray(ref.screen_is_active).astype(bool)
    a = s.active_set[: s.active_set_size]
    s.active_begins = np.cumsum(np.concatenate([[0], group_sizes[s.screen_set[a]]]).astype(int))[:-1]
    s.active_order = np.argsort(s.groups[s.screen_set[a]], kind="stable").astype(int)
    s.betas = ref.betas; s.rsqs = ref.rsqs; s.lmdas = ref.lmdas; s.resid = ref.resid
    REF_STATE.gaussian_pin_naive_base.check(s, method="assert", logger=Silent())


@pytest.mark.parametrize("n, p, G, S", [[10, 100, 20, 13], [100, 23, 4, 3], [100, 100, 50, 20]])
def test_oracle_pin_cov_state_passes_the_reference_check(n, p, G, S):
    """The covariance-method pin state of the oracle (oracle/cov_oracle.hpp) under the reference's OWN `gaussian_pin_cov_base.check`
    (adelie/state.py:723-736 -> gaussian_pin_base.check :179-398: screen / active bookkeeping, screen_vars vs screen_transforms sizes,
    betas / rsqs / lmdas shapes and signs), plus the invariant the cov method maintains: screen_grad = (v - A beta) on the screen values."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cov_data import create_data_gaussian_pin_cov
    args, ex = create_data_gaussian_pin_cov(n, p, G, S)
    a = {k: v for k, v in args.items() if k != "constraints"}
    ref = orc.gaussian_pin_cov(orc.cov_dense(np.asfortranarray(ex["A"])), **a, tol=1e-12)
    assert ref.error == ""
    group_sizes = ex["group_sizes"]
    s = object.__new__(type("PinCovDuck", (REF_STATE.gaussian_pin_cov_base,), {}))
    s.A = REF_MATRIX.MatrixCovBase64()
    s.groups = args["groups"].astype(int); s.group_sizes = group_sizes.astype(int); s.penalty = args["penalty"]
    s.screen_set = args["screen_set"].astype(int); s.screen_begins = ref.screen_begins.astype(int); s.screen_vars = ref.screen_vars
    flat, st, o = ref.screen_transforms_flat, [], 0
    for i in s.screen_set:
        gs = int(group_sizes[i]); st.append(flat[o:o + gs * gs].reshape(gs, gs)); o += gs * gs
    s.screen_transforms = st; s.lmda_path = np.asarray(args["lmda_path"])
    s.active_set_size = int(ref.active_set_size); s.active_set = np.asarray(ref.active_set).astype(int)
    s.screen_is_active = np.asarray(ref.screen_is_active).astype(bool)
    act = s.active_set[: s.active_set_size]
    s.active_begins = np.cumsum(np.concatenate([[0], group_sizes[s.screen_set[act]]]).astype(int))[:-1]
    s.active_order = np.argsort(s.groups[s.screen_set[act]], kind="stable").astype(int)
    s.betas = ref.betas; s.rsqs = ref.rsqs; s.lmdas = ref.lmdas
    REF_STATE.gaussian_pin_cov_base.check(s, method="assert", logger=Silent())
    # screen_grad of the solved state = v - A beta on the screen values (what coordinate_descent keeps up to date, :243-385)
    beta = ref.betas.toarray()[-1]
    grad = ex["v"] - ex["A"] @ beta
    sg = np.concatenate([grad[g:g + gs] for g, gs in zip(args["groups"][args["screen_set"]], group_sizes[args["screen_set"]])])
    np.testing.assert_allclose(ref.screen_grad, sg, atol=1e-10)
    # eigen-decomposition invariant of the screen groups: V diag(vars) V^T = A_gg
    for i, ss in enumerate(s.screen_set):
        g, gs = args["groups"][ss], group_sizes[ss]
        b = s.screen_begins[i]
        np.testing.assert_allclose(st[i] @ np.diag(ref.screen_vars[b:b + gs]) @ st[i].T, ex["A"][g:g + gs, g:g + gs], atol=1e-10)
