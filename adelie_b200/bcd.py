"""Block-coordinate-descent proximal sub-problem (reference: adelie/bcd.py:123-343,
adelie/src/py_bcd.cpp:15-243; double precision only, like the reference).

The solve runs the same warp-cooperative device function the fused sweep kernel uses
(csrc/sweep.cuh: warp_prox_newton)."""
import ctypes as C
import numpy as np
from . import _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def objective(beta, *, quad, linear, l1, l2):
    """0.5 b^T diag(quad) b - linear^T b + l1 ||b|| + 0.5 l2 ||b||^2 (adelie/bcd.py:18-70)."""
    beta = np.asarray(beta)
    return 0.5 * np.sum(quad * beta ** 2) - linear @ beta + l1 * np.linalg.norm(beta) + 0.5 * l2 * np.sum(beta ** 2)


def root_lower_bound(*, quad, linear, l1):
    q, v = _d(quad), _d(linear)
    out = C.c_double()
    _lib.check(_lib.load().ab_bcd_root_lower_bound(q.size, _lib.ptr(q), _lib.ptr(v), l1, C.byref(out)))
    return out.value


def root_upper_bound(*, quad, linear, l1, zero_tol=1e-14):
    q, v = _d(quad), _d(linear)
    out = C.c_double()
    _lib.check(_lib.load().ab_bcd_root_upper_bound(q.size, _lib.ptr(q), _lib.ptr(v), l1, zero_tol, C.byref(out)))
    return out.value


def root_function(h, *, D, v, l1):
    D, v = _d(D), _d(v)
    out = C.c_double()
    _lib.check(_lib.load().ab_bcd_root_function(D.size, float(h), _lib.ptr(D), _lib.ptr(v), l1, C.byref(out)))
    return out.value


def solve(*, quad, linear, l1, l2, tol=1e-12, max_iters=1000, solver="newton_abs", **kwargs):
    """Solves the group prox; returns ``{"beta", "iters"}`` (adelie/bcd.py:182-261)."""
    solvers = {"newton": 0, "newton_abs": 1}
    if solver not in solvers:
        raise RuntimeError(f"adelie_b200.bcd.solve: solver '{solver}' is out of scope (newton / newton_abs only).")
    q, v = _d(quad), _d(linear)
    x = np.empty_like(q)
    iters = C.c_int64()
    _lib.check(_lib.load().ab_bcd_solve(solvers[solver], q.size, _lib.ptr(q), _lib.ptr(v), l1, l2, tol, max_iters,
                                        _lib.ptr(x), C.byref(iters)))
    return {"beta": x, "iters": iters.value}


def root(*, quad, linear, l1, l2, tol=1e-12, max_iters=1000, solver="newton_abs"):
    """Root h = ||beta|| of the prox (adelie/bcd.py:264-343)."""
    out = solve(quad=quad, linear=linear, l1=l1, l2=l2, tol=tol, max_iters=max_iters, solver=solver)
    return {"root": float(np.linalg.norm(out["beta"])), "iters": out["iters"]}
