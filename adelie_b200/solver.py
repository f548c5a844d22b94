"""``grpnet`` -- the user entry point (reference: adelie/solver.py:354-958).

Builds the initial invariants exactly as the reference does (two ``X.mul`` passes and NumPy
scalars, solver.py:848-950), constructs the matching naive state and solves it on the device.
"""
from __future__ import annotations

from typing import Callable, Union

import numpy as np

from . import dist as _dist
from . import matrix
from .state import gaussian_naive as state_gaussian_naive
from .state import glm_naive as state_glm_naive
from .state import gaussian_cov as state_gaussian_cov


def _init_gaussian(X, y, weights, offsets, intercept, dtype):
    """Initial invariants of the Gaussian state (adelie/solver.py:887-904).  In row-sharded mode every sum over
    observations is all-reduced over the ranks (identity otherwise); X only needs ``rows/cols/mul``."""
    n, p = X.rows(), X.cols()
    ones = np.ones(n, dtype=dtype)
    X_means = np.empty(p, dtype=dtype)
    X.mul(ones, weights, X_means)
    X_means = np.asarray(_dist.allreduce(X_means), dtype=dtype)
    y_off = y - offsets
    y_mean = _dist.allreduce(np.sum(y_off * weights))
    yc = y_off
    if intercept:
        yc = yc - y_mean
    y_var = _dist.allreduce(np.sum(weights * yc ** 2))
    resid = np.ascontiguousarray(yc, dtype=dtype)
    resid_sum = _dist.allreduce(np.sum(weights * resid))
    grad = np.empty(p, dtype=dtype)
    X.mul(resid, weights, grad)
    grad = np.asarray(_dist.allreduce(grad), dtype=dtype)
    return dict(X_means=X_means, y_mean=y_mean, y_var=y_var, rsq=0, resid=resid, resid_sum=resid_sum, grad=grad)


def grpnet(
    X, glm, *, constraints: list = None, groups: np.ndarray = None, alpha: float = 1, penalty: np.ndarray = None,
    offsets: np.ndarray = None, lmda_path: np.ndarray = None, irls_max_iters: int = int(1e4), irls_tol: float = 1e-7,
    max_iters: int = int(1e5), tol: float = 1e-7, adev_tol: float = 0.9, ddev_tol: float = 0, newton_tol: float = 1e-12,
    newton_max_iters: int = 1000, n_threads: int = 1, early_exit: bool = True, intercept: bool = True,
    screen_rule: str = "pivot", min_ratio: float = 1e-2, lmda_path_size: int = 100, max_screen_size: int = None,
    max_active_size: int = None, pivot_subset_ratio: float = 0.1, pivot_subset_min: int = 1, pivot_slack_ratio: float = 1.25,
    check_state: bool = False, progress_bar: bool = True, warm_start=None, exit_cond: Callable = None,
):
    """Solves the group elastic net via the naive method (adelie/solver.py:354-958)."""
    X_raw = X
    if isinstance(X, np.ndarray):
        X = matrix.dense(X, method="naive", n_threads=n_threads)
    assert isinstance(X, matrix.MatrixNaiveBase)
    dtype = X.dtype
    n, p = X.rows(), X.cols()

    if offsets is not None:
        if offsets.shape != glm.y.shape:
            raise RuntimeError("offsets must be same shape as y if not None.")
        offsets = np.asarray(offsets, order="C", dtype=dtype)
    else:
        offsets = np.zeros(glm.y.shape, dtype=dtype)

    if lmda_path is not None:
        lmda_path = np.array(np.flip(np.sort(lmda_path)), dtype=dtype)

    solver_args = dict(
        X=X, constraints=constraints, alpha=alpha, offsets=offsets, lmda_path=lmda_path, max_iters=max_iters, tol=tol,
        adev_tol=adev_tol, ddev_tol=ddev_tol, newton_tol=newton_tol, newton_max_iters=newton_max_iters, n_threads=n_threads,
        early_exit=early_exit, intercept=intercept, screen_rule=screen_rule, min_ratio=min_ratio, lmda_path_size=lmda_path_size,
        max_screen_size=max_screen_size, max_active_size=max_active_size, pivot_subset_ratio=pivot_subset_ratio,
        pivot_subset_min=pivot_subset_min, pivot_slack_ratio=pivot_slack_ratio,
    )
    is_gaussian_opt = (glm.name in ["gaussian", "multigaussian"]) and glm.opt
    if not is_gaussian_opt:
        solver_args["glm"] = glm
        solver_args["irls_max_iters"] = irls_max_iters
        solver_args["irls_tol"] = irls_tol
    else:
        solver_args["y"] = glm.y
        solver_args["weights"] = glm.weights

    if groups is None:
        groups = np.arange(p, dtype=int)
    groups = np.asarray(groups)

    if glm.is_multi:
        from .solver_multi import grpnet_multi
        return grpnet_multi(X=X, X_raw=X_raw, glm=glm, groups=groups, penalty=penalty, warm_start=warm_start,
                            solver_args=solver_args, is_gaussian_opt=is_gaussian_opt, check_state=check_state,
                            progress_bar=progress_bar, exit_cond=exit_cond, dtype=dtype)

    group_sizes = np.concatenate([groups, [p]], dtype=int)
    group_sizes = group_sizes[1:] - group_sizes[:-1]
    G = len(groups)
    if penalty is None:
        penalty = np.sqrt(group_sizes).astype(dtype)

    if warm_start is None:
        lmda = np.inf
        lmda_max = None
        screen_set = np.arange(G)[(penalty <= 0) | (alpha <= 0)]
        screen_beta = np.zeros(np.sum(group_sizes[screen_set]), dtype=dtype)
        screen_is_active = np.ones(screen_set.shape[0], dtype=bool)
        active_set_size = screen_set.shape[0]
        active_set = np.empty(groups.shape[0], dtype=int)
        active_set[:active_set_size] = np.arange(active_set_size)
    else:
        lmda = warm_start.lmda
        lmda_max = warm_start.lmda_max
        screen_set = warm_start.screen_set
        screen_beta = warm_start.screen_beta
        screen_is_active = warm_start.screen_is_active
        active_set_size = warm_start.active_set_size
        active_set = warm_start.active_set

    solver_args.update(groups=groups, group_sizes=group_sizes, penalty=penalty, lmda=lmda, lmda_max=lmda_max,
                       screen_set=screen_set, screen_beta=screen_beta, screen_is_active=screen_is_active,
                       active_set_size=active_set_size, active_set=active_set)

    if is_gaussian_opt:
        y = glm.y
        weights = glm.weights
        if warm_start is None:                                       # solver.py:887-904
            inv = _init_gaussian(X, y, weights, offsets, intercept, dtype)
            X_means, y_mean, y_var, rsq = inv["X_means"], inv["y_mean"], inv["y_var"], inv["rsq"]
            resid, resid_sum, grad = inv["resid"], inv["resid_sum"], inv["grad"]
        else:
            X_means = warm_start.X_means
            y_mean = warm_start.y_mean
            y_var = warm_start.y_var
            rsq = warm_start.rsq
            resid = warm_start.resid
            resid_sum = warm_start.resid_sum
            grad = warm_start.grad
        solver_args.update(X_means=X_means, y_mean=y_mean, y_var=y_var, rsq=rsq, resid=resid, resid_sum=resid_sum, grad=grad)
        state = state_gaussian_naive(**solver_args)
    else:
        if warm_start is None:                                       # solver.py:926-950
            ones = np.ones(n, dtype=dtype)
            beta0 = 0
            eta = offsets
            resid = np.empty(n, dtype=dtype)
            glm.gradient(eta, resid)
            grad = np.empty(p, dtype=dtype)
            X.mul(resid, ones, grad)
            grad = np.asarray(_dist.allreduce(grad), dtype=dtype)
            loss_null = None
            loss_full = glm.loss_full()
        else:
            beta0 = warm_start.beta0
            eta = warm_start.eta
            resid = warm_start.resid
            grad = warm_start.grad
            loss_null = warm_start.loss_null
            loss_full = warm_start.loss_full
        solver_args.update(beta0=beta0, grad=grad, eta=eta, resid=resid, loss_null=loss_null, loss_full=loss_full)
        state = state_glm_naive(**solver_args)

    if check_state:
        state.check(method="assert")
    return state.solve(progress_bar=progress_bar, exit_cond=exit_cond)


def gaussian_cov(
    A, v: np.ndarray, *, constraints: list = None, groups: np.ndarray = None, alpha: float = 1, penalty: np.ndarray = None,
    lmda_path: np.ndarray = None, max_iters: int = int(1e5), tol: float = 1e-7, rdev_tol: float = 1e-3, newton_tol: float = 1e-12,
    newton_max_iters: int = 1000, n_threads: int = 1, early_exit: bool = True, screen_rule: str = "pivot", min_ratio: float = 1e-2,
    lmda_path_size: int = 100, max_screen_size: int = None, max_active_size: int = None, pivot_subset_ratio: float = 0.1,
    pivot_subset_min: int = 1, pivot_slack_ratio: float = 1.25, check_state: bool = False, progress_bar: bool = True,
    warm_start=None, exit_cond: Callable = None,
):
    """Solves the Gaussian group elastic net via the covariance method (adelie/solver.py:39-352):
    minimize 1/2 b^T A b - v^T b + lmda * sum_g penalty_g (alpha ||b_g|| + (1 - alpha)/2 ||b_g||^2) for a positive semi-definite A."""
    if isinstance(A, np.ndarray):
        A = matrix.dense(A, method="cov", n_threads=n_threads)
    assert isinstance(A, (matrix.MatrixCovBase64, matrix.MatrixCovBase32))
    dtype = np.float64 if isinstance(A, matrix.MatrixCovBase64) else np.float32
    p = A.cols()
    if constraints is not None and any(c is not None for c in constraints):
        raise RuntimeError("adelie_b200: constraints are out of scope for the B200 path (pass constraints=None).")
    if lmda_path is not None:
        lmda_path = np.array(np.flip(np.sort(lmda_path)))
    if groups is None:
        groups = np.arange(p, dtype=int)
    groups = np.asarray(groups, dtype=int)
    group_sizes = np.concatenate([groups, [p]], dtype=int)
    group_sizes = group_sizes[1:] - group_sizes[:-1]
    G = len(groups)
    if penalty is None:
        penalty = np.sqrt(group_sizes)
    penalty = np.asarray(penalty)
    v = np.asarray(v, dtype=dtype)

    if warm_start is None:
        lmda = np.inf
        lmda_max = None
        screen_set = np.arange(G)[(penalty <= 0) | (alpha <= 0)]
        screen_beta = np.zeros(np.sum(group_sizes[screen_set]), dtype=dtype)
        screen_is_active = np.ones(screen_set.shape[0], dtype=bool)
        active_set_size = screen_set.shape[0]
        active_set = np.empty(G, dtype=int)
        active_set[:active_set_size] = np.arange(active_set_size)
        rsq = 0
        subset = (np.concatenate([np.arange(groups[ss], groups[ss] + group_sizes[ss]) for ss in screen_set])
                  if len(screen_set) else np.zeros(0, dtype=int))
        order = np.argsort(subset)
        grad = np.empty(p, dtype=dtype)
        A.mul(subset[order], np.ascontiguousarray(screen_beta[order], dtype=dtype), grad)
        grad = v - grad
    else:
        lmda = warm_start.lmda
        lmda_max = warm_start.lmda_max
        screen_set = warm_start.screen_set
        screen_beta = warm_start.screen_beta
        screen_is_active = warm_start.screen_is_active
        active_set_size = warm_start.active_set_size
        active_set = warm_start.active_set
        rsq = warm_start.rsq
        grad = warm_start.grad

    state = state_gaussian_cov(
        A=A, v=v, constraints=constraints, groups=groups, group_sizes=group_sizes, alpha=alpha, penalty=penalty,
        screen_set=screen_set, screen_beta=screen_beta, screen_is_active=screen_is_active, active_set_size=active_set_size,
        active_set=active_set, rsq=rsq, lmda=lmda, grad=grad, lmda_path=lmda_path, lmda_max=lmda_max, max_iters=max_iters, tol=tol,
        rdev_tol=rdev_tol, newton_tol=newton_tol, newton_max_iters=newton_max_iters, n_threads=n_threads, early_exit=early_exit,
        screen_rule=screen_rule, min_ratio=min_ratio, lmda_path_size=lmda_path_size, max_screen_size=max_screen_size,
        max_active_size=max_active_size, pivot_subset_ratio=pivot_subset_ratio, pivot_subset_min=pivot_subset_min,
        pivot_slack_ratio=pivot_slack_ratio,
    )
    if check_state:
        state.check(method="assert")
    return state.solve(progress_bar=progress_bar, exit_cond=exit_cond)
