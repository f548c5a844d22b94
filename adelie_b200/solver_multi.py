"""Multi-response branch of ``grpnet`` (reference: adelie/solver.py:699-846).

The single-response reformulation [kron(1, I_K) | kron(X, I_K)] is only a *layout* here: the device solver keeps the (n, p) matrix
and addresses (feature, class) pairs inside the kernels, so X is never expanded in HBM.
"""
from __future__ import annotations

import numpy as np

from . import dist as _dist
from . import matrix


def _mul_aug(X, V, intercept, dtype):
    """[kron(1,I_K) | kron(X,I_K)]^T vec(V) for V (n, K): K passes of X.mul, summed over ranks when row-sharded."""
    n, K = V.shape
    ones = np.ones(n, dtype=dtype)
    G = np.empty((X.cols(), K), dtype=dtype)
    tmp = np.empty(X.cols(), dtype=dtype)
    for l in range(K):
        X.mul(np.ascontiguousarray(V[:, l], dtype=dtype), ones, tmp)
        G[:, l] = tmp
    out = G.ravel()
    if intercept:
        out = np.concatenate([V.sum(axis=0).astype(dtype), out])
    return np.asarray(_dist.allreduce(np.ascontiguousarray(out, dtype=dtype)), dtype=dtype)


def grpnet_multi(*, X, X_raw, glm, groups, penalty, warm_start, solver_args, is_gaussian_opt, check_state, progress_bar,
                 exit_cond, dtype):
    from .state import multigaussian_naive as state_multigaussian_naive
    from .state import multiglm_naive as state_multiglm_naive

    n, p = X.rows(), X.cols()
    intercept = solver_args["intercept"]
    alpha = solver_args["alpha"]
    offsets = solver_args["offsets"]
    K = glm.y.shape[-1]

    groups = np.asarray(groups) * K                                      # solver.py:705
    if intercept:
        groups = np.concatenate([np.arange(K), K + groups]).astype(int)
    group_sizes = np.concatenate([groups, [(p + intercept) * K]]).astype(int)
    group_sizes = group_sizes[1:] - group_sizes[:-1]
    if penalty is None:
        penalty = np.sqrt(group_sizes).astype(dtype)
        if intercept:
            penalty[:K] = 0
    elif intercept:
        penalty = np.concatenate([np.zeros(K), penalty]).astype(dtype)

    if warm_start is None:
        lmda = np.inf
        lmda_max = None
        screen_set = np.arange(groups.shape[0])[(penalty <= 0) | (alpha <= 0)]
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

    if is_gaussian_opt:                                                  # solver.py:766-816
        y = glm.y
        weights = glm.weights
        wms = weights / K
        if warm_start is None:
            ones = np.ones(n, dtype=dtype)
            X_means = np.empty(p, dtype=dtype)
            X.mul(ones, np.ascontiguousarray(wms, dtype=dtype), X_means)
            X_means = np.repeat(np.asarray(_dist.allreduce(X_means), dtype=dtype), K)
            if intercept:
                X_means = np.concatenate([np.full(K, 1 / K), X_means]).astype(dtype)
            y_off = y - offsets
            y_var = _dist.allreduce(np.sum(wms[:, None] * y_off ** 2))
            if intercept:
                ybar = np.asarray(_dist.allreduce(y_off.T @ weights))        # NOT wms: matches the reference
                y_off_c = y_off - ybar[None]
                yc_var = _dist.allreduce(np.sum(wms[:, None] * y_off_c ** 2))
                rsq = yc_var - y_var
                y_var = yc_var
            else:
                rsq = 0
            resid = np.ascontiguousarray(y_off.ravel(), dtype=dtype)
            resid_sum = _dist.allreduce(np.sum(wms[:, None] * y_off))
            grad = _mul_aug(X, (y_off * wms[:, None]).astype(dtype), intercept, dtype)
        else:
            X_means = warm_start.X_means
            y_var = warm_start.y_var
            rsq = warm_start.rsq
            resid = warm_start.resid
            resid_sum = warm_start.resid_sum
            grad = warm_start.grad
        solver_args.update(X_means=X_means, y_var=y_var, rsq=rsq, resid=resid, resid_sum=resid_sum, grad=grad)
        state = state_multigaussian_naive(**solver_args)
    else:                                                                # solver.py:818-846
        if warm_start is None:
            eta = offsets
            resid = np.empty(eta.shape, dtype=dtype)
            glm.gradient(eta, resid)
            grad = _mul_aug(X, resid, intercept, dtype)
            resid = resid.ravel()
            loss_null = None
            loss_full = glm.loss_full()
            eta = eta.ravel()
        else:
            eta = warm_start.eta
            resid = warm_start.resid
            grad = warm_start.grad
            loss_null = warm_start.loss_null
            loss_full = warm_start.loss_full
        solver_args.update(grad=grad, eta=eta, resid=resid, loss_null=loss_null, loss_full=loss_full)
        state = state_multiglm_naive(**solver_args)

    if check_state:
        state.check(method="assert")
    return state.solve(progress_bar=progress_bar, exit_cond=exit_cond)
