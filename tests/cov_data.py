"""Shared generators of the covariance-method tests (CPU oracle tests and GPU parity tests)."""
import numpy as np


def create_data_gaussian_pin_cov(n, p, G, S, alpha=1, sparsity=0.95, seed=0, min_ratio=0.5, n_lmdas=20):
    """tests/test_solver.py:214-335 of the reference with pin=True, method="cov", intercept=False (same draws in the same order)."""
    np.random.seed(seed)
    groups = np.sort(np.concatenate([[0], np.random.choice(np.arange(1, p), size=G - 1, replace=False)])).astype(int)
    group_sizes = np.diff(np.concatenate([groups, [p]])).astype(int)
    X = np.random.normal(0, 1, (n, p))
    beta = np.random.normal(0, 1, p)
    beta[np.random.choice(p, int(sparsity * p), replace=False)] = 0
    y = X @ beta + np.random.normal(0, 1, n)
    X /= np.sqrt(n); y /= np.sqrt(n)
    penalty = np.random.uniform(0, 1, G)
    penalty[np.random.choice(G, int(0.05 * G), replace=False)] = 0
    penalty /= np.linalg.norm(penalty) / np.sqrt(p)
    weights = np.random.uniform(1, 2, n)
    weights /= np.sum(weights)
    grad = X.T @ (weights * y)
    screen_set = np.random.choice(G, S, replace=False)
    abs_grad = np.array([np.linalg.norm(grad[g:g + gs]) for g, gs in zip(groups, group_sizes)])
    nz = penalty > 0
    lmda_max = np.max(abs_grad[nz] / (alpha * penalty[nz]))
    lmda_path = lmda_max * min_ratio ** (np.arange(n_lmdas) / (n_lmdas - 1))
    WsqrtX = np.sqrt(weights)[:, None] * X
    A = WsqrtX.T @ WsqrtX
    screen_grad = np.concatenate([grad[g:g + gs] for g, gs in zip(groups[screen_set], group_sizes[screen_set])])
    args = dict(constraints=None, groups=groups, alpha=alpha, penalty=penalty, rsq=0, active_set_size=0, active_set=np.zeros(G, dtype=int),
                lmda_path=lmda_path, screen_set=screen_set, screen_is_active=np.zeros(S, dtype=bool),
                screen_beta=np.zeros(np.sum(group_sizes[screen_set])), screen_grad=screen_grad)
    return args, dict(A=A, WsqrtX=np.asfortranarray(WsqrtX), v=grad, group_sizes=group_sizes)


def kkt_cov(A, v, groups, group_sizes, penalty, alpha, betas, lmdas, restrict=None, atol=2e-6):
    """Stationarity of 1/2 b^T A b - v^T b + lmda sum_g pen_g (alpha ||b_g|| + (1 - alpha)/2 ||b_g||^2), over `restrict` groups (all if None)."""
    B = np.asarray(betas)
    idx = range(len(groups)) if restrict is None else restrict
    for l, lmda in enumerate(lmdas):
        grad = v - A @ B[l]
        for i in idx:
            g, gs, pen = groups[i], group_sizes[i], penalty[i]
            b = B[l, g:g + gs]
            nb = np.linalg.norm(b)
            if nb > 0:
                np.testing.assert_allclose(grad[g:g + gs], lmda * pen * (alpha * b / nb + (1 - alpha) * b), atol=atol * max(1.0, lmda))
            else:
                assert np.linalg.norm(grad[g:g + gs]) <= lmda * pen * alpha * (1 + 1e-6) + atol
