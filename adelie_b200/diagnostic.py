"""Host-side diagnostics used by tests and callers (reference: adelie/diagnostic.py:30-276)."""
import numpy as np


def predict(X, betas, intercepts, offsets=None):
    """Linear predictions eta = X beta + intercept (+ offsets) for every lambda (diagnostic.py:30-121)."""
    X = np.asarray(X)
    B = np.asarray(betas.todense()) if hasattr(betas, "todense") else np.asarray(betas)
    eta = B @ X.T + np.asarray(intercepts)[:, None]
    if offsets is not None:
        eta = eta + offsets[None]
    return eta


def objective_gaussian(X, y, weights, beta, intercept, lmda, alpha, groups, group_sizes, penalty):
    """0.5 sum w (y - X beta - b0)^2 + lmda sum_g p_g (alpha ||b_g|| + 0.5 (1-alpha) ||b_g||^2)."""
    r = y - X @ beta - intercept
    loss = 0.5 * np.sum(weights * r ** 2)
    pen = 0.0
    for g, gs, pk in zip(groups, group_sizes, penalty):
        bn = np.linalg.norm(beta[g:g + gs])
        pen += pk * (alpha * bn + 0.5 * (1 - alpha) * bn ** 2)
    return loss + lmda * pen
