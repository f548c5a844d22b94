"""Host-side diagnostics used by tests and callers (reference: adelie/diagnostic.py:30-276)."""
import numpy as np


def predict(X, betas, intercepts, offsets=None, n_threads=1):
    """Linear predictions eta = X beta + intercept (+ offsets) for every lambda (adelie/diagnostic.py:30-121).  ``X`` is a NumPy array or
    a device matrix (then ``sp_tmul`` runs on the device); single-response."""
    from . import matrix as _matrix
    intercepts = np.atleast_1d(np.asarray(intercepts))
    if intercepts.ndim == 2:
        # multi-response: coefficient (feature j, class l) sits at column j * K + l of kron(X, I_K); class l uses the columns l::K
        from scipy.sparse import csr_matrix
        K = intercepts.shape[1]
        B = csr_matrix(betas)
        per_class = [predict(X, B[:, l::K], intercepts[:, l]) for l in range(K)]
        eta = np.stack(per_class, axis=-1)                    # (L, n, K)
        if offsets is not None:
            eta = eta + np.asarray(offsets)[None]
        return eta
    if isinstance(X, _matrix.MatrixNaiveBase):
        from scipy.sparse import csr_matrix
        B = csr_matrix(betas)
        eta = np.empty((B.shape[0], X.rows()), dtype=X.dtype)
        X.sp_tmul(B.astype(X.dtype), eta)
        eta = eta + intercepts[:, None].astype(X.dtype)
    else:
        X = np.asarray(X)
        B = np.asarray(betas.todense()) if hasattr(betas, "todense") else np.asarray(betas)
        eta = B @ X.T + intercepts[:, None]
    if offsets is not None:
        eta = eta + np.asarray(offsets)[None]
    return eta


def coefficient(*, lmda, betas, intercepts, lmdas):
    """Coefficients at ``lmda`` by linear interpolation between the two neighbouring solutions of a (decreasing) path; the boundary
    solution outside the range of ``lmdas`` (adelie/diagnostic.py:560-646)."""
    lmdas = np.asarray(lmdas)
    if lmdas.shape[0] == 0:
        raise RuntimeError("lmdas must be non-empty!")
    if lmdas.shape[0] == 1:
        return betas, intercepts
    order = np.argsort(lmdas)
    idx = lmdas.shape[0] - int(np.searchsorted(lmdas, lmda, sorter=order))
    if idx == 0 or idx == lmdas.shape[0]:
        idx = int(np.clip(idx, 0, lmdas.shape[0] - 1))
        return betas[idx], intercepts[idx]
    weight = (lmda - lmdas[idx]) / (lmdas[idx - 1] - lmdas[idx])
    beta = betas[idx - 1].multiply(weight) + betas[idx].multiply(1 - weight)
    intercept = weight * intercepts[idx - 1] + (1 - weight) * intercepts[idx]
    return beta, intercept


def objective_gaussian(X, y, weights, beta, intercept, lmda, alpha, groups, group_sizes, penalty):
    """0.5 sum w (y - X beta - b0)^2 + lmda sum_g p_g (alpha ||b_g|| + 0.5 (1-alpha) ||b_g||^2)."""
    r = y - X @ beta - intercept
    loss = 0.5 * np.sum(weights * r ** 2)
    pen = 0.0
    for g, gs, pk in zip(groups, group_sizes, penalty):
        bn = np.linalg.norm(beta[g:g + gs])
        pen += pk * (alpha * bn + 0.5 * (1 - alpha) * bn ** 2)
    return loss + lmda * pen


def compute_penalty(groups, group_sizes, penalty, alpha, betas):
    """``sum_g penalty_g (alpha ||beta_g||_2 + (1 - alpha) / 2 ||beta_g||_2^2)`` for every row of ``betas`` (dense or CSR)
    (reference: solver.compute_penalty_{dense,sparse}, adelie/src/py_solver.cpp:81-88, adelie_core/solver/utils.hpp)."""
    B = np.asarray(betas.todense()) if hasattr(betas, "todense") else np.atleast_2d(np.asarray(betas))
    out = np.zeros(B.shape[0], dtype=B.dtype)
    for g, gs, pk in zip(groups, group_sizes, penalty):
        nrm = np.linalg.norm(B[:, g:g + gs], axis=1)
        out += pk * (alpha * nrm + 0.5 * (1 - alpha) * nrm ** 2)
    return out


def objective(X, glm, betas, intercepts, lmdas, *, groups=None, alpha=1, penalty=None, offsets=None, relative=True, add_penalty=True,
              n_threads=1):
    """Group elastic net objective ``loss(eta) [- loss_full] + lmda * penalty(beta)`` at every lambda (adelie/diagnostic.py:124-276);
    ``X`` is a NumPy array or a device matrix, the loss is evaluated by the GLM object (on the device)."""
    from . import matrix as _matrix
    intercepts = np.atleast_1d(np.asarray(intercepts))
    K = intercepts.shape[1] if intercepts.ndim == 2 else 1
    p = (X.shape[1] if isinstance(X, np.ndarray) else X.cols()) * K
    if groups is None:
        groups = np.arange(p // K) if K == 1 else K * np.arange(p // K)
    elif K > 1:
        groups = np.asarray(groups) * K
    groups = np.asarray(groups, dtype=int)
    group_sizes = np.diff(np.concatenate([groups, [p]]))
    if penalty is None:
        penalty = np.sqrt(group_sizes)
    Xd = _matrix.dense(np.asfortranarray(X, dtype=glm.dtype), n_threads=n_threads) if isinstance(X, np.ndarray) else X
    etas = predict(Xd, betas, intercepts, offsets=offsets, n_threads=n_threads)
    objs = np.array([glm.loss(np.ascontiguousarray(eta, dtype=glm.dtype)) for eta in etas], dtype=np.float64)
    if relative:
        objs -= glm.loss_full()
    if add_penalty:
        objs += np.asarray(lmdas) * compute_penalty(groups, group_sizes, penalty, alpha, betas)
    return objs
