"""K-fold cross-validated group elastic net (reference: adelie/cv.py:130-325).

The procedure is the reference's: a common lambda grid from the full-data lambda_max, one path per training fold (fold = zero
weights on the held-out observations, ``glm.reweight``), augmented by the fold's own larger lambdas, coefficients interpolated back onto
the common grid, held-out loss = (full-data loss - weights_sum * training loss) / held-out weight.  What differs is where X lives: a
NumPy ``X`` is uploaded ONCE and the ``n_folds + 1`` lambda_max solves, the ``n_folds`` paths and the predictions all run on the same
device-resident matrix (the reference re-wraps ``X_raw`` for every call).  Single-response families.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse

from . import matrix as _matrix
from .diagnostic import coefficient, predict
from .solver import grpnet


@dataclass
class CVGrpnetResult:
    """Result of K-fold CV group elastic net (adelie/cv.py:25-127, without the plotting helpers)."""
    lmdas: np.ndarray        # common regularization path used for all folds
    losses: np.ndarray       # losses[k, i]: CV loss when validating on fold k at lmdas[i]
    avg_losses: np.ndarray   # average CV loss at lmdas[i]
    best_idx: int            # argmin of avg_losses

    def fit(self, X, glm, *, lmda_path_size: int = 100, **grpnet_params):
        """Refit on the full data down to the best lambda (adelie/cv.py:96-127)."""
        lm = self.lmdas[0] * np.logspace(0, np.log10(self.lmdas[self.best_idx] / self.lmdas[0]), lmda_path_size) if self.best_idx > 0 \
            else self.lmdas[:1]
        return grpnet(X=X, glm=glm, lmda_path=lm, early_exit=False, **grpnet_params)


def cv_grpnet(X, glm, *, n_threads: int = 1, early_exit: bool = False, min_ratio: float = 1e-1, lmda_path_size: int = 100,
              n_folds: int = 5, seed: int = None, **grpnet_params):
    # multi-response families (adelie/cv.py:92, is_multi): betas are (L, p K), intercepts (L, K), etas (L, n, K); everything below is
    # shape-agnostic (diagnostic.predict / coefficient handle both layouts)
    if isinstance(X, np.ndarray):
        X = _matrix.dense(X, method="naive", n_threads=n_threads)        # uploaded once, shared by every solve below
    assert isinstance(X, _matrix.MatrixNaiveBase)
    n = X.rows()
    if seed is not None:
        np.random.seed(seed)
    order = np.random.choice(n, n, replace=False)
    fold_size, remaining = divmod(n, n_folds)
    grpnet_params.pop("progress_bar", None)
    init_kw = {k: grpnet_params[k] for k in ("groups", "alpha", "penalty", "offsets", "intercept") if k in grpnet_params}

    state = grpnet(X=X, glm=glm, n_threads=n_threads, lmda_path_size=0, progress_bar=False, **init_kw)
    full_lmdas = state.lmda_max * np.logspace(0, np.log10(min_ratio), lmda_path_size)
    cv_losses = np.empty((n_folds, full_lmdas.shape[0]))
    for fold in range(n_folds):
        begin = (fold_size + 1) * min(fold, remaining) + max(fold - remaining, 0) * fold_size
        held = order[begin:begin + fold_size + (fold < remaining)]
        weights = glm.weights.copy()
        weights[held] = 0
        weights_sum = np.sum(weights)
        glm_c = glm.reweight(weights / weights_sum)
        state = grpnet(X=X, glm=glm_c, n_threads=n_threads, lmda_path_size=0, progress_bar=False, **init_kw)
        curr = state.lmda_max * np.logspace(0, np.log10(min_ratio), lmda_path_size)
        aug_lmdas = np.sort(np.concatenate([full_lmdas, curr[curr > full_lmdas[0]]]))[::-1]
        state = grpnet(X=X, glm=glm_c, ddev_tol=0, n_threads=n_threads, early_exit=early_exit, lmda_path=aug_lmdas, progress_bar=False,
                       **grpnet_params)
        held_weight = np.sum(glm.weights[held])
        pairs = [coefficient(lmda=lm, betas=state.betas, intercepts=state.intercepts, lmdas=state.lmdas) for lm in full_lmdas]
        etas = predict(X=X, betas=scipy.sparse.vstack([b for b, _ in pairs]), intercepts=np.array([b0 for _, b0 in pairs]),
                       offsets=state._offsets, n_threads=n_threads)
        etas = np.ascontiguousarray(etas, dtype=glm.dtype)                 # (L, n) or (L, n, K): one contiguous eta per lambda
        full_losses = np.array([glm.loss(eta) for eta in etas])
        train_losses = weights_sum * np.array([glm_c.loss(eta) for eta in etas])
        cv_losses[fold] = (full_losses - train_losses) / held_weight if held_weight > 0 else 0
    avg_losses = np.mean(cv_losses, axis=0)
    return CVGrpnetResult(lmdas=full_lmdas, losses=cv_losses, avg_losses=avg_losses, best_idx=int(np.argmin(avg_losses)))
