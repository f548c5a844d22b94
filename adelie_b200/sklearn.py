"""scikit-learn style estimator over ``grpnet`` / ``cv_grpnet`` (reference: adelie/sklearn.py:41-262, ``GroupElasticNet``).
Same constructor arguments, fitted attributes (``state_``, ``glm_``, ``coef_``, ``intercept_``, ``lambda_``) and prediction rules; the
covariance-based ``CSSModelSelection`` of the reference module is out of scope."""
from __future__ import annotations

import numpy as np
from scipy.special import expit, softmax
from sklearn.base import BaseEstimator, RegressorMixin

from . import glm as _glm
from .cv import CVGrpnetResult, cv_grpnet
from .diagnostic import predict
from .solver import grpnet

_FAMILIES = {"gaussian": _glm.gaussian, "binomial": _glm.binomial, "poisson": _glm.poisson, "multigaussian": _glm.multigaussian,
             "multinomial": _glm.multinomial}


class GroupElasticNet(BaseEstimator, RegressorMixin):
    """Group elastic net estimator.  ``solver`` in {"grpnet", "cv_grpnet"}; ``family`` in {"gaussian", "binomial", "poisson",
    "multigaussian", "multinomial"}."""
    def __init__(self, solver: str = "grpnet", family: str = "gaussian"):
        self.solver = solver
        self.family = family

    def fit(self, X, y, **kwargs):
        if self.solver not in ("grpnet", "cv_grpnet"):
            raise ValueError("solver must be one of 'grpnet', 'cv_grpnet'.")
        if self.family not in _FAMILIES:
            raise ValueError(f"family must be one of {sorted(_FAMILIES)}.")
        self.glm_ = _FAMILIES[self.family](np.asarray(y, dtype=np.float64) if np.asarray(y).dtype.kind in "iub" else y)
        kwargs.setdefault("progress_bar", False)
        if self.solver == "cv_grpnet":
            cv = cv_grpnet(X=X, glm=self.glm_, **kwargs)
            assert isinstance(cv, CVGrpnetResult)
            fit_kw = {k: v for k, v in kwargs.items() if k not in ("n_folds", "seed", "min_ratio", "lmda_path_size", "early_exit")}
            self.cv_result_ = cv
            self.state_ = cv.fit(X=X, glm=self.glm_, **fit_kw)           # refit down to the best lambda (adelie/sklearn.py:131-142)
            self.coef_ = self.state_.betas[-1]
            self.intercept_ = np.array([self.state_.intercepts[-1]])
            self.lambda_ = np.array([self.state_.lmdas[-1]])
        else:
            self.state_ = grpnet(X=X, glm=self.glm_, **kwargs)
            self.coef_ = self.state_.betas
            self.intercept_ = self.state_.intercepts
            self.lambda_ = self.state_.lmdas
        return self

    def _check_fitted(self):
        if not hasattr(self, "state_"):
            raise RuntimeError("The model has not been fitted yet. Call fit() first.")

    def predict_proba(self, X):
        self._check_fitted()
        if self.family not in ("binomial", "multinomial"):
            raise ValueError("predict_proba is only available for \"binomial\" and \"multinomial\" families.")
        linear_pred = predict(X, self.coef_, self.intercept_)
        if self.family == "binomial":
            proba = expit(linear_pred)
            return np.stack((1 - proba, proba), axis=-1).squeeze()
        return softmax(linear_pred, axis=-1).squeeze()

    def predict(self, X):
        self._check_fitted()
        if self.family in ("binomial", "multinomial"):
            return np.argmax(self.predict_proba(X), axis=-1).squeeze()
        return predict(X, self.coef_, self.intercept_).squeeze()

    def score(self, X, y, sample_weight=None):
        """R^2 of the (last-lambda) linear predictions for the regression families; accuracy for the classification families."""
        self._check_fitted()
        pred = self.predict(X)
        y = np.asarray(y)
        if self.family in ("binomial", "multinomial"):
            labels = y if y.ndim == 1 else np.argmax(y, axis=-1)
            pred = pred[-1] if pred.ndim > labels.ndim else pred       # grpnet solver: one prediction per lambda, score the last
            return float(np.mean(pred == labels))
        pred = pred[-1] if pred.ndim > y.ndim else pred
        w = np.ones(y.shape[0]) if sample_weight is None else np.asarray(sample_weight)
        w = w.reshape((-1,) + (1,) * (y.ndim - 1))
        ss_res = np.sum(w * (y - pred) ** 2); ss_tot = np.sum(w * (y - np.average(y, axis=0, weights=w.ravel())) ** 2)
        return float(1 - ss_res / ss_tot)
