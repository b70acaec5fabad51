"""Host-side metric helpers with the reference's names and signatures (src/evaluation/metrics.py:6-66).

The reference scripts call these on NumPy arrays that ``predict`` already returned
(``from src.evaluation.metrics import rmse, macro_mae``, compare_models.py:23), so they stay plain NumPy;
the fused device versions used inside ``fit`` / ``evaluate_*`` are ``pmf_eval_stats``.
"""
from __future__ import annotations

from math import lgamma

import numpy as np


def rmse(y_true, y_pred):
    return np.sqrt(np.mean((y_true - y_pred) ** 2))


def mae(y_true, y_pred):
    return np.mean(np.abs(y_true - y_pred))


def macro_mae(y_true, y_pred):
    """Mean over the distinct true values of the per-value MAE (metrics.py:37-51)."""
    per = [np.mean(np.abs(y_true[y_true == lab] - y_pred[y_true == lab])) for lab in np.unique(y_true)]
    return np.mean(per)


def GaussianLogPredictiveLikelihood(df, theta, beta, sigma):
    """metrics.py:18-35 -- note the reference squares ``sigma`` although callers pass a variance."""
    predictions = np.sum(theta[df.u] * beta[df.i], axis=1)
    squared_errors = (df.rating - predictions) ** 2
    variance = sigma ** 2
    return np.sum(-0.5 * np.log(2 * np.pi * variance) - squared_errors / (2 * variance))


def PoissonLogPredictiveLikelihood(df, theta, beta, epsilon=1e-10):
    """metrics.py:53-66."""
    lambdas = np.maximum(np.sum(theta[df.u] * beta[df.i], axis=1), epsilon)
    lg = np.vectorize(lgamma, otypes=[float])(np.asarray(df.rating, dtype=float) + 1.0)
    return np.sum(df.rating * np.log(lambdas) - lambdas - lg)
