"""
Oracle: log-densities (TEST INFRASTRUCTURE — see oracle/__init__.py).

The response and prior densities of the reference are evaluated by
``tensorflow_probability.substrates.jax.distributions`` (tfp-nightly 0.26.0.dev20260528,
``uv.lock:2323-2325``), which is NOT under /root/reference. Call sites:
``model/nodes.py:1117-1122,1145`` (``Dist.update``), ``model/distreg.py:102,162,270``.
The closed forms below are TFP's published definitions; tests check them against
``scipy.stats`` and against every number the reference's docs/tests print.
"""

from __future__ import annotations

import numpy as np
from scipy.special import gammaln

LOG_2PI = np.log(2.0 * np.pi)


def normal_log_prob(y, loc, scale):
    """tfd.Normal: -0.5*(y/scale - loc/scale)^2 - 0.5*log(2*pi) - log(scale)."""
    y, loc, scale = np.broadcast_arrays(np.asarray(y), np.asarray(loc), np.asarray(scale))
    z = y / scale - loc / scale
    return -0.5 * z * z - (0.5 * LOG_2PI + np.log(scale))


def softplus(x):
    x = np.asarray(x)
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def bernoulli_logits_log_prob(y, logits):
    """tfd.Bernoulli(logits): y*logits - softplus(logits)."""
    return np.asarray(y) * np.asarray(logits) - softplus(logits)


def poisson_log_rate_log_prob(y, log_rate):
    """tfd.Poisson(log_rate): y*log_rate - exp(log_rate) - lgamma(y + 1)."""
    y = np.asarray(y)
    return y * np.asarray(log_rate) - np.exp(log_rate) - gammaln(y + 1.0)


def inverse_gamma_log_prob(x, a, b):
    """tfd.InverseGamma(concentration=a, scale=b): a*log b - lgamma(a) - (a+1)*log x - b/x."""
    x = np.asarray(x)
    return a * np.log(b) - gammaln(a) - (a + 1.0) * np.log(x) - b / x


# --- rank-deficient multivariate normal (reference distributions/mvn_degen.py) ----------


def rank_from_evals(evals, tol: float = 1e-6):
    """mvn_degen.py:20-29."""
    return np.sum(np.asarray(evals) > tol, axis=-1)


def log_pdet_from_evals(evals, rank=None, tol: float = 1e-6):
    """
    mvn_degen.py:32-56. With ``rank`` given: the entries i >= len - rank of the ascending
    eigenvalue array (:47-52), no tolerance test. Without: eigenvalues > tol (:45).
    """
    evals = np.asarray(evals)
    if rank is None:
        mask = evals > tol
    else:
        max_index = evals.shape[-1] - rank
        mask = np.arange(evals.shape[-1]) >= max_index
    selected = np.where(mask, evals, 1.0)
    return np.sum(np.log(selected), axis=-1)


def mvnd_from_penalty_log_prob(x, var, pen, rank=None, log_pdet=None, loc=0.0):
    """
    ``MultivariateNormalDegenerate.from_penalty(loc, var, pen, rank, log_pdet).log_prob(x)``:
    prec = pen / var (:268); eigvalsh(pen) when rank or log_pdet is missing (:270-273);
    log_pdet_prec = log_pdet - rank*log(var) (:275);
    log_prob = 0.5 * (-(x-loc)^T prec (x-loc) - (rank*log(2*pi) - log_pdet_prec)) (:436-444),
    product order (x @ prec) @ x^T (:442).
    """
    pen = np.asarray(pen)
    x = np.asarray(x) - loc
    prec = pen / var
    if rank is None or log_pdet is None:
        evals = np.linalg.eigvalsh(pen)
        rank = rank_from_evals(evals) if rank is None else rank
        log_pdet = log_pdet_from_evals(evals, rank=rank) if log_pdet is None else log_pdet
    log_pdet_prec = log_pdet - rank * np.log(var)
    prob1 = -((x @ prec) @ x)
    prob2 = rank * np.log(2 * np.pi) - log_pdet_prec
    return 0.5 * (prob1 - prob2)
