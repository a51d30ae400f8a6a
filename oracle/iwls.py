"""
Oracle: IWLS proposal algebra and the Metropolis-Hastings step
(TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows reference ``goose/iwls.py:218-345``, ``goose/iwls_utils.py:14-70`` and
``goose/mh.py:17-71``. Random draws are inputs (standard normals ``z``, uniforms ``u``):
JAX's threefry PRNG is third-party and not reproduced.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg

from . import sam

LOG_2PI = np.log(2.0 * np.pi)


def add_ridge(info):
    """iwls.py:233-237: info + 1e-6 * mean(diag(info)) * I."""
    info = np.asarray(info)
    p = info.shape[-1]
    ridge = 1e-6 * np.mean(np.diagonal(info, axis1=-2, axis2=-1), axis=-1)
    return info + ridge[..., None, None] * np.eye(p)


def cholesky_nan(a):
    """``jnp.linalg.cholesky``: lower factor, all-NaN when the matrix is not positive definite."""
    a = np.asarray(a)
    out = np.empty_like(a)
    flat_a = a.reshape((-1,) + a.shape[-2:])
    flat_o = out.reshape((-1,) + a.shape[-2:])
    for i in range(flat_a.shape[0]):
        try:
            if not np.all(np.isfinite(flat_a[i])):
                raise np.linalg.LinAlgError
            flat_o[i] = np.linalg.cholesky(flat_a[i])
        except np.linalg.LinAlgError:
            flat_o[i] = np.nan
    return out


def chol_info(info):
    """iwls.py:232-238: ridge then Cholesky."""
    return cholesky_nan(add_ridge(info))


def safe_chol(chol, fallback: str | None = "identity"):
    """
    iwls.py:245-290 for the two fallbacks that do not need the information matrix:
    returns (chol, error_code) per chain; code 2 = identity fallback, 1 = none.
    """
    chol = np.array(chol, copy=True)
    bad = np.isnan(chol).any(axis=(-2, -1))
    codes = np.zeros(bad.shape, dtype=np.int32)
    if fallback is None:
        codes[bad] = 1
    elif fallback == "identity":
        chol[bad] = np.eye(chol.shape[-1])
        codes[bad] = 2
    else:
        raise ValueError(fallback)
    return chol, codes


def solve(chol_lhs, rhs):
    """iwls_utils.py:14-28: x with (L L^T) x = rhs via forward then backward substitution."""
    tmp = scipy.linalg.solve_triangular(chol_lhs, rhs, lower=True)
    return scipy.linalg.solve_triangular(chol_lhs.T, tmp, lower=False)


def mvn_log_prob(x, mean, chol_inv_cov):
    """iwls_utils.py:31-49."""
    standardized = (x - mean) @ chol_inv_cov
    log_prob = np.sum(-0.5 * standardized**2 - 0.5 * LOG_2PI)
    adjustment = np.sum(np.log(np.diag(chol_inv_cov)))
    return log_prob + adjustment


def mvn_sample(z, mean, chol_inv_cov):
    """
    iwls_utils.py:52-70 with the standard-normal draw ``z`` supplied:
    ``triangular_solve(chol_inv_cov, z, lower=True)`` with JAX's default ``left_side=False``
    solves x L = z, i.e. x = L^{-T} z.
    """
    centered = scipy.linalg.solve_triangular(chol_inv_cov.T, z, lower=False)
    return centered + mean


def iwls_proposal(spec, block_name, params, aux, step_size, z, fallback="identity"):
    """
    Forward half of ``IWLSKernel._standard_transition`` (iwls.py:313-327) for every chain:
    returns dict(prop [C,p], fwd_log_prob [C], mu [C,p], chol [C,p,p], error_code [C]).
    """
    blk = next(b for b in spec.blocks if b.name == block_name)
    sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    step = np.broadcast_to(np.asarray(step_size, dtype=np.float64), (params.shape[0],))
    score, info = sam.iwls(spec, block_name, params, aux)
    chol, codes = safe_chol(chol_info(info), fallback)
    C, p = score.shape
    prop = np.zeros((C, p))
    mu = np.zeros((C, p))
    fwd = np.zeros(C)
    for c in range(C):
        mu[c] = params[c, sl] + (step[c] ** 2 / 2.0) * solve(chol[c], score[c])
        prop[c] = mvn_sample(z[c], mu[c], chol[c] / step[c])
        fwd[c] = mvn_log_prob(prop[c], mu[c], chol[c] / step[c])
    return {"prop": prop, "fwd_log_prob": fwd, "mu": mu, "chol": chol, "error_code": codes}


def iwls_backward(spec, block_name, params_prop, aux, step_size, flat_pos, fallback="identity"):
    """Backward half (iwls.py:331-336): log q(pos | prop) per chain."""
    blk = next(b for b in spec.blocks if b.name == block_name)
    sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
    params_prop = np.atleast_2d(np.asarray(params_prop, dtype=np.float64))
    step = np.broadcast_to(np.asarray(step_size, dtype=np.float64), (params_prop.shape[0],))
    score, info = sam.iwls(spec, block_name, params_prop, aux)
    chol, _ = safe_chol(chol_info(info), fallback)
    C = score.shape[0]
    bwd = np.zeros(C)
    for c in range(C):
        mu = params_prop[c, sl] + (step[c] ** 2 / 2.0) * solve(chol[c], score[c])
        bwd[c] = mvn_log_prob(flat_pos[c], mu, chol[c] / step[c])
    return bwd


def mh_step(u, current_log_prob, proposed_log_prob, log_correction=0.0):
    """
    mh.py:48-68 with the uniform draw ``u`` supplied. Returns
    (error_code, acceptance_prob, do_accept), vectorised over chains.
    """
    log_acc = np.asarray(proposed_log_prob) - np.asarray(current_log_prob) + log_correction
    nan = np.isnan(log_acc)
    log_acc = np.where(nan, -np.inf, log_acc)
    code = np.where(nan, 90, 0).astype(np.int32)
    with np.errstate(over="ignore"):
        acc = np.minimum(np.exp(log_acc), 1.0)
    return code, acc, np.asarray(u) <= acc
