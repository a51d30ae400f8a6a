"""
Oracle: structured-additive-model log-posterior, gradient, IWLS score/information
(TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates, for the DistReg model family, what the reference computes when a kernel calls
``ModelMixin.log_prob_fn`` (``goose/kernel.py:149-158``) on a ``LieselInterface``
(``goose/interface.py:284-313``):

  * ``Model.update_state`` / ``Model.update`` (``model/model.py:2786-2865,2501-2526``)
    recompute, in topological order, the smooth ``Calc``s ``jnp.dot(X, beta)``
    (``model/distreg.py:105,175``), the predictor sum (``distreg.py:238-240``), the
    inverse link (``distreg.py:242``), every ``Dist`` node (``model/nodes.py:1139-1152``),
  * ``_model_log_prob = sum(node.value.sum() for node in Dist nodes)``
    (``model/model.py:50-53,232-252``).

Everything is evaluated chain by chain in float64 (or the dtype passed as ``work_dtype``,
to study float32 ordering effects). ``spec`` is duck-typed: any object with the attributes
of ``liesel_b200.spec.ModelSpec`` works; nothing from the product package is imported.
"""

from __future__ import annotations

import numpy as np

from . import densities as D
from . import splines as S

KIND_DENSE, KIND_SPLINE, KIND_GROUP, KIND_INTERCEPT = 0, 1, 2, 3
FAMILY_NORMAL, FAMILY_BERNOULLI, FAMILY_POISSON = 0, 1, 2


def _is_mvnd(prior) -> bool:
    return prior is not None and hasattr(prior, "K")


def block_band(blk):
    """(start, weights) of a spline block: precomputed if present, else the oracle's own."""
    if getattr(blk, "start", None) is not None and getattr(blk, "weights", None) is not None:
        return np.asarray(blk.start), np.asarray(blk.weights)
    return S.banded_basis(blk.x, blk.knots, blk.order)


class Prepared:
    """Per-model constants, up-cast once."""

    def __init__(self, spec, work_dtype=np.float64):
        self.spec = spec
        self.dt = np.dtype(work_dtype)
        self.y = np.asarray(spec.y, dtype=self.dt)
        self.n = self.y.shape[0]
        self.blocks = []
        for blk in spec.blocks:
            ent = {"blk": blk}
            if blk.kind == KIND_DENSE:
                ent["X"] = np.asarray(blk.X, dtype=self.dt)
            elif blk.kind == KIND_SPLINE:
                start, w = block_band(blk)
                ent["start"] = start.astype(np.int64)
                ent["w"] = w.astype(self.dt)
                ent["cols"] = ent["start"][:, None] + np.arange(w.shape[1])[None, :]
                # column-centred design B - 1 cbar^T (the dense matrix a user hands to
                # add_np_smooth, distreg.py:175); cbar as stored in the spec, else the column
                # means of B in fp64 rounded to the data dtype like every other data array
                if getattr(blk, "center", False):
                    cm = getattr(blk, "colmean", None)
                    if cm is None:
                        cm = S.band_to_dense(ent["start"], np.asarray(w, dtype=np.float64),
                                             blk.ncoef).mean(axis=0).astype(np.asarray(w).dtype)
                    ent["cbar"] = np.asarray(cm, dtype=self.dt)
            elif blk.kind == KIND_GROUP:
                ent["idx"] = np.asarray(blk.idx, dtype=np.int64)
            self.blocks.append(ent)

    # eta contribution of one block for one chain -----------------------------------
    def smooth(self, ent, beta):
        blk = ent["blk"]
        if blk.kind == KIND_DENSE:
            return ent["X"] @ beta  # jnp.dot(X, beta), distreg.py:105,175
        if blk.kind == KIND_SPLINE:
            eta = np.sum(ent["w"] * beta[ent["cols"]], axis=1)
            if "cbar" in ent:
                eta = eta - ent["cbar"] @ beta
            return eta
        if blk.kind == KIND_INTERCEPT:
            return np.full(self.n, beta[0], dtype=self.dt)  # jnp.dot(ones[n,1], beta)
        return beta[ent["idx"]]  # b[idx], SURVEY a14

    def smooth_T(self, ent, r):
        """X^T r for one block and one chain."""
        blk = ent["blk"]
        if blk.kind == KIND_DENSE:
            return ent["X"].T @ r
        if blk.kind == KIND_SPLINE:
            out = np.zeros(blk.ncoef, dtype=self.dt)
            np.add.at(out, ent["cols"].ravel(), (ent["w"] * r[:, None]).ravel())
            if "cbar" in ent:
                out = out - ent["cbar"] * r.sum()
            return out
        if blk.kind == KIND_INTERCEPT:
            return np.array([r.sum()], dtype=self.dt)
        return np.bincount(ent["idx"], weights=r, minlength=blk.ncoef).astype(self.dt)

    def dense_design(self, ent):
        blk = ent["blk"]
        if blk.kind == KIND_DENSE:
            return ent["X"]
        if blk.kind == KIND_SPLINE:
            B = S.band_to_dense(ent["start"], ent["w"], blk.ncoef)
            return B - ent["cbar"][None, :] if "cbar" in ent else B
        if blk.kind == KIND_INTERCEPT:
            return np.ones((self.n, 1), dtype=self.dt)
        raise ValueError("no dense design for group blocks")

    def predictors(self, theta):
        eta = [np.zeros(self.n, dtype=self.dt), np.zeros(self.n, dtype=self.dt)]
        for ent in self.blocks:
            blk = ent["blk"]
            beta = theta[blk.param_offset: blk.param_offset + blk.ncoef]
            # predictor Calc: sum(args), distreg.py:238-240
            eta[blk.predictor] = eta[blk.predictor] + self.smooth(ent, beta)
        return eta

    def lik_terms(self, eta):
        """
        Per-observation log-density and its first/second derivatives w.r.t. each predictor.
        Returns (ll[n], d1[2][n], w[2][n]) with w = -d2 ll / d eta^2 (observed information).
        """
        spec, y = self.spec, self.y
        zero = np.zeros(self.n, dtype=self.dt)
        if spec.family == FAMILY_NORMAL:
            has_scale = any(e["blk"].predictor == 1 for e in self.blocks)
            if has_scale:
                sigma = np.exp(eta[1])  # tfb.Exp().forward, distreg.py:242
            else:
                sigma = np.full(self.n, spec.fixed_scale, dtype=self.dt)
            ll = D.normal_log_prob(y, eta[0], sigma)
            z = (y - eta[0]) / sigma
            d_loc = z / sigma
            w_loc = 1.0 / (sigma * sigma)
            if has_scale:
                d_sc = z * z - 1.0
                w_sc = 2.0 * z * z  # observed information of the log-scale predictor
            else:
                d_sc, w_sc = zero, zero
            return ll, [d_loc, d_sc], [w_loc, w_sc]
        if spec.family == FAMILY_BERNOULLI:
            ll = D.bernoulli_logits_log_prob(y, eta[0])
            pi = 1.0 / (1.0 + np.exp(-eta[0]))
            return ll, [y - pi, zero], [pi * (1.0 - pi), zero]
        if spec.family == FAMILY_POISSON:
            ll = D.poisson_log_rate_log_prob(y, eta[0])
            mu = np.exp(eta[0])
            return ll, [y - mu, zero], [mu, zero]
        raise ValueError(spec.family)


def _tau2(blk, aux_c):
    pr = blk.prior
    if blk.tau2_slot >= 0:
        return aux_c[blk.tau2_slot]
    return pr.tau2_init


def log_prob_parts(spec, params, aux=None, work_dtype=np.float64, prep: Prepared | None = None):
    """
    Per-``Dist``-node sums, one dict entry per node, each [C]:
    response, every coefficient prior, every tau2 hyperprior (``model.py:246-251``).
    """
    prep = prep or Prepared(spec, work_dtype)
    dt = prep.dt
    params = np.atleast_2d(np.asarray(params, dtype=dt))
    C = params.shape[0]
    A = sum(1 for b in spec.blocks if b.tau2_slot >= 0)
    aux = np.zeros((C, A), dtype=dt) if aux is None else np.atleast_2d(np.asarray(aux, dtype=dt))
    parts: dict[str, np.ndarray] = {"response": np.zeros(C, dtype=dt)}
    for blk in spec.blocks:
        if blk.prior is not None:
            parts[blk.name + "_beta"] = np.zeros(C, dtype=dt)
        if _is_mvnd(blk.prior) and blk.tau2_slot >= 0 and blk.prior.ig_a is not None:
            parts[blk.prior.tau2_name] = np.zeros(C, dtype=dt)
    for c in range(C):
        theta = params[c]
        eta = prep.predictors(theta)
        ll, _, _ = prep.lik_terms(eta)
        parts["response"][c] = ll.sum()  # Dist.update per_obs vector, then .sum()
        for blk in spec.blocks:
            beta = theta[blk.param_offset: blk.param_offset + blk.ncoef]
            pr = blk.prior
            if pr is None:
                continue
            if _is_mvnd(pr):
                tau2 = _tau2(blk, aux[c])
                K = np.asarray(pr.K, dtype=dt)
                parts[blk.name + "_beta"][c] = D.mvnd_from_penalty_log_prob(
                    beta, tau2, K, rank=pr.rank, log_pdet=pr.log_pdet
                )
                if blk.tau2_slot >= 0 and pr.ig_a is not None:
                    parts[pr.tau2_name][c] = D.inverse_gamma_log_prob(tau2, pr.ig_a, pr.ig_b)
            else:
                parts[blk.name + "_beta"][c] = D.normal_log_prob(beta, pr.m, pr.s).sum()
    return parts


def log_prob(spec, params, aux=None, work_dtype=np.float64, prep=None):
    """``_model_log_prob``: Python ``sum`` over the Dist-node sums, plus cached constants."""
    parts = log_prob_parts(spec, params, aux, work_dtype, prep)
    total = 0
    for v in parts.values():
        total = total + v
    return total + spec.const_term


def log_lik(spec, params, aux=None, work_dtype=np.float64, prep=None):
    return log_prob_parts(spec, params, aux, work_dtype, prep)["response"]


def grad(spec, params, aux=None, work_dtype=np.float64, prep=None):
    """
    d log p / d params, [C, P]: what ``jax.value_and_grad(logdensity_fn)`` yields at the
    reference's call sites ``goose/nuts.py:193-194,254-260``: sum_s X_s^T (dl/d eta) plus
    prior gradients -(beta-m)/s^2 and -K beta / tau2.
    """
    prep = prep or Prepared(spec, work_dtype)
    dt = prep.dt
    params = np.atleast_2d(np.asarray(params, dtype=dt))
    C, P = params.shape
    A = sum(1 for b in spec.blocks if b.tau2_slot >= 0)
    aux = np.zeros((C, A), dtype=dt) if aux is None else np.atleast_2d(np.asarray(aux, dtype=dt))
    out = np.zeros((C, P), dtype=dt)
    for c in range(C):
        theta = params[c]
        eta = prep.predictors(theta)
        _, d1, _ = prep.lik_terms(eta)
        for ent in prep.blocks:
            blk = ent["blk"]
            sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
            g = prep.smooth_T(ent, d1[blk.predictor])
            pr = blk.prior
            if pr is not None:
                if _is_mvnd(pr):
                    g = g - (np.asarray(pr.K, dtype=dt) @ theta[sl]) / _tau2(blk, aux[c])
                else:
                    g = g - (theta[sl] - pr.m) / (pr.s * pr.s)
            out[c, sl] = g
    return out


def iwls(spec, block_name, params, aux=None, work_dtype=np.float64, prep=None):
    """
    Score and observed information of ONE coefficient block (what
    ``goose/iwls.py:315-321`` obtains from ``grad`` / ``-jacfwd(grad)`` of the flat
    log-prob of the kernel's position keys): score [C,p], info [C,p,p] = X^T W X + prior
    precision, WITHOUT the ridge of ``iwls.py:233-237``.
    """
    prep = prep or Prepared(spec, work_dtype)
    dt = prep.dt
    params = np.atleast_2d(np.asarray(params, dtype=dt))
    C = params.shape[0]
    A = sum(1 for b in spec.blocks if b.tau2_slot >= 0)
    aux = np.zeros((C, A), dtype=dt) if aux is None else np.atleast_2d(np.asarray(aux, dtype=dt))
    ent = next(e for e in prep.blocks if e["blk"].name == block_name)
    blk = ent["blk"]
    p = blk.ncoef
    sl = slice(blk.param_offset, blk.param_offset + p)
    score = np.zeros((C, p), dtype=dt)
    info = np.zeros((C, p, p), dtype=dt)
    Xd = None if blk.kind == KIND_GROUP else prep.dense_design(ent)
    for c in range(C):
        theta = params[c]
        eta = prep.predictors(theta)
        _, d1, w = prep.lik_terms(eta)
        s = prep.smooth_T(ent, d1[blk.predictor])
        if blk.kind == KIND_GROUP:
            H = np.diag(np.bincount(ent["idx"], weights=w[blk.predictor], minlength=p))
        else:
            H = Xd.T @ (Xd * w[blk.predictor][:, None])
        pr = blk.prior
        if pr is not None:
            if _is_mvnd(pr):
                K = np.asarray(pr.K, dtype=dt)
                tau2 = _tau2(blk, aux[c])
                s = s - (K @ theta[sl]) / tau2
                H = H + K / tau2
            else:
                s = s - (theta[sl] - pr.m) / (pr.s * pr.s)
                H = H + np.eye(p, dtype=dt) / (pr.s * pr.s)
        score[c], info[c] = s, H
    return score, info


def iwls_group_diag(spec, block_name, params, aux=None, work_dtype=np.float64, prep=None):
    """
    ``iwls`` for a random-intercept block with the information returned as its DIAGONAL [C, G] (the dense
    matrix ``iwls`` builds for the same block is diagonal; at G = 1e5 it cannot be formed). Also returns each
    group's share of the log-posterior [C, G]: its rows' log-likelihood terms (without the row constants)
    minus (b_g - m)^2 / (2 s^2).
    """
    prep = prep or Prepared(spec, work_dtype)
    dt = prep.dt
    params = np.atleast_2d(np.asarray(params, dtype=dt))
    C = params.shape[0]
    ent = next(e for e in prep.blocks if e["blk"].name == block_name)
    blk = ent["blk"]
    assert blk.kind == KIND_GROUP
    G = blk.ncoef
    sl = slice(blk.param_offset, blk.param_offset + G)
    score = np.zeros((C, G), dtype=dt)
    diag = np.zeros((C, G), dtype=dt)
    share = np.zeros((C, G), dtype=dt)
    for c in range(C):
        theta = params[c]
        eta = prep.predictors(theta)
        ll, d1, w = prep.lik_terms(eta)
        score[c] = np.bincount(ent["idx"], weights=d1[blk.predictor], minlength=G)
        diag[c] = np.bincount(ent["idx"], weights=w[blk.predictor], minlength=G)
        share[c] = np.bincount(ent["idx"], weights=ll, minlength=G)
        pr = blk.prior
        if pr is not None:
            d = theta[sl] - pr.m
            score[c] -= d / (pr.s * pr.s)
            diag[c] += 1.0 / (pr.s * pr.s)
            share[c] -= 0.5 * d * d / (pr.s * pr.s)
    return score, diag, share


def quadform(spec, block_name, params, work_dtype=np.float64):
    """beta^T K beta per chain: the Gibbs update's ``beta @ K @ beta`` (distreg.py:291)."""
    blk = next(b for b in spec.blocks if b.name == block_name)
    params = np.atleast_2d(np.asarray(params, dtype=work_dtype))
    beta = params[:, blk.param_offset: blk.param_offset + blk.ncoef]
    K = np.asarray(blk.prior.K, dtype=work_dtype)
    return np.einsum("ci,ij,cj->c", beta, K, beta)


def tau2_gibbs_params(spec, block_name, params, work_dtype=np.float64):
    """``tau2_gibbs_kernel`` (distreg.py:277-297): a' = a + rank/2, b' = b + 0.5 beta^T K beta."""
    blk = next(b for b in spec.blocks if b.name == block_name)
    a = blk.prior.ig_a + 0.5 * blk.prior.rank
    b = blk.prior.ig_b + 0.5 * quadform(spec, block_name, params, work_dtype)
    return a, b


def grad_scale(spec, params, aux=None, prep=None):
    """
    Rounding scale of each gradient entry: sum_i |x_ij| * |dl_i/d eta| + |prior gradient|, [C, P].
    A float32 evaluation in ANY summation order differs from the exact value by a small multiple
    of 2^-24 times this number; parity tests state their absolute tolerance in units of it.
    """
    prep = prep or Prepared(spec)
    params = np.atleast_2d(np.asarray(params, dtype=prep.dt))
    C, P = params.shape
    A = sum(1 for b in spec.blocks if b.tau2_slot >= 0)
    aux = np.zeros((C, A)) if aux is None else np.atleast_2d(np.asarray(aux, dtype=prep.dt))
    out = np.zeros((C, P))
    for c in range(C):
        theta = params[c]
        eta = prep.predictors(theta)
        _, d1, _ = prep.lik_terms(eta)
        for ent in prep.blocks:
            blk = ent["blk"]
            sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
            r = np.abs(d1[blk.predictor])
            if blk.kind == KIND_DENSE:
                g = np.abs(ent["X"]).T @ r
            elif blk.kind == KIND_SPLINE:
                g = np.zeros(blk.ncoef)
                np.add.at(g, ent["cols"].ravel(), (np.abs(ent["w"]) * r[:, None]).ravel())
                if "cbar" in ent:
                    g = g + np.abs(ent["cbar"]) * r.sum()
            elif blk.kind == KIND_INTERCEPT:
                g = np.array([r.sum()])
            else:
                g = np.bincount(ent["idx"], weights=r, minlength=blk.ncoef)
            pr = blk.prior
            if pr is not None:
                if _is_mvnd(pr):
                    g = g + (np.abs(np.asarray(pr.K)) @ np.abs(theta[sl])) / _tau2(blk, aux[c])
                else:
                    g = g + np.abs(theta[sl] - pr.m) / (pr.s * pr.s)
            out[c, sl] = g
    return out


def log_prob_scale(spec, params, aux=None, prep=None):
    """sum_i |ll_i| per chain: the rounding scale of the log-likelihood sum."""
    prep = prep or Prepared(spec)
    params = np.atleast_2d(np.asarray(params, dtype=prep.dt))
    out = np.zeros(params.shape[0])
    for c in range(params.shape[0]):
        ll, _, _ = prep.lik_terms(prep.predictors(params[c]))
        out[c] = np.abs(ll).sum()
    return out
