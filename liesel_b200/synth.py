"""
Synthetic workloads S1-S5 / H of SURVEY §8(d), generated with ``numpy.random.default_rng``
(PCG64) in float64 and then cast, so any party can regenerate them bit-for-bit without JAX.
``n`` (and ``G``) can be scaled down for parity tests; the defaults are BASELINE.json's sizes.
NumPy only — nothing here touches the GPU or evaluates a likelihood.
"""

from __future__ import annotations

import math

import numpy as np

from .spec import ModelSpec
from .splines import equidistant_knots, pspline_penalty


def _ig_log_prob(x: float, a: float, b: float) -> float:
    """Constant log-density of a hyperprior on a value that is held fixed (cached Dist node)."""
    return a * math.log(b) - math.lgamma(a) - (a + 1.0) * math.log(x) - b / x


def _greville(knots: np.ndarray, order: int = 3) -> np.ndarray:
    k = knots.shape[0] - order - 1
    return np.array([knots[i + 1: i + order + 1].mean() for i in range(k)])


def s1_blr(n: int = 1000, p: int = 10, dtype=np.float32, seed: int = 101):
    """S1: Bayesian linear regression, Normal response, sigma fixed at 1, beta ~ N(0, 100^2)."""
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(n), rng.uniform(size=(n, p - 1))])
    beta = np.linspace(-1.0, 1.0, p)
    y = X @ beta + rng.normal(size=n)
    spec = ModelSpec(dtype=np.dtype(dtype))
    spec.add_response(y, "normal").add_predictor("loc")
    spec.fixed_scale = 1.0
    spec.const_term = _ig_log_prob(1.0, 0.01, 0.01)  # sigma^2 ~ IG(0.01, 0.01), held at 1.0
    spec.add_p_smooth(X, 0.0, 100.0, "loc", "loc_p0")
    return spec, beta


def s2_pspline_gam(n: int = 100_000, n_smooth: int = 5, k: int = 20, dtype=np.float32,
                   seed: int = 102, center: bool = True, intercept: bool = False):
    """
    S2: additive model with n_smooth cubic P-splines, sigma fixed at 0.5, tau2 ~ IG(1, 0.005).
    ``center=True`` (the evaluation benchmark's form): each basis is column-centred (B - 1 cbar^T,
    SURVEY 8d). Note that column-centring alone leaves the constant coefficient vector in the null space
    of BOTH the design and the difference penalty, so that form is for timing evaluations, not for
    sampling. ``center=False, intercept=True``: plain bases next to an intercept — the model to SAMPLE,
    under the sum-to-zero constraint of ``reparam.sum_to_zero`` (the constrained design ``B Z``).
    """
    rng = np.random.default_rng(seed)
    xs = rng.uniform(size=(n_smooth, n))
    f = sum(np.sin(2 * np.pi * xs[j] * (j + 1) / 2) for j in range(n_smooth))
    y = f + rng.normal(scale=0.5, size=n)
    spec = ModelSpec(dtype=np.dtype(dtype))
    spec.add_response(y, "normal").add_predictor("loc")
    spec.fixed_scale = 0.5
    K = pspline_penalty(k, 2)
    centre = []
    if intercept:
        spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "loc", "loc_p0")
        centre.append(np.array([float(np.mean(f))]))
    for j in range(n_smooth):
        xj = xs[j].astype(dtype)
        knots = equidistant_knots(xj, k, 3)
        spec.add_spline_smooth(xj, knots, K, 1.0, 0.005, "loc", f"loc_np{j}", tau2_init=1.0,
                               center=center)
        cj = np.sin(2 * np.pi * _greville(knots.astype(np.float64)) * (j + 1) / 2)
        centre.append(cj - cj.mean() if intercept else cj)
    return spec, np.concatenate(centre)


def s3_gamlss(n: int = 1_000_000, k: int = 20, dtype=np.float32, seed: int = 103,
              center: bool = False):
    """
    S3 / H: Gaussian location-scale model; mean and log-sd are each intercept + cubic
    P-spline of the same covariate. Block order: loc_p0, loc_np0, scale_p0, scale_np0
    (P = 2k + 2).
    """
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=n)
    loc = 1.0 + np.sin(2 * np.pi * x)
    log_sd = -0.5 + 0.5 * np.cos(2 * np.pi * x)
    y = rng.normal(loc, np.exp(log_sd))
    spec = ModelSpec(dtype=np.dtype(dtype))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    K = pspline_penalty(k, 2)
    xd = x.astype(dtype)
    knots = equidistant_knots(xd, k, 3)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "loc", "loc_p0")
    spec.add_spline_smooth(xd, knots, K, 1.0, 0.005, "loc", "loc_np0", tau2_init=1.0, center=center)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "scale", "scale_p0")
    spec.add_spline_smooth(xd, knots, K, 1.0, 0.005, "scale", "scale_np0", tau2_init=1.0,
                           center=center)
    g = _greville(knots.astype(np.float64))
    centre = np.concatenate([[1.0], np.sin(2 * np.pi * g), [-0.5], 0.5 * np.cos(2 * np.pi * g)])
    return spec, centre


def spline_glm(family: str, n: int = 50_000, k: int = 20, n_smooth: int = 2, dtype=np.float32,
               seed: int = 106):
    """
    Poisson ("poisson") or Bernoulli ("bernoulli") response with an intercept and ``n_smooth`` cubic
    P-splines of one covariate each in the single predictor (north star: banded P-splines under all
    three response families). Rows can be span-sorted, so the TMA run-length kernels apply.
    """
    rng = np.random.default_rng(seed)
    spec = ModelSpec(dtype=np.dtype(dtype))
    xs = rng.uniform(size=(n_smooth, n))
    eta = 0.3 + sum(0.8 * np.sin(2 * np.pi * xs[j] * (j + 1)) for j in range(n_smooth))
    if family == "poisson":
        y = rng.poisson(np.exp(eta)).astype(np.float64)
        pred = "log_rate"
    else:
        y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
        pred = "logits"
    spec.add_response(y, family).add_predictor(pred)
    K = pspline_penalty(k, 2)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, pred, f"{pred}_p0")
    centre = [[0.3]]
    for j in range(n_smooth):
        xj = xs[j].astype(dtype)
        knots = equidistant_knots(xj, k, 3)
        spec.add_spline_smooth(xj, knots, K, 1.0, 0.005, pred, f"{pred}_np{j}", tau2_init=1.0)
        centre.append(0.8 * np.sin(2 * np.pi * _greville(knots.astype(np.float64)) * (j + 1)))
    return spec, np.concatenate(centre)


def s4_logistic(n: int = 10_000_000, p: int = 256, dtype=np.float32, seed: int = 104,
                rank: int = 0, world_size: int = 1):
    """
    S4: logistic regression, rows sharded: rank r generates rows [r*n/W, (r+1)*n/W) from
    seed + 1 + r; beta* comes from ``seed`` on every rank. Prior beta ~ N(0, 10^2).
    """
    beta = np.random.default_rng(seed).normal(size=p)
    n_loc = n // world_size
    rng = np.random.default_rng(seed + 1 + rank)
    X = np.empty((n_loc, p), dtype=dtype)
    X[:, 0] = 1.0
    X[:, 1:] = (rng.standard_normal(size=(n_loc, p - 1), dtype=np.float32) / math.sqrt(p)).astype(dtype)
    eta = X.astype(np.float64) @ beta
    y = (rng.uniform(size=n_loc) < 1.0 / (1.0 + np.exp(-eta))).astype(dtype)
    spec = ModelSpec(dtype=np.dtype(dtype))
    spec.add_response(y, "bernoulli").add_predictor("logits")
    spec.add_p_smooth(X, 0.0, 10.0, "logits", "logits_p0")
    return spec, beta


def s5_poisson_ri(n: int = 5_000_000, G: int = 100_000, p: int = 8, dtype=np.float32,
                  seed: int = 105, balanced: bool = True):
    """
    S5: Poisson regression with a random intercept per group; rows sorted by group.
    beta ~ N(0, 10^2), b_g ~ N(0, 0.5^2) with sigma_b held fixed.
    """
    rng = np.random.default_rng(seed)
    if balanced:
        idx = np.repeat(np.arange(G), n // G)
        n = idx.shape[0]
    else:
        idx = np.sort(rng.integers(0, G, size=n))
    X = np.column_stack([np.ones(n), rng.normal(size=(n, p - 1))]) * 0.3
    beta = np.linspace(-0.5, 0.5, p)
    b = rng.normal(scale=0.5, size=G)
    y = rng.poisson(np.exp(X @ beta + b[idx])).astype(np.float64)
    spec = ModelSpec(dtype=np.dtype(dtype))
    spec.add_response(y, "poisson").add_predictor("log_rate")
    spec.add_p_smooth(X, 0.0, 10.0, "log_rate", "log_rate_p0")
    spec.add_random_intercept(idx, G, 0.0, 0.5, "log_rate", "log_rate_ri0")
    return spec, np.concatenate([beta, b])


def evaluation_points(spec: ModelSpec, centre: np.ndarray, C: int, seed: int = 777,
                      tau2: str = "one"):
    """
    params [C, P] = centre + 0.1 N(0,1); aux [C, A]: tau2 = 1 ("one"), the reference's start
    value 10000 ("start", distreg.py:163) or LogUniform(1e-3, 1e4) per chain ("loguniform").
    """
    rng = np.random.default_rng(seed)
    params = centre[None, :] + 0.1 * rng.normal(size=(C, spec.num_params))
    A = spec.num_aux
    if tau2 == "one":
        aux = np.ones((C, A))
    elif tau2 == "start":
        aux = np.full((C, A), 10000.0)
    elif tau2 == "loguniform":
        aux = np.exp(rng.uniform(math.log(1e-3), math.log(1e4), size=(C, A)))
    else:
        raise ValueError(tau2)
    return params.astype(spec.dtype), aux.astype(spec.dtype)
