"""
A Gaussian location-scale model with two P-splines, sampled on one B200 — the Liesel workflow
(`DistRegBuilder` -> `EngineBuilder` with a `NUTSKernel` + `tau2_gibbs_kernel`s,
docs/source/tutorials of the reference) on this repo's host mirror:

    spec  = ModelSpec ...            # same call shape as DistRegBuilder (model/distreg.py:64-274)
    plan  = Plan(spec)               # uploads the data once, builds the banded bases on the device
    iface = B200Interface(plan)      # goose ModelInterface protocol (goose/types.py:72-102)
    run   = AsyncNutsGibbs(...)      # one NUTS block + Gibbs tau2, every chain on its own tree

    python examples/gamlss_nuts.py [--n 100000] [--chains 256]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import diagnostics, reparam  # noqa: E402
from liesel_b200.interface import B200Interface  # noqa: E402
from liesel_b200.nuts_gibbs import AsyncNutsGibbs, PlanEvaluator, summarise_async  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402
from liesel_b200.spec import ModelSpec  # noqa: E402
from liesel_b200.splines import equidistant_knots, pspline_penalty  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--chains", type=int, default=256)
ap.add_argument("--warmup", type=int, default=300)
ap.add_argument("--samples", type=int, default=300)
args = ap.parse_args()

# ---- data: y ~ N(1 + sin(2 pi x), exp(-0.5 + 0.5 cos(2 pi x))^2) -------------------------------------
rng = np.random.default_rng(1)
x = rng.uniform(size=args.n).astype(np.float32)
y = rng.normal(1.0 + np.sin(2 * np.pi * x), np.exp(-0.5 + 0.5 * np.cos(2 * np.pi * x))).astype(np.float32)

# ---- model: intercept + cubic P-spline (20 coefficients, 2nd-order penalty, IG(1, 0.005) on tau2) for
#      the mean and for the log standard deviation ---------------------------------------------------------
k = 20
knots = equidistant_knots(x, k, 3)
K = pspline_penalty(k, 2)
spec = ModelSpec(dtype=np.dtype(np.float32))
spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale", inverse_link="Exp")
spec.add_p_smooth(np.ones((args.n, 1), dtype=np.float32), 0.0, 100.0, "loc", "loc_p0")
spec.add_spline_smooth(x, knots, K, 1.0, 0.005, "loc", "loc_np0", tau2_init=1.0)
spec.add_p_smooth(np.ones((args.n, 1), dtype=np.float32), 0.0, 100.0, "scale", "scale_p0")
spec.add_spline_smooth(x, knots, K, 1.0, 0.005, "scale", "scale_np0", tau2_init=1.0)

plan = Plan(spec)
iface = B200Interface(plan)
print("kernel variant:", plan.variant, "| parameters:", plan.P, "| position keys:", spec.position_keys())

# ---- the goose protocol: state dict in, log-posterior (and gradient) out --------------------------------
state = iface.initial_state(args.chains)
logp, grad = iface.log_prob_and_grad(state)
print("log-posterior at the start state:", float(logp[0]))

# ---- sampling: sum-to-zero constraint on the smooths (they share a level with the intercepts) ----------
cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1}
rep = reparam.sum_to_zero(spec, cm, device=plan.device)
params, aux = iface.pack(state)
run = AsyncNutsGibbs(spec, PlanEvaluator(plan), params, aux, args.warmup, args.samples, reparam=rep, seed=1,
                     epoch_kw=dict(term_duration=args.warmup // 2))
timing = run.run()
summary = summarise_async(run, timing)
print(json.dumps({k_: summary[k_] for k_ in ("ess_bulk_min", "split_rhat_max", "mean_accept", "mean_treedepth",
                                            "divergent_frac", "posterior_seconds", "ess_per_sec_posterior")}))

# ---- posterior mean of the fitted mean curve at a few covariate values ----------------------------------
draws = run.full(run.draws().reshape(-1, run.draws().shape[2])).double().mean(dim=0).cpu().numpy()
from liesel_b200.splines import basis_matrix  # noqa: E402

xs = np.linspace(0.05, 0.95, 10)
blk = spec.block("loc_np0")
start, w = basis_matrix(xs.astype(np.float32), knots, 3)  # banded: first column + 4 weights per row
start, w = start.cpu().numpy(), w.double().cpu().numpy()
beta = draws[blk.param_offset: blk.param_offset + blk.ncoef]
fit = draws[spec.block("loc_p0").param_offset] + sum(w[:, t] * beta[start + t] for t in range(4))
print("fitted mean :", np.round(fit, 3))
print("truth       :", np.round(1.0 + np.sin(2 * np.pi * xs), 3))
