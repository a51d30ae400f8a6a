"""The SAMPLING workloads of BASELINE.json's configs 1-3 on one GPU (the per-call times of the same
configs are tools/bench_configs.py's): S1 4-chain NUTS on the Bayesian linear regression, S2 NUTS +
Gibbs tau2 on the 5-smooth P-spline GAM with 64 chains, S3 IWLS blocks + Gibbs tau2 with 256 chains.
Asynchronous NUTS driver (liesel_b200/nuts_async.py), reference kernel defaults. One JSON line each.

  python tools/config_samplers.py [--configs s1,s2,s3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import diagnostics, reparam, synth  # noqa: E402
from liesel_b200.nuts_gibbs import AsyncNutsGibbs, PlanEvaluator, summarise_async  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="s1,s2,s3")
args = ap.parse_args()
want = args.configs.split(",")


def nuts(name, spec, centre, C, W, S, term):
    plan = Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1 and not b.center}
    rep = reparam.sum_to_zero(spec, cm, device=plan.device) if cm else None
    a = None if aux is None or not spec.num_aux else torch.as_tensor(aux, device=plan.device)
    run = AsyncNutsGibbs(spec, PlanEvaluator(plan), torch.as_tensor(th, device=plan.device), a, W, S, reparam=rep,
                         seed=1, epoch_kw=dict(term_duration=term))
    t = run.run()
    out = summarise_async(run, t)
    out.update({"config": name, "n": spec.n, "chains": C, "warmup": W, "samples": S, "variant": plan.variant,
                "gibbs_blocks": len(run.nuts.gibbs), "ms_per_transition_posterior": 1e3 * t["seconds_posterior"] / S})
    print(json.dumps(out), flush=True)


if "s1" in want:
    spec, centre = synth.s1_blr(n=1000)
    nuts("s1: Bayesian linear regression n=1000 p=10, 4-chain NUTS", spec, centre, 4, 1000, 1000, 300)
if "s2" in want:
    # the model to sample: plain bases next to an intercept under the sum-to-zero constraint (column-centring
    # alone, the evaluation benchmark's form, leaves the constant coefficient vector unidentified)
    spec, centre = synth.s2_pspline_gam(n=100_000, center=False, intercept=True)
    nuts("s2: P-spline GAM, intercept + 5 smooths x 20 (sum-to-zero), n=1e5, NUTS + Gibbs tau2, 64 chains",
         spec, centre, 64, 400, 400, 200)
if "s3" in want:
    from liesel_b200.iwls_sampler import IWLSGibbsSampler

    spec, centre = synth.s3_gamlss(n=1_000_000)
    plan = Plan(spec)
    C = 256
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    # spline blocks sampled under the sum-to-zero constraint (beta = Z gamma): next to the intercepts the
    # unconstrained coefficients random-walk along the level ridge in ANY sampler (R-hat 14 without it)
    cons = {b.name: reparam.sum_to_zero_basis(plan.column_means(b.name)) for b in spec.blocks if b.kind == 1}
    smp = IWLSGibbsSampler(plan, torch.as_tensor(th, device=plan.device), torch.as_tensor(aux, device=plan.device), seed=1,
                           constraints=cons)
    t0 = time.perf_counter()
    smp.warmup(200)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    draws, _, infos = smp.sample(300)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ess = diagnostics.ess_bulk_device(draws)
    acc = {b.name: float(np.mean([float(i[b.name]["acceptance_prob"].mean()) for i in infos])) for b in smp.blocks}
    print(json.dumps({"config": "s3: location-scale GAMLSS n=1e6, IWLS blocks + Gibbs tau2, 256 chains", "chains": C,
                      "warmup": 200, "samples": 300, "iterations_per_sec": 300 / (t2 - t1),
                      "posterior_seconds": t2 - t1, "warmup_seconds": t1 - t0,
                      "ess_bulk_min": float(ess.min()), "ess_bulk_median": float(np.median(ess)),
                      "split_rhat_max": float(np.max(diagnostics.split_rhat_device(draws))),
                      "ess_per_sec_posterior": float(ess.min() / (t2 - t1)), "mean_accept": acc}), flush=True)
