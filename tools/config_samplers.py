"""The SAMPLING workloads of BASELINE.json's configs 1-3 and 5 on one GPU (the per-call times of the same
configs are tools/bench_configs.py's): S1 4-chain NUTS on the Bayesian linear regression, S2 NUTS +
Gibbs tau2 on the 5-smooth P-spline GAM with 64 chains, S3 IWLS blocks + Gibbs tau2 with 256 chains.
S5 (--configs s5): IWLS blocks for the fixed effects (dense 8 x 8 information) and the random intercept
(G = 1e5, diagonal information: lslb_iwls_group) of the Poisson model, n = 5e6.
Asynchronous NUTS driver (liesel_b200/nuts_async.py), reference kernel defaults. One JSON line each.

  python tools/config_samplers.py [--configs s1,s2,s3,s5] [--s5-chains 256]
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
ap.add_argument("--s5-chains", type=int, default=256)
ap.add_argument("--s5-scale", type=float, default=1.0, help="scales n and G of S5")
ap.add_argument("--s5-group-accept", default="per_group", choices=["joint", "per_group"])
ap.add_argument("--s5-warmup", type=int, default=100)
ap.add_argument("--s5-samples", type=int, default=150)
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

if "s5" in want:
    from liesel_b200.iwls_sampler import IWLSGibbsSampler

    spec, centre = synth.s5_poisson_ri(n=int(5_000_000 * args.s5_scale), G=int(100_000 * args.s5_scale))
    plan = Plan(spec)
    C = args.s5_chains
    th, _ = synth.evaluation_points(spec, centre, C, seed=1)
    smp = IWLSGibbsSampler(plan, torch.as_tensor(th, device=plan.device), None, seed=1, initial_step_size=0.1,
                           group_accept=args.s5_group_accept)
    W, S = args.s5_warmup, args.s5_samples
    # P = 100 008 coordinates per chain: keep the fixed effects and 64 evenly spaced random intercepts
    G = spec.blocks[-1].ncoef
    keep = torch.as_tensor(np.concatenate([np.arange(8), 8 + np.linspace(0, G - 1, 64).astype(np.int64)]), device=plan.device)
    t0 = time.perf_counter()
    smp.warmup(W)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    kept = torch.empty((S, C, keep.numel()), dtype=smp.params.dtype, device=plan.device)
    accs = {b.name: [] for b in smp.blocks}
    for t in range(S):
        infos = smp.sweep()
        kept[t] = smp.params[:, keep]
        for k, v in infos.items():
            accs[k].append(float(v["acceptance_prob"].mean()))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ess = diagnostics.ess_bulk_device(kept)
    print(json.dumps({"config": f"s5: Poisson random intercept n={spec.n} G={G}, IWLS blocks (dense fixed effects + diagonal random intercept), {C} chains",
                      "chains": C, "warmup": W, "samples": S, "iterations_per_sec": S / (t2 - t1),
                      "posterior_seconds": t2 - t1, "warmup_seconds": t1 - t0,
                      "coordinates_summarised": int(keep.numel()), "group_accept": args.s5_group_accept,
                      "ess_bulk_min": float(ess.min()), "ess_bulk_median": float(np.median(ess)),
                      "ess_bulk_min_fixed_effects": float(ess[:8].min()), "ess_bulk_min_random_intercepts": float(ess[8:].min()),
                      "split_rhat_max": float(np.max(diagnostics.split_rhat_device(kept))),
                      "ess_per_sec_posterior": float(ess.min() / (t2 - t1)),
                      "step_size_median": {k: float(v.median()) for k, v in smp.step.items()},
                      "mean_accept": {k: float(np.mean(v)) for k, v in accs.items()}}), flush=True)
