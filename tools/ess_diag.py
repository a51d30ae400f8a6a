"""Where the NUTS-within-Gibbs run on H spends its transitions: tree-depth histogram over chains x
iterations, per-chain step sizes after warm-up, sequential batched evaluations per transition and the
time per evaluation — the numbers behind DESIGN.md's discussion of the ESS/sec run.

  python tools/ess_diag.py --chains 256 --warmup 300 --samples 100
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import reparam, synth  # noqa: E402
from liesel_b200.nuts_gibbs import AsyncNutsGibbs, NutsGibbs, PlanEvaluator, summarise, summarise_async  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--chains", type=int, default=256)
ap.add_argument("--warmup", type=int, default=300)
ap.add_argument("--samples", type=int, default=100)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--no-compact", action="store_true")
ap.add_argument("--term", type=int, default=50)
ap.add_argument("--async", dest="use_async", action="store_true",
                help="asynchronous driver (nuts_async.AsyncNUTS): per-chain trees, the batch shares the evaluation")
args = ap.parse_args()

spec, centre = synth.s3_gamlss(n=args.n)
plan = Plan(spec)
dev = plan.device
cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1}
rep = reparam.sum_to_zero(spec, cm, device=dev)
C = args.chains
th0, aux0 = synth.evaluation_points(spec, centre, C, seed=args.seed)
if args.use_async:
    run = AsyncNutsGibbs(spec, PlanEvaluator(plan), torch.as_tensor(th0, device=dev), torch.as_tensor(aux0, device=dev),
                         args.warmup, args.samples, reparam=rep, seed=args.seed, epoch_kw=dict(term_duration=args.term))
    timing = run.run()
    out = summarise_async(run, timing)
    n = run.nuts
    W, S = args.warmup, args.samples
    out.update({
        "driver": "async", "chains": C, "warmup": W, "samples": S,
        "ms_per_transition_posterior": 1e3 * timing["seconds_posterior"] / S,
        "ms_per_transition_warmup": 1e3 * timing["seconds_warmup"] / W,
        "ms_per_batched_eval": 1e3 * timing["seconds_posterior"] / timing["steps_posterior"],
        "batched_evals_per_transition": timing["steps_posterior"] / S,
        "step_size_quantiles": dict(zip(["0", "0.5", "1"], np.quantile(n.st["step"].cpu().numpy(), [0, 0.5, 1]).tolist())),
    })
    print(json.dumps(out), flush=True)
    sys.exit(0)
run = NutsGibbs(spec, PlanEvaluator(plan), torch.as_tensor(th0, device=dev), torch.as_tensor(aux0, device=dev),
                reparam=rep, seed=args.seed, compact=not args.no_compact)
t0 = time.perf_counter()
winfo = run.warmup(args.warmup, term_duration=args.term)
torch.cuda.synchronize()
t1 = time.perf_counter()
draws, infos = run.sample(args.samples)
torch.cuda.synchronize()
t2 = time.perf_counter()
out = summarise(draws, infos, t2 - t1, t1 - t0)
nuts = run.nuts
depth = torch.stack([i.treedepth for i in infos]).cpu().numpy()  # [S, C]
eps = nuts.step_size.cpu().numpy()
q = [0, 0.01, 0.1, 0.5, 0.9, 0.99, 1]
mean_depth_chain = depth.mean(axis=0)
out.update({
    "chains": C, "warmup": args.warmup, "samples": args.samples,
    "depth_hist": np.bincount(depth.ravel(), minlength=11).tolist(),
    "step_size_quantiles": dict(zip(map(str, q), np.quantile(eps, q).tolist())),
    "mean_depth_per_chain_quantiles": dict(zip(map(str, q), np.quantile(mean_depth_chain, q).tolist())),
    "corr_log_eps_depth": float(np.corrcoef(np.log(eps), mean_depth_chain)[0, 1]),
    "inv_mass_quantiles": dict(zip(map(str, q), np.quantile(nuts.inv_mass.cpu().numpy(), q).tolist())),
    "batched_evals_per_transition": out["batched_evals_posterior"] / args.samples,
    "ms_per_batched_eval": 1e3 * (t2 - t1) / out["batched_evals_posterior"],
    "ms_per_transition": 1e3 * (t2 - t1) / args.samples,
    "mean_batch_fill": float(nuts.chain_evals) / max(1.0, float(nuts.evals) * C),
    "accept_per_chain_quantiles": dict(zip(map(str, q), np.quantile(
        torch.stack([i.acceptance_prob for i in infos]).mean(0).cpu().numpy(), q).tolist())),
})
print(json.dumps(out), flush=True)
