"""NUTS ESS/sec on the S3/H model (Gaussian location-scale GAMLSS, 2 cubic P-splines + 2 intercepts):
all 42 coefficients in one NUTS block (reference defaults: max_treedepth 10, target accept 0.8,
diagonal mass matrix, Stan-style warm-up), tau2 by its Gibbs full conditional after every transition
(reference model/distreg.py:277-297). Reports min-over-coefficients bulk-ESS of all posterior draws
divided by the wall time of the posterior epoch, plus the figure including warm-up.

  python tools/ess_bench.py --n 1000000 --chains 1024 --warmup 300 --samples 200
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import diagnostics, synth  # noqa: E402
from liesel_b200.nuts import BatchedNUTS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--warmup", type=int, default=300)
ap.add_argument("--samples", type=int, default=200)
ap.add_argument("--cpu", action="store_true", help="CPU arm: oracle/cpu_port.py as the evaluator")
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--graphs", action="store_true", help="replay NUTS leaf ranges from CUDA graphs")
ap.add_argument("--compact", action="store_true",
                help="drop chains whose tree has terminated from the batched evaluation")
ap.add_argument("--sampler", default="nuts", choices=["nuts", "iwls"],
                help="nuts: one NUTS block + Gibbs tau2; iwls: IWLS per coefficient block + Gibbs tau2 "
                     "(what dist_reg_mcmc configures, model/distreg.py:300-349)")
args = ap.parse_args()

spec, centre = synth.s3_gamlss(n=args.n)
C = args.chains
th0, aux0 = synth.evaluation_points(spec, centre, C, seed=args.seed)
gen = torch.Generator().manual_seed(args.seed)
gib = [(b, b.prior) for b in spec.blocks if b.tau2_slot >= 0]

f_into = None
if args.cpu:
    from oracle.cpu_port import CpuPort

    torch.set_num_threads(os.cpu_count() or 1)
    port = CpuPort(spec)
    dev = torch.device("cpu")
    aux = torch.as_tensor(aux0)

    def f(q):
        lp, g = port.logp_grad(q.numpy(), aux.numpy())
        return lp, g

    def quad(blk, q):
        b = q[:, blk.param_offset: blk.param_offset + blk.ncoef].double()
        K = torch.as_tensor(np.asarray(blk.prior.K))
        return ((b @ K) * b).sum(dim=1)
    sync = lambda: None
else:
    from liesel_b200.plan import Plan

    plan = Plan(spec)
    dev = plan.device
    aux = torch.as_tensor(aux0, device=dev)

    def f(q):
        return plan.logp_grad(q.contiguous(), aux)

    _sel = {"chains": None, "aux": None}

    def f_into(q, lp, g, chains=None):
        """``chains``: indices of the chains in this (compacted) batch — picks their tau2 rows."""
        if chains is None:
            plan.logp_grad(q, aux, out=(lp, g))
            return
        if _sel["chains"] is not chains:
            _sel["chains"], _sel["aux"] = chains, aux.index_select(0, chains).contiguous()
        plan.logp_grad(q, _sel["aux"], out=(lp, g))

    def quad(blk, q):
        return plan.quadform(blk.name, q.contiguous()).double()
    sync = torch.cuda.synchronize


def gibbs(s):
    """tau2 | beta ~ InvGamma(a + rank/2, b + beta'K beta / 2), then refresh logp/grad."""
    for blk, pr in gib:
        a = pr.ig_a + 0.5 * pr.rank
        b = pr.ig_b + 0.5 * quad(blk, s.q)
        gam = torch._standard_gamma(torch.full((C,), a, dtype=torch.float64, device=dev))
        aux[:, blk.tau2_slot] = (b / gam).to(aux.dtype)
    s.refresh()


def functional_ess(d):
    """Bulk-ESS of IDENTIFIED quantities: both predictors at 10 probe points of the covariate.
    (An intercept plus an uncentred P-spline share a level that only the N(0, 100^2) prior pins —
    raw coefficients random-walk along that ridge in ANY sampler; the fitted curves do not.)"""
    from oracle_free_basis import probe_design

    Bp = probe_design(spec)  # [npred][10, P]
    vals = torch.cat([d.double() @ torch.as_tensor(B, device=d.device).T for B in Bp], dim=2)  # [draws, chains, 20]
    return diagnostics.ess_bulk_device(vals)


q0 = torch.as_tensor(th0, device=dev)
if args.sampler == "iwls":
    from liesel_b200.iwls_sampler import IWLSGibbsSampler

    t0 = time.perf_counter()
    smp = IWLSGibbsSampler(plan, q0, aux, seed=args.seed)
    smp.warmup(args.warmup)
    sync()
    t1 = time.perf_counter()
    draws, _, infos = smp.sample(args.samples)
    sync()
    t2 = time.perf_counter()
    ess = diagnostics.ess_bulk_device(draws)
    ess_f = functional_ess(draws)
    acc = {b.name: float(np.mean([float(i[b.name]["acceptance_prob"].mean()) for i in infos])) for b in smp.blocks}
    print(json.dumps({
        "model": "S3", "sampler": "IWLS blocks + Gibbs tau2", "n": args.n, "chains": C, "arm": "b200",
        "warmup_iters": args.warmup, "posterior_iters": args.samples,
        "ess_bulk_min": float(ess.min()), "ess_bulk_median": float(np.median(ess)),
        "posterior_seconds": t2 - t1, "warmup_seconds": t1 - t0,
        "ess_per_sec_posterior": float(ess.min() / (t2 - t1)),
        "ess_per_sec_incl_warmup": float(ess.min() / (t2 - t0)),
        "ess_bulk_min_identified_functionals": float(ess_f.min()),
        "ess_per_sec_posterior_identified": float(ess_f.min() / (t2 - t1)),
        "iterations_per_sec": args.samples / (t2 - t1), "mean_accept": acc,
        "evals": smp.evals,
    }), flush=True)
    sys.exit(0)
t0 = time.perf_counter()
s = BatchedNUTS(f, q0, seed=args.seed, logp_grad_into=f_into, graphs=args.graphs and not args.cpu,
                compact=args.compact and not args.cpu)
winfo = s.warmup(args.warmup, hook=gibbs)
sync()
t1 = time.perf_counter()
ev_w = s.evals
draws, infos = s.sample(args.samples, hook=gibbs)
sync()
t2 = time.perf_counter()
td = time.perf_counter()
ess = diagnostics.ess_bulk_device(draws)  # on the device that stores the chains (cpu arm: host)
ess_f = functional_ess(draws)
rhat = diagnostics.split_rhat_device(draws)
sync()
out = {
    "split_rhat_max": float(rhat.max()), "diagnostics_seconds": time.perf_counter() - td,
    "ess_bulk_min_identified_functionals": float(ess_f.min()),
    "ess_per_sec_posterior_identified": float(ess_f.min() / (t2 - t1)),
    "model": "S3/H", "n": args.n, "chains": C, "P": int(q0.shape[1]), "arm": "cpu_port" if args.cpu else "b200",
    "warmup_iters": args.warmup, "posterior_iters": args.samples,
    "ess_bulk_min": float(ess.min()), "ess_bulk_median": float(np.median(ess)),
    "posterior_seconds": t2 - t1, "warmup_seconds": t1 - t0,
    "ess_per_sec_posterior": float(ess.min() / (t2 - t1)),
    "ess_per_sec_incl_warmup": float(ess.min() / (t2 - t0)),
    "evals_posterior": int(s.evals - ev_w), "evals_warmup": int(ev_w),
    "batched_evals_per_sec_posterior": (s.evals - ev_w) / (t2 - t1),
    "compact": bool(s.compact), "chain_evals_total": int(s.chain_evals),
    "mean_batch_fill": float(s.chain_evals) / max(1.0, float(s.evals) * C),
    "mean_treedepth": float(np.mean([float(i.treedepth.float().mean()) for i in infos])),
    "max_treedepth_seen": int(max(int(i.treedepth.max()) for i in infos)),
    "mean_accept": float(np.mean([float(i.acceptance_prob.mean()) for i in infos])),
    "divergent_frac": float(np.mean([float(i.divergent.float().mean()) for i in infos])),
    "step_size_median": float(s.step_size.median()),
    "cores": os.cpu_count() if args.cpu else None,
}
print(json.dumps(out), flush=True)
