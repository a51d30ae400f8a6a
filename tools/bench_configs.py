"""Per-config timings of the hot path on ONE GPU (S1-S5 of SURVEY §8, BASELINE.json `configs`):
logp+grad evaluations/s, value-only evaluations/s and, where the config uses IWLS, IWLS passes/s.
Inputs resident in HBM, L2 flushed between timed calls, CUDA events. Prints one JSON object per line.

  python tools/bench_configs.py [--scale 1.0] [--configs s1,s2,s3,s4,s5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import synth  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0, help="scales n (and G)")
ap.add_argument("--configs", default="s1,s2,s3,h,s4,s5")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--chain-sweep", default="", help="comma-separated chain counts: time the H model at each")
args = ap.parse_args()

CFG = {
    "s1": (lambda s: synth.s1_blr(n=1000), 4, None),
    "s2": (lambda s: synth.s2_pspline_gam(n=int(100_000 * s)), 64, "loc_np0"),
    "s3": (lambda s: synth.s3_gamlss(n=int(1_000_000 * s)), 256, "loc_np0"),
    "h": (lambda s: synth.s3_gamlss(n=int(1_000_000 * s)), 1024, None),
    # S4 is n=1e7 over 8 GPUs: one GPU's shard
    "s4": (lambda s: synth.s4_logistic(n=int(10_000_000 * s), world_size=8), 1024, None),
    "s5": (lambda s: synth.s5_poisson_ri(n=int(5_000_000 * s), G=int(100_000 * s)), 1024, None),
}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, reps):
    for _ in range(3):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


if args.chain_sweep:
    spec, centre = synth.s3_gamlss(n=int(1_000_000 * args.scale))
    plan = Plan(spec)
    rows = []
    for C in [int(c) for c in args.chain_sweep.split(",")]:
        th, aux = synth.evaluation_points(spec, centre, C)
        t, a = torch.as_tensor(th, device="cuda"), torch.as_tensor(aux, device="cuda")
        ms = timeit(lambda: plan.logp_grad(t, a), args.reps)
        rows.append({"chains": C, "logp_grad_ms": ms, "us_per_chain": 1e3 * ms / C})
    print(json.dumps({"config": "h_chain_sweep", "n": spec.n, "variant": plan.variant, "sweep": rows}), flush=True)
    sys.exit(0)

for name in args.configs.split(","):
    make, C, iwls_block = CFG[name]
    spec, centre = make(args.scale)
    plan = Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, C)
    t = torch.as_tensor(th, device="cuda")
    a = torch.as_tensor(aux, device="cuda") if spec.num_aux else None
    out = {"config": name, "n": spec.n, "chains": C, "P": plan.P, "variant": plan.variant}
    ms = timeit(lambda: plan.logp_grad(t, a), args.reps)
    out["logp_grad_ms"] = ms
    out["logp_grad_evals_per_s"] = C / (ms * 1e-3)
    ms = timeit(lambda: plan.logp_grad(t, a, want_grad=False), args.reps)
    out["logp_only_ms"] = ms
    if iwls_block:
        ms = timeit(lambda: plan.iwls(iwls_block, t, a), args.reps)
        out["iwls_block"] = iwls_block
        out["iwls_ms"] = ms
        out["iwls_passes_per_s"] = C / (ms * 1e-3)
    print(json.dumps(out), flush=True)
    del plan
    torch.cuda.empty_cache()
