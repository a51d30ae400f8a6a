"""Small invocations of the TMA-ring kernels (banded_x2, dgroup_x2, banded_iwls_x2), the tensor-core
pair and the peer exchange, for `compute-sanitizer --tool memcheck|racecheck|synccheck`:

  compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import dist as ld  # noqa: E402
from liesel_b200 import synth  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(tag, spec, centre, C, iwls_block=None, shard=False):
    plan = Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, C)
    t = torch.as_tensor(th, device="cuda")
    a = torch.as_tensor(aux, device="cuda") if spec.num_aux else None
    lp, g = plan.logp_grad(t, a)
    lp2, _ = plan.logp_grad(t, a, want_grad=False)
    if iwls_block:
        plan.iwls(iwls_block, t, a)
    if shard:
        peer = ld.PeerGroup(C * (plan.P + 1), plan.device)
        for _ in range(3):
            lp3, g3 = plan.logp_grad(t, a, shard=peer)
        torch.cuda.synchronize()
        assert torch.equal(lp3, lp) and torch.equal(g3, g) and not peer.timed_out()
        peer.close()
    torch.cuda.synchronize()
    assert torch.isfinite(lp).all() and torch.isfinite(g).all()
    print(tag, plan.variant, "ok", float(lp[0]), flush=True)


if which in ("all", "banded"):
    spec, centre = synth.s3_gamlss(n=20_000 + 77)  # ragged last tile
    run("s3 C=320 (5 warps per CTA)", spec, centre, 320, iwls_block="loc_np0", shard=True)
    run("s3 C=1024 (16 warps per CTA)", spec, centre, 1024)
if which in ("all", "group"):
    spec, centre = synth.s5_poisson_ri(n=20_000, G=700, balanced=False)
    run("s5 C=256", spec, centre, 256)
    run("s5 C=704", spec, centre, 704)
if which in ("all", "tc"):
    spec, centre = synth.s4_logistic(n=33_000, p=128)
    run("s4 tensor cores C=256", spec, centre, 256)
if which in ("all", "round2"):
    # round-2 kernels: small-chain ring depths (1- and 2-warp CTAs), multi-smooth kernels (msmooth / msort),
    # finalize with the last-CTA-done log-posterior, IWLS with weight products + parallel chunk reduce,
    # warp-cooperative centring rows, the asynchronous NUTS state machine with Gibbs blocks
    spec, centre = synth.s3_gamlss(n=9_000 + 13)
    run("s3 C=64 (1-warp CTAs, 2 stages)", spec, centre, 64, iwls_block="scale_np0")
    run("s3 C=128 (2-warp CTAs, 3 stages)", spec, centre, 128, iwls_block="loc_p0")
    spec, centre = synth.s2_pspline_gam(n=6_000 + 5)
    run("s2 msmooth C=64", spec, centre, 64, iwls_block="loc_np3")
    run("s2 msmooth C=96", spec, centre, 96)
    os.environ["LSLB_MSORT"] = "1"
    run("s2 msort C=64", spec, centre, 64)
    os.environ.pop("LSLB_MSORT")
    from liesel_b200 import reparam
    from liesel_b200.nuts_gibbs import AsyncNutsGibbs, PlanEvaluator

    spec, centre = synth.s3_gamlss(n=3_000)
    plan = Plan(spec)
    cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1}
    rep = reparam.sum_to_zero(spec, cm, device=plan.device)
    th, aux = synth.evaluation_points(spec, centre, 40, seed=1)
    run_a = AsyncNutsGibbs(spec, PlanEvaluator(plan), torch.as_tensor(th, device="cuda"),
                           torch.as_tensor(aux, device="cuda"), 40, 6, reparam=rep, seed=1, max_treedepth=4,
                           epoch_kw=dict(init_duration=10, term_duration=10, base_duration=10))
    t = run_a.run(check_every=16)
    assert torch.isfinite(run_a.draws()).all()
    print("async nuts-gibbs ok", t["steps_warmup"], t["steps_posterior"], flush=True)
