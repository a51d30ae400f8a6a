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
