"""The run-length kernel on smooths of DIFFERENT covariates ("state the limit and measure it").
H is the location-scale model with both smooths on ONE covariate: rows sorted by its knot span give runs of
n / 17 rows. With the log-sd smooth on a second, independent covariate the plan sorts rows lexicographically by
(span of smooth 0, span of smooth 1): smooth 0 keeps its long runs, smooth 1 changes span every n / 17^2 rows,
and the tiles that contain a change take the kernel's per-row path. One JSON line.

  python tools/two_covariates.py [--n 1000000] [--chains 1024]
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
from liesel_b200.spec import ModelSpec  # noqa: E402
from liesel_b200.splines import equidistant_knots, pspline_penalty  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--chains", type=int, default=1024)
args = ap.parse_args()
n, C, k = args.n, args.chains, 20


def model(two):
    rng = np.random.default_rng(103)
    x1 = rng.uniform(size=n)
    x2 = rng.uniform(size=n) if two else x1
    y = rng.normal(1.0 + np.sin(2 * np.pi * x1), np.exp(-0.5 + 0.5 * np.cos(2 * np.pi * x2)))
    spec = ModelSpec(dtype=np.dtype(np.float32))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    K = pspline_penalty(k, 2)
    a, b = x1.astype(np.float32), x2.astype(np.float32)
    ka, kb = equidistant_knots(a, k, 3), equidistant_knots(b, k, 3)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "loc", "loc_p0")
    spec.add_spline_smooth(a, ka, K, 1.0, 0.005, "loc", "loc_np0", tau2_init=1.0)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "scale", "scale_p0")
    spec.add_spline_smooth(b, kb, K, 1.0, 0.005, "scale", "scale_np0", tau2_init=1.0)
    g1, g2 = synth._greville(ka.astype(np.float64)), synth._greville(kb.astype(np.float64))
    return spec, np.concatenate([[1.0], np.sin(2 * np.pi * g1), [-0.5], 0.5 * np.cos(2 * np.pi * g2)])


flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
out = {"n": n, "chains": C}
for two in (False, True):
    spec, centre = model(two)
    plan = Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, C)
    t, a = torch.as_tensor(th, device="cuda"), torch.as_tensor(aux, device="cuda")
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.logp_grad(t, a)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    out["two_covariates" if two else "one_covariate"] = {"variant": plan.variant, "kernel": plan.kernel_name(C),
                                                         "ms_per_call": float(np.median(ts))}
print(json.dumps(out))
