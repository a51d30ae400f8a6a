"""Dense design rows at probe points for the ESS tool (host NumPy, setup only)."""
import numpy as np


def _bspline_row(x, knots, order=3):
    nk = len(knots)
    bv = np.zeros(nk - 1)
    for i in range(nk - 1):
        bv[i] = float(knots[i] <= x < knots[i + 1])
    for m in range(1, order + 1):
        new = np.zeros(nk - 1)
        for i in range(nk - 1 - m):
            new[i] = ((x - knots[i]) / (knots[i + m] - knots[i]) * bv[i]
                      + (knots[i + m + 1] - x) / (knots[i + m + 1] - knots[i + 1]) * bv[i + 1])
        bv = new
    return bv[: nk - order - 1]


def probe_design(spec, npts=10):
    """For every predictor: [npts, P] rows r with eta(x_k) = r @ params."""
    P = spec.num_params
    xs = None
    out = {}
    for b in spec.blocks:
        if b.kind == 1:
            xs = np.quantile(np.asarray(b.x, dtype=np.float64), np.linspace(0.05, 0.95, npts))
            break
    for b in spec.blocks:
        M = out.setdefault(b.predictor, np.zeros((npts, P)))
        if b.kind == 3:
            M[:, b.param_offset] = 1.0
        elif b.kind == 1:
            kn = np.asarray(b.knots, dtype=np.float64)
            for k, x in enumerate(xs):
                M[k, b.param_offset: b.param_offset + b.ncoef] = _bspline_row(x, kn, b.order)
    return [out[k] for k in sorted(out)]
