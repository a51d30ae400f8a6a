"""
B-spline setup for the B200 path.

Mirrors the public functions of the reference's ``liesel.contrib.splines``
(``contrib/splines.py``): ``equidistant_knots`` (:13-70), ``basis_matrix`` (:109-198, here in
BANDED form and evaluated by a CUDA kernel) and ``pspline_penalty`` (:201-220). Knot
placement and the penalty are O(k) / O(k^2) host set-up; the O(n) basis evaluation runs on
the device (``lslb_spline_basis_*`` in ``include/liesel_b200.h``).
"""

from __future__ import annotations

import numpy as np


def equidistant_knots(x, n_param: int, order: int = 3, eps: float = 0.01, dtype=None) -> np.ndarray:
    """
    Equidistant knots, ``len == n_param + order + 1``; the interior range is stretched by
    ``eps`` (reference ``contrib/splines.py:13-70``). Computed in the dtype of ``x`` like the
    reference. The exact last-bit value of each knot depends on the linspace arithmetic
    (JAX's in the reference, NumPy's here); downstream code treats the returned array as
    data and decides knot spans by comparison against it, never by recomputation.
    """
    if order < 0:
        raise ValueError(f"Invalid {order=}.")
    if n_param < order:
        raise ValueError(f"{n_param=} must not be smaller than {order=}.")
    x = np.asarray(x)
    dt = np.dtype(dtype or (x.dtype if x.dtype.kind == "f" else np.float64))
    n_internal = n_param - order + 1
    lo, hi = dt.type(x.min()), dt.type(x.max())
    span = hi - lo
    min_k = lo - span * dt.type(eps / 2)
    max_k = hi + span * dt.type(eps / 2)
    internal = np.linspace(min_k, max_k, n_internal, dtype=dt)
    step = internal[1] - internal[0]
    left = np.linspace(min_k - step * order, min_k - step, order, dtype=dt)
    right = np.linspace(max_k + step, max_k + step * order, order, dtype=dt)
    return np.concatenate((left, internal, right)).astype(dt)


def pspline_penalty(d: int, diff: int = 2) -> np.ndarray:
    """``D^T D`` with ``D`` the ``diff``-th difference of the identity (``splines.py:201-220``)."""
    D = np.diff(np.identity(d), diff, axis=0)
    return D.T @ D


def basis_matrix(x, knots, order: int = 3, outer_ok: bool = False, device=None):
    """
    Banded B-spline basis on the device: returns ``(start, weights)`` torch tensors with
    ``start`` int32 [n] and ``weights`` [n, 4] such that the reference's dense
    ``basis_matrix(x, knots, order)[i, start[i] + t] == weights[i, t]`` and all other
    entries of row i are zero (for ``order < 3`` some entries of the window are zero too).

    ``outer_ok=False`` raises if any x lies outside the interior knot range
    ``[knots[order], knots[-order-1]]`` (the documented behaviour of the reference's flag;
    its actual check at ``splines.py:190`` has an operator-precedence quirk that is
    deliberately not reproduced).
    """
    import torch

    from . import _lib

    if order not in (0, 1, 2, 3):
        raise ValueError("the device basis kernel implements B-spline orders 0..3")
    knots = np.sort(np.asarray(knots))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.ascontiguousarray(np.atleast_1d(x), dtype=knots.dtype))
    dev = torch.device(device) if device is not None else (
        x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    x = x.to(dev).contiguous()
    if not outer_ok and x.numel():
        lo, hi = float(x.min()), float(x.max())
        if lo < knots[order] or hi > knots[knots.shape[0] - order - 1]:
            raise ValueError(
                "Values of x are not inside the range of interior knots, "
                f"[{knots[order]}, {knots[knots.shape[0] - order - 1]}]"
            )
    return _lib.spline_basis(x, torch.as_tensor(knots, device=dev), order)
