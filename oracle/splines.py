"""
Oracle: B-spline setup (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows reference ``src/liesel/contrib/splines.py``:
  * ``equidistant_knots``  :13-70
  * ``_build_basis_vector`` :73-106 (Cox-de Boor, half-open base case)
  * ``basis_matrix``        :109-198
  * ``pspline_penalty``     :201-220
"""

from __future__ import annotations

import numpy as np


def equidistant_knots(x, n_param: int, order: int = 3, eps: float = 0.01, dtype=None):
    """
    splines.py:13-70. The reference calls ``jnp.linspace`` three times (:62,:66-67) in
    the dtype of ``x``; JAX's linspace arithmetic is third-party (jax 0.10.1, not under
    /root/reference), so the last bit of each knot is "parity unpinned". Knots are
    therefore always exchanged as DATA between oracle and device code.
    """
    if order < 0:
        raise ValueError(f"Invalid {order=}.")
    if n_param < order:
        raise ValueError(f"{n_param=} must not be smaller than {order=}.")
    x = np.asarray(x)
    dtype = np.dtype(dtype or (x.dtype if x.dtype.kind == "f" else np.float64))
    n_internal_knots = n_param - order + 1
    a = dtype.type(np.min(x))
    b = dtype.type(np.max(x))
    range_ = b - a
    min_k = a - range_ * dtype.type(eps / 2)
    max_k = b + range_ * dtype.type(eps / 2)
    internal = np.linspace(min_k, max_k, n_internal_knots, dtype=dtype)
    step = internal[1] - internal[0]
    left = np.linspace(min_k - step * order, min_k - step, order, dtype=dtype)
    right = np.linspace(max_k + step, max_k + step * order, order, dtype=dtype)
    return np.concatenate((left, internal, right)).astype(dtype)


def basis_matrix(x, knots, order: int = 3):
    """
    Dense [n, k] basis, k = len(knots) - order - 1, by the recursion of
    splines.py:73-106 evaluated in the dtype of ``knots``:

        level 0:  bv[i] = (x >= t[i]) * (x < t[i+1])                        (:86-90)
        level m:  b1 = (x - t[i]) / (t[i+m] - t[i]) * bv[i]
                  b2 = (t[i+m+1] - x) / (t[i+m+1] - t[i+1]) * bv[i+1]
                  bv[i] = b1 + b2                                           (:93-98)

    The reference fills a working vector of length len(knots)-1 for i < k+order at every
    level and slices ``[:k]`` (:104-106); entries with i >= len(knots)-1-m read clamped
    out-of-range indices but never feed the returned slice, so only the valid triangle
    i < len(knots)-1-m is computed here.
    """
    knots = np.sort(np.asarray(knots))  # splines.py:183
    dt = knots.dtype
    x = np.atleast_1d(np.asarray(x, dtype=dt))
    nk = knots.shape[0]
    k = nk - order - 1
    bv = np.zeros((x.shape[0], nk - 1), dtype=dt)
    for i in range(nk - 1):
        bv[:, i] = ((x >= knots[i]) & (x < knots[i + 1])).astype(dt)
    for m in range(1, order + 1):
        new = np.zeros_like(bv)
        for i in range(nk - 1 - m):
            b1 = (x - knots[i]) / (knots[i + m] - knots[i]) * bv[:, i]
            b2 = (knots[i + m + 1] - x) / (knots[i + m + 1] - knots[i + 1]) * bv[:, i + 1]
            new[:, i] = b1 + b2
        bv = new
    return bv[:, :k]


def span_index(x, knots):
    """
    The integer that must be bit-exact: j with t[j] <= x < t[j+1], decided by the same
    two comparisons as splines.py:86-90 on the stored knot array. -1 when x is outside
    [t[0], t[-1]).
    """
    knots = np.sort(np.asarray(knots))
    x = np.atleast_1d(np.asarray(x, dtype=knots.dtype))
    ind = (x[:, None] >= knots[None, :-1]) & (x[:, None] < knots[None, 1:])
    j = np.where(ind.any(axis=1), ind.argmax(axis=1), -1)
    return j.astype(np.int32)


def banded_basis(x, knots, order: int = 3):
    """
    Banded form of ``basis_matrix``: ``start`` int32 [n] (first column of the window) and
    ``w`` [n, order+1] with dense[i, start[i]+t] == w[i, t] and zeros elsewhere.
    start = clip(span - order, 0, k - order - 1); rows outside the knot range get zero
    weights (the reference's dense row is all zero there).
    """
    knots = np.sort(np.asarray(knots))
    dense = basis_matrix(x, knots, order)
    n, k = dense.shape
    j = span_index(x, knots)
    start = np.clip(j - order, 0, k - order - 1).astype(np.int32)
    start[j < 0] = 0
    cols = start[:, None] + np.arange(order + 1)[None, :]
    w = np.take_along_axis(dense, cols, axis=1)
    # the band must carry every non-zero of the dense row
    check = np.zeros_like(dense)
    np.put_along_axis(check, cols, w, axis=1)
    if not np.array_equal(check, dense):
        raise AssertionError("banded form lost non-zeros of the dense basis")
    return start, w


def band_to_dense(start, w, k: int):
    n, width = w.shape
    dense = np.zeros((n, k), dtype=w.dtype)
    cols = start[:, None] + np.arange(width)[None, :]
    np.put_along_axis(dense, cols, w, axis=1)
    return dense


def pspline_penalty(d: int, diff: int = 2):
    """splines.py:201-220: D = diff(I_d, diff, axis=0); K = D^T D."""
    D = np.diff(np.identity(d), diff, axis=0)
    return D.T @ D
