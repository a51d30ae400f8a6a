"""
Sum-to-zero reparametrisation of P-spline blocks (host-side linear map around the evaluation).

Why. A B-spline basis is a partition of unity (``B 1 = 1``), so in ``eta = beta0 + B beta`` the
intercept and the level of the smooth are one direction of the likelihood; ``DistRegBuilder``
(``model/distreg.py:126-217``) takes the design matrix as given, and a Liesel user therefore passes
a CONSTRAINED basis ``B Z`` with penalty ``Z^T K Z`` — ``Z`` an orthonormal basis of
``{beta : cbar^T beta = 0}``, ``cbar`` the column means of ``B`` (the mgcv / liesel_gam
"sum-to-zero" constraint ``sum_i f(x_i) = 0``). Without it the posterior has a ridge of standard
deviation ~100 (only the intercept's ``N(0, 100^2)`` prior pins it) next to directions of standard
deviation 1e-3, which no diagonal-mass-matrix NUTS resolves (round 1: split-R-hat 5.05 on H).

How. The kernels keep evaluating the BANDED basis at the full coefficient vector: the constrained
model is the same function composed with ``beta = Z gamma``,

    logp_r(gamma) = logp(T gamma),   grad_r = T^T grad,   T = blockdiag(I, Z_1, I, Z_2, ...),

two ``[C x P] x [P x P_r]`` products per evaluation (P = 42, P_r = 40 for H) next to a pass over
1e6 rows. ``gamma^T (Z^T K Z) gamma = beta^T K beta`` and ``rank(Z^T K Z) = rank(K)`` (the constant
vector is not in ``cbar``'s complement), so the smoothing prior and the ``tau2`` Gibbs update
(``model/distreg.py:277-297``) are unchanged up to the constant ``log pdet``.
"""

from __future__ import annotations

import numpy as np
import torch

from .spec import KIND_INTERCEPT, KIND_SPLINE, ModelSpec


def sum_to_zero_basis(cbar: np.ndarray) -> np.ndarray:
    """Orthonormal ``Z [k x (k-1)]`` with ``cbar^T Z = 0`` (Householder reflector of ``cbar``,
    the construction mgcv uses for its identifiability constraints)."""
    c = np.asarray(cbar, dtype=np.float64)
    k = c.shape[0]
    v = c.copy()
    v[0] += np.copysign(np.linalg.norm(c), c[0])
    H = np.eye(k) - 2.0 * np.outer(v, v) / (v @ v)  # H c = -+|c| e_1
    return H[:, 1:]


class LinearReparam:
    """``q = q_r T^T`` for ``T [P x P_r]`` with orthonormal columns; all tensors carry the chain axis."""

    def __init__(self, T: np.ndarray, device=None, dtype=torch.float32):
        self.T = torch.as_tensor(np.asarray(T), dtype=dtype, device=device).contiguous()
        self.Tt = self.T.t().contiguous()
        self.P, self.Pr = self.T.shape

    def expand(self, q_r: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        return torch.matmul(q_r, self.Tt, out=out)

    def reduce(self, q: torch.Tensor) -> torch.Tensor:
        """Coordinates of the projection of ``q`` onto the constrained space (start values)."""
        return torch.matmul(q, self.T)

    def reduce_grad(self, g: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        return torch.matmul(g, self.T, out=out)


def sum_to_zero(spec: ModelSpec, colmeans: dict[str, np.ndarray], device=None,
                dtype=torch.float32) -> LinearReparam:
    """
    Constrains every spline block that shares its predictor with an intercept; ``colmeans`` maps the
    block name to the column means of its basis (``Plan.column_means``). Other blocks map through.
    """
    has_intercept = {b.predictor for b in spec.blocks if b.kind == KIND_INTERCEPT}
    P = spec.num_params
    cols = []
    for b in spec.blocks:
        E = np.zeros((P, b.ncoef))
        E[b.param_offset: b.param_offset + b.ncoef] = np.eye(b.ncoef)
        if b.kind == KIND_SPLINE and b.predictor in has_intercept and not b.center:
            E = E @ sum_to_zero_basis(colmeans[b.name])
        cols.append(E)
    return LinearReparam(np.concatenate(cols, axis=1), device=device, dtype=dtype)
