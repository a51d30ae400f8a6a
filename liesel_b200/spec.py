"""
Host-side description of the model family the B200 hot path evaluates.

The reference builds this family with ``DistRegBuilder`` (reference
``src/liesel/model/distreg.py:64-274``): smooths ``jnp.dot(X, beta)`` are summed into
predictors, pushed through an inverse link and fed into a response distribution, with
a ``tfd.Normal(m, s)`` prior for parametric smooths (``distreg.py:102``) and a
``MultivariateNormalDegenerate.from_penalty`` prior with an ``InverseGamma`` hyperprior
on ``tau2`` for non-parametric smooths (``distreg.py:162-172``).

This module is NumPy-only bookkeeping: it holds the arrays and constants, fixes the
flat parameter layout ``[P]`` (coefficient blocks concatenated in insertion order) and
the auxiliary layout ``[A]`` (one ``tau2`` slot per penalised block, in insertion
order). It performs no likelihood arithmetic; that lives in the CUDA library.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

FAMILY_NORMAL = 0  # tfd.Normal(loc, scale): predictor 0 = loc (identity), 1 = log scale
FAMILY_BERNOULLI = 1  # tfd.Bernoulli(logits): predictor 0 = logits
FAMILY_POISSON = 2  # tfd.Poisson(log_rate): predictor 0 = log_rate

KIND_DENSE = 0
KIND_SPLINE = 1
KIND_GROUP = 2
KIND_INTERCEPT = 3  # add_p_smooth with a single all-ones column: no data is streamed

PRIOR_NONE = 0
PRIOR_NORMAL = 1
PRIOR_MVND = 2

_FAMILIES = {"normal": FAMILY_NORMAL, "bernoulli": FAMILY_BERNOULLI, "poisson": FAMILY_POISSON}
_PREDICTORS = {
    FAMILY_NORMAL: {"loc": 0, "scale": 1},
    FAMILY_BERNOULLI: {"logits": 0},
    FAMILY_POISSON: {"log_rate": 0},
}


@dataclass
class NormalPrior:
    """Elementwise ``tfd.Normal(m, s)`` prior (reference ``distreg.py:102``)."""

    m: float
    s: float


@dataclass
class MVNDPrior:
    """
    ``MultivariateNormalDegenerate.from_penalty(loc=0, var=tau2, pen=K, rank=rank)``
    (reference ``distreg.py:166-172``, ``mvn_degen.py:209-287``) with an optional
    ``tfd.InverseGamma(a, b)`` prior on ``tau2`` (``distreg.py:162``).

    ``log_pdet`` is the log pseudo-determinant of ``K``; when ``None`` it is computed
    once on the host by the rule of ``mvn_degen.py:32-56`` (top-``rank`` eigenvalues).
    """

    K: np.ndarray
    rank: float | None = None
    log_pdet: float | None = None
    tau2_name: str | None = None  # None: tau2 is a fixed constant
    tau2_init: float = 10000.0  # reference default, distreg.py:163
    ig_a: float | None = None
    ig_b: float | None = None


@dataclass
class Block:
    """One additive term of a predictor."""

    name: str
    kind: int
    predictor: int
    ncoef: int
    prior: NormalPrior | MVNDPrior | None
    # dense
    X: np.ndarray | None = None
    # spline: raw covariate + knots (basis built on device) and/or precomputed band
    x: np.ndarray | None = None
    knots: np.ndarray | None = None
    order: int = 3
    start: np.ndarray | None = None  # int32 [n], first non-zero basis column
    weights: np.ndarray | None = None  # [n, order+1]
    # spline: column-centred design B - 1 colmean^T (colmean filled by Plan when center is True
    # and no array was given)
    center: bool = False
    colmean: np.ndarray | None = None
    # group
    idx: np.ndarray | None = None  # int32 [n], sorted ascending
    # filled by ModelSpec
    param_offset: int = 0
    tau2_slot: int = -1


@dataclass
class ModelSpec:
    """
    Builder with the call shape of the reference's ``DistRegBuilder``
    (``add_response`` → ``add_predictor`` → ``add_p_smooth`` / ``add_np_smooth``), plus
    banded-spline and random-intercept terms which the reference spells as user
    ``Calc`` nodes (``distreg.py:105,175`` and SURVEY §8 a14).
    """

    dtype: np.dtype = field(default_factory=lambda: np.dtype(np.float32))
    family: int = FAMILY_NORMAL
    y: np.ndarray | None = None
    fixed_scale: float = 1.0
    const_term: float = 0.0
    blocks: list[Block] = field(default_factory=list)
    _predictors: dict[str, int] = field(default_factory=dict)

    # -- DistRegBuilder-shaped API ----------------------------------------------------
    def add_response(self, response: np.ndarray, distribution: str) -> "ModelSpec":
        """Mirror of ``DistRegBuilder.add_response`` (``distreg.py:253-274``)."""
        if distribution not in _FAMILIES:
            raise ValueError(f"Unsupported response distribution {distribution!r}")
        self.family = _FAMILIES[distribution]
        self.y = np.ascontiguousarray(np.asarray(response), dtype=self.dtype)
        if self.y.ndim != 1:
            raise ValueError("response must be a 1d array")
        return self

    def add_predictor(self, name: str, inverse_link: str | None = None) -> "ModelSpec":
        """
        Mirror of ``DistRegBuilder.add_predictor`` (``distreg.py:219-251``). The inverse
        link is implied by the family/parameter pair (identity for ``loc``, ``Exp`` for
        ``scale``, none for ``logits``/``log_rate``); passing another link raises.
        """
        if self.y is None:
            raise RuntimeError("No response found. Add a response first.")
        table = _PREDICTORS[self.family]
        if name not in table:
            raise ValueError(f"Predictor {name!r} is not a parameter of this family")
        implied = {"loc": "Identity", "scale": "Exp", "logits": "Identity", "log_rate": "Identity"}
        if inverse_link is not None and inverse_link != implied[name]:
            raise ValueError(
                f"inverse_link={inverse_link!r} is not supported for {name!r}; "
                f"the B200 path implements {implied[name]!r}"
            )
        self._predictors[name] = table[name]
        return self

    def _check_predictor(self, predictor: str) -> int:
        if predictor not in self._predictors:
            raise RuntimeError(
                f"No predictor '{predictor}' found. You need to add this predictor to"
                " the builder first."
            )
        return self._predictors[predictor]

    def _check_name(self, name: str) -> None:
        if any(b.name == name for b in self.blocks):
            raise RuntimeError(f"Smooth {name!r} already exists")

    @property
    def n(self) -> int:
        if self.y is None:
            raise RuntimeError("No response found. Add a response first.")
        return int(self.y.shape[0])

    def add_p_smooth(self, X, m: float, s: float | None, predictor: str, name: str) -> "ModelSpec":
        """
        Mirror of ``DistRegBuilder.add_p_smooth`` (``distreg.py:64-124``). ``s=None`` adds the
        term without a prior (a flat prior, as in the reference's ``tests/goose/kernels/model_lm.py``).
        """
        pred = self._check_predictor(predictor)
        self._check_name(name)
        X = np.ascontiguousarray(np.asarray(X), dtype=self.dtype)
        if X.ndim == 1:
            X = X[:, None]
        if X.shape[0] != self.n:
            raise ValueError("X has the wrong number of rows")
        prior = None if s is None else NormalPrior(float(m), float(s))
        if X.shape[1] == 1 and np.all(X == 1):
            self.blocks.append(Block(name, KIND_INTERCEPT, pred, 1, prior))
        else:
            self.blocks.append(Block(name, KIND_DENSE, pred, X.shape[1], prior, X=X))
        return self._relayout()

    def add_np_smooth(
        self, X, K, a: float | None, b: float | None, predictor: str, name: str,
        tau2_name: str | None = "auto", tau2_init: float = 10000.0,
    ) -> "ModelSpec":
        """
        Mirror of ``DistRegBuilder.add_np_smooth`` (``distreg.py:126-217``) for a DENSE
        design matrix. ``rank`` is ``float(np.linalg.matrix_rank(K))`` as in
        ``distreg.py:23,161``.
        """
        pred = self._check_predictor(predictor)
        self._check_name(name)
        X = np.ascontiguousarray(np.asarray(X), dtype=self.dtype)
        if X.shape[0] != self.n:
            raise ValueError("X has the wrong number of rows")
        prior = self._mvnd_prior(K, a, b, name, tau2_name, tau2_init)
        self.blocks.append(Block(name, KIND_DENSE, pred, X.shape[1], prior, X=X))
        return self._relayout()

    # -- banded / grouped terms -------------------------------------------------------
    def add_spline_smooth(
        self, x, knots, K, a: float | None, b: float | None, predictor: str, name: str,
        order: int = 3, tau2_name: str | None = "auto", tau2_init: float = 10000.0,
        start: np.ndarray | None = None, weights: np.ndarray | None = None,
        center: bool | np.ndarray = False,
    ) -> "ModelSpec":
        """
        ``add_np_smooth`` whose design matrix is ``basis_matrix(x, knots, order)``
        (reference ``contrib/splines.py:109-198``; orders 0..3), kept in banded form: per row the
        first column ``start`` of a window of FOUR weights (orders below 3 leave the rest of the window
        zero).

        ``center``: the design matrix is the column-centred basis ``B - 1 cbar^T`` (what a Liesel
        user passes to ``add_np_smooth`` so that the smooth is identified next to an intercept).
        ``True`` takes ``cbar`` = column means of ``B`` (computed when the plan is built and stored
        in ``Block.colmean``); an array of ``k`` values is used as given. The band stays sparse;
        the rank-one part is handled in closed form.
        """
        pred = self._check_predictor(predictor)
        self._check_name(name)
        if order not in (0, 1, 2, 3):
            raise ValueError(f"B-spline orders 0..3 are implemented on the device, got order={order}")
        knots = np.sort(np.asarray(knots, dtype=self.dtype))  # splines.py:183
        k = knots.shape[0] - order - 1
        if k < 4:
            raise ValueError("too few knots: the banded layout needs at least 4 basis functions")
        x = None if x is None else np.ascontiguousarray(np.asarray(x), dtype=self.dtype)
        if x is not None and x.shape != (self.n,):
            raise ValueError("x must be a 1d array with one entry per observation")
        if x is None and (start is None or weights is None):
            raise ValueError("either x or a precomputed (start, weights) band is required")
        prior = self._mvnd_prior(K, a, b, name, tau2_name, tau2_init) if K is not None else None
        blk = Block(name, KIND_SPLINE, pred, k, prior, x=x, knots=knots, order=order)
        if isinstance(center, np.ndarray):
            if center.shape != (k,):
                raise ValueError("center must have one entry per basis function")
            blk.center, blk.colmean = True, np.ascontiguousarray(center, dtype=self.dtype)
        else:
            blk.center = bool(center)
        if start is not None:
            if weights is None:
                raise ValueError("a precomputed band needs both start and weights")
            blk.start = np.ascontiguousarray(start, dtype=np.int32)
            blk.weights = np.ascontiguousarray(weights, dtype=self.dtype)
            if blk.start.shape != (self.n,) or blk.weights.shape != (self.n, 4):
                raise ValueError("band must be start [n] and weights [n, 4] (orders < 3: zero-padded window)")
            if self.n and (blk.start.min() < 0 or blk.start.max() > k - 4):
                raise ValueError(f"band start out of range: need 0 <= start <= {k - 4}")
        self.blocks.append(blk)
        return self._relayout()

    def add_random_intercept(self, idx, num_groups: int, m: float, s: float, predictor: str,
                             name: str) -> "ModelSpec":
        """
        Random intercept ``eta_i += b[idx_i]`` with ``b_g ~ Normal(m, s)``.
        Rows must be sorted by group so that the gradient is a deterministic segment sum.
        """
        pred = self._check_predictor(predictor)
        self._check_name(name)
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int32)
        if idx.shape != (self.n,):
            raise ValueError("idx must be a 1d array with one entry per observation")
        if idx.size and (idx.min() < 0 or idx.max() >= num_groups):
            raise ValueError("group index out of range")
        if np.any(np.diff(idx) < 0):
            raise ValueError("rows must be sorted by group index (use sort_rows_by_group)")
        if any(b.kind == KIND_GROUP for b in self.blocks):
            raise ValueError("at most one random-intercept term is supported")
        self.blocks.append(
            Block(name, KIND_GROUP, pred, int(num_groups), NormalPrior(float(m), float(s)), idx=idx)
        )
        return self._relayout()

    # -- helpers ----------------------------------------------------------------------
    def _mvnd_prior(self, K, a, b, name, tau2_name, tau2_init) -> MVNDPrior:
        K = np.ascontiguousarray(np.asarray(K, dtype=np.float64))
        rank = float(np.linalg.matrix_rank(K))  # distreg.py:161
        if tau2_name == "auto":
            tau2_name = name + "_tau2"
        return MVNDPrior(K=K, rank=rank, log_pdet=log_pdet_top_rank(K, rank),
                         tau2_name=tau2_name, tau2_init=float(tau2_init),
                         ig_a=None if a is None else float(a), ig_b=None if b is None else float(b))

    def _relayout(self) -> "ModelSpec":
        off, slot = 0, 0
        for blk in self.blocks:
            blk.param_offset = off
            off += blk.ncoef
            if isinstance(blk.prior, MVNDPrior) and blk.prior.tau2_name is not None:
                blk.tau2_slot = slot
                slot += 1
            else:
                blk.tau2_slot = -1
        return self

    def rows(self, start: int, stop: int) -> "ModelSpec":
        """
        The same model on rows ``[start, stop)`` only — one rank's share under row sharding
        (``dist.row_shard``). Coefficient layout, priors, knots and group numbering are unchanged, so
        per-rank partial results add up to the full evaluation. A centred spline needs its column
        means fixed beforehand (they are means over ALL rows, not over the shard).
        """
        import copy

        if not 0 <= start <= stop <= self.n:
            raise ValueError("row range out of bounds")
        part = copy.copy(self)
        part.y = np.ascontiguousarray(self.y[start:stop])
        part._predictors = dict(self._predictors)
        part.blocks = []
        for blk in self.blocks:
            if blk.center and blk.colmean is None:
                raise ValueError(f"block {blk.name!r}: fix the column means (center=<array>) before sharding rows")
            nb = copy.copy(blk)
            for attr in ("X", "x", "start", "weights", "idx"):
                arr = getattr(blk, attr)
                if arr is not None:
                    setattr(nb, attr, np.ascontiguousarray(arr[start:stop]))
            part.blocks.append(nb)
        return part._relayout()

    @property
    def num_params(self) -> int:
        return sum(b.ncoef for b in self.blocks)

    @property
    def num_aux(self) -> int:
        return sum(1 for b in self.blocks if b.tau2_slot >= 0)

    @property
    def has_scale_predictor(self) -> bool:
        return any(b.predictor == 1 for b in self.blocks)

    def block(self, name: str) -> Block:
        for b in self.blocks:
            if b.name == name:
                return b
        raise KeyError(name)

    def block_index(self, name: str) -> int:
        for i, b in enumerate(self.blocks):
            if b.name == name:
                return i
        raise KeyError(name)

    def position_keys(self) -> list[str]:
        """Names of the coefficient vectors, ``<block>_beta`` as in ``distreg.py:103,173``."""
        return [b.name + "_beta" for b in self.blocks]

    def aux_keys(self) -> list[str]:
        return [b.prior.tau2_name for b in self.blocks if b.tau2_slot >= 0]  # type: ignore[union-attr]

    def default_aux(self) -> np.ndarray:
        vals = [b.prior.tau2_init for b in self.blocks if b.tau2_slot >= 0]  # type: ignore[union-attr]
        return np.asarray(vals, dtype=self.dtype)


def log_pdet_top_rank(K: np.ndarray, rank: float) -> float:
    """
    Log pseudo-determinant by the rule the reference applies when ``rank`` is given
    (``mvn_degen.py:47-55``): the LAST ``rank`` entries of the ascending ``eigvalsh``
    output, with no tolerance test. Host-side, once per model (K is constant).
    """
    evals = np.linalg.eigvalsh(np.asarray(K, dtype=np.float64))
    r = int(rank)
    return float(np.sum(np.log(evals[evals.shape[0] - r:]))) if r > 0 else 0.0


def sort_rows_by_group(idx: np.ndarray, *arrays: Sequence[np.ndarray]):
    """Stable permutation that sorts rows by group; returns (perm, idx_sorted, arrays...)."""
    perm = np.argsort(np.asarray(idx), kind="stable")
    out = [np.asarray(a)[perm] for a in arrays]
    return (perm, np.asarray(idx)[perm], *out)
