"""
Plan builder from a ``liesel.model.Model``: pattern-matches the graph that the
reference's ``DistRegBuilder`` produces (``src/liesel/model/distreg.py:64-274``) into a
:class:`liesel_b200.spec.ModelSpec`, reading every array ONCE out of the model
(``X``, ``y``, ``K``, ``rank``, ``a``, ``b``, ``m``, ``s``, initial ``tau2``).

Only public attributes of the reference's objects are touched, so the function is duck-typed and
does not import Liesel (JAX is not installable in this image; the tests drive it with stand-ins
that carry exactly these attributes):

  * ``model.vars``                      mapping name -> Var            (``model/model.py:2534``)
  * ``model.groups()``                  mapping name -> Group          (``model/model.py:1941``)
  * ``var.value``, ``var.name``, ``var.dist_node``, ``var.value_node``
  * ``dist_node.distribution``          the TFP class handed to ``Dist`` (``model/nodes.py:1108``)
  * ``dist_node.kwinputs``              mapping parameter name -> node  (``model/nodes.py:336``)
  * ``node.inputs``, ``node.var``       (``model/nodes.py:331,469``)
  * ``group[...]``, ``key in group``    (``model/nodes.py:3561-3566``)

What is matched (anything else raises :class:`UnsupportedModelError`, and the caller keeps
``gs.LieselInterface``):

  response   ``Var.new_obs(y, Dist(tfd.Normal | tfd.Bernoulli | tfd.Poisson), "response")`` (:253-274)
  predictor  ``<param>_pdt`` = sum of its smooths, ``<param>`` = inverse link of it (:219-251);
             links: Identity for loc / logits / log_rate, Exp for scale
  p smooth   Group(smooth, beta, X, m, s) with ``beta ~ Normal(m, s)`` (:64-124)
  np smooth  Group(smooth, beta, tau2, rank, X, K, a, b) with the rank-deficient normal prior and
             ``tau2 ~ InverseGamma(a, b)`` (:126-217)

A dense design matrix that IS a cubic B-spline basis (every row: at most four non-zeros in four
consecutive columns, as ``basis_matrix`` produces, ``contrib/splines.py:109-198``) is converted to
the banded form the fast kernels use — the weights are copied out of the matrix, not recomputed, so
nothing changes numerically. Coefficient blocks keep Liesel's names (``<smooth>_beta``,
``<smooth>_tau2``): positions and ``tau2_gibbs_kernel`` (:277-297) see the keys they expect.
"""

from __future__ import annotations

import numpy as np

from .spec import ModelSpec

_FAMILY_BY_CLASS = {"Normal": "normal", "Bernoulli": "bernoulli", "Poisson": "poisson"}
_LINK_OF = {"loc": "Identity", "scale": "Exp", "logits": "Identity", "log_rate": "Identity"}


class UnsupportedModelError(ValueError):
    """The model graph is not of the shape the B200 path evaluates."""


def band_from_dense(X: np.ndarray, width: int = 4):
    """
    ``(start int32 [n], weights [n, width])`` if every row of ``X`` has all its non-zeros inside
    ``width`` consecutive columns, else ``None``. ``start`` is the first non-zero column, moved left
    where the window would run past the last column; the weights are copied verbatim.
    """
    X = np.asarray(X)
    if X.ndim != 2 or X.shape[1] < width:
        return None
    n, k = X.shape
    nz = X != 0
    any_nz = nz.any(axis=1)
    first = np.where(any_nz, nz.argmax(axis=1), 0)
    last = np.where(any_nz, k - 1 - nz[:, ::-1].argmax(axis=1), 0)
    if np.any(last - first >= width):
        return None
    start = np.minimum(first, k - width).astype(np.int32)
    cols = start[:, None].astype(np.int64) + np.arange(width)[None, :]
    weights = np.take_along_axis(X, cols, axis=1)
    if np.count_nonzero(weights) != np.count_nonzero(X):  # (defensive: every non-zero is inside the window)
        return None
    return start, np.ascontiguousarray(weights)


def _class_name(obj) -> str:
    return getattr(obj, "__name__", None) or type(obj).__name__


def _inverse_link_name(param_var) -> str | None:
    """Name of the bijector class behind ``Var.new_calc(inverse_link().forward, ...)`` (:242)."""
    fn = getattr(param_var.value_node, "function", None)
    owner = getattr(fn, "__self__", None)
    return None if owner is None else type(owner).__name__


def spec_from_model(model, dtype=None, detect_bands: bool = True, trust_links: bool = False) -> ModelSpec:
    """
    Build the :class:`ModelSpec` of a DistReg-shaped Liesel model. ``dtype`` defaults to the dtype
    of the response array (Liesel stores float32 unless x64 is on). ``trust_links`` skips
    the check of the inverse links for objects that do not expose the bijector.
    """
    vars_ = model.vars
    if "response" not in vars_:
        raise UnsupportedModelError("no variable named 'response' (distreg.py:270)")
    response = vars_["response"]
    dist = getattr(response, "dist_node", None)
    if dist is None:
        raise UnsupportedModelError("the response has no distribution")
    fam = _FAMILY_BY_CLASS.get(_class_name(dist.distribution))
    if fam is None:
        raise UnsupportedModelError(f"response distribution {_class_name(dist.distribution)!r} is not supported")
    y = np.asarray(response.value)
    dt = np.dtype(dtype) if dtype is not None else (y.dtype if y.dtype.kind == "f" else np.dtype(np.float32))
    spec = ModelSpec(dtype=dt)
    spec.add_response(y, fam)

    params = list(dist.kwinputs.keys())
    if not params:
        raise UnsupportedModelError("the response distribution has no predictors")
    groups = list(model.groups().values())
    by_smooth = {}
    for g in groups:
        if "smooth" in g:
            by_smooth[id(g["smooth"])] = g

    fixed_scale = None
    if fam == "normal" and "scale" not in params:
        raise UnsupportedModelError("a Normal response needs a 'scale' predictor or constant")
    for pname in params:
        if pname not in _LINK_OF or (pname + "_pdt") not in vars_:
            # a constant distribution parameter (e.g. scale=Var.new_value(1.0)) is allowed for 'scale'
            node = dist.kwinputs[pname]
            val = getattr(getattr(node, "var", None), "value", getattr(node, "value", None))
            if pname == "scale" and val is not None and np.ndim(val) == 0:
                fixed_scale = float(val)
                continue
            raise UnsupportedModelError(f"distribution parameter {pname!r} is not a DistReg predictor")
        link = _inverse_link_name(vars_[pname])
        if link is None and not trust_links:
            raise UnsupportedModelError(
                f"cannot read the inverse link of {pname!r}; pass trust_links=True if it is {_LINK_OF[pname]}")
        if link is not None and link != _LINK_OF[pname]:
            raise UnsupportedModelError(f"inverse link {link} of {pname!r} is not supported (need {_LINK_OF[pname]})")
        spec.add_predictor(pname)
    if fixed_scale is not None:
        spec.fixed_scale = fixed_scale

    nsmooths = 0
    for pname in params:
        if (pname + "_pdt") not in vars_ or pname not in spec._predictors:
            continue
        pdt = vars_[pname + "_pdt"]
        for node in pdt.value_node.inputs:
            svar = getattr(node, "var", None)
            g = by_smooth.get(id(svar))
            if g is None:
                raise UnsupportedModelError(
                    f"input {getattr(node, 'name', node)!r} of predictor {pname!r} is not a DistReg smooth")
            name = svar.name
            if g["beta"].name != name + "_beta":
                raise UnsupportedModelError(f"unexpected coefficient name {g['beta'].name!r}")
            X = np.asarray(g["X"].value)
            if X.ndim != 2 or X.shape[0] != y.shape[0]:
                raise UnsupportedModelError(f"design matrix of {name!r} has shape {X.shape}")
            if "K" in g:  # non-parametric smooth
                K = np.asarray(g["K"].value, dtype=np.float64)
                a, b = float(np.asarray(g["a"].value)), float(np.asarray(g["b"].value))
                tau2 = g["tau2"]
                if tau2.name != name + "_tau2":
                    raise UnsupportedModelError(f"unexpected variance name {tau2.name!r}")
                tau2_init = float(np.asarray(tau2.value))
                band = band_from_dense(X) if detect_bands else None
                if band is not None:
                    start, w = band
                    k = X.shape[1]
                    spec.add_spline_smooth(None, np.arange(k + 4, dtype=dt), K, a, b, pname, name,
                                           tau2_init=tau2_init, start=start, weights=w.astype(dt))
                else:
                    spec.add_np_smooth(X, K, a, b, pname, name, tau2_init=tau2_init)
                blk = spec.blocks[-1]
                rank = float(np.asarray(g["rank"].value))
                if rank != blk.prior.rank:  # the model's stored rank wins (distreg.py:161)
                    from .spec import log_pdet_top_rank

                    blk.prior.rank, blk.prior.log_pdet = rank, log_pdet_top_rank(K, rank)
            elif "m" in g and "s" in g:
                spec.add_p_smooth(X, float(np.asarray(g["m"].value)), float(np.asarray(g["s"].value)), pname, name)
            else:
                raise UnsupportedModelError(f"group {g.name!r} is neither a p smooth nor an np smooth")
            nsmooths += 1
    if nsmooths == 0:
        raise UnsupportedModelError("no smooths found")
    return spec


def initial_position(model, spec: ModelSpec) -> dict:
    """Current values of the model's coefficient and variance variables under the spec's keys."""
    out = {}
    for key in spec.position_keys() + spec.aux_keys():
        out[key] = np.asarray(model.vars[key].value, dtype=spec.dtype)
    return out
