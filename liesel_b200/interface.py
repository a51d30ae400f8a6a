"""
``B200Interface`` — a ``ModelInterface`` (reference ``src/liesel/goose/types.py:72-102``)
backed by the CUDA plan.

The reference's protocol has three methods: ``extract_position(position_keys, model_state)``,
``update_state(position, model_state)`` and ``log_prob(model_state)``. Its existing
implementations are ``DictInterface`` (``goose/interface.py:19-112``, the state is a plain dict
and ``update_state`` is ``model_state | position``, :101) and ``LieselInterface`` (:229-313,
the state is every node of the graph). This class keeps a PARAMETER-ONLY state like
``DictInterface`` — ``{name: array}`` for the coefficient vectors and the ``tau2`` values —
while the data lives in the plan; ``log_prob`` calls the batched CUDA evaluation.

Chain axis. Under the reference's engine every protocol method runs inside
``jax.vmap`` over chains (``goose/engine.py:442-448``); JAX is not installed here, so the chain
axis is explicit: every state leaf carries a leading dimension ``C`` and ``log_prob`` returns
``[C]``. This is exactly the operand layout ``jax.ffi.ffi_call(..., vmap_method="broadcast_all")``
hands to the C ABI.
"""

from __future__ import annotations

from typing import Callable, NamedTuple, Sequence

import torch

from .plan import Plan

Position = dict
ModelState = dict


class NodeState(NamedTuple):
    """Same two fields as the reference's ``NodeState`` (``model/nodes.py:144-154``): code written
    against a Liesel model state reads ``model_state[node_name].value``."""

    value: object
    outdated: bool = False


def _val(entry):
    return entry.value if isinstance(entry, NodeState) else entry


class B200Interface:
    def __init__(self, plan: Plan, shard=None):
        """``shard``: a ``dist.PeerGroup`` / ``dist.NcclComm`` when ``plan`` holds ONE RANK'S ROWS of
        the data (``ModelSpec.rows``): every evaluation is then summed over the ranks inside the C
        call, and each rank's interface returns the log-posterior of the full data."""
        self.plan = plan
        self.shard = shard
        spec = plan.spec
        self._pos_keys = spec.position_keys()
        self._aux_keys = spec.aux_keys()
        self._slices = {
            b.name + "_beta": slice(b.param_offset, b.param_offset + b.ncoef) for b in spec.blocks
        }

    @classmethod
    def from_model(cls, model, device=None, **match_kw) -> "B200Interface":
        """
        From a DistReg-shaped ``liesel.model.Model`` (``from_liesel.spec_from_model`` pattern-matches
        the graph; raises ``UnsupportedModelError`` for anything else, in which case the caller keeps
        ``gs.LieselInterface(model)``). The data is read out of the model once; afterwards only the
        parameter-only state travels through the engine.
        """
        from .from_liesel import spec_from_model

        spec = spec_from_model(model, **match_kw)
        return cls(Plan(spec, device=device) if device is not None else Plan(spec))

    # -- ModelInterface protocol (types.py:72-102) ------------------------------------
    # Two state layouts are understood. (1) Plain: ``{var_name: array}`` like ``DictInterface``.
    # (2) Liesel-named (``liesel_state``): ``{node_name: NodeState(value, outdated)}`` keyed by the
    # VALUE-NODE names of the reference's graph (``<var>_value``, ``model/nodes.py:2611-2613``), with
    # positions keyed by variable names — what ``LieselInterface`` exposes
    # (``goose/interface.py:267-302``, ``model/model.py`` ``extract_position``), so that code written
    # against a Liesel model state (``tau2_gibbs_kernel``: ``group.value_from(model_state, "a")``,
    # ``model/distreg.py:282-289``, ``model/nodes.py:3522-3545``) finds what it reads.
    @staticmethod
    def _resolve(key: str, model_state: ModelState) -> str | None:
        if key in model_state:
            return key
        if key + "_value" in model_state:
            return key + "_value"
        return None

    def extract_position(self, position_keys: Sequence[str], model_state: ModelState) -> Position:
        """Same contract as ``DictInterface.extract_position`` (``interface.py:75-88``); variable
        names resolve to their value node in a Liesel-named state."""
        out = Position()
        for key in position_keys:
            k = self._resolve(key, model_state)
            if k is None:
                raise KeyError(key)
            out[key] = _val(model_state[k])
        return out

    def update_state(self, position: Position, model_state: ModelState) -> ModelState:
        """``model_state | position`` (``interface.py:90-101``); evaluation is lazy."""
        unknown = [k for k in position if self._resolve(k, model_state) is None]
        if unknown:
            raise KeyError(f"unknown position keys {unknown}")
        new = dict(model_state)
        for key, v in position.items():
            k = self._resolve(key, model_state)
            new[k] = NodeState(v, False) if isinstance(model_state[k], NodeState) else v
        return new

    def log_prob(self, model_state: ModelState) -> torch.Tensor:
        """Unnormalised log-posterior per chain, ``[C]`` (``interface.py:304-313``)."""
        params, aux = self.pack(model_state)
        logp, _ = self.plan.logp_grad(params, aux, want_grad=False, shard=self.shard)
        return logp

    # -- what jax.value_and_grad(log_prob_fn) yields in NUTS/HMC (nuts.py:193-194) ----
    def log_prob_and_grad(self, model_state: ModelState, position_keys: Sequence[str] | None = None):
        params, aux = self.pack(model_state)
        logp, grad = self.plan.logp_grad(params, aux, want_grad=True, shard=self.shard)
        keys = self._pos_keys if position_keys is None else list(position_keys)
        bad = [k for k in keys if k not in self._slices]
        if bad:
            raise KeyError(f"no gradient available for {bad} (not a coefficient block)")
        return logp, {k: grad[:, self._slices[k]] for k in keys}

    # -- IWLSKernel(chol_info_fn=...) hook (iwls.py:130-145,241-243) ------------------
    def chol_info_fn(self, position_key: str) -> Callable[[ModelState], torch.Tensor]:
        block = self._block_of(position_key)

        def fn(model_state: ModelState) -> torch.Tensor:
            params, aux = self.pack(model_state)
            _, _, chol = self.plan.iwls(block, params, aux, want_chol=True, shard=self.shard)
            return chol

        return fn

    def score_and_info(self, position_key: str, model_state: ModelState, want_chol: bool = True):
        params, aux = self.pack(model_state)
        return self.plan.iwls(self._block_of(position_key), params, aux, want_chol=want_chol, shard=self.shard)

    # -- state helpers -----------------------------------------------------------------
    def initial_state(self, num_chains: int, position: dict | None = None) -> ModelState:
        """
        Replicated start state: zeros for coefficients (``distreg.py:101,165``) and the blocks'
        ``tau2_init`` (reference default 10000.0, ``distreg.py:163``), overridable per key.
        """
        spec, dev, dt = self.plan.spec, self.plan.device, self.plan.dtype
        state: ModelState = {}
        for b in spec.blocks:
            state[b.name + "_beta"] = torch.zeros((num_chains, b.ncoef), dtype=dt, device=dev)
            if b.tau2_slot >= 0:
                state[b.prior.tau2_name] = torch.full((num_chains,), b.prior.tau2_init, dtype=dt, device=dev)
        for k, v in (position or {}).items():
            if k not in state:
                raise KeyError(k)
            v = torch.as_tensor(v, dtype=dt, device=dev)
            state[k] = v.expand(state[k].shape).contiguous() if v.dim() < state[k].dim() else v.contiguous()
        return state

    def liesel_state(self, num_chains: int, position: dict | None = None) -> ModelState:
        """
        The start state under the reference's node names: ``<smooth>_beta_value``,
        ``<smooth>_tau2_value`` plus the small constants an unmodified ``tau2_gibbs_kernel`` reads by
        value-node name — ``<smooth>_a_value``, ``_b_value``, ``_rank_value``, ``_K_value``
        (``model/distreg.py:155-161,282-289``) — every entry a ``NodeState``. The constants carry no
        chain axis (under the engine's ``jax.vmap`` they are the per-chain scalars the kernel sees).
        """
        plain = self.initial_state(num_chains, position)
        dev, dt = self.plan.device, self.plan.dtype
        state: ModelState = {k + "_value": NodeState(v, False) for k, v in plain.items()}
        for b in self.plan.spec.blocks:
            if b.tau2_slot < 0:
                continue
            pr = b.prior
            state[b.name + "_K_value"] = NodeState(torch.as_tensor(pr.K, dtype=dt, device=dev), False)
            state[b.name + "_rank_value"] = NodeState(torch.tensor(float(pr.rank), dtype=dt, device=dev), False)
            if pr.ig_a is not None:
                state[b.name + "_a_value"] = NodeState(torch.tensor(float(pr.ig_a), dtype=dt, device=dev), False)
                state[b.name + "_b_value"] = NodeState(torch.tensor(float(pr.ig_b), dtype=dt, device=dev), False)
        return state

    def _get(self, model_state: ModelState, key: str):
        k = self._resolve(key, model_state)
        if k is None:
            raise KeyError(key)
        return _val(model_state[k])

    def pack(self, model_state: ModelState):
        """State dict -> (params [C, P], aux [C, A]) in the plan's flat layout."""
        params = torch.cat([self._get(model_state, k) for k in self._pos_keys], dim=1).contiguous()
        aux = None
        if self._aux_keys:
            aux = torch.stack([self._get(model_state, k).reshape(-1) for k in self._aux_keys], dim=1).contiguous()
        return params, aux

    def unpack(self, params: torch.Tensor, aux: torch.Tensor | None = None) -> ModelState:
        state = {k: params[:, s] for k, s in self._slices.items()}
        for i, k in enumerate(self._aux_keys):
            state[k] = aux[:, i]
        return state

    def _block_of(self, position_key: str) -> int:
        if position_key not in self._slices:
            raise KeyError(position_key)
        return self._pos_keys.index(position_key)
