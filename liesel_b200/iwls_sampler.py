"""
Batched IWLS-within-Gibbs sampler for distributional regression models: the device-side
counterpart of what ``dist_reg_mcmc`` configures in the reference (``model/distreg.py:300-349``):
an ``IWLSKernel`` per coefficient block (``distreg.py:113-121,194-201``) and a Gibbs kernel per
smoothing variance (``tau2_gibbs_kernel``, ``distreg.py:277-297``).

One IWLS transition follows ``IWLSKernel._standard_transition`` (``goose/iwls.py:299-345``):

    score, chol(info)        at the current position                         (:320-321)
    mu = beta + step^2/2 * solve(chol, score);  prop ~ N(mu, (chol/step)^-T (chol/step)^-1)  (:323-324)
    forward log q                                                             (:327)
    score, chol(info)        at the proposal;  backward log q                (:331-336)
    Metropolis-Hastings step with correction = bwd - fwd  (``goose/mh.py:17-71``)
    error codes: 2 = indefinite information, identity fallback (``_safe_chol``, :245-290); 90 = NaN
    acceptance probability (``mh.py:53-57``); step size by dual averaging during warm-up
    (defaults ``initial_step_size=0.01``, ``da_target_accept=0.8``, :134-145).

Score, information and Cholesky come from ONE pass over the rows per evaluation point
(``Plan.iwls``), the proposal algebra from ``lslb_iwls_propose``; the log-posterior of the current
state is carried along (the reference reads it from the cached model state, ``mh.py:48``).
All chains advance in lock-step; every tensor has a leading chain axis.
"""

from __future__ import annotations

import math

import torch

from .plan import Plan, chol as chol_info, iwls_propose, iwls_propose_diag, mh_step


class IWLSGibbsSampler:
    def __init__(self, plan: Plan, params: torch.Tensor, aux: torch.Tensor | None, seed: int = 0,
                 blocks: list[str] | None = None, initial_step_size: float = 0.01,
                 da_target_accept: float = 0.8, da_gamma: float = 0.05, da_kappa: float = 0.75,
                 da_t0: int = 10, gibbs: bool = True, constraints: dict | None = None,
                 group_accept: str = "joint"):
        """``constraints``: block name -> ``Z [k, k_r]`` with orthonormal columns (e.g.
        ``reparam.sum_to_zero_basis`` of the basis' column means): the block is sampled in the
        coordinates ``gamma`` of ``beta = Z gamma`` — the constrained design ``B Z`` with penalty
        ``Z' K Z`` that a Liesel user passes to ``add_np_smooth`` for identifiability next to an
        intercept. Score and information transform as ``Z' s`` and ``Z' I Z``; the start values are
        projected onto the constrained space.

        ``group_accept``: how a random-intercept block is accepted. ``"joint"``: one Metropolis-Hastings
        decision for the whole block, the law of the reference's ``IWLSKernel`` on that block (its acceptance
        rate, hence the adapted step size, falls with the number of groups). ``"per_group"`` (an extension,
        not a reference mode): given the other blocks the posterior of b factorises over the groups, so every
        group is accepted on its own ratio - G independent one-dimensional IWLS kernels in one pass; the step
        size then does not depend on G."""
        if group_accept not in ("joint", "per_group"):
            raise ValueError("group_accept must be 'joint' or 'per_group'")
        self.group_accept = group_accept
        self.plan, self.spec = plan, plan.spec
        self.params = params.clone().contiguous()
        self.Z = {k: torch.as_tensor(v, dtype=params.dtype, device=params.device).contiguous()
                  for k, v in (constraints or {}).items()}
        for name, Z in self.Z.items():
            b = self.spec.block(name)
            sl = slice(b.param_offset, b.param_offset + b.ncoef)
            self.params[:, sl] = (self.params[:, sl] @ Z) @ Z.t()
        self.aux = None if aux is None else aux.clone().contiguous()
        self.C = params.shape[0]
        self.dev, self.dt = params.device, params.dtype
        # a random-intercept block (kind 2) is sampled with the same transition on its DIAGONAL information
        # (``group_transition``); by default it is part of the sweep like every other coefficient block
        self.blocks = [b for b in self.spec.blocks if blocks is None or b.name in blocks]
        self.gibbs_blocks = [b for b in self.spec.blocks if gibbs and b.tau2_slot >= 0 and b.prior.ig_a is not None]
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed)
        self.target, self.gamma, self.kappa, self.t0 = da_target_accept, da_gamma, da_kappa, da_t0
        self.step = {b.name: torch.full((self.C,), initial_step_size, dtype=self.dt, device=self.dev)
                     for b in self.blocks}
        self._da_init()
        self.logp, _ = plan.logp_grad(self.params, self.aux, want_grad=False)
        self.logp = self.logp.clone()
        self.evals = {"iwls": 0, "logp": 1}

    # -- dual averaging (goose/da.py) ----------------------------------------------------------
    def _da_init(self):
        self.da = {k: {"err": torch.zeros_like(v), "avg": torch.log(v), "mu": torch.log(10.0 * v)}
                   for k, v in self.step.items()}

    def _da_step(self, name, acc, t_in_epoch):
        t = t_in_epoch + 1
        d = self.da[name]
        d["err"] = d["err"] + (self.target - acc)
        log_step = d["mu"] - d["err"] * math.sqrt(t) / (self.gamma * (self.t0 + t))
        self.step[name] = torch.exp(log_step)
        eta = t ** (-self.kappa)
        d["avg"] = (1 - eta) * d["avg"] + eta * log_step

    def _da_finalize(self):
        for k in self.step:
            self.step[k] = torch.exp(self.da[k]["avg"])

    # -- pieces ----------------------------------------------------------------------------------
    def _safe_chol(self, chol):
        """``_safe_chol`` with ``fallback_chol_info="identity"`` (iwls.py:245-290): code 2."""
        bad = torch.isnan(chol).flatten(1).any(dim=1)
        eye = torch.eye(chol.shape[-1], dtype=chol.dtype, device=chol.device)
        chol = torch.where(bad[:, None, None], eye, chol)
        return chol.contiguous(), bad.to(torch.int32) * 2

    def _score_chol(self, blk, params):
        """Score and Cholesky factor of the (ridged) information in the SAMPLED coordinates."""
        Z = self.Z.get(blk.name)
        if Z is None:
            score, _, ch = self.plan.iwls(blk.name, params, self.aux)
            return score, ch
        score, info, _ = self.plan.iwls(blk.name, params, self.aux, want_chol=False)
        info_r = torch.matmul(Z.t(), torch.matmul(info, Z)).contiguous()
        return (score @ Z).contiguous(), chol_info(info_r)

    def iwls_transition(self, blk, adapt_t: int | None = None):
        sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
        step = self.step[blk.name]
        Z = self.Z.get(blk.name)
        beta = self.params[:, sl]
        pos = (beta if Z is None else beta @ Z).contiguous()
        score, chol = self._score_chol(blk, self.params)
        chol, code = self._safe_chol(chol)
        z = torch.randn(pos.shape, generator=self.gen, device=self.dev, dtype=self.dt)
        prop, fwd = iwls_propose(pos, score, chol, step, z=z)
        params_prop = self.params.clone()
        params_prop[:, sl] = prop if Z is None else prop @ Z.t()
        score_p, chol_p = self._score_chol(blk, params_prop)
        chol_p, _ = self._safe_chol(chol_p)
        _, bwd = iwls_propose(prop, score_p, chol_p, step, x_eval=pos)
        lp_prop, _ = self.plan.logp_grad(params_prop, self.aux, want_grad=False)
        self.evals["iwls"] += 2
        self.evals["logp"] += 1
        u = torch.rand(self.C, generator=self.gen, device=self.dev, dtype=self.dt)
        # mh_step (goose/mh.py:48-68: NaN -> -inf + code 90, acc = min(exp, 1), accept <=> u <= acc) and
        # the selection of the accepted state, one launch
        mh_code, acc, accept = mh_step(u, self.logp, lp_prop.contiguous(), (bwd - fwd).contiguous(),
                                       self.params, params_prop)
        code = code + mh_code
        if adapt_t is not None:
            self._da_step(blk.name, acc, adapt_t)
        return {"error_code": code, "acceptance_prob": acc, "position_moved": accept}

    def group_transition(self, blk, adapt_t: int | None = None):
        """The IWLS transition for the random-intercept block: score and the diagonal of the information from
        ``Plan.iwls_group`` (one pass over the rows), ridge / factor / mean / draw / log-density elementwise
        (``lslb_iwls_propose_diag``). Same law as ``IWLSKernel._standard_transition`` on that block, whose
        information matrix the reference would form as a dense G x G Hessian."""
        if self.group_accept == "per_group":
            return self._group_transition_per_group(blk, adapt_t)
        sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
        step = self.step[blk.name]
        pos = self.params[:, sl].contiguous()
        score, info = self.plan.iwls_group(self.params, self.aux)
        z = torch.randn(pos.shape, generator=self.gen, device=self.dev, dtype=self.dt)
        prop, fwd, code = iwls_propose_diag(pos, score, info, step, z=z)
        params_prop = self.params.clone()
        params_prop[:, sl] = prop
        score_p, info_p = self.plan.iwls_group(params_prop, self.aux)
        _, bwd, _ = iwls_propose_diag(prop, score_p, info_p, step, x_eval=pos)
        lp_prop, _ = self.plan.logp_grad(params_prop, self.aux, want_grad=False)
        self.evals["iwls"] += 2
        self.evals["logp"] += 1
        u = torch.rand(self.C, generator=self.gen, device=self.dev, dtype=self.dt)
        mh_code, acc, accept = mh_step(u, self.logp, lp_prop.contiguous(), (bwd - fwd).contiguous(),
                                       self.params, params_prop)
        if adapt_t is not None:
            self._da_step(blk.name, acc, adapt_t)
        return {"error_code": code + mh_code, "acceptance_prob": acc, "position_moved": accept}

    def _group_transition_per_group(self, blk, adapt_t):
        """Every group on its own Metropolis-Hastings ratio: log-posterior share of the group (its rows'
        log-likelihood + its prior term, ``lslb_iwls_group``'s third output) and the univariate proposal
        densities (``lslb_iwls_propose_diag`` with the entry-local ridge). The step size stays one number per
        chain, adapted on the mean acceptance probability over the groups."""
        sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
        step = self.step[blk.name]
        pos = self.params[:, sl].contiguous()
        score, info, ll = self.plan.iwls_group(self.params, self.aux, want_loglik=True)
        z = torch.randn(pos.shape, generator=self.gen, device=self.dev, dtype=self.dt)
        prop, _, code, fwd = iwls_propose_diag(pos, score, info, step, z=z, elementwise=True)
        params_prop = self.params.clone()
        params_prop[:, sl] = prop
        score_p, info_p, ll_p = self.plan.iwls_group(params_prop, self.aux, want_loglik=True)
        _, _, _, bwd = iwls_propose_diag(prop, score_p, info_p, step, x_eval=pos, elementwise=True)
        log_acc = ll_p - ll + bwd - fwd
        nan = torch.isnan(log_acc)
        log_acc = torch.where(nan, torch.full_like(log_acc, -float("inf")), log_acc)  # mh.py:53-57, per group
        acc = torch.exp(torch.clamp(log_acc, max=0.0))
        u = torch.rand(pos.shape, generator=self.gen, device=self.dev, dtype=self.dt)
        accept = u <= acc
        self.params[:, sl] = torch.where(accept, prop, pos)
        self.logp, _ = self.plan.logp_grad(self.params, self.aux, want_grad=False)
        self.logp = self.logp.clone()
        self.evals["iwls"] += 2
        self.evals["logp"] += 1
        acc_chain = acc.mean(dim=1)
        if adapt_t is not None:
            self._da_step(blk.name, acc_chain, adapt_t)
        return {"error_code": code + 90 * nan.any(dim=1).to(torch.int32), "acceptance_prob": acc_chain,
                "position_moved": accept.any(dim=1)}

    def gibbs_transition(self, blk):
        """tau2 | beta ~ InverseGamma(a + rank/2, b + beta'K beta/2) (distreg.py:277-297)."""
        pr = blk.prior
        a = pr.ig_a + 0.5 * pr.rank
        b = pr.ig_b + 0.5 * self.plan.quadform(blk.name, self.params).double()
        gam = torch._standard_gamma(torch.full((self.C,), a, dtype=torch.float64, device=self.dev),
                                    generator=self.gen)
        self.aux[:, blk.tau2_slot] = (b / gam).to(self.dt)

    def sweep(self, adapt_t: int | None = None):
        infos = {}
        for blk in self.blocks:
            infos[blk.name] = (self.group_transition if blk.kind == 2 else self.iwls_transition)(blk, adapt_t)
        if self.gibbs_blocks:
            for blk in self.gibbs_blocks:
                self.gibbs_transition(blk)
            self.logp, _ = self.plan.logp_grad(self.params, self.aux, want_grad=False)
            self.logp = self.logp.clone()
            self.evals["logp"] += 1
        return infos

    def warmup(self, num: int):
        self._da_init()
        out = [self.sweep(adapt_t=t) for t in range(num)]
        self._da_finalize()
        return out

    def sample(self, num: int):
        P = self.params.shape[1]
        A = 0 if self.aux is None else self.aux.shape[1]
        draws = torch.empty((num, self.C, P), dtype=self.dt, device=self.dev)
        aux_draws = torch.empty((num, self.C, A), dtype=self.dt, device=self.dev)
        infos = []
        for t in range(num):
            infos.append(self.sweep())
            draws[t] = self.params
            if A:
                aux_draws[t] = self.aux
        return draws, aux_draws, infos
