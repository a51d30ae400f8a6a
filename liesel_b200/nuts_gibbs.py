"""
NUTS-within-Gibbs run for distributional regression models: ONE NUTS block over all coefficient
vectors (reference ``NUTSKernel`` defaults, ``goose/nuts.py:165-176``) and, after every transition,
a Gibbs draw of every smoothing variance from its full conditional
``tau2 | beta ~ InverseGamma(a + rank/2, b + beta^T K beta / 2)`` (``tau2_gibbs_kernel``,
``model/distreg.py:277-297``). This is the sampling scheme behind the "NUTS ESS/sec" half of the
metric (SURVEY §8 d-ii); ``bench.py`` and ``tools/ess_bench.py`` run it.

The evaluator is duck-typed (``logp_grad(q, aux, out=None)``, ``quadform(block_name, q)``): the
B200 plan in production (:class:`PlanEvaluator`); ``bench.py`` substitutes the CPU port for the
reported CPU arm. An optional :class:`~liesel_b200.reparam.LinearReparam` (sum-to-zero constraint)
sits between the sampler's coordinates and the evaluator's.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from .nuts import BatchedNUTS


class PlanEvaluator:
    def __init__(self, plan):
        self.plan = plan
        self.device, self.dtype = plan.device, plan.dtype

    def logp_grad(self, q, aux, out=None):
        return self.plan.logp_grad(q, aux, out=out)

    def quadform(self, name, q):
        return self.plan.quadform(name, q)


class NutsGibbs:
    def __init__(self, spec, evaluator, q0: torch.Tensor, aux0: torch.Tensor | None, reparam=None,
                 seed: int = 0, **nuts_kw):
        """``q0 [C, P]`` in the evaluator's (full) coordinates; ``nuts_kw`` goes to ``BatchedNUTS``
        (``compact=True``, ``graphs=True`` need an evaluator that writes into ``out=``)."""
        self.spec, self.ev, self.rep = spec, evaluator, reparam
        self.C = q0.shape[0]
        self.dev, self.dt = q0.device, q0.dtype
        self.aux = None if aux0 is None else aux0.clone().contiguous()
        self.gibbs_blocks = [b for b in spec.blocks if b.tau2_slot >= 0 and b.prior.ig_a is not None]
        self._full_q = torch.empty_like(q0)
        self._full_g = torch.empty_like(q0)
        self._aux_sel = (None, None)
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed + 7919)
        start = q0 if reparam is None else reparam.reduce(q0)
        into = self._f_into if q0.is_cuda else None
        self.nuts = BatchedNUTS(self._f, start.contiguous(), seed=seed, logp_grad_into=into, **nuts_kw)

    # -- evaluation in the sampler's coordinates -------------------------------------------------
    def full(self, q_r: torch.Tensor) -> torch.Tensor:
        return q_r if self.rep is None else self.rep.expand(q_r)

    def _f(self, q_r):
        if self.rep is None:
            return self.ev.logp_grad(q_r.contiguous(), self.aux)
        lp, g = self.ev.logp_grad(self.rep.expand(q_r), self.aux)
        return lp, self.rep.reduce_grad(g)

    def _f_into(self, q_r, lp, g_r, chains=None):
        """``chains``: indices of the chains of a compacted batch (selects their tau2 rows)."""
        aux = self.aux
        if chains is not None and aux is not None:
            if self._aux_sel[0] is not chains:
                self._aux_sel = (chains, aux.index_select(0, chains).contiguous())
            aux = self._aux_sel[1]
        if self.rep is None:
            self.ev.logp_grad(q_r, aux, out=(lp, g_r))
            return
        Cb = q_r.shape[0]
        qf, gf = self._full_q[:Cb], self._full_g[:Cb]
        self.rep.expand(q_r, out=qf)
        self.ev.logp_grad(qf, aux, out=(lp, gf))
        self.rep.reduce_grad(gf, out=g_r)

    # -- Gibbs update of the smoothing variances -------------------------------------------------
    def _gibbs(self, s: BatchedNUTS):
        if not self.gibbs_blocks:
            return
        q = self.full(s.q).contiguous()
        for blk in self.gibbs_blocks:
            pr = blk.prior
            a = pr.ig_a + 0.5 * pr.rank
            b = pr.ig_b + 0.5 * self.ev.quadform(blk.name, q).double()
            gam = torch._standard_gamma(torch.full((self.C,), a, dtype=torch.float64, device=self.dev),
                                        generator=self.gen)
            self.aux[:, blk.tau2_slot] = (b / gam).to(self.aux.dtype)
        self._aux_sel = (None, None)
        s.refresh()

    # -- epochs --------------------------------------------------------------------------------
    def warmup(self, warmup_duration: int = 1000, **epoch_kw):
        return self.nuts.warmup(warmup_duration, hook=self._gibbs, **epoch_kw)

    def sample(self, num_samples: int):
        """(draws [S, C, P_r] in the sampler's coordinates, infos)."""
        return self.nuts.sample(num_samples, hook=self._gibbs)


class AsyncNutsGibbs:
    """The same sampling scheme on the asynchronous driver (``nuts_async.AsyncNUTS``): every chain runs
    its own trees and its own Gibbs updates, the batch shares one evaluation per step. The Gibbs blocks
    are handed to the driver in the SAMPLER's coordinates: ``beta' K beta = gamma' (T_b' K T_b) gamma``
    with ``T_b`` the block's rows of the reparametrisation (identity rows without one)."""

    def __init__(self, spec, evaluator, q0: torch.Tensor, aux0: torch.Tensor | None, warmup: int, samples: int,
                 reparam=None, seed: int = 0, backend: str | None = None, epoch_kw: dict | None = None, **nuts_kw):
        from .nuts_async import AsyncNUTS

        self.spec, self.ev, self.rep = spec, evaluator, reparam
        self.C = q0.shape[0]
        self.dev, self.dt = q0.device, q0.dtype
        self.aux = None if aux0 is None else aux0.clone().contiguous()
        self._full_q = torch.empty_like(q0)
        self._full_g = torch.empty_like(q0)
        self._out_ok = bool(q0.is_cuda)
        start = (q0 if reparam is None else reparam.reduce(q0)).contiguous()
        Pr = start.shape[1]
        T = (torch.eye(q0.shape[1], dtype=torch.float64) if reparam is None
             else reparam.T.detach().double().cpu())
        gibbs = []
        for b in spec.blocks:
            if b.tau2_slot < 0 or b.prior.ig_a is None:
                continue
            K = torch.as_tensor(np.asarray(b.prior.K), dtype=torch.float64)
            Tb = T[b.param_offset: b.param_offset + b.ncoef, :]
            gibbs.append(dict(M=Tb.t() @ K @ Tb, b0=float(b.prior.ig_b),
                              a_post=float(b.prior.ig_a) + 0.5 * float(b.prior.rank), slot=int(b.tau2_slot)))
        assert all(g["M"].shape == (Pr, Pr) for g in gibbs)
        nuts_kw.setdefault("compact_tail", True)
        self.nuts = AsyncNUTS(self._f_into, start, warmup, samples, seed=seed, aux=self.aux, gibbs=gibbs,
                              backend=backend, epoch_kw=epoch_kw, **nuts_kw)

    def full(self, q_r: torch.Tensor) -> torch.Tensor:
        return q_r if self.rep is None else self.rep.expand(q_r)

    def _f_into(self, q_r, lp, g_r, chains=None):
        """``chains``: indices of the chains of a gathered batch (their CURRENT tau2 rows are selected:
        the Gibbs updates change ``aux`` from step to step)."""
        n = q_r.shape[0]
        aux = self.aux
        if chains is not None and aux is not None:
            aux = aux.index_select(0, chains)
        if self.rep is None:
            qf, gf = q_r, g_r
        else:
            qf, gf = self._full_q[:n], self._full_g[:n]
            self.rep.expand(q_r, out=qf)
        if self._out_ok:
            self.ev.logp_grad(qf, aux, out=(lp, gf))
        else:
            l, g = self.ev.logp_grad(qf, aux)
            lp.copy_(l)
            gf.copy_(g)
        if self.rep is not None:
            self.rep.reduce_grad(gf, out=g_r)

    def run(self, **kw) -> dict:
        return self.nuts.run(**kw)

    def draws(self) -> torch.Tensor:
        return self.nuts.draws()


def summarise_async(run: "AsyncNutsGibbs", timing: dict) -> dict:
    """The columns of :func:`summarise` for an asynchronous run."""
    from . import diagnostics

    d = run.draws()
    t0 = time.perf_counter()
    ess = diagnostics.ess_bulk_device(d)
    rhat = diagnostics.split_rhat_device(d)
    sm = run.nuts.summary()
    sp, sw = timing["seconds_posterior"], timing["seconds_warmup"]
    return {
        "ess_bulk_min": float(np.min(ess)), "ess_bulk_median": float(np.median(ess)),
        "split_rhat_max": float(np.max(rhat)),
        "draws": int(d.shape[0]), "chains": int(d.shape[1]), "coordinates": int(d.shape[2]),
        "posterior_seconds": sp, "ess_per_sec_posterior": float(np.min(ess) / sp),
        "mean_accept": sm["mean_accept"], "mean_treedepth": sm["mean_treedepth"],
        "divergent_frac": sm["divergent_frac"], "mean_leapfrogs": sm["mean_leapfrogs"],
        "batched_evals_posterior": int(timing["steps_posterior"]),
        "batched_evals_warmup": int(timing["steps_warmup"]),
        "diagnostics_seconds": time.perf_counter() - t0,
        "warmup_seconds": sw, "ess_per_sec_incl_warmup": float(np.min(ess) / (sp + sw)),
        "step_size_median": sm["step_size_median"],
    }


def summarise(draws: torch.Tensor, infos, seconds_posterior: float, seconds_warmup: float | None = None) -> dict:
    """Bulk-ESS / split-R-hat of all coordinates (``goose/summary_m.py:756-760`` columns) + sampler
    statistics; ``draws [S, C, P]`` stay on their device."""
    from . import diagnostics

    t0 = time.perf_counter()
    ess = diagnostics.ess_bulk_device(draws)
    rhat = diagnostics.split_rhat_device(draws)
    out = {
        "ess_bulk_min": float(np.min(ess)), "ess_bulk_median": float(np.median(ess)),
        "split_rhat_max": float(np.max(rhat)),
        "draws": int(draws.shape[0]), "chains": int(draws.shape[1]), "coordinates": int(draws.shape[2]),
        "posterior_seconds": seconds_posterior,
        "ess_per_sec_posterior": float(np.min(ess) / seconds_posterior),
        "mean_accept": float(np.mean([float(i.acceptance_prob.mean()) for i in infos])),
        "mean_treedepth": float(np.mean([float(i.treedepth.float().mean()) for i in infos])),
        "max_treedepth_seen": int(max(int(i.treedepth.max()) for i in infos)),
        "divergent_frac": float(np.mean([float(i.divergent.float().mean()) for i in infos])),
        "batched_evals_posterior": int(sum(i.evaluations for i in infos)),
        "diagnostics_seconds": time.perf_counter() - t0,
    }
    if seconds_warmup is not None:
        out["warmup_seconds"] = seconds_warmup
        out["ess_per_sec_incl_warmup"] = float(np.min(ess) / (seconds_posterior + seconds_warmup))
    return out
