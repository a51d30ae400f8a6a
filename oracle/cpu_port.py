"""
CPU port of the reference's algorithm for the hot path, used ONLY as the timed CPU baseline
(bench.py ``cpu_baseline`` and ``--impl reference``) — TEST/BENCH INFRASTRUCTURE, see
oracle/__init__.py.

It follows the reference's formulation, not ours: the B-spline design matrix is DENSE [n x k]
(``contrib/splines.py:196``), every smooth is ``jnp.dot(X, beta)`` (``model/distreg.py:105,175``),
the predictor is their sum (:238-240), the response is ``tfd.Normal(loc, exp(.)).log_prob(y)``
summed over rows (``model/nodes.py:1139-1152``, ``model/model.py:50-53``), the gradient is the
reverse-mode transpose ``X^T r``. Chains are batched the way ``jax.vmap`` + XLA would batch them:
one GEMM ``X @ B`` with ``B`` = [k x C], executed by torch's CPU BLAS on all host threads in
float32 (JAX's default dtype). This is kinder to the CPU than the reference's engine, which
replicates X per chain (``goose/builder.py:304-307``); it is labelled "port", not "reference".
"""

from __future__ import annotations

import math
import time

import numpy as np
import torch

from . import sam
from . import splines as S


class CpuPort:
    def __init__(self, spec, threads: int | None = None):
        if threads:
            torch.set_num_threads(threads)
        self.threads = torch.get_num_threads()
        self.spec = spec
        prep = sam.Prepared(spec, work_dtype=np.float32)
        self.y = torch.from_numpy(prep.y.astype(np.float32))
        self.designs = []
        for ent in prep.blocks:
            blk = ent["blk"]
            if blk.kind == sam.KIND_GROUP:
                raise NotImplementedError("cpu_port covers the dense/spline/intercept terms")
            Xd = np.ascontiguousarray(prep.dense_design(ent), dtype=np.float32)
            self.designs.append((blk, torch.from_numpy(Xd)))
        if spec.family != sam.FAMILY_NORMAL or not spec.has_scale_predictor:
            raise NotImplementedError("cpu_port times the Gaussian location-scale workload")

    def logp_grad(self, params: np.ndarray, aux: np.ndarray):
        """params [C, P], aux [C, A] float32 -> (logp [C], grad [C, P]) as torch tensors."""
        spec = self.spec
        th = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float32))
        ax = torch.from_numpy(np.ascontiguousarray(aux, dtype=np.float32))
        C = th.shape[0]
        eta = [torch.zeros((self.y.shape[0], C)), torch.zeros((self.y.shape[0], C))]
        for blk, X in self.designs:
            B = th[:, blk.param_offset: blk.param_offset + blk.ncoef].t().contiguous()
            eta[blk.predictor] = eta[blk.predictor] + X @ B
        sigma = torch.exp(eta[1])
        z = (self.y[:, None] - eta[0]) / sigma
        ll = -0.5 * z * z - eta[1] - 0.5 * math.log(2 * math.pi)
        logp = ll.sum(dim=0)
        r = [z / sigma, z * z - 1.0]
        grad = torch.zeros_like(th)
        for blk, X in self.designs:
            sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
            beta = th[:, sl]
            g = (X.t() @ r[blk.predictor]).t()
            pr = blk.prior
            if pr is not None and hasattr(pr, "K"):
                K = torch.from_numpy(np.asarray(pr.K, dtype=np.float32))
                tau2 = ax[:, blk.tau2_slot] if blk.tau2_slot >= 0 else torch.full((C,), pr.tau2_init)
                kb = beta @ K
                g = g - kb / tau2[:, None]
                logp = logp + 0.5 * (-(kb * beta).sum(dim=1) / tau2 - pr.rank * math.log(2 * math.pi)
                                     + pr.log_pdet - pr.rank * torch.log(tau2))
                if pr.ig_a is not None and blk.tau2_slot >= 0:
                    logp = logp + (pr.ig_a * math.log(pr.ig_b) - math.lgamma(pr.ig_a)
                                   - (pr.ig_a + 1) * torch.log(tau2) - pr.ig_b / tau2)
            elif pr is not None:
                g = g - (beta - pr.m) / (pr.s * pr.s)
                logp = logp + (-0.5 * ((beta - pr.m) / pr.s) ** 2 - math.log(pr.s)
                               - 0.5 * math.log(2 * math.pi)).sum(dim=1)
            grad[:, sl] = g
        return logp + spec.const_term, grad

    def time(self, params, aux, repeats: int = 3, warmup: int = 1):
        for _ in range(warmup):
            self.logp_grad(params, aux)
        best = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            self.logp_grad(params, aux)
            best.append(time.perf_counter() - t0)
        return float(np.median(best))


class CpuEvaluator:
    """Duck-typed evaluator for ``liesel_b200.nuts_gibbs.NutsGibbs`` on the CPU port (the CPU arm of
    the ESS measurement in ``bench.py`` and the CPU tests of the sampler wiring)."""

    def __init__(self, spec, threads: int | None = None):
        self.port = CpuPort(spec, threads)
        self.device, self.dtype = torch.device("cpu"), torch.float32
        self._K = {b.name: torch.from_numpy(np.asarray(b.prior.K, dtype=np.float64))
                   for b in spec.blocks if b.prior is not None and hasattr(b.prior, "K")}
        self._sl = {b.name: slice(b.param_offset, b.param_offset + b.ncoef) for b in spec.blocks}

    def logp_grad(self, q, aux, out=None):
        lp, g = self.port.logp_grad(q.numpy(), aux.numpy())
        if out is not None:
            out[0].copy_(lp)
            out[1].copy_(g)
            return out
        return lp, g

    def quadform(self, name, q):
        b = q[:, self._sl[name]].double()
        return ((b @ self._K[name]) * b).sum(dim=1)

    def column_means(self, spec) -> dict:
        """Column means of every spline block's DENSE basis (what ``Plan.column_means`` returns)."""
        return {blk.name: X.double().mean(dim=0).numpy() for blk, X in self.port.designs if blk.kind == sam.KIND_SPLINE}
