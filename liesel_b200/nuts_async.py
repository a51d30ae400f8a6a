"""
Asynchronous batched NUTS: every chain runs its OWN tree, transition after transition, and the batch
only shares the log-posterior+gradient evaluation — one batched call per step, every slot busy.

Why: under the lock-step driver (``nuts.BatchedNUTS``; the reference under ``jax.vmap`` behaves the
same, ``goose/engine.py:442-448``) a transition costs as many batched evaluations as its DEEPEST
tree. On H (n = 1e6, 1024 chains) the depth spreads from 5 to 9: ~528 sequential evaluations per
transition with a quarter of the batch alive on average (``profiles/r02_ess_diag_*.json``), and the
evaluation kernel is far less efficient on small batches (0.10 ms for 64 chains, 0.69 ms for 1024).
Here a chain that finishes its tree starts its next transition in the very next step, so the batch
stays full: the same chain-evaluations run at the full-batch rate.

One step = [state-machine kernel] -> [batched evaluation at ``q_n``]. The state machine (per chain:
leaf bookkeeping, subtree merge, transition end with dual averaging / mass-matrix windows / the draw /
the Gibbs update of smoothing variances, transition start, next leapfrog position) is ONE CUDA kernel
(``lslb_nuts_async_step``, one warp per chain, ``csrc/nuts_async.cu``); ``_step_torch`` is the same
program written with masked torch ops — the CPU path and the reference the kernel is tested against
step by step (random numbers are INPUTS of a step: one ``randn``/``rand``/``standard_gamma`` launch
per step for the whole batch, consumed by whichever chains need them).

Semantics per chain are those of ``BatchedNUTS.transition`` / ``warmup`` (reference ``NUTSKernel``
defaults, dual averaging ``goose/da.py``, mass-matrix windows ``goose/mm.py``, Stan-style epochs
``goose/warmup.py``) and of ``tau2_gibbs_kernel`` (``model/distreg.py:277-297``) for the optional
Gibbs blocks; only the interleaving of the chains differs, so the draws of one chain have the law of
the lock-step driver's draws (tests: posterior recovery pins of ``tests/goose/kernels/model_lm.py``).
"""

from __future__ import annotations

import math
from typing import Callable

import torch

from .nuts import BatchedNUTS, stan_epochs

REFRESH, LEAF, DONE = 0, 1, 2
F_DIVERGED, F_TURNED, F_TURN_S, F_DIV_S = 1, 2, 4, 8


def _popcount(x: torch.Tensor) -> torch.Tensor:
    x = x.to(torch.int64)
    c = torch.zeros_like(x)
    for b in range(12):
        c += (x >> b) & 1
    return c


def _trailing_ones(x: torch.Tensor) -> torch.Tensor:
    x = x.to(torch.int64)
    c = torch.zeros_like(x)
    run = torch.ones_like(x, dtype=torch.bool)
    for b in range(12):
        bit = ((x >> b) & 1).bool()
        run = run & bit
        c += run.to(torch.int64)
    return c


class AsyncNUTS:
    def __init__(self, logp_grad_into: Callable, q0: torch.Tensor, warmup: int, samples: int, *,
                 max_treedepth: int = 10, da_target_accept: float = 0.8, da_gamma: float = 0.05,
                 da_kappa: float = 0.75, da_t0: int = 10, divergence_threshold: float = 1000.0, seed: int = 0,
                 initial_step_size: float | None = None, aux: torch.Tensor | None = None, gibbs: list | None = None,
                 backend: str | None = None, epoch_kw: dict | None = None, compact_tail: bool = False):
        """
        ``logp_grad_into(q [C, P], logp_out [C], grad_out [C, P])`` evaluates the batch into the given
        buffers (it may read ``aux``, which this driver updates in place when ``gibbs`` is given).
        ``gibbs``: list of ``dict(M=[P, P] tensor, b0=float, a_post=float, slot=int)`` — after every
        transition ``aux[c, slot] = (b0 + q' M q / 2) / Gamma(a_post, 1)`` and the chain re-evaluates
        its position before the next transition.
        ``backend``: "cuda" (kernel; default on CUDA tensors) or "torch" (masked torch ops).
        ``compact_tail``: ``logp_grad_into`` also accepts ``chains=idx`` (indices of the chains of a
        gathered batch); chains that finished a stage are then dropped from the evaluation.
        """
        self.f_into = logp_grad_into
        self.C, self.P = q0.shape
        self.dev, self.dt = q0.device, q0.dtype
        self.W, self.S = int(warmup), int(samples)
        self.maxd = int(max_treedepth)
        self.D1 = self.maxd + 1
        self.target, self.gamma, self.kappa, self.t0 = da_target_accept, da_gamma, da_kappa, da_t0
        self.thr = float(divergence_threshold)
        self.backend = backend or ("cuda" if q0.is_cuda else "torch")
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed)
        self.aux = aux
        self.gibbs = gibbs or []
        ep = stan_epochs(self.W, **(epoch_kw or {})) if self.W > 0 else []
        self.ep_dur = torch.tensor([d for _, d in ep] + [0], dtype=torch.int32, device=self.dev)
        self.ep_kind = torch.tensor([1 if k == "slow" else 0 for k, _ in ep] + [0], dtype=torch.int32, device=self.dev)
        self.nep = len(ep)
        self.steps = 0
        self.chain_evals = 0
        self._active_idx = None
        self.compact_tail = bool(compact_tail)
        self.compact_below = 0.85
        self._active_S = 0  # posterior draws of the running stage (0 during the warm-up stage)
        self._init_state(q0, initial_step_size, seed)

    # ------------------------------------------------------------------------------------------
    def _init_state(self, q0, initial_step_size, seed):
        C, P, dev, dt = self.C, self.P, self.dev, self.dt
        z2 = lambda: torch.zeros((C, P), dtype=dt, device=dev)
        z1 = lambda: torch.zeros(C, dtype=dt, device=dev)
        zi = lambda: torch.zeros(C, dtype=torch.int32, device=dev)
        s = {}
        for k in ("q", "g", "inv_mass", "qL", "rL", "gL", "qR", "rR", "gR", "qp", "gp", "rsum", "qe", "re", "ge",
                  "qs", "gs", "rsum_s", "qn", "r_half", "g_n"):
            s[k] = z2()
        for k in ("lp", "step", "h0", "lpp", "log_w", "sum_acc", "n_leaf", "lps", "logw_s", "lp_n",
                  "da_err", "da_logavg", "da_mu"):
            s[k] = z1()
        for k in ("phase", "depth", "dir", "i_leaf", "n_leaves", "it", "ep", "t_ep", "flags"):
            s[k] = zi()
        s["ck_r"] = torch.zeros((C, self.D1, P), dtype=dt, device=dev)
        s["ck_s"] = torch.zeros_like(s["ck_r"])
        s["ws1"] = torch.zeros((C, P), dtype=torch.float64, device=dev)
        s["ws2"] = torch.zeros_like(s["ws1"])
        s["draws"] = torch.zeros((max(self.S, 1), C, P), dtype=dt, device=dev)
        s["stats"] = torch.zeros((C, 4), dtype=torch.float64, device=dev)  # posterior: sum acc, sum depth, divergences, leaves
        s["n_done"] = torch.zeros(1, dtype=torch.int32, device=dev)
        s["q"].copy_(q0)
        s["inv_mass"].fill_(1.0)
        self.st = s
        # log-posterior at the start and the initial step size: the lock-step driver's search
        def f(q):
            lp = torch.empty(C, dtype=dt, device=dev)
            g = torch.empty((C, P), dtype=dt, device=dev)
            self.f_into(q.contiguous(), lp, g)
            return lp, g


        init = BatchedNUTS(f, q0, seed=seed, initial_step_size=initial_step_size, fused=False,
                           da_target_accept=self.target)
        s["step"].copy_(init.step_size)
        self.init_evals = init.evals
        s["da_logavg"].copy_(torch.log(s["step"]))
        s["da_mu"].copy_(torch.log(10.0 * s["step"]))
        s["phase"].fill_(REFRESH)
        s["qn"].copy_(s["q"])
        nb = len(self.gibbs)
        if nb:
            self.gib_M = torch.stack([g["M"].to(device=dev, dtype=dt) for g in self.gibbs]).contiguous()
            self.gib_par = torch.tensor([[g["a_post"], g["b0"], float(g["slot"])] for g in self.gibbs],
                                        dtype=torch.float64, device=dev)
        else:
            self.gib_M = torch.zeros((1, 1, 1), dtype=dt, device=dev)
            self.gib_par = torch.zeros((1, 3), dtype=torch.float64, device=dev)
        self._Z = torch.empty((C, P), dtype=dt, device=dev)
        self._U = torch.empty((C, 3), dtype=dt, device=dev)
        self._G = torch.ones((C, max(nb, 1)), dtype=torch.float64, device=dev)
        self._ga = (torch.tensor([g["a_post"] for g in self.gibbs], dtype=torch.float64, device=dev)[None, :]
                    .expand(C, nb).contiguous() if nb else None)
        self._ptrs = None

    # ------------------------------------------------------------------------------------------
    def _randoms(self):
        self._Z.normal_(generator=self.gen)
        self._U.uniform_(generator=self.gen)
        if self.gibbs:
            self._G.copy_(torch._standard_gamma(self._ga, generator=self.gen))

    def _evaluate(self):
        s = self.st
        idx = self._active_idx
        if idx is None:
            self.f_into(s["qn"], s["lp_n"], s["g_n"])
        else:  # tail of a stage: only the chains that are still running are evaluated
            n = idx.numel()
            qc, lpc, gc = self._cq[:n], self._clp[:n], self._cg[:n]
            torch.index_select(s["qn"], 0, idx, out=qc)
            self.f_into(qc, lpc, gc, chains=idx)
            s["lp_n"].index_copy_(0, idx, lpc)
            s["g_n"].index_copy_(0, idx, gc)
        self.chain_evals += self.C if idx is None else int(idx.numel())

    def _maybe_compact(self, n_done: int):
        """Chains finish a stage at different steps (per-chain step sizes give per-chain tree sizes): once
        a share of them is done, the evaluation runs on the gathered batch of the others (the evaluation
        kernel's time is proportional to the chains it is given). Needs a callback that accepts ``chains=``."""
        if not self.compact_tail:
            return
        active = self.C - n_done
        cur = self.C if self._active_idx is None else int(self._active_idx.numel())
        if active <= self.compact_below * cur and active > 0:
            self._active_idx = (self.st["phase"] != DONE).nonzero().squeeze(1)
            if not hasattr(self, "_cq"):
                self._cq = torch.empty_like(self.st["qn"])
                self._clp = torch.empty_like(self.st["lp_n"])
                self._cg = torch.empty_like(self.st["g_n"])

    @torch.no_grad()
    def run(self, check_every: int = 64):
        """Runs every chain through the warm-up, waits for the slowest chain, then runs every chain
        through the posterior epoch (two stages, so that the two are timed separately as in the lock-step
        driver). Returns ``dict(steps_warmup, steps_posterior, seconds_warmup, seconds_posterior)`` —
        a step is one batched evaluation."""
        import time

        sync = torch.cuda.synchronize if self.dev.type == "cuda" else (lambda: None)
        out = {}
        for stage, target in (("warmup", 0), ("posterior", self.S)):
            if stage == "warmup" and self.W == 0:
                out["steps_warmup"], out["seconds_warmup"] = 0, 0.0
                continue
            self._active_S = target
            self._ptrs = None  # the stage's sample count is a kernel argument
            s = self.st
            s["phase"].fill_(REFRESH)
            s["n_done"].zero_()
            s["qn"].copy_(s["q"])
            self._active_idx = None
            sync()
            t0 = time.perf_counter()
            n0 = self.steps
            self._evaluate()
            self.steps += 1
            while stage == "warmup" or self.S > 0:
                for _ in range(check_every):
                    self._randoms()
                    self._step()
                    self._evaluate()
                    self.steps += 1
                nd = int(s["n_done"].item())
                if nd >= self.C:
                    break
                self._maybe_compact(nd)
                # a chain needs at most 2^max_treedepth evaluations (+ the refresh) per transition
                if self.steps - n0 > (target if stage == "posterior" else self.W) * (2 ** self.maxd + 2) + check_every:
                    raise RuntimeError("asynchronous NUTS: a chain did not finish its stage (state machine stuck)")
            sync()
            out["steps_" + stage] = self.steps - n0
            out["seconds_" + stage] = time.perf_counter() - t0
        return out

    def _step(self):
        if self.backend == "cuda":
            self._step_cuda()
        else:
            self._step_torch()

    # -- results ---------------------------------------------------------------------------------
    def draws(self) -> torch.Tensor:
        return self.st["draws"][: self.S]

    def summary(self) -> dict:
        st = self.st["stats"].double()
        n = max(self.S, 1)
        return {
            "mean_accept": float((st[:, 0] / n).mean()), "mean_treedepth": float((st[:, 1] / n).mean()),
            "divergent_frac": float((st[:, 2] / n).mean()), "mean_leapfrogs": float((st[:, 3] / n).mean()),
            "steps": int(self.steps), "step_size_median": float(self.st["step"].median()),
        }

    # ------------------------------------------------------------------------------------------
    def _step_cuda(self):
        import ctypes as Ct

        from . import _lib

        s = self.st
        if self._ptrs is None:
            order = ["q", "lp", "g", "step", "inv_mass", "h0", "qL", "rL", "gL", "qR", "rR", "gR", "qp", "lpp", "gp",
                     "log_w", "rsum", "sum_acc", "n_leaf", "qe", "re", "ge", "qs", "lps", "gs", "logw_s", "rsum_s",
                     "ck_r", "ck_s", "qn", "r_half", "lp_n", "g_n"]
            tens = [s[k] for k in order] + [self._Z, self._U, self._G]
            tens += [s["da_err"], s["da_logavg"], s["da_mu"], s["ws1"], s["ws2"], s["draws"]]
            tens += [self.aux if self.aux is not None else self._U, self.gib_M, self.gib_par]
            tens += [s[k] for k in ("phase", "depth", "dir", "i_leaf", "n_leaves", "it", "ep", "t_ep", "flags")]
            tens += [self.ep_dur, self.ep_kind, s["stats"], s["n_done"]]
            self._keep = tens
            self._ptrs = (Ct.c_void_p * len(tens))(*[t.data_ptr() for t in tens])
            A = 0 if self.aux is None else int(self.aux.shape[1])
            self._ints = (Ct.c_int * 9)(self.C, self.P, self.D1, self.maxd, self.W, self._active_S, self.nep,
                                        len(self.gibbs), A)
            self._dbls = (Ct.c_double * 5)(self.thr, self.target, self.gamma, self.kappa, float(self.t0))
            self._fn = getattr(_lib.lib, "lslb_nuts_async_step_" + _lib.sfx(self.dt))
        _lib.check(self._fn(self._ptrs, self._ints, self._dbls, _lib.stream_ptr()), "lslb_nuts_async_step")

    # ------------------------------------------------------------------------------------------
    def _step_torch(self):
        """The state machine with masked torch ops (same program as ``nuts_async_step_kernel``)."""
        s, C, P, dt = self.st, self.C, self.P, self.dt
        W = torch.where
        Z, U = self._Z, self._U
        u_dir, u_leaf, u_merge = U[:, 0], U[:, 1], U[:, 2]
        lp_n, g_n = s["lp_n"], s["g_n"]
        ph = s["phase"]
        is_ref, is_leaf = ph == REFRESH, ph == LEAF
        col = lambda m: m[:, None]
        ninf = torch.full((C,), -float("inf"), dtype=dt, device=self.dev)
        im = s["inv_mass"]

        def turning(r_left, r_right, r_sum):
            rho = r_sum - 0.5 * (r_right + r_left)
            return (((im * r_left) * rho).sum(dim=1) <= 0) | (((im * r_right) * rho).sum(dim=1) <= 0)

        # ---- REFRESH result ----------------------------------------------------------------------
        s["lp"] = W(is_ref, lp_n, s["lp"])
        s["g"] = W(col(is_ref), g_n, s["g"])
        start_tr = is_ref.clone()

        # ---- LEAF: bookkeeping of the leaf that was just evaluated -------------------------------
        act = is_leaf
        eps = s["dir"].to(dt) * s["step"]
        r_n = s["r_half"] + 0.5 * col(eps) * g_n
        kin = 0.5 * (im * r_n * r_n).sum(dim=1)
        dh = torch.nan_to_num((-lp_n + kin) - s["h0"], nan=float("inf"))
        div_n = dh > self.thr
        logw_n = -dh
        acc_n = torch.exp(torch.clamp(-dh, max=0.0))
        logw_new = torch.logaddexp(s["logw_s"], logw_n)
        take = act & (torch.log(u_leaf) < logw_n - logw_new)
        s["qs"] = W(col(take), s["qn"], s["qs"])
        s["gs"] = W(col(take), g_n, s["gs"])
        s["lps"] = W(take, lp_n, s["lps"])
        s["logw_s"] = W(act, logw_new, s["logw_s"])
        s["rsum_s"] = W(col(act), s["rsum_s"] + r_n, s["rsum_s"])
        s["sum_acc"] = s["sum_acc"] + W(act, acc_n, torch.zeros_like(acc_n))
        s["n_leaf"] = s["n_leaf"] + act.to(dt)
        s["qe"] = W(col(act), s["qn"], s["qe"])
        s["re"] = W(col(act), r_n, s["re"])
        s["ge"] = W(col(act), g_n, s["ge"])
        fl = s["flags"]
        div_s = ((fl & F_DIV_S) != 0) | (act & div_n)
        turn_s = (fl & F_TURN_S) != 0
        i = s["i_leaf"]
        idx_max = _popcount(i >> 1)
        hi = idx_max
        lo = idx_max - _trailing_ones(i) + 1
        even = (i & 1) == 0
        ks = torch.arange(self.D1, device=self.dev)[None, :]
        store = (act & even)[:, None] & (ks == hi[:, None])
        s["ck_r"] = W(store[:, :, None], r_n[:, None, :], s["ck_r"])
        s["ck_s"] = W(store[:, :, None], s["rsum_s"][:, None, :], s["ck_s"])
        chk = (act & ~even)[:, None] & (ks >= lo[:, None]) & (ks <= hi[:, None])  # [C, D1]
        sub = s["rsum_s"][:, None, :] - s["ck_s"] + s["ck_r"]
        rho = sub - 0.5 * (r_n[:, None, :] + s["ck_r"])
        tl = ((im[:, None, :] * s["ck_r"]) * rho).sum(dim=2) <= 0
        tr = ((im[:, None, :] * r_n[:, None, :]) * rho).sum(dim=2) <= 0
        turn_s = turn_s | (act & (chk & (tl | tr)).any(dim=1))
        s["i_leaf"] = i + act.to(torch.int32)
        sub_done = act & (div_s | turn_s | (s["i_leaf"] == s["n_leaves"]))

        # ---- subtree finished: merge it into the tree ---------------------------------------------
        ok = sub_done & ~turn_s & ~div_s
        take2 = ok & (torch.log(u_merge) < s["logw_s"] - s["log_w"])
        s["qp"] = W(col(take2), s["qs"], s["qp"])
        s["gp"] = W(col(take2), s["gs"], s["gp"])
        s["lpp"] = W(take2, s["lps"], s["lpp"])
        s["log_w"] = W(ok, torch.logaddexp(s["log_w"], s["logw_s"]), s["log_w"])
        right = s["dir"] > 0
        okr, okl = col(ok & right), col(ok & ~right)
        for a, b in (("qR", "qe"), ("rR", "re"), ("gR", "ge")):
            s[a] = W(okr, s[b], s[a])
        for a, b in (("qL", "qe"), ("rL", "re"), ("gL", "ge")):
            s[a] = W(okl, s[b], s[a])
        s["rsum"] = W(col(ok), s["rsum"] + s["rsum_s"], s["rsum"])
        turn_full = turning(s["rL"], s["rR"], s["rsum"])
        diverged = ((fl & F_DIVERGED) != 0) | (sub_done & div_s)
        turned = ((fl & F_TURNED) != 0) | (sub_done & turn_s) | (ok & turn_full)
        s["depth"] = s["depth"] + sub_done.to(torch.int32)
        alive = ok & ~turn_full
        new_sub = alive & (s["depth"] < self.maxd)
        end_tr = sub_done & ~new_sub

        # ---- transition finished ---------------------------------------------------------------------
        s["q"] = W(col(end_tr), s["qp"], s["q"])
        s["g"] = W(col(end_tr), s["gp"], s["g"])
        s["lp"] = W(end_tr, s["lpp"], s["lp"])
        acc = s["sum_acc"] / torch.clamp(s["n_leaf"], min=1.0)
        it = s["it"]
        wm = end_tr & (it < self.W)
        if bool(wm.any()):
            t = s["t_ep"].to(dt)
            tt = t + 1.0
            eta = tt ** (-self.kappa)
            err = s["da_err"] + (self.target - acc)
            log_step = s["da_mu"] - (err * torch.sqrt(tt)) / (self.gamma * (self.t0 + tt))
            s["da_err"] = W(wm, err, s["da_err"])
            s["step"] = W(wm, torch.exp(log_step), s["step"])
            s["da_logavg"] = W(wm, (1 - eta) * s["da_logavg"] + eta * log_step, s["da_logavg"])
            qd = s["q"].double()
            s["ws1"] = W(col(wm), s["ws1"] + qd, s["ws1"])
            s["ws2"] = W(col(wm), s["ws2"] + qd * qd, s["ws2"])
            s["t_ep"] = s["t_ep"] + wm.to(torch.int32)
            dur = self.ep_dur[s["ep"].long()]
            kind = self.ep_kind[s["ep"].long()]
            ee = wm & (s["t_ep"] == dur)
            if bool(ee.any()):
                s["step"] = W(ee, torch.exp(s["da_logavg"]), s["step"])
                slow = ee & (kind == 1) & (dur > 1)
                d = dur.double().clamp(min=2)[:, None]
                var = (s["ws2"] - s["ws1"] * s["ws1"] / d) / (d - 1)
                new_inv = (var + 0.001).to(dt)
                adj = torch.sqrt(s["inv_mass"].sum(dim=1) / new_inv.sum(dim=1))
                s["step"] = W(slow, adj * s["step"], s["step"])
                s["inv_mass"] = W(col(slow), new_inv, s["inv_mass"])
                s["ws1"] = W(col(ee), torch.zeros_like(s["ws1"]), s["ws1"])
                s["ws2"] = W(col(ee), torch.zeros_like(s["ws2"]), s["ws2"])
                s["ep"] = s["ep"] + ee.to(torch.int32)
                s["t_ep"] = W(ee, torch.zeros_like(s["t_ep"]), s["t_ep"])
                s["da_err"] = W(ee, torch.zeros_like(s["da_err"]), s["da_err"])
                s["da_logavg"] = W(ee, torch.log(s["step"]), s["da_logavg"])
                s["da_mu"] = W(ee, torch.log(10.0 * s["step"]), s["da_mu"])
        pm = end_tr & (it >= self.W)
        if bool(pm.any()) and self.S > 0:
            row = (it - self.W).clamp(min=0, max=self.S - 1).long()
            cidx = pm.nonzero().squeeze(1)
            s["draws"][row[cidx], cidx] = s["q"][cidx]
            st = s["stats"]
            st[:, 0] += W(pm, acc, torch.zeros_like(acc)).double()
            st[:, 1] += W(pm, s["depth"].double(), torch.zeros(C, dtype=torch.float64, device=self.dev))
            st[:, 2] += (pm & diverged).double()
            st[:, 3] += W(pm, s["n_leaf"], torch.zeros_like(acc)).double()
        s["it"] = it + end_tr.to(torch.int32)
        done = end_tr & (s["it"] == self.W + self._active_S)
        s["n_done"] += int(done.sum())
        cont = end_tr & ~done
        if self.gibbs:
            for b, gb in enumerate(self.gibbs):
                quad = ((s["q"] @ self.gib_M[b]) * s["q"]).sum(dim=1).double()
                tau2 = (gb["b0"] + 0.5 * quad) / self._G[:, b]
                sl = int(gb["slot"])
                self.aux[:, sl] = W(end_tr, tau2.to(self.aux.dtype), self.aux[:, sl])
            to_refresh = cont
        else:
            to_refresh = torch.zeros_like(cont)
            start_tr = start_tr | cont

        # ---- a transition starts -----------------------------------------------------------------------
        r0 = Z / torch.sqrt(s["inv_mass"])
        h0 = -s["lp"] + 0.5 * (s["inv_mass"] * r0 * r0).sum(dim=1)
        cs = col(start_tr)
        s["h0"] = W(start_tr, h0, s["h0"])
        for a, b in (("qL", "q"), ("qR", "q"), ("gL", "g"), ("gR", "g"), ("qp", "q"), ("gp", "g")):
            s[a] = W(cs, s[b], s[a])
        for a in ("rL", "rR", "rsum"):
            s[a] = W(cs, r0, s[a])
        s["lpp"] = W(start_tr, s["lp"], s["lpp"])
        zero = torch.zeros(C, dtype=dt, device=self.dev)
        s["log_w"] = W(start_tr, zero, s["log_w"])
        s["sum_acc"] = W(start_tr, zero, s["sum_acc"])
        s["n_leaf"] = W(start_tr, zero, s["n_leaf"])
        s["depth"] = W(start_tr, torch.zeros_like(s["depth"]), s["depth"])
        diverged = diverged & ~start_tr
        turned = turned & ~start_tr
        new_sub = new_sub | start_tr

        # ---- a subtree starts ----------------------------------------------------------------------------
        ndir = W(u_dir < 0.5, torch.ones_like(s["dir"]), -torch.ones_like(s["dir"]))
        s["dir"] = W(new_sub, ndir, s["dir"])
        rt = col(s["dir"] > 0)
        cn = col(new_sub)
        s["qe"] = W(cn, W(rt, s["qR"], s["qL"]), s["qe"])
        s["re"] = W(cn, W(rt, s["rR"], s["rL"]), s["re"])
        s["ge"] = W(cn, W(rt, s["gR"], s["gL"]), s["ge"])
        s["qs"] = W(cn, s["qe"], s["qs"])
        s["gs"] = W(cn, s["ge"], s["gs"])
        s["lps"] = W(new_sub, s["lpp"], s["lps"])
        s["logw_s"] = W(new_sub, ninf, s["logw_s"])
        s["rsum_s"] = W(cn, torch.zeros_like(s["rsum_s"]), s["rsum_s"])
        turn_s = turn_s & ~new_sub
        div_s = div_s & ~new_sub
        s["i_leaf"] = W(new_sub, torch.zeros_like(s["i_leaf"]), s["i_leaf"])
        s["n_leaves"] = W(new_sub, (torch.ones_like(s["depth"]) << s["depth"]), s["n_leaves"])
        s["flags"] = (diverged.to(torch.int32) * F_DIVERGED + turned.to(torch.int32) * F_TURNED +
                      turn_s.to(torch.int32) * F_TURN_S + div_s.to(torch.int32) * F_DIV_S)
        phase = W(new_sub, torch.full_like(ph, LEAF), ph)
        phase = W(to_refresh, torch.full_like(ph, REFRESH), phase)
        phase = W(done, torch.full_like(ph, DONE), phase)
        s["phase"] = phase

        # ---- where the next evaluation happens ---------------------------------------------------------------
        leaf_now = phase == LEAF
        eps = s["dir"].to(dt) * s["step"]
        r_half = s["re"] + 0.5 * col(eps) * s["ge"]
        s["r_half"] = W(col(leaf_now), r_half, s["r_half"])
        # in-place: the evaluation callback and the kernel path hold on to these buffers
        s["qn"].copy_(W(col(leaf_now), s["qe"] + col(eps) * (s["inv_mass"] * r_half), s["q"]))
