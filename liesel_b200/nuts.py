"""
Batched NUTS driver: all chains advance in lock-step around ONE batched log-posterior+gradient
call per leapfrog (SURVEY §8 f1 — the caller that issues up to 2^max_treedepth - 1 evaluations per
transition, reference ``goose/nuts.py:240-264`` → BlackJAX ``nuts(...).step``).

Semantics follow the reference's ``NUTSKernel`` (``src/liesel/goose/nuts.py``):
  * defaults ``max_treedepth=10``, ``da_target_accept=0.8``, ``da_gamma=0.05``, ``da_kappa=0.75``,
    ``da_t0=10``, diagonal mass matrix (:165-176); divergence threshold 1000 (:64-66);
  * error code = divergent + 2 * (treedepth == max_treedepth) (:79-96);
  * dual averaging exactly as ``da_step`` / ``da_init`` / ``da_finalize`` (``goose/da.py:20-127``);
  * inverse mass vector = per-chain sample variance of the epoch's history + 1e-3 and the step-size
    rescaling sqrt(trace(old)/trace(new)) (``goose/mm.py:19-35``, ``goose/nuts.py:306-334``);
  * Stan-style warm-up epochs as ``stan_epochs`` (``goose/warmup.py:12-88``).
The trajectory itself (iterative tree doubling, multinomial sampling inside a subtree, biased
progressive sampling between subtrees, generalised U-turn criterion on checkpointed momenta) is
BlackJAX 1.5's algorithm; BlackJAX is third-party and not under /root/reference, so this part is
restated from its published algorithm and pinned only statistically (tests: posterior recovery and
acceptance rate as in ``tests/goose/kernels/model_lm.py:70-84``). Deviations, all deliberate:
``find_reasonable_step_size`` uses one leapfrog per trial (Hoffman & Gelman, Alg. 4) instead of a
full NUTS step per trial, and random numbers come from torch, not JAX's threefry.

Like the reference under ``jax.vmap`` (``goose/engine.py:442-448``), every chain steps until the
slowest chain of the batch has finished its tree; finished chains are masked.

The module only needs ``logp_grad_fn(q [C, P]) -> (logp [C], grad [C, P])``; it runs on whatever
device those tensors live on (the B200 plan in production, torch-CPU for the CPU baseline).
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable

import torch


def stan_epochs(warmup_duration=1000, init_duration=75, term_duration=50, base_duration=25):
    """
    Stan-style warm-up schedule as ``[(kind, duration)]`` with kind "fast" (step size only) or "slow"
    (step size + mass matrix), the schedule of ``goose/warmup.py:12-88``: an initial fast window, slow
    windows that double in length for as long as the NEXT doubled window would still fit, one last slow
    window that absorbs the remainder, and a closing fast window.
    """
    if warmup_duration < 20:
        raise ValueError("warmup_duration too short (< 20)")
    if warmup_duration < init_duration + term_duration + base_duration:
        raise ValueError("warmup_duration too short (< init_duration + term_duration + base_duration)")
    slow_budget = warmup_duration - init_duration - term_duration
    slow, window = [], base_duration
    # a window of length w is followed by one of 2w: emit w only while w + 2w still fits
    while slow_budget - sum(slow) >= 3 * window:
        slow.append(window)
        window *= 2
    slow.append(slow_budget - sum(slow))
    return [("fast", init_duration)] + [("slow", d) for d in slow] + [("fast", term_duration)]


def _leaf_idx_to_ckpt_idxs(n: int):
    """Checkpoint range touched by leaf ``n`` of a subtree (iterative NUTS)."""
    idx_max = bin(n >> 1).count("1")
    num_subtrees = 0
    m = n
    while m & 1:
        num_subtrees += 1
        m >>= 1
    return idx_max - num_subtrees + 1, idx_max


@dataclass
class NUTSInfo:
    error_code: torch.Tensor
    acceptance_prob: torch.Tensor
    divergent: torch.Tensor
    turning: torch.Tensor
    treedepth: torch.Tensor
    leapfrog: torch.Tensor
    evaluations: int  # batched logp+grad calls issued by this transition


class BatchedNUTS:
    def __init__(self, logp_grad_fn: Callable, q0: torch.Tensor, max_treedepth: int = 10,
                 da_target_accept: float = 0.8, da_gamma: float = 0.05, da_kappa: float = 0.75,
                 da_t0: int = 10, divergence_threshold: float = 1000.0, seed: int = 0,
                 initial_step_size: float | None = None, fused: bool | None = None,
                 logp_grad_into: Callable | None = None, graphs: bool = False, compact: bool = False,
                 compact_below: float = 0.8):
        """
        ``fused`` (default: on for CUDA tensors): run the per-leaf bookkeeping in the two CUDA kernels
        ``lslb_nuts_leaf_pre/post`` instead of ~30 torch ops. ``logp_grad_into(q, logp_out, grad_out)``
        optionally evaluates into preallocated buffers (``Plan.logp_grad(..., out=...)``).
        ``graphs`` (needs ``fused`` and ``logp_grad_into``): the leaves of a subtree are replayed from
        CUDA graphs, one graph per leaf range [2^(k-1), 2^k) (the ranges between two host-side "is any
        chain still active" checks), captured on first use; every leaf is five kernel launches, so
        at small n the launch latency, not the kernels, bounds a transition.
        ``compact`` (needs ``fused`` and a ``logp_grad_into(q, logp_out, grad_out, chains=idx)`` that
        accepts the indices of the chains it is given): chains whose tree has already terminated are
        dropped from the batch — at every doubling, and inside a subtree at the host checks between
        leaf ranges, the still-running chains are gathered into a dense batch (padded to a multiple of
        64) whenever that batch is at most ``compact_below`` of the current one. The reference cannot do
        this under ``jax.vmap`` (every chain pays for the deepest tree of the batch); here the
        evaluation kernel's time is proportional to the chains it is given, so the saving is real.
        The draws are those of the uncompacted driver up to float reduction order.
        """
        self.f = logp_grad_fn
        self.f_into = logp_grad_into
        self.fused = bool(q0.is_cuda) if fused is None else bool(fused)
        self.use_graphs = bool(graphs) and self.fused and logp_grad_into is not None
        self._graphs: dict = {}
        self._graph_ready = False
        self.compact = bool(compact) and self.fused and logp_grad_into is not None and not self.use_graphs
        self.compact_below = float(compact_below)
        self.chain_evals = 0  # chain-evaluations actually computed inside subtrees (compaction metric)
        self.q = q0.clone()
        self.C, self.P = q0.shape
        self.dev, self.dt = q0.device, q0.dtype
        self.max_treedepth = max_treedepth
        self.target, self.gamma, self.kappa, self.t0 = da_target_accept, da_gamma, da_kappa, da_t0
        self.div_thr = divergence_threshold
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed)
        self.inv_mass = torch.ones_like(self.q)  # identity inverse mass vector (nuts.py:206-213)
        self.logp, self.grad = self.f(self.q)
        self.logp, self.grad = self.logp.clone(), self.grad.clone()
        self.evals = 1
        if initial_step_size is None:
            self.step_size = self._find_reasonable_step_size(0.001)
        else:
            self.step_size = torch.full((self.C,), float(initial_step_size), dtype=self.dt, device=self.dev)
        self._da_init()

    # -- random numbers ---------------------------------------------------------------------
    def _randn(self, *shape):
        return torch.randn(*shape, generator=self.gen, device=self.dev, dtype=self.dt)

    def _rand(self, *shape):
        return torch.rand(*shape, generator=self.gen, device=self.dev, dtype=self.dt)

    # -- dynamics ---------------------------------------------------------------------------
    def _kinetic(self, r):
        return 0.5 * (self.inv_mass * r * r).sum(dim=1)

    def _leapfrog(self, q, r, g, eps):
        e = eps[:, None]
        r_half = r + 0.5 * e * g
        q_new = q + e * (self.inv_mass * r_half)
        logp, grad = self.f(q_new)
        self.evals += 1
        r_new = r_half + 0.5 * e * grad
        return q_new, r_new, logp.clone(), grad.clone()

    def _is_turning(self, r_left, r_right, r_sum):
        rho = r_sum - 0.5 * (r_right + r_left)
        tl = ((self.inv_mass * r_left) * rho).sum(dim=1) <= 0
        tr = ((self.inv_mass * r_right) * rho).sum(dim=1) <= 0
        return tl | tr

    def _find_reasonable_step_size(self, eps0: float):
        """One-leapfrog doubling/halving search per chain for an acceptance near the target."""
        eps = torch.full((self.C,), eps0, dtype=self.dt, device=self.dev)
        r = self._randn(self.C, self.P) / self.inv_mass.sqrt()
        h0 = -self.logp + self._kinetic(r)
        _, r1, lp1, _ = self._leapfrog(self.q, r, self.grad, eps)
        dh = (-lp1 + self._kinetic(r1)) - h0
        acc = torch.exp(torch.clamp(-torch.nan_to_num(dh, nan=float("inf")), max=0.0))
        direction = torch.where(acc > self.target, 1.0, -1.0).to(self.dt)
        active = torch.ones(self.C, dtype=torch.bool, device=self.dev)
        for _ in range(60):
            if not bool(active.any()):
                break
            eps = torch.where(active, eps * torch.pow(torch.tensor(2.0, dtype=self.dt, device=self.dev), direction), eps)
            r = self._randn(self.C, self.P) / self.inv_mass.sqrt()
            h0 = -self.logp + self._kinetic(r)
            _, r1, lp1, _ = self._leapfrog(self.q, r, self.grad, eps)
            dh = (-lp1 + self._kinetic(r1)) - h0
            acc = torch.exp(torch.clamp(-torch.nan_to_num(dh, nan=float("inf")), max=0.0))
            crossed = torch.where(direction > 0, acc <= self.target, acc > self.target)
            active = active & ~crossed
        return eps

    # -- dual averaging (goose/da.py) ---------------------------------------------------------
    def _da_init(self):
        self.da_error_sum = torch.zeros(self.C, dtype=self.dt, device=self.dev)
        self.da_log_avg = torch.log(self.step_size)
        self.da_mu = torch.log(10.0 * self.step_size)

    def _da_step(self, acceptance_prob, time_in_epoch: int):
        t = time_in_epoch + 1
        eta = t ** (-self.kappa)
        self.da_error_sum = self.da_error_sum + (self.target - acceptance_prob)
        log_step = self.da_mu - (self.da_error_sum * math.sqrt(t)) / (self.gamma * (self.t0 + t))
        self.step_size = torch.exp(log_step)
        self.da_log_avg = (1 - eta) * self.da_log_avg + eta * log_step

    def _da_finalize(self):
        self.step_size = torch.exp(self.da_log_avg)

    # -- the two per-leaf kernels (hooks: the CPU tests substitute torch restatements to exercise the
    #    subtree bookkeeping, compaction included, without a GPU) ------------------------------------
    def _leaf_pre(self, fn, st, ptrs, ints, stream, leaf):
        from . import _lib

        _lib.check(fn(ptrs, ints, stream), "lslb_nuts_leaf_pre")

    def _leaf_post(self, fn, st, ptrs, ints, stream, leaf):
        from . import _lib

        _lib.check(fn(ptrs, ints, self.div_thr, stream), "lslb_nuts_leaf_post")

    # -- fused subtree (CUDA kernels for the per-leaf bookkeeping) ----------------------------------
    def _subtree_fused(self, j, q_e, r_e, g_e, eps, h0, alive, lp_prop, U, ck_r, ck_s):
        import ctypes as Ct

        from . import _lib

        C, P, dev, dt = self.C, self.P, self.dev, self.dt
        st = self._fused_state if hasattr(self, "_fused_state") else None
        if st is None:
            z2 = lambda: torch.empty((C, P), dtype=dt, device=dev)
            z1 = lambda: torch.empty(C, dtype=dt, device=dev)
            st = {k: z2() for k in ("q_e", "r_e", "g_e", "q_n", "r_half", "g_n", "q_s", "g_s", "rsum_s", "inv_mass")}
            st.update({k: z1() for k in ("lp_n", "lp_s", "logw_s", "sum_acc", "n_leaf", "eps", "h0")})
            st.update({k: torch.empty(C, dtype=torch.uint8, device=dev) for k in ("act", "turn_s", "div_s")})
            # persistent copies of what the caller allocates per transition / per subtree, so that
            # captured graphs keep pointing at live memory
            st["U"] = torch.empty((2 ** max(self.max_treedepth - 1, 0), C), dtype=dt, device=dev)
            st["ck_r"] = torch.empty((C, self.max_treedepth + 1, P), dtype=dt, device=dev)
            st["ck_s"] = torch.empty_like(st["ck_r"])
            self._fused_state = st
        n_leaves = 2 ** j
        # -- which chains enter the batch ----------------------------------------------------------
        idx, n_real, Cb = None, C, C  # idx: batch row -> chain; rows >= n_real are inactive padding
        if self.compact:
            sel = alive.nonzero().squeeze(1)
            n_alive = int(sel.numel())
            padded = -(-max(n_alive, 1) // 64) * 64
            if padded <= self.compact_below * C:
                idx, n_real, Cb = sel, n_alive, padded
        out_keys2 = ("q_e", "r_e", "g_e", "q_s", "g_s", "rsum_s")
        out_keys1 = ("lp_s", "logw_s", "sum_acc", "n_leaf", "turn_s", "div_s")
        if idx is None:
            st["q_e"].copy_(q_e); st["r_e"].copy_(r_e); st["g_e"].copy_(g_e)
            st["q_s"].copy_(q_e); st["g_s"].copy_(g_e); st["lp_s"].copy_(lp_prop)
            st["eps"].copy_(eps); st["h0"].copy_(h0)
            st["act"].copy_(alive.to(torch.uint8))
            st["inv_mass"].copy_(self.inv_mass)
            st["U"][:n_leaves].copy_(U)
            st["ck_r"].copy_(ck_r); st["ck_s"].copy_(ck_s)
            full = None
        else:
            gidx = idx if n_real == Cb else torch.cat([idx, idx[:1].expand(Cb - n_real)])
            for k, src in (("q_e", q_e), ("r_e", r_e), ("g_e", g_e), ("q_s", q_e), ("g_s", g_e),
                           ("lp_s", lp_prop), ("eps", eps), ("h0", h0), ("inv_mass", self.inv_mass)):
                torch.index_select(src, 0, gidx, out=st[k][:Cb])
            st["act"][:n_real].fill_(1)
            st["act"][n_real:Cb].zero_()
            st["U"][:n_leaves, :Cb] = U.index_select(1, gidx)
            # results of chains that are not in the batch: untouched edge state, empty subtree
            full = {"q_e": q_e.clone(), "r_e": r_e.clone(), "g_e": g_e.clone(), "q_s": q_e.clone(),
                    "g_s": g_e.clone(), "rsum_s": torch.zeros_like(q_e), "lp_s": lp_prop.clone(),
                    "logw_s": torch.full((C,), -float("inf"), dtype=dt, device=dev),
                    "sum_acc": torch.zeros(C, dtype=dt, device=dev), "n_leaf": torch.zeros(C, dtype=dt, device=dev),
                    "turn_s": torch.zeros(C, dtype=torch.uint8, device=dev),
                    "div_s": torch.zeros(C, dtype=torch.uint8, device=dev)}
        st["logw_s"][:Cb].fill_(-float("inf")); st["rsum_s"][:Cb].zero_()
        st["sum_acc"][:Cb].zero_(); st["n_leaf"][:Cb].zero_()
        st["turn_s"][:Cb].zero_(); st["div_s"][:Cb].zero_()
        sfx = _lib.sfx(dt)
        pre = getattr(_lib.lib, "lslb_nuts_leaf_pre_" + sfx)
        post = getattr(_lib.lib, "lslb_nuts_leaf_post_" + sfx)
        order = ["q_e", "r_e", "g_e", "q_n", "r_half", "g_n", "lp_n"]
        tens = [st[k] for k in order] + [st["eps"], st["inv_mass"], st["h0"], st["q_s"], st["g_s"], st["lp_s"],
                                         st["logw_s"], st["rsum_s"], st["ck_r"], st["ck_s"], st["sum_acc"],
                                         st["n_leaf"], st["act"], st["turn_s"], st["div_s"], st["U"]]
        ptrs = (Ct.c_void_p * 23)(*[t.data_ptr() for t in tens])
        ints = (Ct.c_int * 6)(Cb, P, st["ck_r"].shape[1], 0, 0, 0)

        def run_leaves(i0, i1):
            stream = _lib.stream_ptr() if dev.type == "cuda" else None
            for i in range(i0, i1):
                lo, hi = _leaf_idx_to_ckpt_idxs(i)
                ints[3], ints[4], ints[5] = i & 1, lo, hi
                ptrs[22] = st["U"][i].data_ptr()
                self._leaf_pre(pre, st, ptrs, ints, stream, i)
                if self.f_into is not None:
                    if idx is None:
                        self.f_into(st["q_n"], st["lp_n"], st["g_n"])
                    else:
                        self.f_into(st["q_n"][:Cb], st["lp_n"][:Cb], st["g_n"][:Cb], chains=gidx)
                else:
                    lp, g = self.f(st["q_n"])
                    st["lp_n"].copy_(lp); st["g_n"].copy_(g)
                self._leaf_post(post, st, ptrs, ints, stream, i)

        def flush():
            """Results of the current batch rows -> full-size outputs (real rows only)."""
            for k in out_keys2 + out_keys1:
                full[k].index_copy_(0, idx, st[k][:n_real])

        # leaf ranges [0,1), [1,2), [2,4), [4,8), ...: between two ranges the host checks whether any
        # chain is still active (every check is a host sync)
        i0 = 0
        while i0 < n_leaves:
            i1 = 1 if i0 == 0 else min(2 * i0, n_leaves)
            if i0 > 0:
                if idx is None and not self.compact:
                    if not bool(st["act"].any()):
                        break
                else:
                    keep = st["act"][:n_real].nonzero().squeeze(1)
                    n_keep = int(keep.numel())
                    if n_keep == 0:
                        break
                    padded = -(-n_keep // 64) * 64
                    if self.compact and i1 - i0 >= 4 and padded <= self.compact_below * Cb:
                        # drop the chains whose subtree already terminated: their results are final
                        if idx is None:
                            idx = torch.arange(C, device=dev)
                            full = {k: st[k].clone() for k in out_keys2 + out_keys1}
                        else:
                            flush()
                        gk = keep if n_keep == padded else torch.cat([keep, keep[:1].expand(padded - n_keep)])
                        for k in ("q_e", "r_e", "g_e", "q_s", "g_s", "rsum_s", "inv_mass", "lp_s", "logw_s",
                                  "sum_acc", "n_leaf", "eps", "h0", "turn_s", "div_s", "ck_r", "ck_s"):
                            st[k][:padded] = st[k].index_select(0, gk)
                        st["U"][i0:n_leaves, :padded] = st["U"][i0:n_leaves].index_select(1, gk)
                        st["act"][:n_keep].fill_(1)
                        st["act"][n_keep:padded].zero_()
                        idx = idx.index_select(0, keep)
                        n_real, Cb = n_keep, padded
                        gidx = idx if n_real == Cb else torch.cat([idx, idx[:1].expand(Cb - n_real)])
                        ints[0] = Cb
            if self.use_graphs and self._graph_ready:
                g = self._graphs.get(i0)
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        run_leaves(i0, i1)
                    self._graphs[i0] = g  # (capturing does not execute: replay below)
                g.replay()
            else:
                run_leaves(i0, i1)
            self.evals += i1 - i0
            self.chain_evals += (i1 - i0) * Cb
            i0 = i1
        self._graph_ready = True  # the first subtree ran eagerly (lazy one-time initialisation done)
        if idx is None:
            ck_r.copy_(st["ck_r"]); ck_s.copy_(st["ck_s"])
            cl = lambda k: st[k].clone()
        else:
            flush()  # (checkpoints are scratch of ONE subtree: nothing to hand back)
            cl = lambda k: full[k]
        return (cl("q_e"), cl("r_e"), cl("g_e"), cl("q_s"), cl("lp_s"), cl("g_s"), cl("logw_s"), cl("rsum_s"),
                cl("turn_s").bool(), cl("div_s").bool(), cl("sum_acc"), cl("n_leaf"))

    # -- one NUTS transition for all chains ---------------------------------------------------
    @torch.no_grad()
    def transition(self) -> NUTSInfo:
        C, P, dev = self.C, self.P, self.dev
        evals0 = self.evals
        ninf = torch.full((C,), -float("inf"), dtype=self.dt, device=dev)
        r0 = self._randn(C, P) / self.inv_mass.sqrt()
        h0 = -self.logp + self._kinetic(r0)
        qL, rL, gL = self.q, r0, self.grad
        qR, rR, gR = self.q, r0, self.grad
        q_prop, lp_prop, g_prop = self.q, self.logp, self.grad
        log_w = torch.zeros(C, dtype=self.dt, device=dev)
        r_sum = r0.clone()
        alive = torch.ones(C, dtype=torch.bool, device=dev)
        depth = torch.zeros(C, dtype=torch.int32, device=dev)
        sum_acc = torch.zeros(C, dtype=self.dt, device=dev)
        n_leaf = torch.zeros(C, dtype=self.dt, device=dev)
        diverged = torch.zeros(C, dtype=torch.bool, device=dev)
        turned = torch.zeros(C, dtype=torch.bool, device=dev)
        ck_r = torch.zeros((C, self.max_treedepth + 1, P), dtype=self.dt, device=dev)
        ck_s = torch.zeros((C, self.max_treedepth + 1, P), dtype=self.dt, device=dev)

        for j in range(self.max_treedepth):
            if not bool(alive.any()):
                break
            right = self._rand(C) < 0.5
            sel = right[:, None]
            q_e = torch.where(sel, qR, qL)
            r_e = torch.where(sel, rR, rL)
            g_e = torch.where(sel, gR, gL)
            eps = torch.where(right, self.step_size, -self.step_size)
            q_s, lp_s, g_s = q_e, lp_prop, g_e
            logw_s = ninf.clone()
            rsum_s = torch.zeros_like(r0)
            turn_s = torch.zeros(C, dtype=torch.bool, device=dev)
            div_s = torch.zeros(C, dtype=torch.bool, device=dev)
            act = alive.clone()
            U = self._rand(2 ** j, C)
            if self.fused:
                q_e, r_e, g_e, q_s, lp_s, g_s, logw_s, rsum_s, turn_s, div_s, sa, nl = self._subtree_fused(
                    j, q_e, r_e, g_e, eps, h0, alive, lp_prop, U, ck_r, ck_s)
                sum_acc = sum_acc + sa
                n_leaf = n_leaf + nl
            for i in range(0 if self.fused else 2 ** j):
                if not bool(act.any()):
                    break
                q_n, r_n, lp_n, g_n = self._leapfrog(q_e, r_e, g_e, eps)
                dh = (-lp_n + self._kinetic(r_n)) - h0
                dh = torch.nan_to_num(dh, nan=float("inf"))
                div_n = dh > self.div_thr
                logw_n = -dh
                acc_n = torch.exp(torch.clamp(-dh, max=0.0))
                m = act[:, None]
                # multinomial (progressive uniform) sampling inside the subtree
                logw_new = torch.logaddexp(logw_s, logw_n)
                take = act & (torch.log(U[i]) < logw_n - logw_new)
                tk = take[:, None]
                q_s = torch.where(tk, q_n, q_s)
                g_s = torch.where(tk, g_n, g_s)
                lp_s = torch.where(take, lp_n, lp_s)
                logw_s = torch.where(act, logw_new, logw_s)
                rsum_s = torch.where(m, rsum_s + r_n, rsum_s)
                sum_acc = sum_acc + torch.where(act, acc_n, torch.zeros_like(acc_n))
                n_leaf = n_leaf + act.to(self.dt)
                q_e = torch.where(m, q_n, q_e)
                r_e = torch.where(m, r_n, r_e)
                g_e = torch.where(m, g_n, g_e)
                div_s = div_s | (act & div_n)
                lo, hi = _leaf_idx_to_ckpt_idxs(i)
                if i % 2 == 0:
                    ck_r[:, hi] = torch.where(m, r_n, ck_r[:, hi])
                    ck_s[:, hi] = torch.where(m, rsum_s, ck_s[:, hi])
                else:
                    t_now = torch.zeros(C, dtype=torch.bool, device=dev)
                    for k in range(hi, lo - 1, -1):
                        sub = rsum_s - ck_s[:, k] + ck_r[:, k]
                        t_now = t_now | self._is_turning(ck_r[:, k], r_n, sub)
                    turn_s = turn_s | (act & t_now)
                act = act & ~div_s & ~turn_s
            ok = alive & ~turn_s & ~div_s
            # biased progressive sampling between the old tree and the new subtree
            take = ok & (torch.log(self._rand(C)) < logw_s - log_w)
            tk = take[:, None]
            q_prop = torch.where(tk, q_s, q_prop)
            g_prop = torch.where(tk, g_s, g_prop)
            lp_prop = torch.where(take, lp_s, lp_prop)
            log_w = torch.where(ok, torch.logaddexp(log_w, logw_s), log_w)
            okr, okl = (ok & right)[:, None], (ok & ~right)[:, None]
            qR, rR, gR = torch.where(okr, q_e, qR), torch.where(okr, r_e, rR), torch.where(okr, g_e, gR)
            qL, rL, gL = torch.where(okl, q_e, qL), torch.where(okl, r_e, rL), torch.where(okl, g_e, gL)
            r_sum = torch.where(ok[:, None], r_sum + rsum_s, r_sum)
            turn_full = self._is_turning(rL, rR, r_sum)
            diverged = diverged | (alive & div_s)
            turned = turned | (alive & turn_s) | (ok & turn_full)
            depth = depth + alive.to(torch.int32)
            alive = ok & ~turn_full

        self.q, self.logp, self.grad = q_prop, lp_prop, g_prop
        acc = sum_acc / torch.clamp(n_leaf, min=1.0)
        code = diverged.to(torch.int32) + 2 * (depth == self.max_treedepth).to(torch.int32)
        return NUTSInfo(code, acc, diverged, turned, depth, n_leaf.to(torch.int32), self.evals - evals0)

    # -- warm-up and sampling -------------------------------------------------------------------
    @torch.no_grad()
    def warmup(self, warmup_duration: int = 1000, hook: Callable | None = None, **epoch_kw):
        """Stan-style warm-up; ``hook(self)`` runs after every transition (e.g. a Gibbs update)."""
        infos = []
        for kind, duration in stan_epochs(warmup_duration, **epoch_kw):
            self._da_init()  # start_epoch, nuts.py:336-348
            s1 = torch.zeros((self.C, self.P), dtype=torch.float64, device=self.dev)
            s2 = torch.zeros_like(s1)
            for t in range(duration):
                info = self.transition()
                self._da_step(info.acceptance_prob, t)
                if hook is not None:
                    hook(self)
                qd = self.q.double()
                s1 += qd
                s2 += qd * qd
                infos.append(info)
            self._da_finalize()  # end_epoch, nuts.py:350-362
            if kind == "slow" and duration > 1:  # _tune_slow, nuts.py:306-334
                var = (s2 - s1 * s1 / duration) / (duration - 1)
                new_inv = (var + 0.001).to(self.dt)
                adj = torch.sqrt(self.inv_mass.sum(dim=1) / new_inv.sum(dim=1))
                self.step_size = adj * self.step_size
                self.inv_mass = new_inv
        return infos

    @torch.no_grad()
    def sample(self, num_samples: int, hook: Callable | None = None):
        """Posterior epoch: returns (draws [num_samples, C, P], infos)."""
        out = torch.empty((num_samples, self.C, self.P), dtype=self.dt, device=self.dev)
        infos = []
        for t in range(num_samples):
            infos.append(self.transition())
            if hook is not None:
                hook(self)
            out[t] = self.q
        return out, infos

    def refresh(self):
        """Re-evaluate logp/grad at the current position (after the target changed, e.g. Gibbs)."""
        self.logp, self.grad = self.f(self.q)
        self.logp, self.grad = self.logp.clone(), self.grad.clone()
        self.evals += 1


# =============================================================================================
# HMC with a fixed number of leapfrog steps (reference: goose/hmc.py)
# =============================================================================================
@dataclass
class HMCInfo:
    error_code: torch.Tensor      # int32 [C]: 1 = divergent transition (hmc.py:67-77, error_book :131)
    acceptance_prob: torch.Tensor
    position_moved: torch.Tensor  # bool [C]
    divergent: torch.Tensor       # bool [C]
    num_evals: int


class BatchedHMC(BatchedNUTS):
    """
    Batched counterpart of the reference's ``HMCKernel`` (``src/liesel/goose/hmc.py``): every
    transition integrates ``num_integration_steps`` leapfrog steps (default 10, :149) from a fresh
    momentum and accepts the end point with probability ``min(1, exp(H0 - H1))``; a transition is
    divergent when ``H1 - H0`` exceeds the threshold 1000 (:62-66; a NaN energy counts as +inf).
    Dual averaging, the Stan-style warm-up epochs and the mass-matrix tuning are shared with
    ``BatchedNUTS`` (the reference shares ``da_step`` / ``tune_inv_mm_diag`` between both kernels,
    hmc.py:255-336). One batched logp+grad call per leapfrog step, all chains in lock-step.
    """

    def __init__(self, logp_grad_fn: Callable, q0: torch.Tensor, num_integration_steps: int = 10, **kw):
        kw.setdefault("fused", False)
        self.num_integration_steps = int(num_integration_steps)
        super().__init__(logp_grad_fn, q0, **kw)

    @torch.no_grad()
    def transition(self) -> HMCInfo:
        evals0 = self.evals
        r0 = self._randn(self.C, self.P) / self.inv_mass.sqrt()
        h0 = -self.logp + self._kinetic(r0)
        q, r, lp, g = self.q, r0, self.logp, self.grad
        for _ in range(self.num_integration_steps):
            q, r, lp, g = self._leapfrog(q, r, g, self.step_size)
        h1 = -lp + self._kinetic(r)
        delta = torch.nan_to_num(h0 - h1, nan=-float("inf"))  # blackjax safe_energy_diff
        divergent = -delta > self.div_thr
        p_accept = torch.exp(torch.clamp(delta, max=0.0))
        accept = self._rand(self.C) < p_accept
        sel = accept[:, None]
        self.q = torch.where(sel, q, self.q)
        self.grad = torch.where(sel, g, self.grad)
        self.logp = torch.where(accept, lp, self.logp)
        return HMCInfo(divergent.to(torch.int32), p_accept, accept, divergent, self.evals - evals0)
