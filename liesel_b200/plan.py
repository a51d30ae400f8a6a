"""
Device plan: uploads a :class:`~liesel_b200.spec.ModelSpec` once, builds banded spline bases
on the device, orders rows for long constant-span runs and creates the immutable
``lslb_plan`` of the C ABI. All likelihood arithmetic happens inside ``libliesel_b200.so``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .spec import (KIND_DENSE, KIND_GROUP, KIND_INTERCEPT, KIND_SPLINE, PRIOR_MVND, PRIOR_NONE,
                   PRIOR_NORMAL, ModelSpec, MVNDPrior, NormalPrior)

_TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


class Plan:
    """
    One model, resident on one GPU. Data (y, X, bands, group index, K) is uploaded once and
    shared by all chains — the reference replicates it per chain instead
    (``goose/builder.py:304-307``).
    """

    def __init__(self, spec: ModelSpec, device=None, sort_rows: bool = True):
        if not torch.cuda.is_available():
            raise _lib.LieselB200Error("liesel_b200 needs a CUDA device (no CPU fallback)")
        self.spec = spec
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = _TORCH[np.dtype(spec.dtype)]
        self._keep: list[torch.Tensor] = []
        self._ws: dict[int, torch.Tensor] = {}
        self._ws_bytes: dict[int, int] = {}
        self._ws_retired: list[torch.Tensor] = []  # outgrown buffers stay alive (captured CUDA graphs)
        self._handle = C.c_void_p()
        dev, dt = self.device, self.dtype

        with torch.cuda.device(dev):
            y = torch.as_tensor(spec.y, device=dev).to(dt).contiguous()
            rows = {"y": y}
            has_dense = False
            for b, blk in enumerate(spec.blocks):
                if blk.kind == KIND_DENSE:
                    rows[f"X{b}"] = torch.as_tensor(blk.X, device=dev).to(dt).contiguous()
                    has_dense = True
                elif blk.kind == KIND_SPLINE:
                    if blk.start is not None:
                        start = torch.as_tensor(blk.start, device=dev).to(torch.int32).contiguous()
                        w = torch.as_tensor(blk.weights, device=dev).to(dt).contiguous()
                    else:
                        x = torch.as_tensor(blk.x, device=dev).to(dt).contiguous()
                        knots = torch.as_tensor(np.asarray(blk.knots), device=dev).to(dt)
                        start, w = _lib.spline_basis(x, knots, blk.order)
                    rows[f"s{b}"], rows[f"w{b}"] = start, w
                elif blk.kind == KIND_GROUP:
                    rows[f"g{b}"] = torch.as_tensor(blk.idx, device=dev).to(torch.int32).contiguous()

            # Row order. Sums over rows are order-independent, so rows may be permuted freely;
            # lexicographic order by spline span gives long runs in which the four active
            # coefficients of a smooth stay in registers. Rows of a group stay contiguous.
            self.perm = None
            spl = [b for b, blk in enumerate(spec.blocks) if blk.kind == KIND_SPLINE]
            grp = [b for b, blk in enumerate(spec.blocks) if blk.kind == KIND_GROUP]
            if sort_rows and spl and not has_dense and y.numel() > 1:
                key = torch.zeros(y.numel(), dtype=torch.int64, device=dev)
                if grp:
                    key = rows[f"g{grp[0]}"].to(torch.int64)
                for b in spl[:6]:
                    key = key * 64 + rows[f"s{b}"].to(torch.int64).clamp_(0, 63)
                self.perm = torch.argsort(key, stable=True)
                rows = {k: v.index_select(0, self.perm).contiguous() for k, v in rows.items()}
            self.rows = rows

            blocks = (_lib.BlockDesc * len(spec.blocks))()
            for b, blk in enumerate(spec.blocks):
                d = blocks[b]
                d.kind, d.predictor, d.ncoef = blk.kind, blk.predictor, blk.ncoef
                d.data, d.index, d.ld, d.K = None, None, 0, None
                d.tau2_slot, d.has_ig, d.center = -1, 0, None
                if blk.kind == KIND_DENSE:
                    d.data = rows[f"X{b}"].data_ptr()
                    d.ld = rows[f"X{b}"].shape[1]
                elif blk.kind == KIND_SPLINE:
                    d.data = rows[f"w{b}"].data_ptr()
                    d.index = rows[f"s{b}"].data_ptr()
                    if blk.center:
                        if blk.colmean is None:  # column means of the basis, accumulated in fp64
                            blk.colmean = self._column_means(rows, b, blk.ncoef).astype(spec.dtype)
                        cmt = torch.as_tensor(np.asarray(blk.colmean), device=dev).to(dt).contiguous()
                        self._keep.append(cmt)
                        d.center = cmt.data_ptr()
                elif blk.kind == KIND_GROUP:
                    d.index = rows[f"g{b}"].data_ptr()
                pr = blk.prior
                if pr is None:
                    d.prior = PRIOR_NONE
                elif isinstance(pr, NormalPrior):
                    d.prior, d.prior_m, d.prior_s = PRIOR_NORMAL, pr.m, pr.s
                elif isinstance(pr, MVNDPrior):
                    Kt = torch.as_tensor(np.asarray(pr.K), device=dev).to(dt).contiguous()
                    self._keep.append(Kt)
                    d.prior, d.K = PRIOR_MVND, Kt.data_ptr()
                    d.rank, d.log_pdet = float(pr.rank), float(pr.log_pdet)
                    d.tau2_slot, d.tau2_fixed = blk.tau2_slot, pr.tau2_init
                    if pr.ig_a is not None and blk.tau2_slot >= 0:
                        d.has_ig, d.ig_a, d.ig_b = 1, pr.ig_a, pr.ig_b
            desc = _lib.ModelDesc()
            desc.dtype = 0 if dt == torch.float32 else 1
            desc.family = spec.family
            desc.n = y.numel()
            desc.y = rows["y"].data_ptr()
            desc.fixed_scale = spec.fixed_scale
            desc.const_term = spec.const_term
            desc.nblocks = len(spec.blocks)
            desc.blocks = blocks
            _lib.check(_lib.lib.lslb_plan_create(C.byref(desc), C.byref(self._handle)),
                       "lslb_plan_create")
        self.P = _lib.lib.lslb_plan_num_params(self._handle)
        self.A = _lib.lib.lslb_plan_num_aux(self._handle)
        self.n = int(y.numel())
        self.variant = _lib.lib.lslb_plan_variant(self._handle).decode()
        assert self.P == spec.num_params and self.A == spec.num_aux

    @staticmethod
    def _column_means(rows, b: int, ncoef: int) -> np.ndarray:
        w64 = rows[f"w{b}"].to(torch.float64)
        cols = rows[f"s{b}"].to(torch.int64)[:, None] + torch.arange(w64.shape[1], device=w64.device)[None, :]
        cm = torch.zeros(ncoef, dtype=torch.float64, device=w64.device)
        cm.index_add_(0, cols.reshape(-1), w64.reshape(-1))
        return (cm / max(w64.shape[0], 1)).cpu().numpy()

    def kernel_name(self, C_: int) -> str:
        """Symbol of the row-streaming kernel a logp+grad call on ``C_`` chains launches (as ncu names it)."""
        return _lib.lib.lslb_plan_kernel_name(self._handle, int(C_)).decode()

    def column_means(self, block: int | str) -> np.ndarray:
        """Column means (fp64) of a spline block's basis matrix, from the resident band."""
        b = self.spec.block_index(block) if isinstance(block, str) else int(block)
        if self.spec.blocks[b].kind != KIND_SPLINE:
            raise ValueError("column_means is defined for spline blocks")
        return self._column_means(self.rows, b, self.spec.blocks[b].ncoef)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value and _lib is not None and getattr(_lib, "lib", None) is not None:
            _lib.lib.lslb_plan_destroy(h)  # (module globals may already be gone at interpreter exit)
            h.value = None

    # -- workspace ------------------------------------------------------------------
    def workspace(self, C_: int) -> torch.Tensor:
        """One scratch buffer per plan, grown to the largest chain count seen (a batched sampler that
        drops finished chains calls with many different C)."""
        need = self._ws_bytes.get(C_)
        if need is None:
            need = max(int(_lib.lib.lslb_workspace_bytes(self._handle, C_)), 256)
            self._ws_bytes[C_] = need
        ws = self._ws.get(0)
        if ws is None or ws.numel() < need:
            if ws is not None:
                self._ws_retired.append(ws)
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[0] = ws
        return ws

    def _check_inputs(self, params: torch.Tensor, aux: torch.Tensor | None):
        _lib.require_cuda(params, aux)
        if params.dtype != self.dtype or params.dim() != 2 or params.shape[1] != self.P:
            raise _lib.LieselB200Error(f"params must be [{'C'}, {self.P}] {self.dtype}")
        if self.A > 0:
            if aux is None or aux.dtype != self.dtype or aux.shape != (params.shape[0], self.A):
                raise _lib.LieselB200Error(f"aux must be [C, {self.A}] {self.dtype}")
        return params.shape[0]

    # -- compute entry points -------------------------------------------------------
    def logp_grad(self, params, aux=None, want_grad: bool = True, flags: int = 0, out=None, shard=None):
        """
        (logp [C], grad [C, P] or None) for C chains: one pass over the rows. ``shard`` is a
        ``dist.PeerGroup`` or ``dist.NcclComm`` when this plan holds one rank's rows: the partial
        sums are then added over the ranks inside the same C call.
        """
        Cn = self._check_inputs(params, aux)
        if out is None:
            logp = torch.empty(Cn, dtype=self.dtype, device=self.device)
            grad = torch.empty((Cn, self.P), dtype=self.dtype, device=self.device) if want_grad else None
        else:
            logp, grad = out
            if not want_grad:
                grad = None
            _lib.require_cuda(logp, grad)
            for name, t, shape in (("logp", logp, (Cn,)), ("grad", grad, (Cn, self.P))):
                if t is not None and (t.dtype != self.dtype or t.device != self.device or tuple(t.shape) != shape):
                    raise _lib.LieselB200Error(f"out {name} must be a contiguous {shape} {self.dtype} tensor on {self.device}")
            if want_grad and grad is None:
                raise _lib.LieselB200Error("want_grad=True needs a grad output tensor")
        ws = self.workspace(Cn)
        with torch.cuda.device(self.device):
            if shard is None:
                fn = getattr(_lib.lib, "lslb_logp_grad_" + _lib.sfx(self.dtype))
                _lib.check(fn(self._handle, Cn, _lib.ptr(params), _lib.ptr(aux), _lib.ptr(logp),
                              _lib.ptr(grad), flags, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                           "lslb_logp_grad")
            else:
                fn = getattr(_lib.lib, f"lslb_logp_grad_{shard.transport}_" + _lib.sfx(self.dtype))
                _lib.check(fn(self._handle, Cn, _lib.ptr(params), _lib.ptr(aux), _lib.ptr(logp),
                              _lib.ptr(grad), flags, _lib.ptr(ws), ws.numel(), shard.handle,
                              _lib.stream_ptr()), "lslb_logp_grad_" + shard.transport)
        return logp, grad

    def iwls(self, block: int | str, params, aux=None, want_chol: bool = True, flags: int = 0, shard=None):
        """(score [C,p], info [C,p,p], chol [C,p,p] or None) of one coefficient block; ``shard`` as in
        :meth:`logp_grad` (the Cholesky factor is then taken of the information summed over ranks)."""
        Cn = self._check_inputs(params, aux)
        b = self.spec.block_index(block) if isinstance(block, str) else int(block)
        p = self.spec.blocks[b].ncoef
        score = torch.empty((Cn, p), dtype=self.dtype, device=self.device)
        info = torch.empty((Cn, p, p), dtype=self.dtype, device=self.device)
        chol = torch.empty((Cn, p, p), dtype=self.dtype, device=self.device) if want_chol else None
        ws = self.workspace(Cn)
        with torch.cuda.device(self.device):
            if shard is None:
                fn = getattr(_lib.lib, "lslb_iwls_" + _lib.sfx(self.dtype))
                _lib.check(fn(self._handle, Cn, b, _lib.ptr(params), _lib.ptr(aux), _lib.ptr(score),
                              _lib.ptr(info), _lib.ptr(chol), flags, _lib.ptr(ws), ws.numel(),
                              _lib.stream_ptr()), "lslb_iwls")
            else:
                fn = getattr(_lib.lib, f"lslb_iwls_{shard.transport}_" + _lib.sfx(self.dtype))
                _lib.check(fn(self._handle, Cn, b, _lib.ptr(params), _lib.ptr(aux), _lib.ptr(score),
                              _lib.ptr(info), _lib.ptr(chol), flags, _lib.ptr(ws), ws.numel(),
                              shard.handle, _lib.stream_ptr()), "lslb_iwls_" + shard.transport)
        return score, info, chol

    def iwls_group(self, params, aux=None, flags: int = 0, want_loglik: bool = False):
        """(score [C,G], info_diag [C,G]) of the random-intercept block: the block's information matrix is
        diagonal (reference: ``-jacfwd(grad(log_prob))``, ``goose/iwls.py:218-243``, a dense G x G matrix).
        ``want_loglik``: also each group's share of the log-posterior as a function of its own effect
        ([C,G]: its rows' log-likelihood + its prior term; the posterior of b factorises over groups)."""
        Cn = self._check_inputs(params, aux)
        grp = [b for b in self.spec.blocks if b.kind == 2]
        if not grp:
            raise ValueError("the model has no random-intercept block")
        G = grp[0].ncoef
        score = torch.empty((Cn, G), dtype=self.dtype, device=self.device)
        info = torch.empty((Cn, G), dtype=self.dtype, device=self.device)
        ll = torch.empty((Cn, G), dtype=self.dtype, device=self.device) if want_loglik else None
        with torch.cuda.device(self.device):
            fn = getattr(_lib.lib, "lslb_iwls_group_" + _lib.sfx(self.dtype))
            _lib.check(fn(self._handle, Cn, _lib.ptr(params), _lib.ptr(aux), _lib.ptr(score), _lib.ptr(info),
                          _lib.ptr(ll), flags, _lib.stream_ptr()), "lslb_iwls_group")
        return (score, info, ll) if want_loglik else (score, info)

    def quadform(self, block: int | str, params):
        """beta^T K beta per chain (``tau2_gibbs_kernel``, reference ``distreg.py:291``)."""
        _lib.require_cuda(params)
        b = self.spec.block_index(block) if isinstance(block, str) else int(block)
        out = torch.empty(params.shape[0], dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            fn = getattr(_lib.lib, "lslb_quadform_" + _lib.sfx(self.dtype))
            _lib.check(fn(self._handle, params.shape[0], b, _lib.ptr(params), _lib.ptr(out),
                          _lib.stream_ptr()), "lslb_quadform")
        return out


def chol(info: torch.Tensor) -> torch.Tensor:
    """Ridge (1e-6 * mean diag) + lower Cholesky per chain; NaN-filled where it fails."""
    _lib.require_cuda(info)
    Cn, p, _ = info.shape
    out = torch.empty_like(info)
    with torch.cuda.device(info.device):
        fn = getattr(_lib.lib, "lslb_chol_" + _lib.sfx(info.dtype))
        _lib.check(fn(Cn, p, _lib.ptr(info), _lib.ptr(out), _lib.stream_ptr()), "lslb_chol")
    return out


def iwls_propose(beta, score, chol_, step, z=None, x_eval=None):
    """
    Proposal tail of the IWLS kernel (reference ``goose/iwls.py:323-327,335-336``):
    returns (prop or None, logq). With ``z``: draw and forward log-density; with ``x_eval``:
    backward log-density of ``x_eval``.
    """
    _lib.require_cuda(beta, score, chol_, step, z, x_eval)
    Cn, p = beta.shape
    prop = torch.empty_like(beta) if z is not None else None
    logq = torch.empty(Cn, dtype=beta.dtype, device=beta.device)
    with torch.cuda.device(beta.device):
        fn = getattr(_lib.lib, "lslb_iwls_propose_" + _lib.sfx(beta.dtype))
        _lib.check(fn(Cn, p, _lib.ptr(beta), _lib.ptr(score), _lib.ptr(chol_), _lib.ptr(step),
                      _lib.ptr(z), _lib.ptr(x_eval), _lib.ptr(prop), _lib.ptr(logq),
                      _lib.stream_ptr()), "lslb_iwls_propose")
    return prop, logq


def iwls_propose_diag(beta, score, info_diag, step, z=None, x_eval=None, elementwise: bool = False):
    """
    :func:`iwls_propose` for a diagonal information matrix (random-intercept block, ``Plan.iwls_group``):
    returns (prop or None, logq, error_code int32 [C]); code 2 = the ridged diagonal was not positive, identity
    fallback as ``IWLSKernel._safe_chol`` (``goose/iwls.py:245-290``). ``elementwise``: ridge every entry by
    1e-6 of itself (its proposal then depends on that entry alone) and return the per-entry log-densities
    [C, p] as a fourth value.
    """
    _lib.require_cuda(beta, score, info_diag, step, z, x_eval)
    Cn, p = beta.shape
    prop = torch.empty_like(beta) if z is not None else None
    logq = torch.empty(Cn, dtype=beta.dtype, device=beta.device)
    code = torch.empty(Cn, dtype=torch.int32, device=beta.device)
    el = torch.empty_like(beta) if elementwise else None
    with torch.cuda.device(beta.device):
        fn = getattr(_lib.lib, "lslb_iwls_propose_diag_" + _lib.sfx(beta.dtype))
        _lib.check(fn(Cn, p, _lib.ptr(beta), _lib.ptr(score), _lib.ptr(info_diag), _lib.ptr(step),
                      _lib.ptr(z), _lib.ptr(x_eval), _lib.ptr(prop), _lib.ptr(logq), _lib.ptr(code),
                      _lib.ptr(el), 1 if elementwise else 0, _lib.stream_ptr()), "lslb_iwls_propose_diag")
    return (prop, logq, code, el) if elementwise else (prop, logq, code)


def mh_step(u, cur_lp, prop_lp, log_corr=None, params=None, params_prop=None):
    """
    Metropolis-Hastings step of the reference (``goose/mh.py:48-68``) for all chains in one launch:
    returns (error_code int32 [C], acceptance_prob [C], moved bool [C]); ``cur_lp`` and (when given)
    ``params`` are updated IN PLACE where the proposal is accepted.
    """
    _lib.require_cuda(u, cur_lp, prop_lp, log_corr, params, params_prop)
    Cn = u.shape[0]
    P = 0 if params is None else params.shape[1]
    code = torch.empty(Cn, dtype=torch.int32, device=u.device)
    acc = torch.empty_like(u)
    moved = torch.empty(Cn, dtype=torch.uint8, device=u.device)
    with torch.cuda.device(u.device):
        fn = getattr(_lib.lib, "lslb_mh_step_" + _lib.sfx(u.dtype))
        _lib.check(fn(Cn, P, _lib.ptr(u), _lib.ptr(cur_lp), _lib.ptr(prop_lp), _lib.ptr(log_corr),
                      _lib.ptr(params), _lib.ptr(params_prop), _lib.ptr(code), _lib.ptr(acc),
                      _lib.ptr(moved), _lib.stream_ptr()), "lslb_mh_step")
    return code, acc, moved.bool()
