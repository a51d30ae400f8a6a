"""
The batched NUTS driver and the ESS estimator on CPU (torch tensors; the log-posterior comes from
the oracle, the B200 library cannot run here). Statistical pins are the reference's own:
tests/goose/kernels/model_lm.py:70-84 — posterior mean of beta within 5 % of OLS, mean log sigma
within 5 % of log sigma_OLS, mean acceptance within 0.05 of the dual-averaging target.
"""
import json
import os

import numpy as np
import pytest
import torch

from liesel_b200 import diagnostics, nuts
from liesel_b200.spec import ModelSpec
from oracle import sam


def test_stan_epochs_match_reference_schedule():
    # goose/warmup.py:12-88 with the defaults: 75 | 25 50 100 200 500 | 50
    assert nuts.stan_epochs(1000) == [("fast", 75), ("slow", 25), ("slow", 50), ("slow", 100),
                                      ("slow", 200), ("slow", 500), ("fast", 50)]
    assert sum(d for _, d in nuts.stan_epochs(300)) == 300
    with pytest.raises(ValueError):
        nuts.stan_epochs(10)
    with pytest.raises(ValueError):
        nuts.stan_epochs(100)


def test_ckpt_indices():
    # leaf 0..7 of an 8-leaf subtree: even leaves store, odd leaves check the ranges below
    idx = [nuts._leaf_idx_to_ckpt_idxs(i) for i in range(8)]
    assert [hi for _, hi in idx[0::2]] == [0, 1, 1, 2]          # where even leaves are stored
    assert idx[1::2] == [(0, 0), (0, 1), (1, 1), (0, 2)]        # ranges odd leaves check


def test_ess_iid_and_ar1():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1000, 8, 2))
    e = diagnostics.ess_bulk(x)
    assert np.all(e > 6000) and np.all(e < 10500)
    # AR(1) with rho = 0.9: ESS ~ N (1 - rho) / (1 + rho) = N / 19
    n, c, rho = 4000, 4, 0.9
    z = rng.normal(size=(n, c))
    y = np.zeros((n, c))
    for t in range(1, n):
        y[t] = rho * y[t - 1] + np.sqrt(1 - rho**2) * z[t]
    e = diagnostics.ess_bulk(y)[0]
    assert 0.6 * n * c / 19 < e < 1.6 * n * c / 19
    # a constant shift per chain (no mixing) must collapse the ESS
    bad = y + np.arange(c)[None, :] * 10.0
    assert diagnostics.ess_bulk(bad)[0] < 20


def test_nuts_standard_normal():
    torch.manual_seed(0)

    def f(q):
        return -0.5 * (q * q).sum(dim=1), -q

    s = nuts.BatchedNUTS(f, torch.zeros(16, 5, dtype=torch.float64), seed=1)
    infos = s.warmup(300)
    draws, infos = s.sample(400)
    d = draws.numpy()
    assert abs(d.mean()) < 0.05
    assert abs(d.var() - 1.0) < 0.08
    acc = np.mean([float(i.acceptance_prob.mean()) for i in infos])
    assert 0.75 < acc < 0.93  # the averaged step size is conservative (Jensen), as in Stan
    assert max(int(i.treedepth.max()) for i in infos) <= 10
    assert np.all(diagnostics.ess_bulk(d) > 1500)


def test_nuts_recovers_model_lm(golden_dir):
    """The reference's kernel integration test, driven by our NUTS + oracle gradients."""
    g = json.load(open(os.path.join(golden_dir, "model_lm.json")))
    X, y = np.array(g["X"]), np.array(g["y"])
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    spec.add_p_smooth(X, 0.0, None, "loc", "beta")
    spec.add_p_smooth(np.ones((30, 1)), 0.0, None, "scale", "log_sigma")
    prep = sam.Prepared(spec)

    def f(q):
        th = q.numpy()
        return (torch.from_numpy(sam.log_prob(spec, th, prep=prep)),
                torch.from_numpy(sam.grad(spec, th, prep=prep)))

    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64).repeat(4, 1)
    s = nuts.BatchedNUTS(f, q0, seed=42)
    # the reference's test uses a long final fast-adaptation epoch (term_duration=4000 of 5000,
    # model_lm.py:56-60) so that the averaged step size has converged; scaled down here
    s.warmup(2000, term_duration=1500)
    draws, infos = s.sample(800)
    d = draws.numpy()
    avg_beta = d[:, :, :2].mean(axis=(0, 1))
    assert avg_beta == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    acc = np.mean([float(i.acceptance_prob.mean()) for i in infos])
    assert acc == pytest.approx(0.8, abs=0.05)
    codes = np.concatenate([i.error_code.numpy() for i in infos])
    assert (codes == 0).mean() > 0.98


def test_device_diagnostics_match_host_versions():
    """The torch (device) bulk-ESS / split-R-hat are the same algorithm as the NumPy ones."""
    rng = np.random.default_rng(3)
    n, c, P = 400, 6, 3
    x = np.zeros((n, c, P))
    e = rng.normal(size=(n, c, P))
    for t in range(1, n):
        x[t] = 0.8 * x[t - 1] + e[t]
    x[:, :, 1] = np.round(x[:, :, 1], 1)  # ties: average ranks
    x[:, 0, 2] += 2.5  # one chain elsewhere: R-hat well above 1
    t = torch.as_tensor(x)
    np.testing.assert_allclose(diagnostics.ess_bulk_device(t), diagnostics.ess_bulk(x), rtol=1e-9)
    r = diagnostics.split_rhat(x)
    np.testing.assert_allclose(diagnostics.split_rhat_device(t), r, rtol=1e-9)
    assert r[0] < 1.05 and r[2] > 1.1


def test_hmc_recovers_model_lm(golden_dir):
    """The reference's HMC kernel test (tests/goose/kernels/test_hmc.py via model_lm.py:70-84)."""
    g = json.load(open(os.path.join(golden_dir, "model_lm.json")))
    X, y = np.array(g["X"]), np.array(g["y"])
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    spec.add_p_smooth(X, 0.0, None, "loc", "beta")
    spec.add_p_smooth(np.ones((30, 1)), 0.0, None, "scale", "log_sigma")
    prep = sam.Prepared(spec)

    def f(q):
        th = q.numpy()
        return (torch.from_numpy(sam.log_prob(spec, th, prep=prep)),
                torch.from_numpy(sam.grad(spec, th, prep=prep)))

    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64).repeat(4, 1)
    s = nuts.BatchedHMC(f, q0, seed=7)
    assert s.num_integration_steps == 10  # hmc.py:149
    s.warmup(5000, term_duration=4000)  # model_lm.py:56-60
    draws, infos = s.sample(800)
    d = draws.numpy()
    assert d[:, :, :2].mean(axis=(0, 1)) == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    acc = np.mean([float(i.acceptance_prob.mean()) for i in infos])
    assert acc == pytest.approx(0.8, abs=0.05)
    assert all(i.num_evals == 10 for i in infos)
    codes = np.concatenate([i.error_code.numpy() for i in infos])
    assert (codes == 0).mean() > 0.98
    moved = np.mean([float(i.position_moved.float().mean()) for i in infos])
    assert 0.6 < moved < 0.95


def test_hmc_nan_energy_is_divergent_and_rejected():
    def f(q):
        lp = torch.where(q[:, 0] > 0.5, torch.full_like(q[:, 0], float("nan")), -0.5 * (q * q).sum(dim=1))
        return lp, -q

    s = nuts.BatchedHMC(f, torch.zeros(8, 2, dtype=torch.float64), seed=3, initial_step_size=5.0,
                        num_integration_steps=3)
    q_before = s.q.clone()
    info = s.transition()
    bad = info.divergent
    assert torch.equal(s.q[bad], q_before[bad])
    assert (info.acceptance_prob[bad] == 0).all() and (info.error_code[bad] == 1).all()


# --- the fused subtree bookkeeping (and chain compaction) on CPU --------------------------------
class TorchLeafNUTS(nuts.BatchedNUTS):
    """BatchedNUTS whose two per-leaf CUDA kernels (csrc/nuts.cu) are replaced by torch restatements
    acting on the same state buffers, so that the fused driver — buffer management, leaf ranges,
    checkpoints, compaction — runs without a GPU."""

    def _leaf_pre(self, fn, st, ptrs, ints, stream, leaf):
        Cb = ints[0]
        e = st["eps"][:Cb, None]
        rh = st["r_e"][:Cb] + 0.5 * e * st["g_e"][:Cb]
        st["r_half"][:Cb] = rh
        st["q_n"][:Cb] = st["q_e"][:Cb] + e * (st["inv_mass"][:Cb] * rh)

    def _leaf_post(self, fn, st, ptrs, ints, stream, leaf):
        Cb, odd, lo, hi = ints[0], ints[3], ints[4], ints[5]
        v = lambda k: st[k][:Cb]
        e, im = v("eps")[:, None], v("inv_mass")
        rn = v("r_half") + 0.5 * e * v("g_n")
        st["r_half"][:Cb] = rn
        act = v("act").bool()
        kin = 0.5 * (im * rn * rn).sum(dim=1)
        dh = torch.nan_to_num((-v("lp_n") + kin) - v("h0"), nan=float("inf"))
        div = dh > self.div_thr
        logw_n = -dh
        acc = torch.exp(torch.clamp(-dh, max=0.0))
        lw_new = torch.logaddexp(v("logw_s"), logw_n)
        take = act & (torch.log(st["U"][leaf][:Cb]) < logw_n - lw_new)
        rs = v("rsum_s") + rn
        turning = torch.zeros(Cb, dtype=torch.bool)
        if odd:
            for k in range(hi, lo - 1, -1):
                rl = v("ck_r")[:, k]
                sub = rs - v("ck_s")[:, k] + rl
                rho = sub - 0.5 * (rn + rl)
                turning |= ((im * rl * rho).sum(dim=1) <= 0) | ((im * rn * rho).sum(dim=1) <= 0)
        m = act[:, None]
        st["rsum_s"][:Cb] = torch.where(m, rs, v("rsum_s"))
        for dst, src in (("q_s", "q_n"), ("g_s", "g_n")):
            st[dst][:Cb] = torch.where(take[:, None], v(src), v(dst))
        for dst, src in (("q_e", v("q_n")), ("r_e", rn), ("g_e", v("g_n"))):
            st[dst][:Cb] = torch.where(m, src, v(dst))
        if not odd:
            st["ck_r"][:Cb, hi] = torch.where(m, rn, v("ck_r")[:, hi])
            st["ck_s"][:Cb, hi] = torch.where(m, rs, v("ck_s")[:, hi])
        st["lp_s"][:Cb] = torch.where(take, v("lp_n"), v("lp_s"))
        st["logw_s"][:Cb] = torch.where(act, lw_new, v("logw_s"))
        st["sum_acc"][:Cb] = v("sum_acc") + torch.where(act, acc, torch.zeros_like(acc))
        st["n_leaf"][:Cb] = v("n_leaf") + act.to(v("n_leaf").dtype)
        d = v("div_s").bool() | (act & div)
        t = v("turn_s").bool() | (act & turning)
        st["div_s"][:Cb] = d.to(torch.uint8)
        st["turn_s"][:Cb] = t.to(torch.uint8)
        st["act"][:Cb] = (act & ~(d | t)).to(torch.uint8)


def test_fused_subtree_and_compaction_on_cpu():
    """The fused driver (with the kernels restated in torch) reproduces the plain torch path draw for
    draw, with and without dropping terminated chains from the batch; per-chain step sizes give the
    heterogeneous tree depths that make compaction kick in."""
    C, P = 200, 6
    scale = torch.linspace(0.5, 3.0, P, dtype=torch.float64)

    def f(q):
        return -0.5 * ((q / scale) ** 2).sum(dim=1), -q / scale**2

    seen = []

    def into(q, lp, g, chains=None):
        seen.append(q.shape[0])
        a, b = f(q)
        lp.copy_(a)
        g.copy_(b)

    q0 = torch.randn(C, P, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    kw = dict(seed=11, initial_step_size=0.1, max_treedepth=7)
    ref = nuts.BatchedNUTS(f, q0, fused=False, **kw)
    fus = TorchLeafNUTS(f, q0, fused=True, logp_grad_into=into, **kw)
    cmp_ = TorchLeafNUTS(f, q0, fused=True, logp_grad_into=into, compact=True, **kw)
    assert cmp_.compact and not fus.compact
    eps = 0.05 * torch.pow(torch.tensor(2.0, dtype=torch.float64), torch.arange(C, dtype=torch.float64) % 5)
    for s in (ref, fus, cmp_):
        s.step_size = eps.clone()
    for _ in range(5):
        i_ref, i_fus = ref.transition(), fus.transition()
        seen.clear()
        i_cmp = cmp_.transition()
        for info, s in ((i_fus, fus), (i_cmp, cmp_)):
            assert torch.equal(info.treedepth, i_ref.treedepth)
            assert torch.equal(info.error_code, i_ref.error_code)
            assert torch.equal(info.leapfrog, i_ref.leapfrog)
            np.testing.assert_allclose(info.acceptance_prob.numpy(), i_ref.acceptance_prob.numpy(), rtol=1e-10)
            np.testing.assert_allclose(s.q.numpy(), ref.q.numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(s.logp.numpy(), ref.logp.numpy(), rtol=1e-10)
        assert min(seen) < C and all(b % 64 == 0 or b == C for b in seen)  # compacted, padded to 64
    assert int(i_ref.treedepth.max()) > int(i_ref.treedepth.min())
    assert cmp_.evals == fus.evals and cmp_.chain_evals < 0.9 * fus.chain_evals
