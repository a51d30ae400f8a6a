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
