"""
The asynchronous NUTS driver (``liesel_b200/nuts_async.py``) on CPU: its state machine written with
masked torch ops (the program the CUDA kernel ``lslb_nuts_async_step`` restates; the GPU suite compares
the two step by step). Pins are the reference's own statistical ones
(tests/goose/kernels/model_lm.py:70-84) plus agreement with the lock-step driver on the
NUTS-within-Gibbs scheme of the ESS metric.
"""
import json
import os

import numpy as np
import pytest
import torch

from liesel_b200 import diagnostics, nuts_async, reparam, synth
from liesel_b200.nuts_gibbs import AsyncNutsGibbs, NutsGibbs, summarise_async
from oracle.cpu_port import CpuEvaluator


def test_checkpoint_index_helpers_match_the_scalar_rule():
    from liesel_b200.nuts import _leaf_idx_to_ckpt_idxs

    i = torch.arange(0, 1024, dtype=torch.int32)
    hi = nuts_async._popcount(i >> 1)
    lo = hi - nuts_async._trailing_ones(i) + 1
    for n in range(1024):
        l, h = _leaf_idx_to_ckpt_idxs(n)
        assert int(hi[n]) == h
        if n & 1:
            assert int(lo[n]) == l


def test_async_standard_normal():
    def f_into(q, lp, g):
        lp.copy_(-0.5 * (q * q).sum(dim=1))
        g.copy_(-q)

    s = nuts_async.AsyncNUTS(f_into, torch.zeros(16, 5, dtype=torch.float64), warmup=300, samples=400, seed=1)
    t = s.run()
    d = s.draws().numpy()
    assert d.shape == (400, 16, 5)
    assert abs(d.mean()) < 0.05
    assert abs(d.var() - 1.0) < 0.08
    sm = s.summary()
    assert 0.75 < sm["mean_accept"] < 0.93 and sm["divergent_frac"] == 0.0
    assert np.all(diagnostics.ess_bulk(d) > 1500)
    # every chain did exactly warmup + samples transitions; a step is one batched evaluation
    assert int(s.st["it"].min()) == int(s.st["it"].max()) == 700
    assert t["steps_warmup"] + t["steps_posterior"] == s.steps
    # the batch stays busy: far fewer sequential evaluations than transitions x deepest tree
    assert s.steps < 700 * 9


def test_async_tail_compaction_does_not_change_the_draws():
    """Dropping the chains that finished a stage from the evaluated batch is bookkeeping only."""
    def make(ct):
        def f_into(q, lp, g, chains=None):
            lp.copy_(-0.5 * (q * q).sum(dim=1))
            g.copy_(-q)

        return nuts_async.AsyncNUTS(f_into, torch.full((16, 5), 0.1, dtype=torch.float64), warmup=160, samples=60,
                                    seed=1, compact_tail=ct)

    a, b = make(False), make(True)
    ta, tb = a.run(check_every=8), b.run(check_every=8)
    assert ta["steps_posterior"] == tb["steps_posterior"] and ta["steps_warmup"] == tb["steps_warmup"]
    assert torch.equal(a.draws(), b.draws())
    assert b.chain_evals < a.chain_evals == a.steps * 16


def test_async_recovers_model_lm(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "model_lm.json")))
    X = torch.tensor(g["X"], dtype=torch.float64)
    y = torch.tensor(g["y"], dtype=torch.float64)

    def f_into(q, lp, gr):
        b, ls = q[:, :2], q[:, 2]
        r = y[None, :] - b @ X.T
        s2 = torch.exp(-2 * ls)
        lp.copy_((-0.5 * r * r * s2[:, None]).sum(dim=1) - 30 * ls - 15 * np.log(2 * np.pi))
        gr.copy_(torch.cat([(r * s2[:, None]) @ X, ((r * r).sum(dim=1) * s2 - 30)[:, None]], dim=1))

    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64).repeat(8, 1)
    s = nuts_async.AsyncNUTS(f_into, q0, warmup=2000, samples=600, seed=42, epoch_kw=dict(term_duration=1500))
    s.run()
    d = s.draws().numpy()
    assert d[:, :, :2].mean(axis=(0, 1)) == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    assert s.summary()["mean_accept"] == pytest.approx(0.8, abs=0.05)
    # the mass matrix was tuned per chain (slow epochs) and the step size by dual averaging
    assert float((s.st["inv_mass"] - 1.0).abs().min()) > 0
    assert int(s.st["ep"].min()) == s.nep


def test_async_nuts_gibbs_agrees_with_the_lock_step_driver():
    """Same model, same scheme (one NUTS block + Gibbs tau2 under the sum-to-zero constraint): the two
    drivers' posterior means of every coordinate agree within Monte-Carlo error, both converge."""
    torch.manual_seed(0)
    spec, centre = synth.s3_gamlss(n=2000)
    ev = CpuEvaluator(spec, threads=4)
    rep = reparam.sum_to_zero(spec, ev.column_means(spec))
    C = 8
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    a = AsyncNutsGibbs(spec, ev, torch.as_tensor(th), torch.as_tensor(aux), 150, 150, reparam=rep, seed=1)
    assert len(a.nuts.gibbs) == 2 and a.nuts.P == 40
    t = a.run(check_every=32)
    sa = summarise_async(a, t)
    assert sa["split_rhat_max"] < 1.06, sa
    # (min bulk-ESS of 40 coordinates from 8 x 150 draws scatters between ~90 and ~320 over seeds)
    assert sa["ess_bulk_min"] > 50 and sa["divergent_frac"] < 0.01 and 0.7 < sa["mean_accept"] < 0.97, sa
    assert float((a.aux - 1.0).abs().min()) > 0  # tau2 was updated by the Gibbs step
    b = NutsGibbs(spec, ev, torch.as_tensor(th), torch.as_tensor(aux), reparam=rep, seed=2)
    b.warmup(150)
    db, _ = b.sample(150)
    da = a.draws()
    ma, mb = da.double().mean(dim=(0, 1)), db.double().mean(dim=(0, 1))
    sd = db.double().std(dim=(0, 1))
    ess = torch.as_tensor(np.minimum(diagnostics.ess_bulk(da.numpy()), diagnostics.ess_bulk(db.numpy())))
    z = (ma - mb).abs() / (sd * torch.sqrt(2.0 / ess))
    assert float(z.max()) < 4.5, z
