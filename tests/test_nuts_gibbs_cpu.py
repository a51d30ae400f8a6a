"""
Host logic of the NUTS-within-Gibbs run behind the ESS metric (``liesel_b200/nuts_gibbs.py``,
``liesel_b200/reparam.py``), on CPU with the oracle's port as the evaluator.
"""
import numpy as np
import torch

from liesel_b200 import reparam, synth
from liesel_b200.nuts_gibbs import NutsGibbs, summarise
from oracle import sam
from oracle.cpu_port import CpuEvaluator


def test_sum_to_zero_basis_properties():
    rng = np.random.default_rng(5)
    c = rng.uniform(0.01, 0.2, size=20)
    c /= c.sum()
    Z = reparam.sum_to_zero_basis(c)
    assert Z.shape == (20, 19)
    np.testing.assert_allclose(c @ Z, 0.0, atol=1e-15)
    np.testing.assert_allclose(Z.T @ Z, np.eye(19), atol=1e-14)
    # the constant vector is NOT in the constrained space: rank(Z^T K Z) = rank(K)
    from liesel_b200.splines import pspline_penalty
    K = pspline_penalty(20, 2)
    assert np.linalg.matrix_rank(Z.T @ K @ Z) == np.linalg.matrix_rank(K) == 18


def test_reparam_is_the_constrained_design():
    """logp(T gamma) with the banded basis == the model a Liesel user builds with the constrained
    DENSE design B Z (``add_np_smooth(X=B Z, K=Z^T K Z)``), up to the constant log-pdet."""
    spec, centre = synth.s3_gamlss(n=3000)
    ev = CpuEvaluator(spec)
    rep = reparam.sum_to_zero(spec, ev.column_means(spec), dtype=torch.float64)
    assert (rep.P, rep.Pr) == (42, 40)
    prep = sam.Prepared(spec, work_dtype=np.float64)
    rng = np.random.default_rng(3)
    g = rng.normal(size=(3, 40)) * 0.3
    q = torch.as_tensor(g) @ rep.Tt
    # constraint holds: fitted smooth sums to zero over the rows
    for ent in prep.blocks:
        blk = ent["blk"]
        if blk.kind == sam.KIND_SPLINE:
            B = prep.dense_design(ent)
            f = B @ q[:, blk.param_offset: blk.param_offset + blk.ncoef].numpy().T
            np.testing.assert_allclose(f.mean(axis=0), 0.0, atol=1e-12)
    aux = np.ones((3, spec.num_aux))
    lp = sam.log_prob(spec, q.numpy(), aux)
    gr = sam.grad(spec, q.numpy(), aux)
    # finite-difference check of the reduced gradient T^T grad
    gr_r = gr @ rep.T.numpy()
    eps = 1e-6
    for j in (0, 5, 22, 39):
        gp = g.copy(); gp[:, j] += eps
        gm = g.copy(); gm[:, j] -= eps
        fd = (sam.log_prob(spec, gp @ rep.Tt.numpy(), aux) - sam.log_prob(spec, gm @ rep.Tt.numpy(), aux)) / (2 * eps)
        np.testing.assert_allclose(fd, gr_r[:, j], rtol=2e-5, atol=1e-4)
    assert np.all(np.isfinite(lp))


def test_nuts_gibbs_converges_under_the_constraint():
    """Small-n H: with the sum-to-zero constraint the chains agree (split-R-hat near 1); the
    unconstrained run of round 1 ended at 5.05."""
    torch.manual_seed(0)
    spec, centre = synth.s3_gamlss(n=2000)
    ev = CpuEvaluator(spec, threads=4)
    rep = reparam.sum_to_zero(spec, ev.column_means(spec))
    C = 8
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    run = NutsGibbs(spec, ev, torch.as_tensor(th), torch.as_tensor(aux), reparam=rep, seed=1)
    run.warmup(150)
    draws, infos = run.sample(120)
    s = summarise(draws, infos, 1.0)
    assert s["split_rhat_max"] < 1.06, s
    assert s["ess_bulk_min"] > 200, s
    assert s["divergent_frac"] < 0.01
    assert 0.7 < s["mean_accept"] < 0.97
    assert s["mean_treedepth"] < 7
    # tau2 was updated by the Gibbs step
    assert float((run.aux - 1.0).abs().min()) > 0
