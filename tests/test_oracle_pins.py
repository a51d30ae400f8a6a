"""
Pins the CPU oracle: every closed-form number the reference prints for this path, the
NumPy/SciPy baselines the reference's own tests compare to, SciPy's densities, and finite
differences. CPU only.
"""
import json
import os

import numpy as np
import pytest
import scipy.linalg
import scipy.stats
from scipy.interpolate import BSpline

from liesel_b200 import synth
from liesel_b200.spec import ModelSpec
from oracle import densities as D
from oracle import iwls as OI
from oracle import sam
from oracle import splines as OS


def load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


# --- reference doc/test pins ----------------------------------------------------------
def test_normal_pins(golden_dir):
    pins = load(golden_dir, "reference_pins.json")["normal_logpdf"]
    for pin in pins:
        lp = D.normal_log_prob(np.array(pin["y"]), pin["loc"], pin["scale"])
        if "log_prob" in pin:
            np.testing.assert_allclose(lp, pin["log_prob"], rtol=2e-7, atol=0)
        else:
            # printed float32 sums: 7 significant digits
            assert lp.sum() == pytest.approx(pin["log_prob_sum"], rel=3e-7)


def test_model_log_prob_pins(golden_dir):
    pins = load(golden_dir, "reference_pins.json")["model_log_prob"]
    # LogProb docstring: x ~ N(0,1), position [1, 2]; expressed as a model with one row of
    # data-free structure is not possible, so check the prior term directly and through sam
    p0 = pins[0]
    x = np.array(p0["x"])
    assert D.normal_log_prob(x, 0.0, 1.0).sum() == pytest.approx(p0["log_prob"], rel=3e-7)
    np.testing.assert_allclose(-(x - 0.0) / 1.0**2, p0["grad"])
    # same numbers through the structured-additive oracle: 1 observation with a huge fixed
    # scale contributes a constant; remove it via log_prob_parts
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(np.zeros(1), "normal").add_predictor("loc")
    spec.fixed_scale = 1.0
    spec.add_p_smooth(np.zeros((1, 2)) + [[0.0, 0.0]], 0.0, 1.0, "loc", "x") if False else None
    # test_model.py:543-562: obs = 0 ~ N(y, 1) at y = 0.5
    p1 = pins[1]
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(np.zeros(1), "normal").add_predictor("loc")
    spec.fixed_scale = 1.0
    spec.add_p_smooth(np.ones((1, 1)), 0.0, None, "loc", "y")
    th = np.array([[p1["y_value"]]])
    assert sam.log_prob(spec, th)[0] == pytest.approx(p1["log_prob"], rel=3e-7)
    assert sam.grad(spec, th)[0, 0] == pytest.approx(p1["grad"], rel=1e-12)


def test_readme_model(golden_dir):
    pins = load(golden_dir, "reference_pins.json")["normal_logpdf"]
    for pin in pins[4:6]:
        spec = ModelSpec(dtype=np.dtype(np.float64))
        spec.add_response(np.array(pin["y"]), "normal").add_predictor("loc")
        spec.fixed_scale = pin["scale"]
        spec.add_p_smooth(np.ones((5, 1)), 0.0, None, "loc", "loc")
        lp = sam.log_prob(spec, np.array([[pin["loc"]]]))[0]
        assert lp == pytest.approx(pin["log_prob_sum"], rel=3e-7)


def test_model_lm_golden(golden_dir):
    """tests/goose/kernels/model_lm.py arrays: logp, gradient and Hessian at the start state."""
    g = load(golden_dir, "model_lm.json")
    X, y = np.array(g["X"]), np.array(g["y"])
    rng = np.random.default_rng(1337)  # the arrays must be regenerable bit for bit
    X2 = np.column_stack([np.ones(30), rng.uniform(size=[30, 1])])
    y2 = rng.normal(X2 @ np.ones(2), 0.1, size=30)
    assert np.array_equal(X, X2) and np.array_equal(y, y2)
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    spec.add_p_smooth(X, 0.0, None, "loc", "beta")
    spec.add_p_smooth(np.ones((30, 1)), 0.0, None, "scale", "log_sigma")
    th = np.array([g["beta"] + [g["log_sigma"]]])
    assert sam.log_prob(spec, th)[0] == pytest.approx(g["log_prob"], rel=1e-12)
    gr = sam.grad(spec, th)[0]
    np.testing.assert_allclose(gr[:2], g["grad_beta"], rtol=1e-11)
    assert gr[2] == pytest.approx(g["grad_log_sigma"], rel=1e-11)
    H = np.array(g["hessian"])
    _, info_b = sam.iwls(spec, "beta", th)
    np.testing.assert_allclose(info_b[0], -H[:2, :2], rtol=1e-11)
    _, info_s = sam.iwls(spec, "log_sigma", th)
    np.testing.assert_allclose(info_s[0], -H[2:, 2:], rtol=1e-11)


# --- third-party densities vs SciPy -----------------------------------------------------
def test_densities_vs_scipy():
    rng = np.random.default_rng(5)
    y = rng.normal(size=50)
    loc, sc = rng.normal(size=50), np.exp(rng.normal(size=50))
    np.testing.assert_allclose(D.normal_log_prob(y, loc, sc), scipy.stats.norm.logpdf(y, loc, sc), rtol=1e-12)
    eta = rng.normal(scale=3, size=50)
    yb = (rng.uniform(size=50) < 0.5).astype(float)
    np.testing.assert_allclose(D.bernoulli_logits_log_prob(yb, eta),
                               scipy.stats.bernoulli.logpmf(yb, 1 / (1 + np.exp(-eta))), rtol=1e-10)
    yp = rng.poisson(3.0, size=50).astype(float)
    np.testing.assert_allclose(D.poisson_log_rate_log_prob(yp, eta / 3),
                               scipy.stats.poisson.logpmf(yp, np.exp(eta / 3)), rtol=1e-10)
    x = np.exp(rng.normal(size=20))
    np.testing.assert_allclose(D.inverse_gamma_log_prob(x, 1.5, 0.7),
                               scipy.stats.invgamma.logpdf(x, 1.5, scale=0.7), rtol=1e-12)


# --- MVND vs the NumPy baseline of tests/distributions/test_mvn_degen.py:27-78 -------------
def _baseline_mvnorm_degen_logpdf(x, mu, precision, log_det, sigma2):
    def determinant_structure_degen(R, tol=1e-6):
        evals = np.linalg.eigvalsh(R)
        mask = evals > tol
        return np.sum(np.log(np.where(mask, evals, 1.0))), mask.sum()

    xmu = x - mu
    t1 = -0.5 * xmu @ precision @ xmu
    if log_det:
        gen_log_det = log_det[0] - log_det[1] * np.log(sigma2)
        rank_prec = log_det[1]
    else:
        gen_log_det, rank_prec = determinant_structure_degen(precision)
    return 0.5 * (gen_log_det - rank_prec * np.log(2.0 * np.pi)) + t1


@pytest.mark.parametrize("k,diff", [(5, 1), (10, 2), (20, 2)])
def test_mvnd_vs_baseline(k, diff):
    rng = np.random.default_rng(k)
    K = OS.pspline_penalty(k, diff)
    beta = rng.normal(size=k)
    for tau2 in [1e-3, 1.0, 10000.0]:
        ours = D.mvnd_from_penalty_log_prob(beta, tau2, K)  # rank/log_pdet from eigenvalues
        rank = float(np.linalg.matrix_rank(K))
        # distreg passes rank but not log_pdet (distreg.py:166-172): top-rank rule
        ours2 = D.mvnd_from_penalty_log_prob(beta, tau2, K, rank=rank)
        assert ours2 == pytest.approx(ours, rel=1e-10)
        evals = np.linalg.eigvalsh(K)
        if evals[evals > 1e-6].min() / tau2 > 1e-5:
            # the baseline thresholds the eigenvalues of the PRECISION at 1e-6, so it only
            # agrees while K/tau2 keeps its non-zero eigenvalues above that tolerance
            base = _baseline_mvnorm_degen_logpdf(beta, np.zeros(k), K / tau2, None, tau2)
            assert ours == pytest.approx(base, rel=1e-10)
        ld = (np.sum(np.log(evals[evals > 1e-6])), rank)
        base3 = _baseline_mvnorm_degen_logpdf(beta, np.zeros(k), K / tau2, ld, tau2)
        assert ours2 == pytest.approx(base3, rel=1e-10)


def test_log_pdet_rules():
    evals = np.array([1e-9, 1e-7, 0.5, 2.0])
    assert D.rank_from_evals(evals) == 2
    assert D.log_pdet_from_evals(evals) == pytest.approx(np.log(0.5) + np.log(2.0))
    # rank given: last `rank` entries, no tolerance test (mvn_degen.py:47-52)
    assert D.log_pdet_from_evals(evals, rank=3) == pytest.approx(np.log(1e-7) + np.log(0.5) + np.log(2.0))


# --- IWLS utilities vs tests/goose/test_iwls_utils.py ---------------------------------------
def test_iwls_utils_golden(golden_dir):
    g = load(golden_dir, "iwls_utils.json")
    lhs, rhs = np.array(g["solve"]["lhs"]), np.array(g["solve"]["rhs"])
    np.testing.assert_allclose(OI.solve(np.linalg.cholesky(lhs), rhs), g["solve"]["x"], rtol=1e-12)
    m = g["mvn_log_prob"]
    cov = np.array(m["cov"])
    lp = OI.mvn_log_prob(np.array(m["x"]), np.array(m["mean"]), np.linalg.cholesky(np.linalg.inv(cov)))
    assert lp == pytest.approx(m["log_prob"], rel=1e-12)


def test_mvn_sample_covariance():
    rng = np.random.default_rng(0)
    n = 5
    cov = rng.uniform(size=(n, n))
    cov = 0.5 * (cov + cov.T) + n * np.identity(n)
    L = np.linalg.cholesky(np.linalg.inv(cov))
    z = rng.normal(size=(200_000, n))
    smp = np.array([OI.mvn_sample(zz, np.zeros(n), L) for zz in z[:20000]])
    np.testing.assert_allclose(np.cov(smp, rowvar=False), cov, atol=0.15)


def test_chol_nan_and_ridge():
    a = np.array([[1.0, 2.0], [2.0, 1.0]])  # indefinite
    assert np.isnan(OI.cholesky_nan(a)).all()
    info = np.array([[[2.0, 0.0], [0.0, 4.0]]])
    np.testing.assert_allclose(OI.add_ridge(info)[0], np.diag([2.0 + 3e-6, 4.0 + 3e-6]))
    chol, code = OI.safe_chol(OI.cholesky_nan(np.stack([a, np.eye(2)])))
    assert code.tolist() == [2, 0] and np.array_equal(chol[0], np.eye(2))


def test_mh_step():
    code, acc, do = OI.mh_step(np.array([0.5, 0.5, 0.5]), np.array([0.0, 0.0, 0.0]),
                               np.array([1.0, -2.0, np.nan]))
    assert code.tolist() == [0, 0, 90]
    np.testing.assert_allclose(acc, [1.0, np.exp(-2.0), 0.0])
    assert do.tolist() == [True, False, False]


# --- splines vs tests/contrib/test_splines.py ------------------------------------------------
def test_knots():
    x = np.arange(-1000, 1000).astype(np.float64)
    for nparam in [10, 100, 2000]:
        for order in [0, 1, 2, 3, 12]:
            if order == 0:
                continue  # np.linspace(num=0) edge of the reference's own loop is not on our path
            if nparam < order:
                with pytest.raises(ValueError):
                    OS.equidistant_knots(x, nparam, order)
                continue
            knots = OS.equidistant_knots(x, nparam, order)
            assert len(knots) == nparam + order + 1
            d = np.diff(knots)
            assert np.allclose(d, d[0], atol=1e-3)
    with pytest.raises(ValueError):
        OS.equidistant_knots(x, 2, 3)
    with pytest.raises(ValueError):
        OS.equidistant_knots(x, 10, -1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_basis_vs_scipy(dtype):
    x = np.arange(0, 200).astype(dtype)
    knots = OS.equidistant_knots(x, 40, 3)
    X = OS.basis_matrix(x, knots, 3)
    assert X.shape == (200, 40)
    beta = np.random.default_rng(1).normal(size=40)
    ref = BSpline(knots.astype(np.float64), beta, 3)(x.astype(np.float64))
    tol = 1e-3 if dtype == np.float32 else 1e-10
    assert np.allclose(X.astype(np.float64) @ beta, ref, tol, tol)  # test_splines.py:48-59
    dm = BSpline.design_matrix(x.astype(np.float64), knots.astype(np.float64), 3, extrapolate=True).toarray()
    np.testing.assert_allclose(X, dm[:, :40], atol=2e-5 if dtype == np.float32 else 1e-12)
    start, w = OS.banded_basis(x, knots, 3)
    assert start.dtype == np.int32 and w.shape == (200, 4)
    np.testing.assert_array_equal(OS.band_to_dense(start, w, 40), X)
    np.testing.assert_allclose(w.sum(axis=1), 1.0, rtol=1e-5)  # partition of unity inside the range


def test_basis_outside_range():
    x = np.arange(0, 200).astype(np.float64)
    knots = OS.equidistant_knots(x, 40, 3)
    B = OS.basis_matrix(np.array([-1.0, knots[0] - 1, knots[-1] + 1]), knots, 3)
    assert B.shape == (3, 40)
    assert np.all(B[1] == 0) and np.all(B[2] == 0)
    start, w = OS.banded_basis(np.array([-1.0, knots[-1] + 1]), knots, 3)
    np.testing.assert_array_equal(OS.band_to_dense(start, w, 40), B[[0, 2]])


def test_penalty_rank():
    for diff in [1, 2, 3]:
        assert np.linalg.matrix_rank(OS.pspline_penalty(10, diff)) == 10 - diff
    assert OS.pspline_penalty(10).shape == (10, 10)


# --- finite differences on the model family (parity unpinned by the reference) --------------
def _fd_grad(spec, th, aux, prep, eps=1e-6):
    g = np.zeros_like(th)
    for j in range(th.shape[1]):
        tp, tm = th.copy(), th.copy()
        tp[:, j] += eps
        tm[:, j] -= eps
        g[:, j] = (sam.log_prob(spec, tp, aux, prep=prep) - sam.log_prob(spec, tm, aux, prep=prep)) / (2 * eps)
    return g


@pytest.mark.parametrize("cfg", ["s1", "s2", "s3", "s4", "s5"])
def test_oracle_grad_and_info_fd(cfg):
    make = {
        "s1": lambda: synth.s1_blr(n=200, dtype=np.float64),
        "s2": lambda: synth.s2_pspline_gam(n=300, n_smooth=2, k=8, dtype=np.float64),
        "s3": lambda: synth.s3_gamlss(n=400, k=8, dtype=np.float64),
        "s4": lambda: synth.s4_logistic(n=300, p=6, dtype=np.float64),
        "s5": lambda: synth.s5_poisson_ri(n=300, G=10, p=3, dtype=np.float64),
    }[cfg]
    spec, centre = make()
    th, aux = synth.evaluation_points(spec, centre, 2, tau2="loguniform")
    th, aux = th.astype(np.float64), aux.astype(np.float64)
    prep = sam.Prepared(spec)
    g = sam.grad(spec, th, aux, prep=prep)
    fd = _fd_grad(spec, th, aux, prep)
    np.testing.assert_allclose(g, fd, rtol=2e-5, atol=2e-5 * np.abs(g).max())
    # log_prob == log_prior + log_lik (model.py:935-943 invariant)
    parts = sam.log_prob_parts(spec, th, aux, prep=prep)
    total = sum(parts.values()) + spec.const_term
    np.testing.assert_allclose(total, sam.log_prob(spec, th, aux, prep=prep), rtol=1e-14)
    # information = -d grad / d beta_block
    for blk in spec.blocks:
        if blk.kind == 2:
            continue
        _, info = sam.iwls(spec, blk.name, th, aux, prep=prep)
        sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
        for j in range(blk.ncoef):
            tp, tm = th.copy(), th.copy()
            tp[:, blk.param_offset + j] += 1e-6
            tm[:, blk.param_offset + j] -= 1e-6
            col = -(sam.grad(spec, tp, aux, prep=prep) - sam.grad(spec, tm, aux, prep=prep))[:, sl] / 2e-6
            np.testing.assert_allclose(info[:, :, j], col, rtol=1e-4, atol=1e-5 * np.abs(info).max())
        score, _ = sam.iwls(spec, blk.name, th, aux, prep=prep)
        np.testing.assert_allclose(score, g[:, sl], rtol=1e-12, atol=1e-12)


def test_tau2_gibbs_params():
    spec, centre = synth.s2_pspline_gam(n=100, n_smooth=1, k=8, dtype=np.float64)
    th, _ = synth.evaluation_points(spec, centre, 3)
    a, b = sam.tau2_gibbs_params(spec, "loc_np0", th)
    K = spec.blocks[0].prior.K
    assert a == pytest.approx(1.0 + 0.5 * 6)
    np.testing.assert_allclose(b, 0.005 + 0.5 * np.einsum("ci,ij,cj->c", th, K, th))


def test_centred_spline_equals_dense_centred_design():
    """
    The oracle's closed form for a column-centred basis (band + rank-one term) equals the plain
    dense path ``jnp.dot(X, beta)`` (distreg.py:175) with X = B - 1 cbar^T, for value, gradient
    and the IWLS quantities.
    """
    from liesel_b200.spec import ModelSpec
    from liesel_b200.splines import equidistant_knots, pspline_penalty

    rng = np.random.default_rng(5)
    n, k = 400, 9
    x = rng.uniform(size=n)
    y = np.sin(3 * x) + rng.normal(scale=0.3, size=n)
    knots = equidistant_knots(x, k, 3)
    K = pspline_penalty(k, 2)
    start, w = OS.banded_basis(x, knots, 3)
    B = OS.band_to_dense(start, w, k)
    cbar = B.mean(axis=0)

    def build(dense):
        spec = ModelSpec(dtype=np.dtype(np.float64))
        spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
        spec.add_p_smooth(np.ones((n, 1)), 0.0, 10.0, "loc", "loc_p0")
        if dense:
            spec.add_np_smooth(B - cbar[None, :], K, 1.0, 0.005, "loc", "loc_np0")
        else:
            spec.add_spline_smooth(x, knots, K, 1.0, 0.005, "loc", "loc_np0", start=start,
                                   weights=w, center=cbar)
        spec.add_p_smooth(np.ones((n, 1)), 0.0, 10.0, "scale", "scale_p0")
        return spec

    a, b = build(False), build(True)
    th = rng.normal(scale=0.3, size=(3, a.num_params))
    aux = np.exp(rng.normal(size=(3, a.num_aux)))
    np.testing.assert_allclose(sam.log_prob(a, th, aux), sam.log_prob(b, th, aux), rtol=1e-12)
    np.testing.assert_allclose(sam.grad(a, th, aux), sam.grad(b, th, aux), rtol=1e-9, atol=1e-9)
    sa, ia = sam.iwls(a, "loc_np0", th, aux)
    sb, ib = sam.iwls(b, "loc_np0", th, aux)
    np.testing.assert_allclose(sa, sb, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ia, ib, rtol=1e-9, atol=1e-9)
    # centred columns sum to zero over the rows: the constant direction is not identified
    assert np.abs((B - cbar).sum(axis=0)).max() < 1e-10


# --- the reference's model-level pin behind jax.random (tests/goose/test_interface_liesel.py) ----
def _interface_liesel_case(golden_dir):
    fx = load(golden_dir, "interface_liesel.json")
    x = np.frombuffer(bytes.fromhex("".join(fx["x_float32_hex"])), dtype=np.float32).copy()
    spec = ModelSpec(dtype=np.float32)
    spec.add_response(x, "normal").add_predictor("loc")          # y = x ~ Normal(mu, sigma = 1)
    spec.add_p_smooth(np.ones((x.size, 1), np.float32), 0.0, None, "loc", "mu")  # mu: a Var without prior
    spec.fixed_scale = 1.0
    return fx, x, spec


def test_threefry_known_answer_and_fixture_regenerates(golden_dir):
    from oracle import jax_prng

    a, b = jax_prng.threefry2x32(0x13198A2E, 0x03707344, np.array([0x243F6A88], np.uint32),
                                 np.array([0x85A308D3], np.uint32))
    assert (int(a[0]), int(b[0])) == (0xC4923A9C, 0x483DF7A0)  # Random123 known-answer vector
    fx, x, _ = _interface_liesel_case(golden_dir)
    np.testing.assert_array_equal(jax_prng.normal(jax_prng.prng_key(1337), 500), x)


def test_interface_liesel_log_prob_pins(golden_dir):
    """The oracle reproduces the two log-prob values the reference's own test asserts, on the inputs
    that test creates with jax.random (same tolerance as there: pytest.approx, rel 1e-6)."""
    fx, x, spec = _interface_liesel_case(golden_dir)
    lp = sam.log_prob(spec, np.array([[0.0], [10.0]]))
    assert lp[0] == pytest.approx(fx["pins"]["log_prob_mu_0"])
    assert lp[1] == pytest.approx(fx["pins"]["log_prob_mu_10"])
    g = sam.grad(spec, np.array([[0.0], [10.0]]))
    np.testing.assert_allclose(g[:, 0], [x.astype(np.float64).sum(), x.astype(np.float64).sum() - 5000.0], rtol=1e-12)


# --- tutorial 01b: a whole model's log-prob, log-prior and log-lik as printed by the reference ----
def tutorial_01b_spec(pin, sigma_sq, dtype=np.float32):
    rng = np.random.default_rng(42)
    n = pin["data"]["n"]
    x0 = rng.uniform(size=n)
    X = np.column_stack([np.ones(n), x0])
    eps = rng.normal(scale=pin["data"]["true_sigma"], size=n)
    y = X @ np.asarray(pin["data"]["true_beta"]) + eps
    spec = ModelSpec(dtype=np.dtype(dtype))
    spec.add_response(y, "normal").add_predictor("loc")
    spec.add_p_smooth(X, 0.0, 100.0, "loc", "beta")
    spec.fixed_scale = float(np.sqrt(sigma_sq))               # sigma = sqrt(sigma_sq), a Calc in the tutorial
    spec.const_term = float(D.inverse_gamma_log_prob(sigma_sq, 0.01, 0.01))  # sigma_sq ~ InverseGamma(a, b)
    return spec


def test_tutorial_01b_model_log_prob_pins(golden_dir):
    """docs/source/tutorials/md/01b-model.md prints model.log_prob / log_prior / log_lik and the
    InverseGamma log-prob of sigma_sq for data from numpy's default_rng(42): float32 printouts,
    compared at 7 significant digits."""
    pin = load(golden_dir, "reference_pins.json")["tutorial_01b_model"]
    th = np.array([pin["at_beta"]])
    for key, s2 in (("sigma_sq_10", 10.0), ("sigma_sq_1", 1.0)):
        spec = tutorial_01b_spec(pin, s2)
        want = pin[key]
        parts = sam.log_prob_parts(spec, th)
        assert sam.log_prob(spec, th)[0] == pytest.approx(want["model_log_prob"], rel=3e-7)
        assert parts["response"][0] == pytest.approx(want["y_log_prob_sum"], rel=3e-7)
        assert D.inverse_gamma_log_prob(s2, 0.01, 0.01) == pytest.approx(want["sigma_sq_log_prob"], rel=1e-6)
        if "log_prior" in want:
            assert sam.log_prob(spec, th)[0] - parts["response"][0] == pytest.approx(want["log_prior"], rel=3e-7)
            assert parts["response"][0] == pytest.approx(want["log_lik"], rel=3e-7)


def test_group_block_information_is_diagonal_and_shares_add_up():
    """The random-intercept block: the dense information `sam.iwls` forms (what the reference's
    `-jacfwd(grad(log_prob))` gives, goose/iwls.py:218-243) is diagonal and equals `sam.iwls_group_diag`;
    the score is the log-posterior's gradient (finite differences); moving ONE group's effect changes the
    log-posterior by exactly the change of that group's share."""
    spec, centre = synth.s5_poisson_ri(n=3000, G=40, dtype=np.float64, balanced=False)
    th, aux = synth.evaluation_points(spec, centre, 2)
    name = "log_rate_ri0"
    blk = spec.block(name)
    s_dense, i_dense = sam.iwls(spec, name, th, aux)
    s_diag, i_diag, share = sam.iwls_group_diag(spec, name, th, aux)
    assert np.array_equal(s_dense, s_diag)
    assert np.array_equal(np.diagonal(i_dense, axis1=1, axis2=2), i_diag)
    assert np.all(i_dense - i_diag[:, :, None] * np.eye(blk.ncoef) == 0)
    g = 7
    h = 1e-5
    tp, tm = th.copy(), th.copy()
    tp[:, blk.param_offset + g] += h
    tm[:, blk.param_offset + g] -= h
    lp_p, lp_m = sam.log_prob(spec, tp, aux), sam.log_prob(spec, tm, aux)
    np.testing.assert_allclose((lp_p - lp_m) / (2 * h), s_diag[:, g], rtol=1e-6)
    lp0 = sam.log_prob(spec, th, aux)
    h2 = 1e-3  # (a wider step for the second difference: rounding of the log-posterior / h^2)
    tp2, tm2 = th.copy(), th.copy()
    tp2[:, blk.param_offset + g] += h2
    tm2[:, blk.param_offset + g] -= h2
    second = -(sam.log_prob(spec, tp2, aux) - 2 * lp0 + sam.log_prob(spec, tm2, aux)) / h2**2
    np.testing.assert_allclose(second, i_diag[:, g], rtol=1e-4)
    _, _, share_p = sam.iwls_group_diag(spec, name, tp, aux)
    np.testing.assert_allclose(share_p[:, g] - share[:, g], lp_p - lp0, rtol=1e-8, atol=1e-9)
    mask = np.ones(blk.ncoef, dtype=bool)
    mask[g] = False
    assert np.array_equal(share_p[:, mask], share[:, mask])
