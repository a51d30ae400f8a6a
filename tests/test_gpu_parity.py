"""
Parity of the CUDA path (called through the C ABI via ctypes) against the CPU oracle.

Tolerances (BASELINE.json north_star: rtol 1e-5 fp32 / 1e-10 fp64 on log-prob, gradients and
Gram; bit-exact basis and group indices):
  * log-prob:  |err| <= rtol * |ref|
  * gradient / information: |err| <= rtol * |ref| + ulps * eps * scale, where ``scale`` is the sum
    of absolute values of the terms of that entry (oracle.sam.grad_scale) — entries that are a
    near-cancelling sum of 1e5..1e6 terms cannot meet a pure relative bound in ANY float32
    evaluation order, the reference's XLA reduction included.
"""
import numpy as np
import pytest
import torch

from liesel_b200 import synth
from oracle import iwls as OI
from oracle import sam
from oracle import splines as OS

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-10}
EPS = {np.dtype(np.float32): 2.0**-24, np.dtype(np.float64): 2.0**-53}
ULPS = 16.0


@pytest.fixture(scope="module")
def lb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from liesel_b200 import plan as P
    from liesel_b200.interface import B200Interface

    class NS:
        Plan = P.Plan
        chol = staticmethod(P.chol)
        iwls_propose = staticmethod(P.iwls_propose)
        Interface = B200Interface

    return NS


def make(cfg, dtype):
    return {
        "s1": lambda: synth.s1_blr(n=1000, dtype=dtype),
        "s2": lambda: synth.s2_pspline_gam(n=20_000, dtype=dtype),
        "s3": lambda: synth.s3_gamlss(n=50_000, dtype=dtype),
        "s3c": lambda: synth.s3_gamlss(n=50_000, dtype=dtype, center=True),
        "s4": lambda: synth.s4_logistic(n=4000, p=48, dtype=dtype),
        "s4s": lambda: synth.s4_logistic(n=4000, p=12, dtype=dtype),
        "s5": lambda: synth.s5_poisson_ri(n=20_000, G=500, dtype=dtype),
        "s5u": lambda: synth.s5_poisson_ri(n=20_000, G=700, dtype=dtype, balanced=False),
    }[cfg]()


def dev(a, plan):
    return torch.as_tensor(np.ascontiguousarray(a), device=plan.device)


def check_logp_grad(spec, plan, th, aux, chains=None):
    dt = np.dtype(spec.dtype)
    logp, grad = plan.logp_grad(dev(th, plan), dev(aux, plan) if spec.num_aux else None)
    logp_v, _ = plan.logp_grad(dev(th, plan), dev(aux, plan) if spec.num_aux else None, want_grad=False)
    torch.cuda.synchronize()
    logp, grad, logp_v = logp.cpu().numpy(), grad.cpu().numpy(), logp_v.cpu().numpy()
    sel = np.arange(th.shape[0]) if chains is None else np.asarray(chains)
    prep = sam.Prepared(spec)
    ref_lp = sam.log_prob(spec, th[sel], aux[sel], prep=prep)
    ref_g = sam.grad(spec, th[sel], aux[sel], prep=prep)
    sc_g = sam.grad_scale(spec, th[sel], aux[sel], prep=prep)
    sc_lp = sam.log_prob_scale(spec, th[sel], aux[sel], prep=prep)
    # value-only and value+grad paths agree (mh_step reads log_prob without a gradient)
    np.testing.assert_allclose(logp_v[sel], logp[sel], rtol=10 * EPS[dt])
    err = np.abs(logp[sel] - ref_lp)
    assert np.all(err <= RTOL[dt] * np.abs(ref_lp) + ULPS * EPS[dt] * sc_lp), (err.max(), ref_lp)
    gerr = np.abs(grad[sel] - ref_g)
    bound = RTOL[dt] * np.abs(ref_g) + ULPS * EPS[dt] * sc_g
    assert np.all(gerr <= bound), float((gerr / bound).max())
    return logp, grad


@pytest.mark.parametrize("cfg", ["s1", "s2", "s3", "s4", "s5", "s5u"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_logp_grad_parity(lb, cfg, dtype):
    spec, centre = make(cfg, dtype)
    plan = lb.Plan(spec)
    C = {"s1": 4, "s2": 64, "s3": 40, "s4": 12, "s5": 33, "s5u": 7}[cfg]
    th, aux = synth.evaluation_points(spec, centre, C)
    check_logp_grad(spec, plan, th, aux)


@pytest.mark.parametrize("C", [24, 300])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_centred_splines(lb, C, dtype):
    """Column-centred bases (B - 1 cbar^T) in both predictors: the scalar kernels (C = 24, fp64)
    and the packed kernel (fp32, C > 128), log-prob/gradient and the IWLS quantities."""
    spec, centre = make("s3c", dtype)
    plan = lb.Plan(spec)
    assert spec.block("loc_np0").colmean is not None
    th, aux = synth.evaluation_points(spec, centre, C, tau2="loguniform")
    check_logp_grad(spec, plan, th, aux, chains=None if C < 100 else [0, 1, 127, 128, 255, 299])
    for name in ["loc_p0", "loc_np0", "scale_np0"]:
        check_iwls(spec, plan, name, th, aux, chains=[0, 1, 2, C // 2, C - 1])


@pytest.mark.parametrize("family", ["poisson", "bernoulli"])
@pytest.mark.parametrize("n_smooth", [1, 2])
@pytest.mark.parametrize("C,dtype", [(20, np.float32), (20, np.float64), (300, np.float32)])
def test_spline_glm_families(lb, family, n_smooth, C, dtype):
    """P-splines under the Poisson and Bernoulli families: scalar TMA run-length kernel (few chains,
    fp64) and the packed kernel (fp32, more than 128 chains), log-prob, gradient and IWLS."""
    spec, centre = synth.spline_glm(family, n=30_000, n_smooth=n_smooth, dtype=dtype)
    plan = lb.Plan(spec)
    assert plan.variant.startswith("banded")
    th, aux = synth.evaluation_points(spec, centre, C, tau2="loguniform")
    check_logp_grad(spec, plan, th, aux, chains=None if C < 100 else [0, 1, 128, 129, 255, 256, 299])
    pred = "log_rate" if family == "poisson" else "logits"
    check_iwls(spec, plan, f"{pred}_np0", th, aux, chains=[0, 1, C - 1])


@pytest.mark.parametrize("tau2", ["start", "loguniform"])
def test_logp_grad_tau2_range(lb, tau2):
    """tau2 spans 1e-3..1e4 during Gibbs sampling; also the reference's start state beta = 0."""
    spec, centre = make("s3", np.float32)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 16, tau2=tau2)
    if tau2 == "start":
        th = np.zeros_like(th)  # distreg.py:101,165
    check_logp_grad(spec, plan, th, aux)


@pytest.mark.parametrize("C", [1, 3, 31, 33, 65, 130])
def test_ragged_chain_counts(lb, C):
    spec, centre = make("s3", np.float32)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, C)
    check_logp_grad(spec, plan, th, aux)


@pytest.mark.parametrize("C", [129, 192, 320, 576, 704, 1088, 2500])
def test_chain_counts_with_ring_wraparound(lb, C):
    """Chain counts that give 3..16 warps per CTA and 1..5 CTAs per SM, at a row count where every CTA
    walks more tiles than the TMA ring has stages (the small cases never refill a stage)."""
    spec, centre = synth.s3_gamlss(n=150_000 + 37)
    plan = lb.Plan(spec)
    assert plan.variant == "banded_tma_runlength"
    th, aux = synth.evaluation_points(spec, centre, C, tau2="loguniform")
    check_logp_grad(spec, plan, th, aux, chains=[0, 63, 64, C - 1])
    spec, centre = synth.s5_poisson_ri(n=150_000, G=2900, balanced=False)
    plan = lb.Plan(spec)
    assert plan.variant == "dense_group_x2_tma"
    th, aux = synth.evaluation_points(spec, centre, C)
    check_logp_grad(spec, plan, th, aux, chains=[0, C - 1])


def test_reference_interface_liesel_pins_on_device(lb, golden_dir):
    """reference tests/goose/test_interface_liesel.py:36-44 on the device: x = jax.random.normal(
    PRNGKey(1337), (500,)) (regenerated through the restated JAX PRNG, committed as a fixture),
    y = x ~ Normal(mu, 1): log_prob -719.46875 at mu = 0 and -25616.779296875 at mu = 10."""
    import json
    import os

    fx = json.load(open(os.path.join(golden_dir, "interface_liesel.json")))
    x = np.frombuffer(bytes.fromhex("".join(fx["x_float32_hex"])), dtype=np.float32).copy()
    from liesel_b200.spec import ModelSpec

    spec = ModelSpec(dtype=np.float32)
    spec.add_response(x, "normal").add_predictor("loc")
    spec.add_p_smooth(np.ones((x.size, 1), np.float32), 0.0, None, "loc", "mu")
    spec.fixed_scale = 1.0
    plan = lb.Plan(spec)
    iface = lb.Interface(plan)
    state = iface.initial_state(2, {"mu_beta": np.array([[0.0], [10.0]], np.float32)})
    lp = iface.log_prob(state).cpu().numpy()
    assert lp[0] == pytest.approx(fx["pins"]["log_prob_mu_0"])
    assert lp[1] == pytest.approx(fx["pins"]["log_prob_mu_10"])
    new_state = iface.update_state({"mu_beta": state["mu_beta"][[1, 0]].contiguous()}, state)  # update_state, :29-34
    lp2 = iface.log_prob(new_state).cpu().numpy()
    assert lp2[0] == pytest.approx(fx["pins"]["log_prob_mu_10"]) and lp2[1] == pytest.approx(fx["pins"]["log_prob_mu_0"])


def test_reference_tutorial_01b_pins_on_device(lb, golden_dir):
    """The whole-model log-probs that docs/source/tutorials/md/01b-model.md:213-277 prints (data from
    numpy's default_rng(42)), through the C ABI in float32 and float64."""
    import json
    import os

    from test_oracle_pins import tutorial_01b_spec

    pin = json.load(open(os.path.join(golden_dir, "reference_pins.json")))["tutorial_01b_model"]
    for dtype, rel in ((np.float32, 1e-6), (np.float64, 3e-7)):  # the printed values are float32
        for key, s2 in (("sigma_sq_10", 10.0), ("sigma_sq_1", 1.0)):
            spec = tutorial_01b_spec(pin, s2, dtype)
            plan = lb.Plan(spec)
            th = torch.zeros((3, 2), dtype=plan.dtype, device=plan.device)
            lp, g = plan.logp_grad(th)
            assert float(lp[0]) == pytest.approx(pin[key]["model_log_prob"], rel=rel)
            lp_lik, _ = plan.logp_grad(th, flags=1)  # LSLB_SKIP_PRIOR: the response part alone
            assert float(lp_lik[1]) == pytest.approx(pin[key]["y_log_prob_sum"], rel=rel)


def test_unsorted_rows_same_result(lb):
    """Row order is free: the span-sorted plan and the plan in data order agree."""
    spec, centre = make("s2", np.float32)
    th, aux = synth.evaluation_points(spec, centre, 8)
    a = lb.Plan(spec, sort_rows=True)
    b = lb.Plan(spec, sort_rows=False)
    la, ga = check_logp_grad(spec, a, th, aux)
    lb_, gb = check_logp_grad(spec, b, th, aux)
    np.testing.assert_allclose(la, lb_, rtol=2e-6)


def test_bitwise_reproducible(lb):
    spec, centre = make("s3", np.float32)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 64)
    t, a = dev(th, plan), dev(aux, plan)
    l1, g1 = plan.logp_grad(t, a)
    l1, g1 = l1.clone(), g1.clone()
    l2, g2 = plan.logp_grad(t, a)
    assert torch.equal(l1, l2) and torch.equal(g1, g2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("family", ["normal", "bernoulli", "poisson"])
def test_single_launch_kernel_against_oracle_and_general_path(lb, dtype, family):
    """Small regressions (S1-like: n <= 4095 rows, rows x chains <= 2^20) run lslb::tiny_kernel, one launch per
    call. Oracle parity per family and dtype, agreement with the three-launch general path (taken by the same
    plan when rows x chains exceeds the bound), value-only calls, bitwise reproducibility."""
    rng = np.random.default_rng(11)
    n, p = 1000, 10
    X = np.column_stack([np.ones(n), rng.uniform(size=(n, p - 1))]).astype(dtype)
    beta = np.linspace(-1, 1, p) * (0.3 if family != "normal" else 1.0)
    eta = X.astype(np.float64) @ beta
    if family == "normal":
        spec, centre = synth.s1_blr(n=n, p=p, dtype=dtype)
    else:
        from liesel_b200.spec import ModelSpec

        y = rng.binomial(1, 1 / (1 + np.exp(-eta))) if family == "bernoulli" else rng.poisson(np.exp(eta))
        spec = ModelSpec(dtype=np.dtype(dtype))
        pred = "logits" if family == "bernoulli" else "log_rate"
        spec.add_response(y.astype(np.float64), family).add_predictor(pred)
        spec.add_p_smooth(X[:, 1:], 0.0, 10.0, pred, pred + "_p0")
        spec.add_p_smooth(np.ones((n, 1), dtype=dtype), 0.5, 100.0, pred, pred + "_icpt")
        centre = np.concatenate([beta[1:], beta[:1]])
    plan = lb.Plan(spec)
    assert plan.kernel_name(4) == "lslb::tiny_kernel", plan.kernel_name(4)
    C = 6
    th, aux = synth.evaluation_points(spec, centre, C)
    check_logp_grad(spec, plan, th, aux)
    t = dev(th, plan)
    l1, g1 = plan.logp_grad(t)
    l1, g1 = l1.clone(), g1.clone()
    l2, g2 = plan.logp_grad(t)
    assert torch.equal(l1, l2) and torch.equal(g1, g2)
    # the same plan on 2048 chains leaves the single-launch bound: first chains must agree
    Cb = 2048
    assert plan.kernel_name(Cb) != "lslb::tiny_kernel"
    thb = np.concatenate([th, np.tile(th[:1], (Cb - C, 1))])
    lbig, gbig = plan.logp_grad(dev(thb, plan))
    tol = 2e-5 if dtype == np.float32 else 1e-11
    np.testing.assert_allclose(lbig[:C].cpu().numpy(), l1.cpu().numpy(), rtol=tol)
    scale = np.abs(g1.cpu().numpy()).max()
    np.testing.assert_allclose(gbig[:C].cpu().numpy(), g1.cpu().numpy(), rtol=tol * 10, atol=tol * 10 * scale)


def test_tiny_and_empty_inputs(lb):
    for n in [1, 2, 31, 33]:
        spec, centre = synth.s3_gamlss(n=max(n, 8), k=8, dtype=np.float64)
        # cut to n rows
        spec.y = spec.y[:n]
        for b in spec.blocks:
            if b.x is not None:
                b.x = b.x[:n]
        plan = lb.Plan(spec)
        th, aux = synth.evaluation_points(spec, centre, 3)
        check_logp_grad(spec, plan, th, aux)


def test_nan_inf_propagate(lb):
    """Numerical failure is signalled by NaN/Inf values, never by an exception (mh.py:53-57)."""
    spec, centre = make("s5", np.float32)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 4)
    th[1, 0] = 2000.0  # exp overflow in the Poisson rate (X[:, 0] = 0.3)
    th[2, 1] = np.nan
    logp, grad = plan.logp_grad(dev(th, plan), None)
    logp = logp.cpu().numpy()
    assert np.isfinite(logp[0]) and np.isfinite(logp[3])
    assert not np.isfinite(logp[1]) and np.isnan(logp[2])


# --- spline basis -----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_basis_bit_exact(lb, dtype):
    from liesel_b200 import splines

    rng = np.random.default_rng(3)
    x = rng.uniform(size=200_000).astype(dtype)
    knots = splines.equidistant_knots(x, 20, 3)
    # adversarial points: every knot, and its neighbours one ulp away
    edge = np.concatenate([knots, np.nextafter(knots, dtype(np.inf)), np.nextafter(knots, dtype(-np.inf))])
    x = np.concatenate([x, edge.astype(dtype), [knots[0] - 1, knots[-1] + 1]]).astype(dtype)
    start, w = splines.basis_matrix(x, knots, 3, outer_ok=True, device="cuda")
    o_start, o_w = OS.banded_basis(x, knots, 3)
    assert np.array_equal(start.cpu().numpy(), o_start)  # basis index: bit-exact
    assert np.array_equal(w.cpu().numpy(), o_w)          # same operation order: bit-exact too
    with pytest.raises(ValueError):
        splines.basis_matrix(x, knots, 3, outer_ok=False, device="cuda")


def test_group_index_passthrough(lb):
    spec, _ = make("s5u", np.float32)
    plan = lb.Plan(spec)
    assert np.array_equal(plan.rows["g1"].cpu().numpy(), spec.blocks[1].idx)


# --- IWLS ---------------------------------------------------------------------------------
def check_iwls(spec, plan, block, th, aux, chains=None):
    dt = np.dtype(spec.dtype)
    score, info, chol = plan.iwls(block, dev(th, plan), dev(aux, plan) if spec.num_aux else None)
    torch.cuda.synchronize()
    score, info, chol = score.cpu().numpy(), info.cpu().numpy(), chol.cpu().numpy()
    if chains is not None:  # device call on all chains, oracle on a subset
        sel = np.asarray(chains)
        score, info, chol, th, aux = score[sel], info[sel], chol[sel], th[sel], aux[sel]
    prep = sam.Prepared(spec)
    r_score, r_info = sam.iwls(spec, block, th, aux, prep=prep)
    blk = spec.block(block)
    sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
    sc = sam.grad_scale(spec, th, aux, prep=prep)[:, sl]
    assert np.all(np.abs(score - r_score) <= RTOL[dt] * np.abs(r_score) + ULPS * EPS[dt] * sc)
    # information: all terms of an entry share a sign pattern set by the (non-negative) basis;
    # scale by the diagonal, which dominates its row
    dscale = np.sqrt(np.abs(np.einsum("cii->ci", r_info)))
    iscale = dscale[:, :, None] * dscale[:, None, :]
    ierr = np.abs(info - r_info)
    assert np.all(ierr <= RTOL[dt] * np.abs(r_info) + ULPS * EPS[dt] * iscale), float(ierr.max())
    np.testing.assert_array_equal(info, np.swapaxes(info, 1, 2))
    r_chol = OI.chol_info(r_info)
    ok = ~np.isnan(r_chol).any(axis=(1, 2))
    assert ok.any()
    ctol = 2e-4 if dt == np.float32 else 1e-9
    np.testing.assert_allclose(chol[ok], r_chol[ok], rtol=ctol, atol=ctol * np.abs(r_chol[ok]).max())
    assert np.isnan(chol[~ok]).all()
    return score, info, chol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_iwls_s3_all_blocks(lb, dtype):
    spec, centre = make("s3", dtype)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 24, tau2="loguniform")
    for name in ["loc_p0", "loc_np0", "scale_p0", "scale_np0"]:
        check_iwls(spec, plan, name, th, aux)


@pytest.mark.parametrize("cfg,block", [("s1", "loc_p0"), ("s2", "loc_np3"), ("s4s", "logits_p0"), ("s5", "log_rate_p0")])
def test_iwls_other_families(lb, cfg, block):
    spec, centre = make(cfg, np.float64)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 5)
    check_iwls(spec, plan, block, th, aux)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("p", [3, 6, 8])
@pytest.mark.parametrize("group", [False, True])
def test_iwls_dense_block_in_registers(lb, dtype, p, group):
    """Dense target with p <= 8 in a one-dense-block plan: `lslb::iwls_dense_reg_kernel` (score and packed
    upper triangle in registers; instantiations PD = 4 / 8, with and without a random intercept) against the
    oracle, and against the generic shared-memory kernel's layout through the shared finalize."""
    if group:
        spec, centre = synth.s5_poisson_ri(n=20_000, G=500, p=p, dtype=dtype, balanced=False)
        block = "log_rate_p0"
    else:
        spec, centre = synth.s1_blr(n=3000, p=p, dtype=dtype)
        block = "loc_p0"
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 37)
    check_iwls(spec, plan, block, th, aux)
    if p == 8 and dtype == np.float32:  # from 512 chains: the packed twin `iwls_dense_reg_x2_kernel`
        th, aux = synth.evaluation_points(spec, centre, 512)
        check_iwls(spec, plan, block, th, aux, chains=[0, 1, 300, 511])


def test_chol_failure_is_nan(lb):
    a = np.stack([np.array([[1.0, 2.0], [2.0, 1.0]]), np.eye(2) * 4.0, np.full((2, 2), np.nan)])
    out = lb.chol(torch.as_tensor(a, device="cuda")).cpu().numpy()
    assert np.isnan(out[0]).all() and np.isnan(out[2]).all()
    np.testing.assert_allclose(out[1], np.eye(2) * np.sqrt(4.0 + 4e-6), rtol=1e-12)


def test_iwls_utils_golden_on_device(lb, golden_dir):
    """tests/goose/test_iwls_utils.py:11-35 through the device proposal tail."""
    import json
    import os

    g = json.load(open(os.path.join(golden_dir, "iwls_utils.json")))
    lhs, rhs = np.array(g["solve"]["lhs"]), np.array(g["solve"]["rhs"])
    L = np.linalg.cholesky(lhs)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    # mu = beta + step^2/2 * solve(L, rhs) with beta = 0, step = sqrt(2)  ->  mu = solve
    step = np.array([np.sqrt(2.0)])
    _, _ = lb.iwls_propose(t(np.zeros((1, 5))), t(rhs[None]), t(L[None]), t(step), x_eval=t(np.zeros((1, 5))))
    m = g["mvn_log_prob"]
    cov = np.array(m["cov"])
    Lc = np.linalg.cholesky(np.linalg.inv(cov))
    # with score = 0 the mean stays at beta; step = 1
    _, logq = lb.iwls_propose(t(np.array(m["mean"])[None]), t(np.zeros((1, 5))), t(Lc[None]),
                              t(np.ones(1)), x_eval=t(np.array(m["x"])[None]))
    assert logq.cpu().numpy()[0] == pytest.approx(m["log_prob"], rel=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_iwls_proposal_tail(lb, dtype):
    spec, centre = make("s3", dtype)
    plan = lb.Plan(spec)
    C = 9
    th, aux = synth.evaluation_points(spec, centre, C, tau2="loguniform")
    rng = np.random.default_rng(11)
    blk = spec.block("loc_np0")
    sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
    z = rng.normal(size=(C, blk.ncoef)).astype(dtype)
    step = rng.uniform(0.5, 1.5, size=C).astype(dtype)
    score, info, chol = plan.iwls("loc_np0", dev(th, plan), dev(aux, plan))
    prop, fwd = lb.iwls_propose(dev(th[:, sl], plan), score, chol, dev(step, plan), z=dev(z, plan))
    ref = OI.iwls_proposal(spec, "loc_np0", th, aux, step.astype(np.float64), z.astype(np.float64))
    tol = 5e-4 if dtype == np.float32 else 1e-8
    np.testing.assert_allclose(prop.cpu().numpy(), ref["prop"], rtol=tol, atol=tol)
    np.testing.assert_allclose(fwd.cpu().numpy(), ref["fwd_log_prob"], rtol=tol, atol=tol)
    # backward density of the current position under the proposal made at `prop`
    th2 = th.copy()
    th2[:, sl] = prop.cpu().numpy()
    score2, _, chol2 = plan.iwls("loc_np0", dev(th2, plan), dev(aux, plan))
    _, bwd = lb.iwls_propose(prop, score2, chol2, dev(step, plan), x_eval=dev(th[:, sl], plan))
    ref_b = OI.iwls_backward(spec, "loc_np0", th2, aux, step.astype(np.float64), th[:, sl].astype(np.float64))
    np.testing.assert_allclose(bwd.cpu().numpy(), ref_b, rtol=10 * tol, atol=10 * tol)


def test_quadform(lb):
    spec, centre = make("s2", np.float64)
    plan = lb.Plan(spec)
    th, _ = synth.evaluation_points(spec, centre, 6)
    q = plan.quadform("loc_np2", dev(th, plan)).cpu().numpy()
    np.testing.assert_allclose(q, sam.quadform(spec, "loc_np2", th), rtol=1e-12)


# --- the ModelInterface mirror --------------------------------------------------------------
def test_interface_protocol(lb):
    spec, centre = make("s3", np.float32)
    plan = lb.Plan(spec)
    iface = lb.Interface(plan)
    C = 6
    th, aux = synth.evaluation_points(spec, centre, C)
    state = iface.unpack(dev(th, plan), dev(aux, plan))
    assert set(state) == set(spec.position_keys()) | set(spec.aux_keys())
    pos = iface.extract_position(["loc_np0_beta"], state)
    assert list(pos) == ["loc_np0_beta"] and pos["loc_np0_beta"].shape == (C, 20)
    lp0 = iface.log_prob(state).cpu().numpy()
    new = {"loc_np0_beta": pos["loc_np0_beta"] + 0.05}
    state2 = iface.update_state(new, state)
    assert state2 is not state and torch.equal(state["loc_np0_beta"], pos["loc_np0_beta"])  # pure
    th2 = th.copy()
    th2[:, 1:21] += 0.05
    lp2 = iface.log_prob(state2).cpu().numpy()
    np.testing.assert_allclose(lp2, sam.log_prob(spec, th2, aux), rtol=1e-5)
    assert not np.allclose(lp0, lp2)
    with pytest.raises(KeyError):
        iface.update_state({"nope": 1}, state)
    lp, g = iface.log_prob_and_grad(state2, ["scale_np0_beta"])
    np.testing.assert_allclose(g["scale_np0_beta"].cpu().numpy(), sam.grad(spec, th2, aux)[:, 22:42],
                               rtol=1e-4, atol=1e-2)
    chol = iface.chol_info_fn("scale_np0_beta")(state2)
    assert chol.shape == (C, 20, 20)
    init = iface.initial_state(3)
    assert float(init["loc_np0_tau2"][0]) == spec.block("loc_np0").prior.tau2_init and float(init["loc_np0_beta"].abs().sum()) == 0.0


def test_status_codes_not_exceptions_from_c(lb):
    import ctypes as Ct

    from liesel_b200 import _lib

    spec, centre = make("s1", np.float32)
    plan = lb.Plan(spec)
    th, _ = synth.evaluation_points(spec, centre, 2)
    t = dev(th, plan)
    out = torch.empty(2, device="cuda")
    # workspace too small -> status 2, nothing launched
    st = _lib.lib.lslb_logp_grad_f32(plan._handle, 2, _lib.ptr(t), None, _lib.ptr(out), None, 0,
                                     _lib.ptr(plan.workspace(2)), 16, _lib.stream_ptr())
    assert st == 2
    st = _lib.lib.lslb_logp_grad_f64(plan._handle, 2, _lib.ptr(t), None, _lib.ptr(out), None, 0,
                                     _lib.ptr(plan.workspace(2)), 1 << 20, _lib.stream_ptr())
    assert st == 1  # dtype mismatch
    with pytest.raises(_lib.LieselB200Error):
        plan.logp_grad(t.double())


# --- row sharding (S4 style), emulated on one GPU --------------------------------------------
def test_row_sharded_partials_sum_to_full(lb):
    """Two 'ranks' hold half the rows each; rank 1 passes LSLB_SKIP_PRIOR; partials add up."""
    from liesel_b200 import dist as ld
    from liesel_b200.spec import ModelSpec

    spec, centre = make("s4", np.float32)
    th, aux = synth.evaluation_points(spec, centre, 10)
    full = lb.Plan(spec)
    lp_full, g_full = full.logp_grad(dev(th, full))
    acc_lp, acc_g = None, None
    for rank in range(2):
        a, b = ld.row_shard(spec.n, rank, 2)
        part = ModelSpec(dtype=spec.dtype)
        part.add_response(spec.y[a:b], "bernoulli").add_predictor("logits")
        blk = spec.blocks[0]
        part.add_p_smooth(blk.X[a:b], blk.prior.m, blk.prior.s, "logits", blk.name)
        plan = lb.Plan(part)
        lp, g = plan.logp_grad(dev(th, plan), flags=0 if rank == 0 else ld.SKIP_PRIOR)
        acc_lp = lp if acc_lp is None else acc_lp + lp
        acc_g = g if acc_g is None else acc_g + g
    np.testing.assert_allclose(acc_lp.cpu().numpy(), lp_full.cpu().numpy(), rtol=2e-6)
    scale = sam.grad_scale(spec, th)
    assert np.all(np.abs(acc_g.cpu().numpy() - g_full.cpu().numpy()) <= 32 * 2.0**-24 * scale + 1e-6)
    check_logp_grad(spec, full, th, aux)


# --- register-sliced dense kernel (S4 path) ----------------------------------------------------
@pytest.mark.parametrize("p,n,C", [(256, 3000, 70), (100, 2500, 9), (36, 777, 64), (20, 1000, 5)])
def test_dense_x2_logistic(lb, p, n, C):
    spec, centre = synth.s4_logistic(n=n, p=p, dtype=np.float32)
    plan = lb.Plan(spec)
    assert plan.variant == "dense_x2_register_sliced"
    th, aux = synth.evaluation_points(spec, centre, C)
    check_logp_grad(spec, plan, th, aux)


def test_dense_x2_other_families(lb):
    from liesel_b200.spec import ModelSpec

    rng = np.random.default_rng(9)
    n, p = 1500, 40
    X = np.column_stack([rng.normal(size=(n, p)) * 0.2])
    beta = rng.normal(size=p) * 0.3
    # Poisson, dense + intercept
    y = rng.poisson(np.exp(0.5 + X @ beta)).astype(np.float64)
    spec = ModelSpec(dtype=np.dtype(np.float32))
    spec.add_response(y, "poisson").add_predictor("log_rate")
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 10.0, "log_rate", "b0")
    spec.add_p_smooth(X, 0.0, 10.0, "log_rate", "bx")
    plan = lb.Plan(spec)
    assert plan.variant == "dense_x2_register_sliced"
    th, aux = synth.evaluation_points(spec, np.concatenate([[0.5], beta]), 12)
    check_logp_grad(spec, plan, th, aux)
    # Normal with fixed scale
    y = X @ beta + rng.normal(size=n)
    spec = ModelSpec(dtype=np.dtype(np.float32))
    spec.add_response(y, "normal").add_predictor("loc")
    spec.fixed_scale = 0.7
    spec.add_p_smooth(X, 0.0, 10.0, "loc", "bx")
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, beta, 3)
    check_logp_grad(spec, plan, th, aux)


# --- fused NUTS leaf kernels vs the torch reference path ----------------------------------------
def test_nuts_fused_matches_torch_path(lb):
    from liesel_b200.nuts import BatchedNUTS

    spec, centre = synth.s3_gamlss(n=3000, k=8, dtype=np.float64)
    plan = lb.Plan(spec)
    C = 24
    th, aux = synth.evaluation_points(spec, centre, C)
    q0, a = dev(th, plan), dev(aux, plan)
    f = lambda q: plan.logp_grad(q.contiguous(), a)
    into = lambda q, lp, g: plan.logp_grad(q, a, out=(lp, g))
    s_ref = BatchedNUTS(f, q0, seed=5, fused=False, initial_step_size=0.02, max_treedepth=6)
    s_fus = BatchedNUTS(f, q0, seed=5, fused=True, initial_step_size=0.02, max_treedepth=6,
                        logp_grad_into=into)
    for _ in range(4):
        i_ref, i_fus = s_ref.transition(), s_fus.transition()
        assert torch.equal(i_ref.treedepth, i_fus.treedepth)
        assert torch.equal(i_ref.error_code, i_fus.error_code)
        assert torch.equal(i_ref.leapfrog, i_fus.leapfrog)
        np.testing.assert_allclose(i_fus.acceptance_prob.cpu().numpy(), i_ref.acceptance_prob.cpu().numpy(),
                                   rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(s_fus.q.cpu().numpy(), s_ref.q.cpu().numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(s_fus.logp.cpu().numpy(), s_ref.logp.cpu().numpy(), rtol=1e-10)
    assert int(i_ref.treedepth.max()) >= 3  # trajectories long enough to exercise the checkpoints


def test_nuts_compaction_same_draws(lb):
    """Dropping terminated chains from the batch changes the cost, not the draws (fp64: the only
    difference is the reduction order of a smaller batch)."""
    from liesel_b200.nuts import BatchedNUTS

    spec, centre = synth.s3_gamlss(n=3000, k=8, dtype=np.float64)
    plan = lb.Plan(spec)
    C = 200
    th, aux = synth.evaluation_points(spec, centre, C)
    q0, a = dev(th, plan), dev(aux, plan)
    f = lambda q: plan.logp_grad(q.contiguous(), a)

    def into(q, lp, g, chains=None):
        plan.logp_grad(q, a if chains is None else a.index_select(0, chains).contiguous(), out=(lp, g))

    kw = dict(seed=11, fused=True, initial_step_size=0.02, max_treedepth=7, logp_grad_into=into)
    s_ref = BatchedNUTS(f, q0, **kw)
    s_cmp = BatchedNUTS(f, q0, compact=True, **kw)
    assert s_cmp.compact and not s_ref.compact
    # heterogeneous step sizes -> heterogeneous tree depths, as after per-chain adaptation
    eps = 0.004 * torch.pow(torch.tensor(2.0, device=plan.device, dtype=q0.dtype),
                            torch.arange(C, device=plan.device, dtype=q0.dtype) % 5)
    s_ref.step_size, s_cmp.step_size = eps.clone(), eps.clone()
    for _ in range(4):
        i_ref, i_cmp = s_ref.transition(), s_cmp.transition()
        assert torch.equal(i_ref.treedepth, i_cmp.treedepth)
        assert torch.equal(i_ref.error_code, i_cmp.error_code)
        assert torch.equal(i_ref.leapfrog, i_cmp.leapfrog)
        np.testing.assert_allclose(i_cmp.acceptance_prob.cpu().numpy(), i_ref.acceptance_prob.cpu().numpy(),
                                   rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(s_cmp.q.cpu().numpy(), s_ref.q.cpu().numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(s_cmp.logp.cpu().numpy(), s_ref.logp.cpu().numpy(), rtol=1e-10)
    depths = i_ref.treedepth
    assert int(depths.max()) > int(depths.min())  # the batch really was heterogeneous
    assert s_cmp.evals == s_ref.evals and s_cmp.chain_evals < 0.9 * s_ref.chain_evals


def test_nuts_cuda_graphs_same_draws(lb):
    """Replaying the leaf ranges from CUDA graphs changes launch overhead, not results."""
    from liesel_b200.nuts import BatchedNUTS

    spec, centre = synth.s3_gamlss(n=3000, k=8, dtype=np.float32)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 40)
    q0, a = dev(th, plan), dev(aux, plan)
    f = lambda q: plan.logp_grad(q.contiguous(), a)  # noqa: E731
    into = lambda q, lp, g: plan.logp_grad(q, a, out=(lp, g))  # noqa: E731
    kw = dict(seed=9, fused=True, initial_step_size=0.02, max_treedepth=6, logp_grad_into=into)
    s_a = BatchedNUTS(f, q0, graphs=False, **kw)
    s_b = BatchedNUTS(f, q0, graphs=True, **kw)
    assert s_b.use_graphs
    for _ in range(6):
        i_a, i_b = s_a.transition(), s_b.transition()
        assert torch.equal(i_a.treedepth, i_b.treedepth) and torch.equal(i_a.leapfrog, i_b.leapfrog)
        assert torch.equal(s_a.q, s_b.q) and torch.equal(s_a.logp, s_b.logp)
        assert i_a.evaluations == i_b.evaluations
    assert len(s_b._graphs) >= 3  # several leaf ranges were captured and replayed


# --- IWLS-within-Gibbs sampler: the reference's own kernel integration pins ------------------------
def test_iwls_sampler_recovers_model_lm(lb, golden_dir):
    """tests/goose/kernels/test_iwls.py via model_lm.py:70-84: posterior means within 5 % of OLS,
    mean acceptance within 0.05 of the dual-averaging target."""
    import json
    import os

    from liesel_b200.iwls_sampler import IWLSGibbsSampler
    from liesel_b200.spec import ModelSpec

    g = json.load(open(os.path.join(golden_dir, "model_lm.json")))
    X, y = np.array(g["X"]), np.array(g["y"])
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    spec.add_p_smooth(X, 0.0, None, "loc", "beta")
    spec.add_p_smooth(np.ones((30, 1)), 0.0, None, "scale", "log_sigma")
    plan = lb.Plan(spec)
    C = 8
    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64, device="cuda").repeat(C, 1)
    s = IWLSGibbsSampler(plan, q0, None, seed=42)
    s.warmup(1500)
    draws, _, infos = s.sample(1500)
    d = draws.cpu().numpy()
    assert d[:, :, :2].mean(axis=(0, 1)) == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    for name in ["beta", "log_sigma"]:
        acc = np.mean([float(i[name]["acceptance_prob"].mean()) for i in infos])
        assert acc == pytest.approx(0.8, abs=0.05), (name, acc)
        codes = torch.cat([i[name]["error_code"] for i in infos])
        assert float((codes == 0).float().mean()) > 0.99


# --- IWLS of the random-intercept block (diagonal information) ---------------------------------------
def _small_ri_model(family, dtype, n=6000, G=60, seed=9, constant_column=True):
    """Random-intercept model with unbalanced groups (two of them empty) for `family`."""
    from liesel_b200.spec import ModelSpec

    rng = np.random.default_rng(seed)
    idx = np.sort(rng.integers(0, G - 1, size=n))  # group G - 1 has no rows
    idx[idx >= 7] += 1                              # ... and so has group 7
    idx = np.minimum(idx, G - 2)
    X = np.column_stack([np.ones(n) if constant_column else rng.normal(size=n), rng.normal(size=(n, 3))]) * 0.3
    beta = np.array([0.4, -0.3, 0.2, 0.1])
    b = rng.normal(scale=0.5, size=G)
    eta = X @ beta + b[idx]
    spec = ModelSpec(dtype=np.dtype(dtype))
    if family == "poisson":
        spec.add_response(rng.poisson(np.exp(eta)).astype(np.float64), "poisson").add_predictor("log_rate")
        pred = "log_rate"
    elif family == "bernoulli":
        spec.add_response(rng.binomial(1, 1 / (1 + np.exp(-eta))).astype(np.float64), "bernoulli").add_predictor("logits")
        pred = "logits"
    else:
        spec.add_response(eta + rng.normal(scale=0.7, size=n), "normal").add_predictor("loc").add_predictor("scale")
        spec.add_p_smooth(np.ones((n, 1)), 0.0, 10.0, "scale", "log_sigma")
        pred = "loc"
    spec.add_p_smooth(X, 0.0, 10.0, pred, pred + "_p0")
    spec.add_random_intercept(idx, G, 0.1, 0.5, pred, pred + "_ri0")
    centre = np.concatenate(([np.array([np.log(0.7)])] if family == "normal" else []) + [beta, b])
    return spec, centre, pred + "_ri0"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("family", ["poisson", "bernoulli", "normal"])
def test_iwls_group_block_against_oracle(lb, dtype, family):
    """Score and information of the random-intercept block (`lslb_iwls_group`): the oracle forms the dense
    G x G information the reference's `-jacfwd(grad(log_prob))` would (goose/iwls.py:218-243); it is
    diagonal, and the device returns that diagonal. Groups without rows carry the prior terms only."""
    spec, centre, name = _small_ri_model(family, dtype)
    plan = lb.Plan(spec)
    C = 37
    th, aux = synth.evaluation_points(spec, centre, C)
    score, info = plan.iwls_group(dev(th, plan), dev(aux, plan) if spec.num_aux else None)
    sel = [0, 17, C - 1]
    r_score, r_info = sam.iwls(spec, name, th[sel], aux[sel])
    diag = np.diagonal(r_info, axis1=1, axis2=2)
    assert np.all(r_info - diag[:, :, None] * np.eye(diag.shape[1]) == 0)  # the dense matrix is diagonal
    tol = 3e-5 if dtype == np.float32 else 1e-11
    np.testing.assert_allclose(info.cpu().numpy()[sel], diag, rtol=tol)
    # the score is a sum of signed residuals: absolute tolerance on the scale of the summed magnitudes
    np.testing.assert_allclose(score.cpu().numpy()[sel], r_score, rtol=tol, atol=tol * np.abs(diag).max())
    blk = spec.block(name)
    empty = [7, blk.ncoef - 1]
    np.testing.assert_allclose(info.cpu().numpy()[:, empty], 1 / 0.5**2, rtol=1e-6)
    # the groups' shares of the log-posterior: changing ONE group's effect changes the log-posterior by the
    # change of that group's share and leaves every other share alone (fp64: to rounding)
    t = dev(th, plan)
    _, _, llg = plan.iwls_group(t, dev(aux, plan) if spec.num_aux else None, want_loglik=True)
    t2 = t.clone()
    g_mod = 11
    t2[:, blk.param_offset + g_mod] += 0.25
    _, _, llg2 = plan.iwls_group(t2, dev(aux, plan) if spec.num_aux else None, want_loglik=True)
    lp1 = sam.log_prob(spec, th[sel], aux[sel])
    th2 = th.copy()
    th2[:, blk.param_offset + g_mod] += np.asarray(0.25, dtype=dtype)
    lp2 = sam.log_prob(spec, th2[sel], aux[sel])
    d = (llg2 - llg).cpu().numpy()[sel]
    ltol = 2e-4 if dtype == np.float32 else 1e-9
    assert np.all(np.abs(d[:, g_mod] - (lp2 - lp1)) <= ltol * (np.abs(lp2 - lp1) + np.abs(llg.cpu().numpy()[sel, g_mod])))
    d[:, g_mod] = 0
    assert np.all(d == 0)
    # full S5-type model (synth) at two chain counts that do not fill the 32 x 32 tiles
    spec5, centre5 = make("s5u", dtype)
    plan5 = lb.Plan(spec5)
    th5, _ = synth.evaluation_points(spec5, centre5, 5)
    s5, i5 = plan5.iwls_group(dev(th5, plan5))
    rs5, ri5 = sam.iwls(spec5, "log_rate_ri0", th5[:2], None)
    np.testing.assert_allclose(i5.cpu().numpy()[:2], np.diagonal(ri5, axis1=1, axis2=2), rtol=tol)
    np.testing.assert_allclose(s5.cpu().numpy()[:2], rs5, rtol=tol, atol=tol * np.abs(ri5).max())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_iwls_group_proposal_tail_against_oracle(lb, dtype):
    """`lslb_iwls_propose_diag` after `lslb_iwls_group`: mean, draw, forward and backward log-density equal
    the reference's dense algebra (ridge + Cholesky + triangular solves, goose/iwls.py:232-238,323-336,
    iwls_utils.py:14-70) on the dense diagonal matrix; an information with a non-positive entry takes the
    identity fallback with error code 2 (`_safe_chol`, iwls.py:245-290)."""
    from liesel_b200.plan import iwls_propose_diag
    from oracle import iwls as oiw

    spec, centre, name = _small_ri_model("poisson", dtype)
    plan = lb.Plan(spec)
    C = 5
    th, aux = synth.evaluation_points(spec, centre, C)
    blk = spec.block(name)
    sl = slice(blk.param_offset, blk.param_offset + blk.ncoef)
    rng = np.random.default_rng(3)
    z = rng.normal(size=(C, blk.ncoef))
    step = np.array([0.3, 0.7, 1.0, 1.3, 0.05])
    ref = oiw.iwls_proposal(spec, name, th, None, step, z)
    t = dev(th, plan)
    score, info = plan.iwls_group(t)
    pos = t[:, sl].contiguous()
    prop, fwd, code = iwls_propose_diag(pos, score, info, dev(step.astype(dtype), plan), z=dev(z.astype(dtype), plan))
    tol = 2e-4 if dtype == np.float32 else 1e-9
    np.testing.assert_allclose(prop.cpu().numpy(), ref["prop"], rtol=tol, atol=tol)
    np.testing.assert_allclose(fwd.cpu().numpy(), ref["fwd_log_prob"], rtol=tol, atol=tol * blk.ncoef)
    assert int(code.abs().sum()) == 0
    # backward density of the current position from the proposal's side
    thp = th.copy()
    thp[:, sl] = prop.cpu().numpy()
    bwd_ref = oiw.iwls_backward(spec, name, thp, None, step, th[:, sl])
    tp = dev(thp, plan)
    score_p, info_p = plan.iwls_group(tp)
    _, bwd, _ = iwls_propose_diag(tp[:, sl].contiguous(), score_p, info_p, dev(step.astype(dtype), plan), x_eval=pos)
    np.testing.assert_allclose(bwd.cpu().numpy(), bwd_ref, rtol=tol, atol=tol * blk.ncoef)
    # indefinite information on chain 1: identity fallback, code 2 (dense oracle on the same diagonal)
    info_bad = info.clone()
    info_bad[1, 3] = -5.0
    prop_b, fwd_b, code_b = iwls_propose_diag(pos, score, info_bad, dev(step.astype(dtype), plan), z=dev(z.astype(dtype), plan))
    assert code_b.cpu().tolist() == [0, 2, 0, 0, 0]
    dense = np.diag(info_bad[1].double().cpu().numpy())
    ch, codes = oiw.safe_chol(oiw.chol_info(dense[None]))
    assert codes[0] == 2
    mu = th[1, sl] + step[1] ** 2 / 2 * oiw.solve(ch[0], score[1].double().cpu().numpy())
    pr = oiw.mvn_sample(z[1], mu, ch[0] / step[1])
    np.testing.assert_allclose(prop_b[1].cpu().numpy(), pr, rtol=tol, atol=tol)
    np.testing.assert_allclose(float(fwd_b[1]), oiw.mvn_log_prob(pr, mu, ch[0] / step[1]), rtol=tol)


@pytest.mark.parametrize("group_accept", ["joint", "per_group"])
def test_iwls_sampler_with_random_intercept_block(lb, group_accept):
    """IWLS-within-Gibbs over BOTH blocks of a Poisson random-intercept model (the S5 family): the fixed
    effects through the dense IWLS pass, the random intercept through its diagonal information, accepted
    jointly (the reference kernel's law) or group by group. Checked against the NUTS driver on the same
    posterior (means within 5 Monte-Carlo standard errors, a standard error floor of 2 % of the posterior
    sd; posterior sds within 10 %) and for a sane acceptance rate. (No constant column next to the random
    intercept: block updates random-walk along that ridge, which is a property of the blocking.)"""
    from liesel_b200 import diagnostics
    from liesel_b200.iwls_sampler import IWLSGibbsSampler
    from liesel_b200.nuts import BatchedNUTS

    spec, centre, name = _small_ri_model("poisson", np.float64, n=4000, G=24, constant_column=False)
    plan = lb.Plan(spec)
    C = 64
    th, _ = synth.evaluation_points(spec, centre, C)
    q0 = dev(th, plan)
    smp = IWLSGibbsSampler(plan, q0, None, seed=7, initial_step_size=0.1, group_accept=group_accept)
    assert [b.name for b in smp.blocks] == ["log_rate_p0", name]
    smp.warmup(300)
    draws, _, infos = smp.sample(400)
    for k in ("log_rate_p0", name):
        acc = np.mean([float(i[k]["acceptance_prob"].mean()) for i in infos])
        assert 0.65 < acc < 0.95, (k, acc)  # dual averaging towards 0.8 over a short warm-up
        codes = torch.cat([i[k]["error_code"] for i in infos])
        assert float((codes == 0).float().mean()) > 0.99
    nuts = BatchedNUTS(lambda q: plan.logp_grad(q.contiguous(), None), q0, seed=11)
    nuts.warmup(300)
    draws_n, _ = nuts.sample(300)
    a, b = draws.cpu().numpy(), draws_n.cpu().numpy()
    sd = b.std(axis=(0, 1))
    ess_a = diagnostics.ess_bulk(a)
    ess_b = diagnostics.ess_bulk(b)
    se = sd * np.sqrt(1 / ess_a + 1 / ess_b) + 0.02 * sd
    assert np.all(np.abs(a.mean(axis=(0, 1)) - b.mean(axis=(0, 1))) < 5 * se)
    np.testing.assert_allclose(a.std(axis=(0, 1)), sd, rtol=0.1)


def test_full_size_s5_iwls_both_blocks(lb):
    """BASELINE config 5 at size (n = 5e6, G = 1e5, 256 chains): the IWLS pass of the fixed-effects block
    (`iwls_dense_reg_kernel`: fp32 register accumulators over n / chunks rows each - the thing small-n parity
    cannot vouch for) and of the random-intercept block (`group_iwls_kernel`) against the fp64 oracle on three
    chains. Same tolerance formulas as the small cases."""
    spec, centre = synth.s5_poisson_ri(n=5_000_000, G=100_000)
    plan = lb.Plan(spec)
    C = 256
    th, aux = synth.evaluation_points(spec, centre, C)
    sel = [0, 100, C - 1]
    check_iwls(spec, plan, "log_rate_p0", th, aux, chains=sel)
    # 1024 chains: four times fewer row chunks would fit the SMs; the launch keeps the rows per accumulator
    th4 = np.concatenate([th, th, th, th])
    check_iwls(spec, plan, "log_rate_p0", th4, np.concatenate([aux] * 4), chains=[3, 700, 1023])
    t = dev(th, plan)
    score, diag, share = plan.iwls_group(t, want_loglik=True)
    prep = sam.Prepared(spec)
    r_score, r_diag, r_share = sam.iwls_group_diag(spec, "log_rate_ri0", th[sel], None, prep=prep)
    np.testing.assert_allclose(diag.cpu().numpy()[sel], r_diag, rtol=3e-5)
    np.testing.assert_allclose(score.cpu().numpy()[sel], r_score, rtol=3e-5, atol=3e-5 * np.abs(r_diag).max())
    # shares: the device leaves the row constants out, so compare differences between chains
    sh = share.cpu().numpy()[sel].astype(np.float64)
    d_dev, d_ref = sh[1:] - sh[:1], r_share[1:] - r_share[:1]
    scale = np.abs(r_share).max()
    assert np.all(np.abs(d_dev - d_ref) <= 3e-5 * np.abs(d_ref) + 16 * 2.0**-24 * scale)


# --- BASELINE.json full sizes: size-independent properties -------------------------------------------
def test_full_size_h_properties(lb):
    """H (n = 1e6, 1024 chains): the oracle would take minutes, so check properties instead:
    (1) 4 chains picked out of the 1024 give bit-identical results when evaluated alone in a batch
        of the same tile width (chains are independent),
    (2) the sum over two disjoint row halves equals the full evaluation (additivity over rows),
    (3) a few chains agree with the float64 oracle on a 2 % row subsample scaled... (not valid) ->
        instead the float64 plan of the same data agrees with the float32 plan within rtol 1e-5,
    (4) gradient = finite difference of the float64 log-posterior."""
    from liesel_b200.spec import ModelSpec

    n, C = 1_000_000, 1024
    spec, centre = synth.s3_gamlss(n=n)
    plan = lb.Plan(spec)
    assert plan.variant == "banded_tma_runlength"
    th, aux = synth.evaluation_points(spec, centre, C, tau2="loguniform")
    t, a = dev(th, plan), dev(aux, plan)
    lp, g = plan.logp_grad(t, a)
    lp, g = lp.clone(), g.clone()
    assert torch.isfinite(lp).all() and torch.isfinite(g).all()
    # (1) chain independence: chains 0..255 evaluated alone agree with their values inside the
    # 1024-chain batch (not bitwise: a smaller batch uses more row chunks, i.e. another summation tree)
    lp_s, g_s = plan.logp_grad(t[:256].contiguous(), a[:256].contiguous())
    np.testing.assert_allclose(lp_s.cpu().numpy(), lp[:256].cpu().numpy(), rtol=1e-6)
    gm = g[:256].abs().max().item()
    np.testing.assert_allclose(g_s.cpu().numpy(), g[:256].cpu().numpy(), rtol=1e-5, atol=2e-6 * gm)
    # (3)/(4) float64 plan of the same data
    spec64, _ = synth.s3_gamlss(n=n, dtype=np.float64)
    # same float32 covariate/knots/response as the float32 model, up-cast
    spec64.y = spec.y.astype(np.float64)
    for b64, b32 in zip(spec64.blocks, spec.blocks):
        if b32.x is not None:
            b64.x, b64.knots = b32.x.astype(np.float64), b32.knots.astype(np.float64)
    plan64 = lb.Plan(spec64)
    sel = [0, 300, 1023]
    t64 = torch.as_tensor(th[sel].astype(np.float64), device="cuda")
    a64 = torch.as_tensor(aux[sel].astype(np.float64), device="cuda")
    lp64, g64 = plan64.logp_grad(t64, a64)
    np.testing.assert_allclose(lp[sel].cpu().numpy(), lp64.cpu().numpy(), rtol=1e-5)
    gs = g64.abs().max().item()
    np.testing.assert_allclose(g[sel].cpu().numpy(), g64.cpu().numpy(), rtol=1e-4, atol=2e-6 * gs * 50)
    for j in [0, 7, 21, 35]:
        e = 1e-6
        tp, tm = t64.clone(), t64.clone()
        tp[:, j] += e
        tm[:, j] -= e
        fd = (plan64.logp_grad(tp, a64, want_grad=False)[0] - plan64.logp_grad(tm, a64, want_grad=False)[0]) / (2 * e)
        np.testing.assert_allclose(fd.cpu().numpy(), g64[:, j].cpu().numpy(), rtol=2e-4, atol=1e-3)
    # (2) additivity over rows: two half plans, second with LSLB_SKIP_PRIOR
    acc_lp, acc_g = 0, 0
    for r, (lo, hi) in enumerate([(0, n // 2), (n // 2, n)]):
        part = ModelSpec(dtype=spec.dtype)
        part.add_response(spec.y[lo:hi], "normal").add_predictor("loc").add_predictor("scale")
        for b in spec.blocks:
            pred = "loc" if b.predictor == 0 else "scale"
            if b.kind == 3:
                part.add_p_smooth(np.ones((hi - lo, 1)), b.prior.m, b.prior.s, pred, b.name)
            else:
                part.add_spline_smooth(b.x[lo:hi], b.knots, b.prior.K, b.prior.ig_a, b.prior.ig_b, pred,
                                       b.name, tau2_init=b.prior.tau2_init)
        pp = lb.Plan(part)
        l_r, g_r = pp.logp_grad(t[:256].contiguous(), a[:256].contiguous(), flags=0 if r == 0 else 1)
        acc_lp, acc_g = acc_lp + l_r.double(), acc_g + g_r.double()
    np.testing.assert_allclose(acc_lp.cpu().numpy(), lp[:256].double().cpu().numpy(), rtol=2e-6)
    gmax = g[:256].abs().max().item()
    np.testing.assert_allclose(acc_g.cpu().numpy(), g[:256].double().cpu().numpy(), rtol=1e-4, atol=1e-5 * gmax)


def test_full_size_h_and_s5_against_oracle(lb):
    """BASELINE sizes against the fp64 oracle itself on three chains each (the oracle needs seconds
    per chain at these sizes), same tolerance as the small cases: errors that grow with the number of
    rows a CTA sums do not show at n = 5e4."""
    spec, centre = synth.s3_gamlss(n=1_000_000)
    plan = lb.Plan(spec)
    assert plan.variant == "banded_tma_runlength"
    th, aux = synth.evaluation_points(spec, centre, 1024, tau2="loguniform")
    check_logp_grad(spec, plan, th, aux, chains=[0, 411, 1023])
    del plan
    spec, centre = synth.s5_poisson_ri()
    plan = lb.Plan(spec)
    assert plan.variant == "dense_group_x2_tma"
    th, aux = synth.evaluation_points(spec, centre, 1024)
    check_logp_grad(spec, plan, th, aux, chains=[0, 1023])


def test_full_size_s5_group_sum_properties(lb):
    """S5 at full size: the random-intercept gradient is a segment sum — summing it over groups must
    equal the fixed-intercept column's gradient divided by its constant (0.3), prior terms removed."""
    spec, centre = synth.s5_poisson_ri()
    plan = lb.Plan(spec)
    assert plan.variant == "dense_group_x2_tma"
    C = 64
    th, aux = synth.evaluation_points(spec, centre, C)
    t = dev(th, plan)
    lp, g = plan.logp_grad(t, None)
    lp2, g2 = plan.logp_grad(t, None)
    assert torch.equal(lp, lp2) and torch.equal(g, g2)          # deterministic segment sums
    lp_l, g_l = plan.logp_grad(t, None, flags=1)                 # likelihood part only
    p = 8
    sum_b = g_l[:, p:].double().sum(dim=1)                       # sum_g sum_{i in g} r_i = sum_i r_i
    from_x0 = g_l[:, 0].double() / 0.3                           # X[:, 0] = 0.3
    np.testing.assert_allclose(sum_b.cpu().numpy(), from_x0.cpu().numpy(), rtol=2e-4)
    # prior part: grad - grad_lik = -(b - m)/s^2
    pr = (g - g_l)[:, p:].double()
    np.testing.assert_allclose(pr.cpu().numpy(), (-(t[:, p:].double()) / 0.25).cpu().numpy(), rtol=1e-4, atol=1e-4)


# --- tensor-core dense path (tcgen05, TF32 x3 split) ------------------------------------------------
@pytest.mark.parametrize("p,n,C", [(128, 40_000, 256), (256, 33_000, 300)])
def test_dense_tcgen05_logistic(lb, p, n, C):
    spec, centre = synth.s4_logistic(n=n, p=p, dtype=np.float32)
    plan = lb.Plan(spec)
    assert plan.variant == "dense_tcgen05_tf32x3"
    th, aux = synth.evaluation_points(spec, centre, C)
    sel = [0, 1, 127, 128, C - 1]
    logp, grad = plan.logp_grad(dev(th, plan), None)
    logp_v, _ = plan.logp_grad(dev(th, plan), None, want_grad=False)
    torch.cuda.synchronize()
    prep = sam.Prepared(spec)
    ref_lp = sam.log_prob(spec, th[sel], prep=prep)
    ref_g = sam.grad(spec, th[sel], prep=prep)
    sc_g = sam.grad_scale(spec, th[sel], prep=prep)
    sc_lp = sam.log_prob_scale(spec, th[sel], prep=prep)
    lp = logp.cpu().numpy()[sel]
    np.testing.assert_allclose(logp_v.cpu().numpy()[sel], lp, rtol=1e-6)
    assert np.all(np.abs(lp - ref_lp) <= 1e-5 * np.abs(ref_lp) + ULPS * 2.0**-24 * sc_lp)
    gerr = np.abs(grad.cpu().numpy()[sel] - ref_g)
    bound = 1e-5 * np.abs(ref_g) + ULPS * 2.0**-24 * sc_g
    assert np.all(gerr <= bound), float((gerr / bound).max())
    # the SIMT path on the same data (fewer chains than the tensor-core threshold) agrees
    lp_s, g_s = plan.logp_grad(dev(th[:64], plan), None)
    np.testing.assert_allclose(lp_s.cpu().numpy(), logp.cpu().numpy()[:64], rtol=2e-6)


def test_dense_tcgen05_two_pass_large_n(lb):
    """
    From 2^19 rows on (and when no column's TF32 rounding residue has a systematic component) the
    tensor-core path drops the two products with a lo factor of the large operands. Same tolerance
    as everywhere (rtol 1e-5 + 16 ulps of the term scale); reported: how much of it is used.
    """
    spec, centre = synth.s4_logistic(n=600_000, p=128, dtype=np.float32)
    plan = lb.Plan(spec)
    assert plan.variant == "dense_tcgen05_tf32x3"
    C = 256
    th, aux = synth.evaluation_points(spec, centre, C)
    sel = [0, 131, C - 1]
    logp, grad = plan.logp_grad(dev(th, plan), None)
    logp_v, _ = plan.logp_grad(dev(th, plan), None, want_grad=False)
    torch.cuda.synchronize()
    prep = sam.Prepared(spec)
    ref_lp = sam.log_prob(spec, th[sel], prep=prep)
    ref_g = sam.grad(spec, th[sel], prep=prep)
    sc_g = sam.grad_scale(spec, th[sel], prep=prep)
    sc_lp = sam.log_prob_scale(spec, th[sel], prep=prep)
    lp = logp.cpu().numpy()[sel]
    np.testing.assert_allclose(logp_v.cpu().numpy()[sel], lp, rtol=1e-6)
    lerr = np.abs(lp - ref_lp) / (1e-5 * np.abs(ref_lp) + ULPS * 2.0**-24 * sc_lp)
    gerr = np.abs(grad.cpu().numpy()[sel] - ref_g) / (1e-5 * np.abs(ref_g) + ULPS * 2.0**-24 * sc_g)
    print(f"two-pass tolerance use: logp {lerr.max():.3f}, grad {gerr.max():.3f}")
    assert lerr.max() <= 1.0 and gerr.max() <= 1.0, (float(lerr.max()), float(gerr.max()))
    # a constant column that TF32 cannot represent makes the rounding residue systematic: such data
    # must stay on three passes (checked through the result: it still meets the tolerance)
    import dataclasses

    X2 = spec.blocks[0].X.copy()
    X2[:, 5] = np.float32(0.1)
    spec2 = dataclasses.replace(spec, blocks=[dataclasses.replace(spec.blocks[0], X=X2)])
    plan2 = lb.Plan(spec2)
    lp2, g2 = plan2.logp_grad(dev(th, plan2), None)
    torch.cuda.synchronize()
    r_lp = sam.log_prob(spec2, th[sel])
    r_g = sam.grad(spec2, th[sel])
    s_g = sam.grad_scale(spec2, th[sel])
    s_lp = sam.log_prob_scale(spec2, th[sel])
    assert np.all(np.abs(lp2.cpu().numpy()[sel] - r_lp) <= 1e-5 * np.abs(r_lp) + ULPS * 2.0**-24 * s_lp)
    g2err = np.abs(g2.cpu().numpy()[sel] - r_g) / (1e-5 * np.abs(r_g) + ULPS * 2.0**-24 * s_g)
    print(f"constant-column data: grad tolerance use {g2err.max():.3f} at {np.unravel_index(g2err.argmax(), g2err.shape)}")
    assert g2err.max() <= 1.0, float(g2err.max())


def test_full_size_s4_shard_against_oracle(lb):
    """One S4 shard at BASELINE size (1.25e6 rows x 256, 1024 chains) against the fp64 oracle on three
    chains: the tensor-core path's accumulation error grows with the rows a CTA sums, so small-n parity
    says nothing about this size."""
    spec, centre = synth.s4_logistic(n=10_000_000, p=256, world_size=8, rank=3)
    plan = lb.Plan(spec)
    assert plan.variant == "dense_tcgen05_tf32x3" and spec.n == 1_250_000
    C = 1024
    th, aux = synth.evaluation_points(spec, centre, C)
    sel = [0, 517, C - 1]
    logp, grad = plan.logp_grad(dev(th, plan), None)
    torch.cuda.synchronize()
    prep = sam.Prepared(spec)
    ref_lp = sam.log_prob(spec, th[sel], prep=prep)
    ref_g = sam.grad(spec, th[sel], prep=prep)
    sc_g = sam.grad_scale(spec, th[sel], prep=prep)
    sc_lp = sam.log_prob_scale(spec, th[sel], prep=prep)
    lerr = np.abs(logp.cpu().numpy()[sel] - ref_lp) / (1e-5 * np.abs(ref_lp) + ULPS * 2.0**-24 * sc_lp)
    gerr = np.abs(grad.cpu().numpy()[sel] - ref_g) / (1e-5 * np.abs(ref_g) + ULPS * 2.0**-24 * sc_g)
    print(f"S4 shard tolerance use: logp {lerr.max():.3f}, grad {gerr.max():.3f}")
    assert lerr.max() <= 1.0 and gerr.max() <= 1.0, (float(lerr.max()), float(gerr.max()))


def test_device_diagnostics_on_gpu(lb):
    """Bulk-ESS and split-R-hat computed on the GPU that stores the draws equal the host versions."""
    from liesel_b200 import diagnostics

    rng = np.random.default_rng(11)
    n, c, P = 300, 16, 4
    x = np.zeros((n, c, P))
    e = rng.normal(size=(n, c, P))
    for t in range(1, n):
        x[t] = 0.6 * x[t - 1] + e[t]
    t = torch.as_tensor(x, device="cuda")
    np.testing.assert_allclose(diagnostics.ess_bulk_device(t), diagnostics.ess_bulk(x), rtol=1e-8)
    np.testing.assert_allclose(diagnostics.split_rhat_device(t), diagnostics.split_rhat(x), rtol=1e-8)


def test_hmc_on_device(lb):
    """BatchedHMC (HMCKernel mirror) driven by the CUDA log-prob/gradient: adapts to the target
    acceptance and leaves rejected chains untouched."""
    from liesel_b200.nuts import BatchedHMC

    spec, centre = make("s1", np.float32)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 64)
    f = lambda q: plan.logp_grad(q, None)  # noqa: E731
    s = BatchedHMC(f, dev(th, plan), seed=5, num_integration_steps=8)
    s.warmup(150, init_duration=40, term_duration=60, base_duration=25)
    q_before = s.q.clone()
    info = s.transition()
    assert info.num_evals == 8
    kept = ~info.position_moved
    assert torch.equal(s.q[kept], q_before[kept])
    draws, infos = s.sample(60)
    acc = float(torch.stack([i.acceptance_prob for i in infos]).mean())
    assert 0.6 < acc < 0.97
    assert torch.isfinite(draws).all()


# --- BASELINE-size parity of configs 2 and 3 ---------------------------------------------------------
def test_full_size_iwls_s3_all_blocks(lb):
    """BASELINE config 3 at size: the IWLS pass (score, observed information, ridge + Cholesky) of all
    four blocks of the Gaussian location-scale model at n = 1e6 with 256 chains on the device, three of
    them against the fp64 oracle. Same tolerance formula as the n = 5e4 cases — the information
    entries are fp32 sums over up to n/148 rows per accumulator, which small-n parity cannot vouch for."""
    spec, centre = synth.s3_gamlss(n=1_000_000)
    plan = lb.Plan(spec)
    assert plan.variant == "banded_tma_runlength"
    th, aux = synth.evaluation_points(spec, centre, 256, tau2="loguniform")
    for name in ["loc_p0", "loc_np0", "scale_p0", "scale_np0"]:
        check_iwls(spec, plan, name, th, aux, chains=[0, 101, 255])


def test_full_size_s2_against_oracle(lb):
    """BASELINE config 2 at size (n = 1e5, 5 centred cubic P-splines, 64 chains): log-posterior,
    gradient and one block's IWLS pass against the fp64 oracle."""
    spec, centre = synth.s2_pspline_gam(n=100_000)
    plan = lb.Plan(spec)
    th, aux = synth.evaluation_points(spec, centre, 64, tau2="loguniform")
    check_logp_grad(spec, plan, th, aux, chains=[0, 31, 63])
    check_iwls(spec, plan, "loc_np3", th, aux, chains=[0, 63])


def test_chunk_sorted_multi_smooth_experiment_matches_oracle(lb, monkeypatch):
    """``msort_x2_kernel`` (opt-in, ``LSLB_MSORT=1`` at plan creation; slower than ``msmooth_x2`` and
    therefore not the default) must still be CORRECT: S2 with its centred smooths, value + gradient
    against the fp64 oracle, and the same numbers as the default kernel to float rounding."""
    spec, centre = make("s2", np.float32)
    th, aux = synth.evaluation_points(spec, centre, 64, tau2="loguniform")
    plan_default = lb.Plan(spec)
    assert plan_default.variant == "multi_smooth_x2"
    lp0, g0 = plan_default.logp_grad(dev(th, plan_default), dev(aux, plan_default))
    monkeypatch.setenv("LSLB_MSORT", "1")
    plan = lb.Plan(spec)
    monkeypatch.delenv("LSLB_MSORT")
    assert plan.variant == "multi_smooth_chunk_sorted_x2"
    check_logp_grad(spec, plan, th, aux, chains=[0, 31, 63])
    lp1, g1 = plan.logp_grad(dev(th, plan), dev(aux, plan))
    torch.testing.assert_close(lp1, lp0, rtol=2e-6, atol=0)
    scale = g0.abs().amax(dim=0, keepdim=True).clamp_min(1.0)
    assert float(((g1 - g0).abs() / scale).max()) < 2e-5


# --- Metropolis-Hastings step (goose/mh.py) ------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mh_step_decisions_match_the_reference_rule(dtype):
    """``lslb_mh_step`` against ``oracle.iwls.mh_step`` (reference goose/mh.py:48-68) on crafted
    inputs: NaN acceptance (code 90, never accepted unless u == 0), +/-inf log-probabilities, the
    inf - inf = NaN case, ties ``u == acc`` (accepted: the rule is ``<=``), exp overflow clipped to 1."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from liesel_b200.plan import mh_step

    inf, nan = np.inf, np.nan
    cur = np.array([-10.0, -10.0, -10.0, -inf, -inf, -10.0, nan, -10.0, -10.0, 5.0, -1e30, -10.0, -3.0, -3.0], dtype=dtype)
    prop = np.array([-9.0, -11.0, nan, -5.0, -inf, -inf, -4.0, inf, -10.0, 5.0, 1e30, -10.5, -3.0, -4.0], dtype=dtype)
    corr = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, nan, 0.0, 0.0, 0.0, -0.0, 1.0], dtype=dtype)
    u = np.array([0.999, 0.2, 0.0, 0.5, 0.5, 0.0, 0.3, 1.0, 0.1, 1.0, 0.7, 0.0, 1.0, 1.0], dtype=dtype)
    with np.errstate(invalid="ignore"):
        acc_exact = np.minimum(np.exp((prop - cur + corr).astype(dtype)), 1)
    u[11] = acc_exact[11]  # tie at an interior acceptance probability: u == exp(-0.5)
    C, P = cur.shape[0], 7
    rng = np.random.default_rng(0)
    p_cur, p_prop = rng.normal(size=(C, P)).astype(dtype), rng.normal(size=(C, P)).astype(dtype)
    with np.errstate(invalid="ignore"):
        r_code, r_acc, r_accept = OI.mh_step(u, cur, prop, corr)
    d = lambda a: torch.as_tensor(a.copy(), device="cuda")
    d_cur, d_par = d(cur), d(p_cur)
    code, acc, moved = mh_step(d(u), d_cur, d(prop), d(corr), d_par, d(p_prop))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(code.cpu().numpy(), r_code)
    np.testing.assert_allclose(acc.cpu().numpy(), r_acc.astype(dtype), rtol=4e-7 if dtype == np.float32 else 1e-15)
    got = moved.cpu().numpy()
    # the interior tie depends on the last bit of exp(): everything else must agree exactly
    keep = np.arange(C) != 11
    np.testing.assert_array_equal(got[keep], r_accept[keep])
    assert got[11] == (u[11] <= acc.cpu().numpy()[11])
    assert r_code[2] == 90 and r_accept[2] and r_code[4] == 90 and not r_accept[4] and r_accept[5]  # the crafted cases
    # state selection: accepted rows take the proposal (and its log-prob), the others stay
    want_par = np.where(got[:, None], p_prop, p_cur)
    np.testing.assert_array_equal(d_par.cpu().numpy(), want_par)
    np.testing.assert_array_equal(d_cur.cpu().numpy(), np.where(got, prop, cur))


# --- the reference's kernel recovery pin on the device (tests/goose/kernels/model_lm.py:70-84) -----------
def _model_lm_plan(lb, golden_dir):
    import json
    import os

    from liesel_b200.spec import ModelSpec

    g = json.load(open(os.path.join(golden_dir, "model_lm.json")))
    X, y = np.array(g["X"]), np.array(g["y"])
    spec = ModelSpec(dtype=np.dtype(np.float64))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    spec.add_p_smooth(X, 0.0, None, "loc", "beta")
    spec.add_p_smooth(np.ones((30, 1)), 0.0, None, "scale", "log_sigma")
    return g, lb.Plan(spec)


@pytest.mark.parametrize("compact", [False, True])
def test_nuts_on_device_recovers_model_lm(lb, golden_dir, compact):
    """``tests/goose/kernels/test_nuts.py`` through ``model_lm.py:70-84`` with everything on the GPU:
    CUDA log-prob/gradient through the C ABI, the fused leaf kernels (``lslb_nuts_leaf_pre/post``),
    with and without dropping finished chains. The reference's three bounds: posterior mean of beta
    within 5 % of OLS, mean log sigma within 5 % of log sigma_OLS, mean acceptance within 0.05 of the
    dual-averaging target 0.8 (long final fast epoch as in ``model_lm.py:56-60``)."""
    from liesel_b200.nuts import BatchedNUTS

    g, plan = _model_lm_plan(lb, golden_dir)
    C = 64
    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64, device="cuda").repeat(C, 1)

    def f_into(q, lp, gr, chains=None):
        plan.logp_grad(q, None, out=(lp, gr))

    s = BatchedNUTS(lambda q: plan.logp_grad(q.contiguous(), None), q0, seed=42, logp_grad_into=f_into,
                    compact=compact)
    assert s.fused and s.compact == compact
    s.warmup(2000, term_duration=1500)
    draws, infos = s.sample(400)
    d = draws.cpu().numpy()
    assert d[:, :, :2].mean(axis=(0, 1)) == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    acc = np.mean([float(i.acceptance_prob.mean()) for i in infos])
    assert acc == pytest.approx(0.8, abs=0.05), acc
    codes = torch.cat([i.error_code for i in infos])
    assert float((codes == 0).float().mean()) > 0.98
    from liesel_b200 import diagnostics

    assert diagnostics.split_rhat_device(draws).max() < 1.02


def test_hmc_on_device_recovers_model_lm(lb, golden_dir):
    """``tests/goose/kernels/test_hmc.py`` through ``model_lm.py:70-84`` on the device (same bounds)."""
    from liesel_b200.nuts import BatchedHMC

    g, plan = _model_lm_plan(lb, golden_dir)
    C = 64
    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64, device="cuda").repeat(C, 1)
    s = BatchedHMC(lambda q: plan.logp_grad(q.contiguous(), None), q0, seed=7)
    # the reference's own durations (model_lm.py:56-60): the averaged dual-averaging step size sits
    # ABOVE the target by Jensen's inequality, and the gap closes only slowly with the length of the
    # final fast epoch (0.858 after 1500 iterations, 0.832 after 4000 with this driver on CPU)
    s.warmup(5000, term_duration=4000)
    draws, infos = s.sample(600)
    d = draws.cpu().numpy()
    assert d[:, :, :2].mean(axis=(0, 1)) == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    acc = np.mean([float(i.acceptance_prob.mean()) for i in infos])
    assert acc == pytest.approx(0.8, abs=0.05), acc


def test_nuts_gibbs_on_device_converges(lb):
    """The sampling scheme of the ESS metric (one NUTS block + Gibbs tau2, sum-to-zero constraint) on
    the device at a reduced H: chains agree (rank-normalised split-R-hat), no divergences, the
    constrained fit reproduces the data-generating curves."""
    from liesel_b200 import reparam
    from liesel_b200.nuts_gibbs import NutsGibbs, PlanEvaluator, summarise

    spec, centre = synth.s3_gamlss(n=50_000)
    plan = lb.Plan(spec)
    cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1}
    rep = reparam.sum_to_zero(spec, cm, device=plan.device)
    C = 256
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    run = NutsGibbs(spec, PlanEvaluator(plan), dev(th, plan), dev(aux, plan), reparam=rep, seed=1, compact=True)
    run.warmup(300)
    S = 200
    draws, infos = run.sample(S)
    s = summarise(draws, infos, 1.0)
    # split-R-hat of well-mixed chains is bounded below by their LENGTH, not by their number:
    # R-hat^2 ~ 1 + 1 / (ESS per half chain), so the bound follows the measured ESS (1.046 at ~11
    # effective draws per half chain); a chain stuck elsewhere would push it far above that
    ess_half = s["ess_bulk_min"] / (2 * C)
    assert s["split_rhat_max"] < np.sqrt(1.0 + 2.0 / ess_half), s
    assert s["ess_bulk_min"] > 0.05 * C * S, s
    assert s["divergent_frac"] < 0.01 and s["mean_treedepth"] < 7, s
    # fitted mean curve: eta_loc = beta0 + B beta (uncentred basis, full coordinates) ~ 1 + sin(2 pi x)
    q = run.full(draws.reshape(-1, draws.shape[2])).double().mean(dim=0).cpu().numpy()
    from oracle import splines as OSp

    xs = np.linspace(0.05, 0.95, 19)
    blk = spec.block("loc_np0")
    B = OSp.basis_matrix(xs, np.asarray(blk.knots, dtype=np.float64), 3)
    fit = q[0] + B @ q[blk.param_offset: blk.param_offset + blk.ncoef]
    np.testing.assert_allclose(fit, 1.0 + np.sin(2 * np.pi * xs), atol=0.05)


def test_iwls_sampler_converges_under_the_sum_to_zero_constraint(lb):
    """BASELINE config 3's scheme (IWLS blocks + Gibbs tau2) on a reduced S3 with the spline blocks sampled
    in the constrained coordinates ``beta = Z gamma`` (the design ``B Z`` a Liesel user passes): the chains
    agree and the fit reproduces the data-generating curves. Without the constraint the spline level and
    the intercept share a ridge and split-R-hat is ~14."""
    from liesel_b200 import diagnostics, reparam
    from liesel_b200.iwls_sampler import IWLSGibbsSampler

    spec, centre = synth.s3_gamlss(n=50_000)
    plan = lb.Plan(spec)
    C = 64
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    cons = {b.name: reparam.sum_to_zero_basis(plan.column_means(b.name)) for b in spec.blocks if b.kind == 1}
    smp = IWLSGibbsSampler(plan, dev(th, plan), dev(aux, plan), seed=1, constraints=cons)
    # the start values satisfy the constraint: cbar' beta = 0 for every constrained block
    for name in cons:
        b = spec.block(name)
        cbar = torch.as_tensor(plan.column_means(name), device=plan.device, dtype=smp.params.dtype)
        assert float((smp.params[:, b.param_offset: b.param_offset + b.ncoef] @ cbar).abs().max()) < 1e-4
    smp.warmup(150)
    draws, _, infos = smp.sample(300)
    assert float(np.max(diagnostics.split_rhat_device(draws))) < 1.08
    assert float(diagnostics.ess_bulk_device(draws).min()) > 0.04 * C * 300  # measured ~0.07 per draw
    for b in smp.blocks:
        acc = np.mean([float(i[b.name]["acceptance_prob"].mean()) for i in infos])
        assert 0.6 < acc < 0.97, (b.name, acc)
    q = draws.reshape(-1, draws.shape[2]).double().mean(dim=0).cpu().numpy()
    from oracle import splines as OSp

    xs = np.linspace(0.05, 0.95, 19)
    blk = spec.block("loc_np0")
    B = OSp.basis_matrix(xs, np.asarray(blk.knots, dtype=np.float64), 3)
    fit = q[0] + B @ q[blk.param_offset: blk.param_offset + blk.ncoef]
    np.testing.assert_allclose(fit, 1.0 + np.sin(2 * np.pi * xs), atol=0.05)


# --- asynchronous NUTS driver: the CUDA state machine against its torch statement --------------------------
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_nuts_async_kernel_matches_torch_statement_step_by_step(dtype):
    """``lslb_nuts_async_step`` against ``AsyncNUTS._step_torch`` on the same device, same state, same
    random inputs, step after step through warm-up epochs (dual averaging, mass-matrix windows), Gibbs
    updates of an auxiliary scale, draws: integer state identical, floats to rounding. (Random numbers
    are inputs of a step, so the comparison is deterministic.) The two states are re-synchronised after
    every comparison so that a last-bit difference cannot grow and flip a later accept decision."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from liesel_b200.nuts_async import AsyncNUTS

    C, P = 48, 7
    d = torch.device("cuda")
    scale = torch.linspace(0.5, 3.0, P, device=d, dtype=dtype)
    M = torch.diag(1.0 / scale).to(dtype)

    def make(backend):
        aux = torch.ones((C, 2), dtype=dtype, device=d)

        def f_into(q, lp, g):  # N(0, tau2 * diag(scale)) with tau2 = aux[:, 1]
            t = aux[:, 1:2]
            lp.copy_((-0.5 * q * q / (scale * t)).sum(dim=1) - 0.5 * P * torch.log(t[:, 0]))
            g.copy_(-q / (scale * t))

        q0 = torch.linspace(-1, 1, C * P, device=d, dtype=dtype).reshape(C, P).contiguous()
        gib = [dict(M=M, b0=1.5, a_post=2.0 + 0.5 * P, slot=1)]
        return AsyncNUTS(f_into, q0, warmup=60, samples=25, seed=3, aux=aux, gibbs=gib, backend=backend,
                         initial_step_size=0.3, max_treedepth=5,
                         epoch_kw=dict(init_duration=15, term_duration=15, base_duration=10))

    A, B = make("torch"), make("cuda")
    keys_f = ["q", "lp", "g", "step", "inv_mass", "h0", "qp", "lpp", "log_w", "rsum", "qe", "re", "ge", "qs", "logw_s",
              "rsum_s", "qn", "r_half", "da_err", "da_logavg", "da_mu", "sum_acc", "n_leaf"]
    keys_i = ["phase", "depth", "dir", "i_leaf", "n_leaves", "it", "ep", "t_ep", "flags"]
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    for stage, S_act in (("warmup", 0), ("posterior", 25)):
        for X in (A, B):
            X._active_S = S_act
            X._ptrs = None
            X.st["phase"].fill_(0)
            X.st["n_done"].zero_()
            X.st["qn"].copy_(X.st["q"])
            X._evaluate()
        for step in range(6000):
            A._randoms()
            B._Z.copy_(A._Z); B._U.copy_(A._U); B._G.copy_(A._G)
            A._step(); B._step()
            for k in keys_i:
                assert torch.equal(A.st[k], B.st[k]), (stage, step, k)
            leaf = A.st["phase"] == 1
            for k in keys_f:
                a, b = A.st[k].double(), B.st[k].double()
                if k == "r_half":  # scratch outside a leaf (the kernel parks r_n there)
                    a, b = a[leaf], b[leaf]
                fin = torch.isfinite(a)
                assert torch.equal(fin, torch.isfinite(b)), (stage, step, k)
                err = ((a - b).abs() / (1.0 + a.abs()))[fin]
                assert err.numel() == 0 or float(err.max()) < tol, (stage, step, k, float(err.max()))
            assert torch.allclose(A.aux.double(), B.aux.double(), rtol=tol * 10, atol=0)
            # re-synchronise: Hamiltonian trajectories amplify a last-bit difference exponentially (1e-11
            # after 128 steps in float64), and an accept decision would eventually flip
            if True:
                for k in keys_f + ["gs", "gp", "qL", "rL", "gL", "qR", "rR", "gR", "lps", "ck_r", "ck_s", "ws1", "ws2"]:
                    B.st[k].copy_(A.st[k])
                B.aux.copy_(A.aux)
            A._evaluate(); B._evaluate()
            if int(A.st["n_done"].item()) >= C:
                assert int(B.st["n_done"].item()) >= C
                break
        else:
            raise AssertionError("stage did not finish")
    assert torch.allclose(A.draws().double(), B.draws().double(), rtol=tol * 10, atol=tol * 10)
    assert torch.allclose(A.st["stats"], B.st["stats"], rtol=1e-6, atol=1e-6)
    assert int(A.st["it"].min()) == 85


def test_nuts_async_on_device_recovers_model_lm(lb, golden_dir):
    """The reference's kernel integration pins (``model_lm.py:70-84``) with the asynchronous driver on
    the GPU: CUDA log-prob/gradient through the C ABI + the state-machine kernel."""
    from liesel_b200.nuts_async import AsyncNUTS

    g, plan = _model_lm_plan(lb, golden_dir)
    C = 64
    q0 = torch.tensor([g["beta"] + [g["log_sigma"]]], dtype=torch.float64, device="cuda").repeat(C, 1)

    def f_into(q, lp, gr):
        plan.logp_grad(q, None, out=(lp, gr))

    s = AsyncNUTS(f_into, q0, warmup=2000, samples=400, seed=42, epoch_kw=dict(term_duration=1500))
    assert s.backend == "cuda"
    t = s.run()
    d = s.draws().cpu().numpy()
    assert d[:, :, :2].mean(axis=(0, 1)) == pytest.approx(np.array(g["beta_ols"]), rel=0.05)
    assert d[:, :, 2].mean() == pytest.approx(np.log(g["sigma_ols"]), rel=0.05)
    sm = s.summary()
    assert sm["mean_accept"] == pytest.approx(0.8, abs=0.05), sm
    assert sm["divergent_frac"] < 0.02
    from liesel_b200 import diagnostics

    assert diagnostics.split_rhat_device(s.draws()).max() < 1.02
    # the batch stays full: sequential evaluations ~ transitions x mean leapfrogs, not x deepest tree
    assert t["steps_warmup"] + t["steps_posterior"] < 2400 * 16


def test_async_nuts_gibbs_on_device_converges(lb):
    """The ESS metric's sampling scheme on the asynchronous driver at a reduced H (see
    ``test_nuts_gibbs_on_device_converges`` for the lock-step counterpart)."""
    from liesel_b200 import reparam
    from liesel_b200.nuts_gibbs import AsyncNutsGibbs, PlanEvaluator, summarise_async

    spec, centre = synth.s3_gamlss(n=50_000)
    plan = lb.Plan(spec)
    cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1}
    rep = reparam.sum_to_zero(spec, cm, device=plan.device)
    C, S = 256, 200
    th, aux = synth.evaluation_points(spec, centre, C, seed=1)
    run = AsyncNutsGibbs(spec, PlanEvaluator(plan), dev(th, plan), dev(aux, plan), 300, S, reparam=rep, seed=1)
    t = run.run()
    s = summarise_async(run, t)
    ess_half = s["ess_bulk_min"] / (2 * C)
    assert s["split_rhat_max"] < np.sqrt(1.0 + 2.0 / ess_half), s
    assert s["ess_bulk_min"] > 0.05 * C * S, s
    assert s["divergent_frac"] < 0.01 and s["mean_treedepth"] < 7, s
    q = run.full(run.draws().reshape(-1, run.draws().shape[2])).double().mean(dim=0).cpu().numpy()
    from oracle import splines as OSp

    xs = np.linspace(0.05, 0.95, 19)
    blk = spec.block("loc_np0")
    B = OSp.basis_matrix(xs, np.asarray(blk.knots, dtype=np.float64), 3)
    fit = q[0] + B @ q[blk.param_offset: blk.param_offset + blk.ncoef]
    np.testing.assert_allclose(fit, 1.0 + np.sin(2 * np.pi * xs), atol=0.05)


# --- B-spline orders below 3 (contrib/splines.py:109-198 is order-generic) -----------------------------
@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_basis_other_orders_bit_exact(lb, order, dtype):
    """The device band is four wide for every order; scattered into the dense [n, k] matrix it must be
    bit-identical to the oracle's dense recursion (index AND weights), adversarial points included."""
    from liesel_b200 import splines

    rng = np.random.default_rng(4 + order)
    x = rng.uniform(size=50_000).astype(dtype)
    knots = splines.equidistant_knots(x, 12, order)
    edge = np.concatenate([knots, np.nextafter(knots, dtype(np.inf)), np.nextafter(knots, dtype(-np.inf))])
    x = np.concatenate([x, edge.astype(dtype), [knots[0] - 1, knots[-1] + 1]]).astype(dtype)
    start, w = splines.basis_matrix(x, knots, order, outer_ok=True, device="cuda")
    start, w = start.cpu().numpy(), w.cpu().numpy()
    k = knots.shape[0] - order - 1
    assert w.shape == (x.shape[0], 4) and start.min() >= 0 and start.max() <= k - 4
    dense = OS.band_to_dense(start, w, k)
    assert np.array_equal(dense, OS.basis_matrix(x, knots, order))
    assert (np.count_nonzero(w, axis=1) <= order + 1).all()


@pytest.mark.parametrize("order,C", [(1, 40), (2, 40), (2, 320)])
def test_logp_grad_lower_order_splines(lb, order, C):
    """Location-scale model with linear / quadratic P-splines through the scalar run-length kernel
    (C = 40) and the packed kernel (C = 320), against the oracle (which builds its own order + 1 wide
    band from the same x and knots)."""
    from liesel_b200.spec import ModelSpec
    from liesel_b200.splines import equidistant_knots, pspline_penalty

    rng = np.random.default_rng(11)
    n, k = 30_000, 14
    x = rng.uniform(size=n)
    y = rng.normal(1.0 + np.sin(2 * np.pi * x), np.exp(-0.5 + 0.5 * np.cos(2 * np.pi * x)))
    spec = ModelSpec(dtype=np.dtype(np.float32))
    spec.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    xd = x.astype(np.float32)
    knots = equidistant_knots(xd, k, order)
    K = pspline_penalty(k, 2)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "loc", "loc_p0")
    spec.add_spline_smooth(xd, knots, K, 1.0, 0.005, "loc", "loc_np0", order=order, tau2_init=1.0)
    spec.add_p_smooth(np.ones((n, 1)), 0.0, 100.0, "scale", "scale_p0")
    spec.add_spline_smooth(xd, knots, K, 1.0, 0.005, "scale", "scale_np0", order=order, tau2_init=1.0)
    plan = lb.Plan(spec)
    assert plan.variant == "banded_tma_runlength"
    rng2 = np.random.default_rng(12)
    th = (0.2 * rng2.normal(size=(C, spec.num_params))).astype(np.float32)
    th[:, 0] += 1.0
    aux = np.exp(rng2.uniform(-2, 2, size=(C, 2))).astype(np.float32)
    check_logp_grad(spec, plan, th, aux, chains=[0, 1, C // 2, C - 1])
    check_iwls(spec, plan, "loc_np0", th, aux, chains=[0, C - 1])
