"""
``liesel_b200.from_liesel.spec_from_model`` on stand-ins that carry exactly the public attributes of
the reference's ``Var`` / ``Dist`` / ``Calc`` / ``Group`` / ``Model`` objects which the matcher reads
(Liesel itself needs JAX and Python >= 3.13, neither is available here). The stand-in builder below
replays ``DistRegBuilder`` (reference ``src/liesel/model/distreg.py:64-274``) call for call: same
variable names, same group members, same default values.
"""
from types import MappingProxyType, SimpleNamespace

import numpy as np
import pytest

from liesel_b200 import from_liesel as FL
from liesel_b200 import splines, synth
from liesel_b200.spec import ModelSpec
from oracle import sam
from oracle import splines as OS


class Node(SimpleNamespace):
    pass


def _var(name, value, dist=None, inputs=(), function=None):
    v = SimpleNamespace(name=name, value=value, dist_node=dist, groups={})
    v.value_node = Node(name=name + "_value", inputs=tuple(inputs), function=function, var=v)
    return v


class Group:
    def __init__(self, name, **members):
        self.name, self._m = name, members

    def __contains__(self, key):
        return key in self._m

    def __getitem__(self, key):
        return self._m[key]


class Identity:
    def forward(self, x):
        return x


class Exp:
    def forward(self, x):
        return np.exp(x)


class Softplus:
    def forward(self, x):
        return np.log1p(np.exp(x))


def _tfd(name):
    return type(name, (), {})


class FakeDistReg:
    """Replays DistRegBuilder: add_response -> add_predictor -> add_p_smooth / add_np_smooth."""

    def __init__(self):
        self.vars, self.groups_, self.pdt, self.par = {}, {}, {}, {}

    def add_response(self, y, dist_name):
        d = Node(distribution=_tfd(dist_name), kwinputs={})
        self.vars["response"] = _var("response", y, dist=d)
        return self

    def add_predictor(self, name, link):
        pdt = _var(name + "_pdt", None)
        par = _var(name, None, inputs=[pdt.value_node], function=link().forward)
        self.vars[pdt.name], self.vars[par.name] = pdt, par
        self.vars["response"].dist_node.kwinputs[name] = par.value_node
        self.pdt[name] = pdt
        return self

    def _attach(self, predictor, smooth):
        pdt = self.pdt[predictor]
        pdt.value_node.inputs = pdt.value_node.inputs + (smooth.value_node,)

    def add_p_smooth(self, X, m, s, predictor, name):
        Xv, mv, sv = _var(name + "_X", X), _var(name + "_m", m), _var(name + "_s", s)
        beta = _var(name + "_beta", np.zeros(X.shape[1], np.float32))
        smooth = _var(name, None, inputs=[Xv.value_node, beta.value_node])
        for v in (Xv, mv, sv, beta, smooth):
            self.vars[v.name] = v
        self._attach(predictor, smooth)
        self.groups_[name] = Group(name, smooth=smooth, beta=beta, X=Xv, m=mv, s=sv)
        return self

    def add_np_smooth(self, X, K, a, b, predictor, name):
        Xv, Kv, av, bv = _var(name + "_X", X), _var(name + "_K", K), _var(name + "_a", a), _var(name + "_b", b)
        rank = _var(name + "_rank", float(np.linalg.matrix_rank(K)))
        tau2 = _var(name + "_tau2", 10000.0)
        beta = _var(name + "_beta", np.zeros(X.shape[1], np.float32))
        smooth = _var(name, None, inputs=[Xv.value_node, beta.value_node])
        for v in (Xv, Kv, av, bv, rank, tau2, beta, smooth):
            self.vars[v.name] = v
        self._attach(predictor, smooth)
        self.groups_[name] = Group(name, smooth=smooth, beta=beta, tau2=tau2, rank=rank, X=Xv, K=Kv, a=av, b=bv)
        return self

    def build_model(self):
        return SimpleNamespace(vars=MappingProxyType(self.vars), groups=lambda: dict(self.groups_))


def _gamlss_model(n=400, k=12, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=n).astype(np.float32)
    knots = splines.equidistant_knots(x, k, 3)
    B = OS.basis_matrix(x.astype(np.float64), knots.astype(np.float64), 3).astype(np.float32)
    K = splines.pspline_penalty(k, 2)
    y = (np.sin(6 * x) + 0.3 * rng.normal(size=n)).astype(np.float32)
    Z = np.column_stack([np.ones(n), rng.normal(size=n)]).astype(np.float32)
    b = FakeDistReg().add_response(y, "Normal").add_predictor("loc", Identity).add_predictor("scale", Exp)
    b.add_p_smooth(Z, 0.0, 100.0, "loc", "loc_p0").add_np_smooth(B, K, 1.0, 0.005, "loc", "loc_np0")
    b.add_p_smooth(np.ones((n, 1), np.float32), 0.0, 100.0, "scale", "scale_p0")
    b.add_np_smooth(B, K, 1.0, 0.005, "scale", "scale_np0")
    return b.build_model(), (x, knots, B, K, y, Z)


def test_band_from_dense_is_exact():
    x = np.linspace(0, 1, 300).astype(np.float32)
    knots = splines.equidistant_knots(x, 10, 3)
    B = OS.basis_matrix(x.astype(np.float64), knots.astype(np.float64), 3).astype(np.float32)
    start, w = FL.band_from_dense(B)
    assert start.dtype == np.int32 and w.shape == (300, 4)
    np.testing.assert_array_equal(OS.band_to_dense(start, w, 10), B)  # bit-exact round trip
    assert start.min() >= 0 and start.max() <= 6
    assert FL.band_from_dense(np.random.default_rng(0).normal(size=(50, 10))) is None
    assert FL.band_from_dense(np.ones((5, 3))) is None


def test_spec_from_model_matches_hand_built_spec():
    model, (x, knots, B, K, y, Z) = _gamlss_model()
    spec = FL.spec_from_model(model)
    assert spec.position_keys() == ["loc_p0_beta", "loc_np0_beta", "scale_p0_beta", "scale_np0_beta"]
    assert spec.aux_keys() == ["loc_np0_tau2", "scale_np0_tau2"]
    kinds = [b.kind for b in spec.blocks]
    assert kinds == [0, 1, 3, 1]  # dense, banded spline, intercept, banded spline
    assert np.allclose(spec.default_aux(), 10000.0)  # distreg.py:163
    hand = ModelSpec(dtype=np.float32)
    hand.add_response(y, "normal").add_predictor("loc").add_predictor("scale")
    hand.add_p_smooth(Z, 0.0, 100.0, "loc", "loc_p0").add_np_smooth(B, K, 1.0, 0.005, "loc", "loc_np0")
    hand.add_p_smooth(np.ones((len(y), 1)), 0.0, 100.0, "scale", "scale_p0")
    hand.add_np_smooth(B, K, 1.0, 0.005, "scale", "scale_np0")
    rng = np.random.default_rng(1)
    th = 0.1 * rng.normal(size=(3, spec.num_params))
    aux = np.exp(rng.normal(size=(3, 2)))
    np.testing.assert_allclose(sam.log_prob(spec, th, aux), sam.log_prob(hand, th, aux), rtol=1e-12)
    np.testing.assert_allclose(sam.grad(spec, th, aux), sam.grad(hand, th, aux), rtol=1e-10, atol=1e-10)
    pos = FL.initial_position(model, spec)
    assert set(pos) == set(spec.position_keys() + spec.aux_keys()) and pos["loc_np0_tau2"] == 10000.0
    dense = FL.spec_from_model(model, detect_bands=False)
    assert [b.kind for b in dense.blocks] == [0, 0, 3, 0]
    np.testing.assert_allclose(sam.log_prob(dense, th, aux), sam.log_prob(hand, th, aux), rtol=1e-12)


def test_spec_from_model_rejects_what_the_path_does_not_evaluate():
    model, _ = _gamlss_model()
    model.vars["response"].dist_node.distribution = _tfd("Gamma")
    with pytest.raises(FL.UnsupportedModelError):
        FL.spec_from_model(model)
    b = FakeDistReg().add_response(np.zeros(10, np.float32), "Poisson").add_predictor("log_rate", Softplus)
    b.add_p_smooth(np.ones((10, 2), np.float32), 0.0, 10.0, "log_rate", "log_rate_p0")
    with pytest.raises(FL.UnsupportedModelError):
        FL.spec_from_model(b.build_model())
    b = FakeDistReg().add_response(np.zeros(10, np.float32), "Poisson").add_predictor("log_rate", Identity)
    with pytest.raises(FL.UnsupportedModelError):  # no smooths
        FL.spec_from_model(b.build_model())
    b.add_p_smooth(np.ones((10, 2), np.float32), 0.0, 10.0, "log_rate", "log_rate_p0")
    extra = _var("user_term", None)
    b._attach("log_rate", extra)  # a user Calc on the predictor that is not a DistReg smooth
    with pytest.raises(FL.UnsupportedModelError):
        FL.spec_from_model(b.build_model())
