"""
The jax.ffi shim (liesel_b200/csrc/xla_ffi_shim.cc) compiled against the test stub of
``xla/ffi/api/ffi.h``: it builds on every build (CPU test: symbols + argument checks that need no
GPU), and on the GPU its handlers are driven through the stub call frame and must return exactly
what the C ABI returns through ctypes — the reference-side binding of INTEGRATION.md §3.
"""
import os

import numpy as np
import pytest
import torch

import ffi_stub_driver as drv
from liesel_b200 import synth


def test_shim_builds_and_exports_every_handler():
    assert os.path.exists(drv.STUB_SO), "run `bash liesel_b200/csrc/build.sh`"
    lib = drv.load()
    for name in drv.HANDLERS:
        assert hasattr(lib, name), name


def test_shim_rejects_bad_calls_without_touching_the_gpu():
    lib = drv.load()
    buf = lambda shape, code=drv.F32: (code, None, shape)
    # null plan handle
    code, msg = drv.call(lib, "LslbLogpGrad", 0, [buf((4, 3)), buf((4, 0))], [buf((4,)), buf((4, 3))])
    assert code == drv.INVALID_ARGUMENT and "plan" in msg
    # operand dtype that does not match the target
    code, msg = drv.call(lib, "LslbLogpGrad", 0, [buf((4, 3), drv.F64), buf((4, 0))], [buf((4,)), buf((4, 3))])
    assert code == drv.INVALID_ARGUMENT and "dtype" in msg
    # wrong operand count
    code, msg = drv.call(lib, "LslbLogp", 0, [buf((4, 3))], [buf((4,))])
    assert code == drv.INVALID_ARGUMENT


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_shim_handlers_equal_the_c_abi(dtype):
    from liesel_b200.plan import Plan

    lib = drv.load()
    sfx = "" if dtype == np.float32 else "F64"
    spec, centre = synth.s3_gamlss(n=20_000, dtype=dtype)
    plan = Plan(spec)
    C_ = 48
    th, aux = synth.evaluation_points(spec, centre, C_, tau2="loguniform")
    t, a = torch.as_tensor(th, device=plan.device), torch.as_tensor(aux, device=plan.device)
    lp0, g0 = plan.logp_grad(t, a)
    s0, i0, c0 = plan.iwls("scale_np0", t, a)
    q0 = plan.quadform("scale_np0", t)
    torch.cuda.synchronize()
    h = plan._handle.value
    lp, g = torch.empty_like(lp0), torch.empty_like(g0)
    code, msg = drv.call(lib, "LslbLogpGrad" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)],
                         [drv.tensor_buf(lp), drv.tensor_buf(g)])
    assert code == drv.OK, msg
    assert torch.equal(lp, lp0) and torch.equal(g, g0)
    lpv = torch.empty_like(lp0)
    code, msg = drv.call(lib, "LslbLogp" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)], [drv.tensor_buf(lpv)])
    assert code == drv.OK, msg
    np.testing.assert_allclose(lpv.cpu().numpy(), lp0.cpu().numpy(), rtol=1e-6 if dtype == np.float32 else 1e-12)
    b = spec.block_index("scale_np0")
    s, i, c = torch.empty_like(s0), torch.empty_like(i0), torch.empty_like(c0)
    code, msg = drv.call(lib, "LslbIwls" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)],
                         [drv.tensor_buf(s), drv.tensor_buf(i), drv.tensor_buf(c)], block=b)
    assert code == drv.OK, msg
    assert torch.equal(s, s0) and torch.equal(i, i0) and torch.equal(c, c0)
    q = torch.empty_like(q0)
    code, msg = drv.call(lib, "LslbQuadform" + sfx, h, [drv.tensor_buf(t)], [drv.tensor_buf(q)], block=b)
    assert code == drv.OK, msg
    assert torch.equal(q, q0)
    # the other dtype's target refuses this plan; wrong shapes are refused; nothing is written
    other = "F64" if dtype == np.float32 else ""
    code, msg = drv.call(lib, "LslbLogpGrad" + other, h,
                         [(drv.F64 if other else drv.F32, t.data_ptr(), t.shape), (drv.F64 if other else drv.F32, a.data_ptr(), a.shape)],
                         [(drv.F64 if other else drv.F32, lp.data_ptr(), lp.shape), (drv.F64 if other else drv.F32, g.data_ptr(), g.shape)])
    assert code == drv.INVALID_ARGUMENT and "dtype" in msg
    code, msg = drv.call(lib, "LslbLogpGrad" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)],
                         [drv.tensor_buf(lp), drv.tensor_buf(g, (C_, plan.P + 1))])
    assert code == drv.INVALID_ARGUMENT and "grad" in msg
    code, msg = drv.call(lib, "LslbIwls" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)],
                         [drv.tensor_buf(s), drv.tensor_buf(i), drv.tensor_buf(c)], block=99)
    assert code == drv.INVALID_ARGUMENT and "block" in msg


@pytest.mark.gpu
def test_shim_model_without_aux():
    """A == 0 (no smoothing variance): XLA hands over a [C, 0] operand whose data pointer may be NULL."""
    from liesel_b200.plan import Plan

    lib = drv.load()
    spec, centre = synth.s1_blr()
    plan = Plan(spec)
    th, _ = synth.evaluation_points(spec, centre, 4)
    t = torch.as_tensor(th, device=plan.device)
    lp0, g0 = plan.logp_grad(t, None)
    lp, g = torch.empty_like(lp0), torch.empty_like(g0)
    code, msg = drv.call(lib, "LslbLogpGrad", plan._handle.value, [drv.tensor_buf(t), (drv.F32, None, (4, 0))],
                         [drv.tensor_buf(lp), drv.tensor_buf(g)])
    assert code == drv.OK, msg
    assert torch.equal(lp, lp0) and torch.equal(g, g0)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_shim_group_iwls_handler_equals_the_c_abi(dtype):
    """`LslbIwlsGroup`: score and diagonal information of the random-intercept block through the stub call
    frame, bit for bit what `Plan.iwls_group` (ctypes) returns; shape checks without touching the GPU."""
    from liesel_b200.plan import Plan

    lib = drv.load()
    sfx = "" if dtype == np.float32 else "F64"
    spec, centre = synth.s5_poisson_ri(n=20_000, G=300, dtype=dtype, balanced=False)
    plan = Plan(spec)
    C_ = 21
    th, _ = synth.evaluation_points(spec, centre, C_)
    t = torch.as_tensor(th, device=plan.device)
    a = torch.empty((C_, 0), dtype=t.dtype, device=plan.device)
    s0, i0 = plan.iwls_group(t)
    torch.cuda.synchronize()
    h = plan._handle.value
    s1, i1 = torch.empty_like(s0), torch.empty_like(i0)
    code, msg = drv.call(lib, "LslbIwlsGroup" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)],
                         [drv.tensor_buf(s1), drv.tensor_buf(i1)])
    assert code == drv.OK, msg
    torch.cuda.synchronize()
    assert torch.equal(s1, s0) and torch.equal(i1, i0)
    code, msg = drv.call(lib, "LslbIwlsGroup" + sfx, h, [drv.tensor_buf(t), drv.tensor_buf(a)],
                         [drv.tensor_buf(s1), drv.tensor_buf(i1, shape=(C_, 299))])
    assert code == drv.INVALID_ARGUMENT and "info_diag" in msg
