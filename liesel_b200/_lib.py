"""
ctypes binding of ``libliesel_b200.so`` (C ABI declared in ``include/liesel_b200.h``).

There is NO fallback: if the shared library is missing, importing this module raises, and
every wrapper raises on a non-zero status code. PyTorch appears only as the owner of device
memory and of the CUDA stream.
"""

from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LSLB_LIB_PATH: A/B measurements of two builds of the same library (tools only)
LIB_PATH = os.environ.get("LSLB_LIB_PATH") or os.path.join(_HERE, "csrc", "libliesel_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `bash liesel_b200/csrc/build.sh` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "liesel_b200 has no CPU or PyTorch fallback."
    )

lib = C.CDLL(LIB_PATH)

MAX_BLOCKS = 16


class BlockDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("predictor", C.c_int32), ("ncoef", C.c_int32), ("prior", C.c_int32),
        ("data", C.c_void_p), ("index", C.c_void_p), ("ld", C.c_int64),
        ("prior_m", C.c_double), ("prior_s", C.c_double), ("K", C.c_void_p),
        ("rank", C.c_double), ("log_pdet", C.c_double),
        ("tau2_slot", C.c_int32), ("has_ig", C.c_int32),
        ("tau2_fixed", C.c_double), ("ig_a", C.c_double), ("ig_b", C.c_double),
        ("center", C.c_void_p),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("family", C.c_int32), ("n", C.c_int64), ("y", C.c_void_p),
        ("fixed_scale", C.c_double), ("const_term", C.c_double),
        ("nblocks", C.c_int32), ("reserved", C.c_int32), ("blocks", C.POINTER(BlockDesc)),
    ]


_vp, _i, _u32, _sz, _i64 = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t, C.c_int64

lib.lslb_version.restype = _i
lib.lslb_status_string.restype = C.c_char_p
lib.lslb_status_string.argtypes = [_i]
lib.lslb_last_cuda_error.restype = C.c_char_p
lib.lslb_plan_create.restype = _i
lib.lslb_plan_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(_vp)]
lib.lslb_plan_destroy.restype = None
lib.lslb_plan_destroy.argtypes = [_vp]
lib.lslb_plan_num_params.restype = _i
lib.lslb_plan_num_params.argtypes = [_vp]
lib.lslb_plan_num_aux.restype = _i
lib.lslb_plan_num_aux.argtypes = [_vp]
lib.lslb_plan_dtype.restype = _i
lib.lslb_plan_dtype.argtypes = [_vp]
lib.lslb_plan_block_ncoef.restype = _i
lib.lslb_plan_block_ncoef.argtypes = [_vp, _i]
lib.lslb_plan_variant.restype = C.c_char_p
lib.lslb_plan_variant.argtypes = [_vp]
lib.lslb_plan_kernel_name.restype = C.c_char_p
lib.lslb_plan_kernel_name.argtypes = [_vp, _i]
lib.lslb_workspace_bytes.restype = _sz
lib.lslb_workspace_bytes.argtypes = [_vp, _i]
for _sfx in ("f32", "f64"):
    f = getattr(lib, "lslb_logp_grad_" + _sfx)
    f.restype = _i
    f.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _u32, _vp, _sz, _vp]
    f = getattr(lib, "lslb_iwls_" + _sfx)
    f.restype = _i
    f.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _sz, _vp]
    f = getattr(lib, "lslb_chol_" + _sfx)
    f.restype = _i
    f.argtypes = [_i, _i, _vp, _vp, _vp]
    f = getattr(lib, "lslb_quadform_" + _sfx)
    f.restype = _i
    f.argtypes = [_vp, _i, _i, _vp, _vp, _vp]
    f = getattr(lib, "lslb_iwls_propose_" + _sfx)
    f.restype = _i
    f.argtypes = [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    f = getattr(lib, "lslb_mh_step_" + _sfx)
    f.restype = _i
    f.argtypes = [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    f = getattr(lib, "lslb_iwls_group_" + _sfx)
    f.restype = _i
    f.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _vp, _u32, _vp]
    f = getattr(lib, "lslb_iwls_propose_diag_" + _sfx)
    f.restype = _i
    f.argtypes = [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]
    f = getattr(lib, "lslb_spline_basis_" + _sfx)
    f.restype = _i
    f.argtypes = [_vp, _i64, _vp, _i, _vp, _vp, _vp]
    f = getattr(lib, "lslb_spline_basis_order_" + _sfx)
    f.restype = _i
    f.argtypes = [_vp, _i64, _vp, _i, _i, _vp, _vp, _vp]

for _sfx in ("f32", "f64"):
    # row-sharded variants: one more pointer (lslb_peer* / ncclComm_t) before the stream
    for _tr in ("peer", "nccl"):
        f = getattr(lib, f"lslb_logp_grad_{_tr}_{_sfx}")
        f.restype = _i
        f.argtypes = [_vp, _i, _vp, _vp, _vp, _vp, _u32, _vp, _sz, _vp, _vp]
        f = getattr(lib, f"lslb_iwls_{_tr}_{_sfx}")
        f.restype = _i
        f.argtypes = [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _sz, _vp, _vp]
lib.lslb_peer_create.restype = _i
lib.lslb_peer_create.argtypes = [_i, _i, _sz, C.POINTER(_vp), _vp]
lib.lslb_peer_connect.restype = _i
lib.lslb_peer_connect.argtypes = [_vp, _vp]
lib.lslb_peer_destroy.restype = None
lib.lslb_peer_destroy.argtypes = [_vp]
lib.lslb_peer_timed_out.restype = _i
lib.lslb_peer_timed_out.argtypes = [_vp, C.POINTER(_i)]
lib.lslb_peer_abort.restype = _i
lib.lslb_peer_abort.argtypes = [_vp, _vp]
lib.lslb_peer_reset.restype = _i
lib.lslb_peer_reset.argtypes = [_vp]
lib.lslb_nccl_unique_id.restype = _i
lib.lslb_nccl_unique_id.argtypes = [_vp]
lib.lslb_nccl_comm_create.restype = _i
lib.lslb_nccl_comm_create.argtypes = [_i, _i, _vp, C.POINTER(_vp)]
lib.lslb_nccl_comm_destroy.restype = None
lib.lslb_nccl_comm_destroy.argtypes = [_vp]
PEER_HANDLE_BYTES = 64
NCCL_ID_BYTES = 128

for _sfx in ("f32", "f64"):
    f = getattr(lib, "lslb_nuts_leaf_pre_" + _sfx)
    f.restype = _i
    f.argtypes = [C.POINTER(_vp), C.POINTER(_i), _vp]
    f = getattr(lib, "lslb_nuts_leaf_post_" + _sfx)
    f.restype = _i
    f.argtypes = [C.POINTER(_vp), C.POINTER(_i), C.c_double, _vp]
    f = getattr(lib, "lslb_nuts_async_step_" + _sfx)
    f.restype = _i
    f.argtypes = [C.POINTER(_vp), C.POINTER(_i), C.POINTER(C.c_double), _vp]

lib.lslb_debug_set_events.restype = _i
lib.lslb_debug_set_events.argtypes = [_vp, _vp]
lib.lslb_debug_event_count.restype = _i

BENCH_LIB_PATH = os.path.join(os.path.dirname(LIB_PATH), "libliesel_b200_bench.so")
BENCH_EXPORTS = ["lslb_debug_pipe_peak", "lslb_debug_tc_probe"]
_bench_lib = None


def bench_lib():
    """The TOOLS library (pipe-peak micro-benchmarks, tcgen05 probe; ``include/liesel_b200_bench.h``),
    loaded on first use by bench.py and tools/ — the product path never touches it."""
    global _bench_lib
    if _bench_lib is None:
        b = C.CDLL(BENCH_LIB_PATH)
        b.lslb_debug_pipe_peak.restype = _i
        b.lslb_debug_pipe_peak.argtypes = [_i, _i, _i, _vp, C.POINTER(C.c_double), _vp]
        b.lslb_debug_tc_probe.restype = _i
        b.lslb_debug_tc_probe.argtypes = [_vp, _vp, _i, _i, _i, _vp, _vp]
        _bench_lib = b
    return _bench_lib

EXPORTS = [
    "lslb_peer_create", "lslb_peer_connect", "lslb_peer_destroy", "lslb_peer_timed_out", "lslb_peer_abort", "lslb_peer_reset",
    "lslb_logp_grad_peer_f32", "lslb_logp_grad_peer_f64", "lslb_iwls_peer_f32", "lslb_iwls_peer_f64",
    "lslb_nccl_unique_id", "lslb_nccl_comm_create", "lslb_nccl_comm_destroy",
    "lslb_logp_grad_nccl_f32", "lslb_logp_grad_nccl_f64", "lslb_iwls_nccl_f32", "lslb_iwls_nccl_f64",
    "lslb_debug_set_events", "lslb_debug_event_count",
    "lslb_nuts_leaf_pre_f32", "lslb_nuts_leaf_pre_f64", "lslb_nuts_leaf_post_f32", "lslb_nuts_leaf_post_f64",
    "lslb_nuts_async_step_f32", "lslb_nuts_async_step_f64",
    "lslb_version", "lslb_status_string", "lslb_last_cuda_error", "lslb_plan_create",
    "lslb_plan_destroy", "lslb_plan_num_params", "lslb_plan_num_aux", "lslb_plan_dtype", "lslb_plan_block_ncoef", "lslb_plan_variant", "lslb_plan_kernel_name",
    "lslb_workspace_bytes", "lslb_logp_grad_f32", "lslb_logp_grad_f64", "lslb_iwls_f32",
    "lslb_iwls_f64", "lslb_chol_f32", "lslb_chol_f64", "lslb_quadform_f32", "lslb_quadform_f64",
    "lslb_iwls_group_f32", "lslb_iwls_group_f64", "lslb_iwls_propose_diag_f32", "lslb_iwls_propose_diag_f64",
    "lslb_iwls_propose_f32", "lslb_iwls_propose_f64", "lslb_mh_step_f32", "lslb_mh_step_f64", "lslb_spline_basis_f32",
    "lslb_spline_basis_f64", "lslb_spline_basis_order_f32", "lslb_spline_basis_order_f64",
]


class LieselB200Error(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.lslb_status_string(status).decode()
        if status in (3, 5):
            msg += ": " + lib.lslb_last_cuda_error().decode()
        raise LieselB200Error(f"{what} failed with status {status} ({msg})")


def sfx(dtype: torch.dtype) -> str:
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"unsupported dtype {dtype}")


def ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LieselB200Error("liesel_b200 operates on CUDA tensors only (no CPU fallback)")
        if t is not None and not t.is_contiguous():
            raise LieselB200Error("tensors passed to the C ABI must be contiguous")


def spline_basis(x: torch.Tensor, knots: torch.Tensor, order: int = 3):
    """Banded B-spline basis of order 0..3 on the device: (start int32 [n], weights [n, 4]); for
    order < 3 the window is zero-padded (see ``lslb_spline_basis_order_*``)."""
    require_cuda(x, knots)
    if x.dtype != knots.dtype:
        raise TypeError("x and knots must share a dtype")
    n = x.numel()
    start = torch.empty(n, dtype=torch.int32, device=x.device)
    w = torch.empty((n, 4), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        fn = getattr(lib, "lslb_spline_basis_order_" + sfx(x.dtype))
        check(fn(ptr(x), n, ptr(knots), knots.numel(), int(order), ptr(start), ptr(w), stream_ptr()),
              "lslb_spline_basis_order")
    return start, w
