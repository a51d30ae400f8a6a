"""Drives the jax.ffi shim's handlers through the TEST STUB call frame (tests/stubs/xla/ffi/api/ffi.h;
``lslb_stub_call`` in liesel_b200/csrc/xla_ffi_shim.cc). Test infrastructure only."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB_SO = os.path.join(ROOT, "liesel_b200", "csrc", "build", "liblslb_xla_ffi_stub.so")
HANDLERS = ["LslbLogpGrad", "LslbLogpGradF64", "LslbLogp", "LslbLogpF64", "LslbIwls", "LslbIwlsF64",
            "LslbQuadform", "LslbQuadformF64", "LslbIwlsGroup", "LslbIwlsGroupF64"]
F32, F64 = 11, 12
OK, INVALID_ARGUMENT, INTERNAL = 0, 3, 13


def load():
    lib = C.CDLL(STUB_SO)
    lib.lslb_stub_call.restype = C.c_int
    lib.lslb_stub_call.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int64), C.c_char_p, C.c_int]
    return lib


def call(lib, target, plan_handle, args, rets, block=0, stream=None):
    """``args`` / ``rets``: lists of (dtype_code, data_ptr or None, shape). Returns (code, message)."""
    bufs = list(args) + list(rets)
    n = len(bufs)
    dtypes = (C.c_int * n)(*[b[0] for b in bufs])
    datas = (C.c_void_p * n)(*[b[1] for b in bufs])
    ranks = (C.c_int * n)(*[len(b[2]) for b in bufs])
    flat = [int(d) for b in bufs for d in b[2]]
    dims = (C.c_int64 * max(len(flat), 1))(*flat)
    err = C.create_string_buffer(256)
    fn = C.cast(getattr(lib, target), C.c_void_p)
    code = lib.lslb_stub_call(fn, stream, int(plan_handle or 0), int(block), len(args), len(rets), dtypes, datas,
                              ranks, dims, err, 256)
    return code, err.value.decode()


def tensor_buf(t, shape=None):
    import torch

    code = F32 if t.dtype == torch.float32 else F64
    return (code, t.data_ptr(), tuple(t.shape) if shape is None else tuple(shape))
