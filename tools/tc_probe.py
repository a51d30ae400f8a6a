"""Numerical check of the tcgen05 building block (lslb_debug_tc_probe) against torch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from liesel_b200 import _lib

torch.manual_seed(0)
for N, K in [(128, 32), (128, 64), (64, 256), (256, 256)]:
    A = torch.randn(128, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    ref = (A.double() @ B.double().T)
    for split in (0, 1, 2, 3, 4):  # bit 0: 3-pass hi/lo split; bit 1: 16-float chunks (64-byte swizzle)
        D = torch.zeros(128, N, device="cuda")
        st = _lib.bench_lib().lslb_debug_tc_probe(_lib.ptr(A), _lib.ptr(B), N, K, split, _lib.ptr(D), _lib.stream_ptr())
        torch.cuda.synchronize()
        err = (D.double() - ref).abs().max().item()
        scale = ref.abs().max().item()
        print(f"N={N} K={K} split={split} status={st} max_abs_err={err:.3e} rel={err/scale:.3e}", flush=True)
