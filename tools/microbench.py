"""Measures the pipe peaks the rooflines are quoted against (run under gpurun):
FP32 FFMA, packed FFMA2, FFMA2+FFMA interleaved, MUFU.EX2 and shared-memory loads."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from liesel_b200 import _lib  # noqa: E402

sm = torch.cuda.get_device_properties(0).multi_processor_count
sink = torch.zeros(4, device="cuda")
out = {"gpu": torch.cuda.get_device_name(0), "sm_count": sm}
for kind, name, iters, scale in [(0, "ffma_tflops", 4000, 2e-12), (1, "ffma2_tflops", 4000, 2e-12),
                                 (4, "ffma2_plus_ffma_tflops", 4000, 2e-12), (2, "mufu_ex2_tops", 2000, 1e-12),
                                 (3, "lds32_tops", 4000, 1e-12),
                                 (5, "ffma2_3distinct_tflops", 4000, 2e-12), (6, "ffma2_shared_operand_tflops", 4000, 2e-12),
                                 (7, "ffma_3distinct_tflops", 4000, 2e-12), (8, "ffma_shared_operand_tflops", 4000, 2e-12)]:
    best = None
    work = C.c_double(0.0)
    for rep in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _lib.check(_lib.bench_lib().lslb_debug_pipe_peak(kind, sm * 8, iters, _lib.ptr(sink), C.byref(work),
                                                 _lib.stream_ptr()), "pipe_peak")
        e.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, s.elapsed_time(e)) if best else s.elapsed_time(e)
    out[name] = work.value / (best * 1e-3) * scale
print(json.dumps(out))
