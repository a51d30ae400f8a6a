"""Investigation tool: per-CTA timeline of banded_x2_kernel on the H workload. Needs a trace build:

  NVCC_EXTRA=-DLSLB_X2_TRACE bash liesel_b200/csrc/build.sh   (into a scratch copy of the library)
  LSLB_LIB_PATH=<that .so> python tools/x2_trace.py
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from liesel_b200 import _lib, synth  # noqa: E402
from liesel_b200.plan import Plan  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spec, centre = synth.s3_gamlss(n=1_000_000)
plan = Plan(spec)
th, aux = synth.evaluation_points(spec, centre, chains)
t, a = torch.as_tensor(th, device="cuda"), torch.as_tensor(aux, device="cuda")
for _ in range(3):
    plan.logp_grad(t, a)
torch.cuda.synchronize()
n = 8192
buf = (C.c_ulonglong * (n * 6))()
_lib.lib.lslb_debug_x2_trace.argtypes = [C.c_void_p, C.c_int]
assert _lib.lib.lslb_debug_x2_trace(buf, n) == 0
tr = np.frombuffer(buf, dtype=np.uint64).reshape(n, 6).astype(np.int64)
tr = tr[tr[:, 2] > 0]
t0 = tr[:, 1].min()
start, end = (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3
dur = end - start
tiles, nonuni = tr[:, 5] >> 32, tr[:, 5] & 0xFFFFFFFF
per_sm = {}
for sm, e in zip(tr[:, 0], end):
    per_sm[int(sm)] = max(per_sm.get(int(sm), 0.0), float(e))
sm_end = np.array(sorted(per_sm.values()))
out = {
    "chains": chains, "ctas": int(len(tr)), "sms_used": len(per_sm),
    "kernel_span_us": float(end.max()),
    "cta_start_us": {"min": float(start.min()), "p50": float(np.median(start)), "max": float(start.max())},
    "cta_dur_us": {"min": float(dur.min()), "p10": float(np.percentile(dur, 10)), "p50": float(np.median(dur)),
                   "p90": float(np.percentile(dur, 90)), "max": float(dur.max())},
    "sm_finish_us": {"min": float(sm_end.min()), "p10": float(np.percentile(sm_end, 10)),
                     "p50": float(np.median(sm_end)), "p90": float(np.percentile(sm_end, 90)), "max": float(sm_end.max())},
    "ctas_per_sm": {"min": int(np.bincount(tr[:, 0]).min()), "max": int(np.bincount(tr[:, 0]).max())},
    "tiles_per_cta": {"min": int(tiles.min()), "max": int(tiles.max())},
    "nonuniform_tiles_per_cta": {"min": int(nonuni.min()), "p50": float(np.median(nonuni)), "max": int(nonuni.max())},
    "corr_dur_nonuni": float(np.corrcoef(dur, nonuni)[0, 1]) if nonuni.std() > 0 else None,
    "us_per_tile_p50": float(np.median(dur / np.maximum(tiles, 1))),
}
# duration by number of co-resident CTAs is not observable directly; report duration by start rank
order = np.argsort(start)
out["dur_first_quarter_us"] = float(np.median(dur[order[: len(order) // 4]]))
out["dur_last_quarter_us"] = float(np.median(dur[order[-len(order) // 4:]]))
print(json.dumps(out))
