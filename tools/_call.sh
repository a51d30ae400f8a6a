set -u
mkdir -p gpurun_out
SW=64,128,256,512,1024
timeout 120 python tools/bench_configs.py --chain-sweep $SW > gpurun_out/ph_sweep_phased.json 2>&1
LSLB_X2_NO_PHASES=1 timeout 120 python tools/bench_configs.py --chain-sweep $SW > gpurun_out/ph_sweep_old.json 2>&1
LSLB_NO_SHARED_BAND=1 timeout 120 python tools/bench_configs.py --chain-sweep $SW > gpurun_out/ph_sweep_phased_nsl2.json 2>&1
timeout 120 python tools/bench_configs.py --configs s3 > gpurun_out/ph_s3.json 2>&1
LSLB_IW_NO_PHASES=1 timeout 120 python tools/bench_configs.py --configs s3 > gpurun_out/ph_s3_iwold.json 2>&1
python - <<'PY'
import json
for f in ["phased","old","phased_nsl2"]:
    try:
        d=json.loads(open(f"gpurun_out/ph_sweep_{f}.json").read().strip().splitlines()[-1])
        print(f, [(r["chains"], round(r["logp_grad_ms"],4)) for r in d["sweep"]])
    except Exception as e: print(f, "ERR", e, open(f"gpurun_out/ph_sweep_{f}.json").read()[-600:])
PY
tail -1 gpurun_out/ph_s3.json; tail -1 gpurun_out/ph_s3_iwold.json
(timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/ph_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ph_pytest.log)
tail -5 gpurun_out/ph_pytest.log
