#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout=600 -x -k "iwls" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log)
(timeout 200 python tools/bench_configs.py --configs s3 > gpurun_out/r2f_cfg.json 2>&1)
(LSLB_NO_IWLS_PAIRS=1 timeout 200 python tools/bench_configs.py --configs s3 > gpurun_out/r2f_cfg_nopairs.json 2>&1)
(timeout 300 ncu -k "regex:msmooth|finalize_all|gather_small" --set full --clock-control none -c 3 -o gpurun_out/r2f_s2_full -f python tools/bench_configs.py --configs s2 --reps 2 > gpurun_out/r2f_s2_ncu.log 2>&1)
tail -3 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_cfg.json gpurun_out/r2f_cfg_nopairs.json; tail -2 gpurun_out/r2f_s2_ncu.log
