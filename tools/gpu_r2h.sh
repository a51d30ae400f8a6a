#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "not nuts and not hmc and not sampler" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log)
(timeout 300 python tools/bench_configs.py --configs s2,s3 > gpurun_out/r2h_cfg.json 2>&1)
KF='regex:iwls|chol|gather_small|finalize'
(timeout 300 ncu -k "$KF" --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2h_s3_launches.csv python tools/bench_configs.py --configs s3 --reps 2 > gpurun_out/r2h_s3_ncu.log 2>&1)
tail -4 gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_cfg.json
