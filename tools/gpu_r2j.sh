#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "not nuts and not hmc and not sampler" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log)
(timeout 300 python tools/bench_configs.py --configs s2,s3,h > gpurun_out/r2j_cfg.json 2>&1)
(timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256 > gpurun_out/r2j_sweep.json 2>&1)
(timeout 300 ncu -k "regex:iwls|chol" --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file gpurun_out/r2j_s3_iwls_launches.csv python tools/bench_configs.py --configs s3 --reps 2 > gpurun_out/r2j_s3_ncu.log 2>&1)
(timeout 300 ncu -k "regex:banded_x2" --set full --clock-control none -c 1 -o gpurun_out/r2j_x2_c64 -f python tools/bench_configs.py --chain-sweep 64 --reps 2 > gpurun_out/r2j_c64_ncu.log 2>&1)
(timeout 300 ncu -k "regex:banded_iwls" --set full --clock-control none -c 1 -o gpurun_out/r2j_iwls_s3 -f python tools/bench_configs.py --configs s3 --reps 2 > gpurun_out/r2j_iwls_ncu.log 2>&1)
tail -3 gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_cfg.json gpurun_out/r2j_sweep.json
