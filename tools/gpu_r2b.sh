#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log)
(timeout 300 python tools/ess_diag.py --chains 256 --warmup 300 --samples 100 > gpurun_out/r2b_diag256.json 2> gpurun_out/r2b_diag256.err)
(timeout 400 python tools/ess_diag.py --chains 1024 --warmup 300 --samples 100 > gpurun_out/r2b_diag1024.json 2> gpurun_out/r2b_diag1024.err)
tail -5 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_diag256.json; tail -3 gpurun_out/r2b_diag256.err
(timeout 120 python tools/bench_configs.py --configs s2 --reps 10 > gpurun_out/r2b_s2.json 2>&1)
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2b_s2_launches.csv python tools/bench_configs.py --configs s2 --reps 3 > gpurun_out/r2b_s2_ncu.log 2>&1)
cat gpurun_out/r2b_s2.json
