#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout=600 -x -k "logp_grad or chain or full_size or nuts_fused or compaction" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log)
(timeout 120 python tools/bench_configs.py --chain-sweep 48,64,96,128,192,256,512,1024 > gpurun_out/r2c_sweep.json 2>&1)
(timeout 300 ncu -k regex:lslb --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_s2_launches.csv python tools/bench_configs.py --configs s2 --reps 3 > gpurun_out/r2c_s2_ncu.log 2>&1)
(timeout 400 python tools/ess_diag.py --chains 1024 --warmup 500 --term 250 --samples 150 > gpurun_out/r2c_diag1024.json 2> gpurun_out/r2c_diag1024.err)
tail -3 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_sweep.json; cat gpurun_out/r2c_diag1024.json; tail -3 gpurun_out/r2c_diag1024.err
