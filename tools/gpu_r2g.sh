#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log)
(timeout 300 python tools/bench_configs.py > gpurun_out/r2g_cfg.json 2>&1)
(timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256,1024 > gpurun_out/r2g_sweep.json 2>&1)
KF='regex:msmooth|finalize|gather_small|banded_x2'
(timeout 300 ncu -k "$KF" --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2g_s2_launches.csv python tools/bench_configs.py --configs s2 --reps 2 > gpurun_out/r2g_s2_ncu.log 2>&1)
(timeout 300 ncu -k "$KF" --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2g_h_launches.csv python tools/bench_configs.py --configs h --reps 2 > gpurun_out/r2g_h_ncu.log 2>&1)
tail -4 gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_cfg.json gpurun_out/r2g_sweep.json
