#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
KF='regex:msmooth|finalize|gather_small|banded_x2|banded_kernel|transpose|expand|reduce'
(timeout 300 ncu -k "$KF" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2d_s2_launches.csv python tools/bench_configs.py --configs s2 --reps 2 > gpurun_out/r2d_s2_ncu.log 2>&1)
(timeout 300 ncu -k "$KF" --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/r2d_sweep_launches.csv python tools/bench_configs.py --chain-sweep 64,128,256 --reps 2 > gpurun_out/r2d_sweep_ncu.log 2>&1)
tail -2 gpurun_out/r2d_s2_ncu.log; tail -2 gpurun_out/r2d_sweep_ncu.log
