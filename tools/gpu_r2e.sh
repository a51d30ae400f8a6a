#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q --timeout=600 -x -k "logp_grad or chain or full_size or s2 or families or reproducible" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log)
(timeout 200 python tools/bench_configs.py --configs s2,s3,h > gpurun_out/r2e_cfg.json 2>&1)
(timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256 > gpurun_out/r2e_sweep.json 2>&1)
KF='regex:msmooth|finalize|gather_small'
(timeout 300 ncu -k "$KF" --metrics gpu__time_duration.sum --clock-control none -c 9 --csv --log-file gpurun_out/r2e_s2_launches.csv python tools/bench_configs.py --configs s2 --reps 2 > gpurun_out/r2e_s2_ncu.log 2>&1)
SECONDS=0
(timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$? wall=${SECONDS}s" >> gpurun_out/r2e_bench.err)
tail -3 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_cfg.json gpurun_out/r2e_sweep.json; tail -2 gpurun_out/r2e_bench.err
