#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "not nuts and not hmc and not sampler and not async" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log)
tail -3 gpurun_out/r2q_pytest.log
(timeout 300 python tools/bench_configs.py --configs s3,h > gpurun_out/r2q_cfg.json 2>&1)
(timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256,512,1024 > gpurun_out/r2q_sweep.json 2>&1)
cat gpurun_out/r2q_cfg.json gpurun_out/r2q_sweep.json
