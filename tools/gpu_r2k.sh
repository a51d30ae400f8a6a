#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "iwls" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log)
(timeout 300 python tools/bench_configs.py --configs s2,s3 > gpurun_out/r2k_cfg.json 2>&1)
tail -3 gpurun_out/r2k_pytest.log; cat gpurun_out/r2k_cfg.json
