#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout=600 -x -k "not nuts and not hmc and not sampler and not async" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log)
tail -3 gpurun_out/r2t_pytest.log
(timeout 120 python tools/bench_configs.py --chain-sweep 64,128,256,512,1024 --reps 20 > gpurun_out/r2t_sweep.json 2>&1)
(timeout 300 python bench.py --steps 20 --warmup 5 --no-ess --no-secondary --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err)
cat gpurun_out/r2t_sweep.json; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"
